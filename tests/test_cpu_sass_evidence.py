"""The shipped library contains what DESIGN.md says it contains: disassemble libparllel_b200.so (cuobjdump, no GPU
needed) and look for the SASS mnemonics of the Blackwell features the kernels are built on."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "parllel_b200", "libparllel_b200.so")


@pytest.fixture(scope="module")
def sass():
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    if not os.path.exists(LIB):
        from parllel_b200 import build
        build.build()
    out = subprocess.run([tool, "-sass", LIB], capture_output=True, text=True, timeout=600).stdout
    per_kernel, name = {}, None
    for line in out.splitlines():
        if "Function :" in line:
            name = line.split("Function :")[1].strip()
            per_kernel[name] = []
        elif name is not None and "/*" in line:
            per_kernel[name].append(line)
    assert per_kernel, "no device code found in the library"
    return per_kernel


def _kernels(sass, fragment):
    return {k: v for k, v in sass.items() if fragment in k}


def _has(lines, mnemonic):
    return any(mnemonic in ln for ln in lines)


def test_built_for_sm_100a_only():
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool) or not os.path.exists(LIB):
        pytest.skip("cuobjdump or the library not available")
    out = subprocess.run([tool, "-lelf", LIB], capture_output=True, text=True, timeout=120).stdout
    archs = {ln.split("sm_")[1].split(".")[0] for ln in out.splitlines() if "sm_" in ln}
    assert archs == {"100a"}, archs


def test_scans_use_tma_loads_stores_and_mbarriers(sass):
    scans = _kernels(sass, "scan_pipelined_kernel")
    assert scans
    for name, lines in scans.items():
        assert _has(lines, "UTMALDG"), f"{name}: no TMA tensor load"
        assert _has(lines, "UTMASTG"), f"{name}: no TMA tensor store"
        assert _has(lines, "SYNCS"), f"{name}: no mbarrier operations"


def test_streaming_kernels_take_part_in_programmatic_dependent_launch(sass):
    """griddepcontrol.launch_dependents = PREEXIT, griddepcontrol.wait = ACQBULK."""
    for fragment in ("scan_pipelined_kernel", "normalize_advantage_partials_kernel", "normalize_advantage_xch_kernel",
                     "replay_indices_cluster_kernel", "replay_indices_kernel"):
        kernels = _kernels(sass, fragment)
        assert kernels, fragment
        for name, lines in kernels.items():
            assert _has(lines, "PREEXIT") and _has(lines, "ACQBULK"), name


def test_resident_kernel_parks_the_advantage_in_tensor_memory(sass):
    """tcgen05.st / tcgen05.ld = STTM / LDTM; only the RESIDENT instantiations (last template argument 1) have them."""
    resident = {k: v for k, v in _kernels(sass, "scan_pipelined_kernel").items() if k.endswith("Lb1EEEvNS0_10TensorMapsENS0_6ParamsE")}
    assert resident
    for name, lines in resident.items():
        assert _has(lines, "STTM") and _has(lines, "LDTM"), name


def test_gather_stages_rows_with_bulk_copies_and_index_generation_is_a_cluster_kernel(sass):
    gather = _kernels(sass, "gather_tree_kernel")
    assert gather and all(_has(lines, "UBLKCP") for lines in gather.values())
    cluster = _kernels(sass, "replay_indices_cluster_kernel")
    assert cluster and all(_has(lines, "UCGABAR") for lines in cluster.values())

"""Write a text summary of an .ncu-rep (raw metrics subset + hottest SASS lines) for profiles/.
    python tools/ncu_report.py gpurun_out/X.ncu-rep profiles/X.txt"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg", "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__shared_mem_per_block_dynamic"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
lines = [f"# ncu summary of {rep}", ""]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    lines.append(f"## kernel: {name}")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            lines.append(f"{w:75s} {r[i]:>16s} {units[i]}")
    lines.append("")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
hidx = [i for i, r in enumerate(srows) if r and r[0] == "Address"]
if hidx:
    h = srows[hidx[0]]
    end = hidx[1] - 1 if len(hidx) > 1 else len(srows)
    body = [r for r in srows[hidx[0] + 1:end] if len(r) >= len(h)]
    col = {n: i for i, n in enumerate(h)}
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    tot = {s: sum(int(r[col[s]]) for r in body) for s in stalls}
    total = sum(int(r[col["# Samples"]]) for r in body)
    instr = sum(int(r[col["Instructions Executed"]]) for r in body)
    ops = {}
    for r in body:
        s = r[col["Source"]].split()
        op = (s[1] if s[0].startswith("@") else s[0]).split(".")[0]
        ops[op] = ops.get(op, 0) + int(r[col["Instructions Executed"]])
    lines += [f"## source page (first captured launch): {total} samples, {instr} warp-instructions",
              "stall samples: " + ", ".join(f"{k[6:]}={v}" for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]),
              "opcode mix:    " + ", ".join(f"{k}={v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:16]),
              "hottest instructions (samples, executions, SASS, stall breakdown):"]
    for r in sorted(body, key=lambda r: -int(r[col["# Samples"]]))[:15]:
        st = {s[6:]: int(r[col[s]]) for s in stalls if int(r[col[s]]) > 0}
        lines.append(f"  {r[col['# Samples']]:>5s} {r[col['Instructions Executed']]:>8s}  {r[col['Source']][:70]:70s} {st}")
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)

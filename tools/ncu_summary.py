"""Summarise an ncu source-page CSV: stall reasons, opcode mix, hottest instructions.
    ncu -i X.ncu-rep --page source --csv --print-source sass > src.csv; python tools/ncu_summary.py src.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hdr_idx[0]]
end = hdr_idx[1] - 1 if len(hdr_idx) > 1 else len(rows)
body = [r for r in rows[hdr_idx[0] + 1:end] if len(r) >= len(h)]
col = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
tot = {s: 0 for s in stalls}
total = instr = 0
ops = {}
for r in body:
    ns = int(r[col["# Samples"]]); total += ns
    ie = int(r[col["Instructions Executed"]]); instr += ie
    src = r[col["Source"]].split()
    op = (src[1] if src[0].startswith("@") else src[0]).split(".")[0]
    d = ops.setdefault(op, [0, 0]); d[0] += ie; d[1] += ns
    for s in stalls:
        tot[s] += int(r[col[s]])
print("samples", total, "warp-instructions", instr)
print("stalls:", ", ".join(f"{s[6:]}={v}" for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]))
print("opcodes:", ", ".join(f"{op}={ie}" for op, (ie, ns) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:18]))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 14
for r in sorted(body, key=lambda r: -int(r[col["# Samples"]]))[:n]:
    st = {s[6:]: int(r[col[s]]) for s in stalls if int(r[col[s]]) > 0}
    print(r[col["# Samples"]], r[col["Instructions Executed"]], r[col["Source"]][:64], st)

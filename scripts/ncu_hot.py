#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from an exported `ncu --page source --csv` capture."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
ci = {n: i for i, n in enumerate(h)}
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
data = []
tot = 0
for r in rows[hi + 1:]:
    if len(r) < len(h):
        continue
    try:
        s = int(r[ci["# Samples"]])
    except ValueError:
        continue
    tot += s
    st = {n: int(r[ci[n]] or 0) for n in stall_cols}
    data.append((s, r[ci["Address"]], r[ci["Source"]], st, int(r[ci["Instructions Executed"]] or 0)))
print("total samples", tot, "instructions", len(data))
agg = {}
for s, a, src, st, ie in data:
    for k, v in st.items():
        agg[k] = agg.get(k, 0) + v
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]})
# by opcode
byop = {}
for s, a, src, st, ie in data:
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    e = byop.setdefault(op, [0, 0])
    e[0] += s
    e[1] += ie
print("by opcode (samples, executed):", sorted(((v[0], k, v[1]) for k, v in byop.items()), reverse=True)[:16])
for s, a, src, st, ie in sorted(data, reverse=True)[:top]:
    main = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{s:7d} {100*s/tot:5.2f}% exec={ie:9d} {a[-5:]} {src[:60]:60s} {main}")

"""Summarise `ncu --page source --csv` output (SASS view): stall reasons, hottest opcodes, hottest instructions.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > src.csv ; python profiles/ncu_source_summary.py src.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if len(r) > 5 and r[0] == "Address"]
for k, start in enumerate(hi[:1]):
    h = rows[start]
    end = hi[k + 1] - 1 if k + 1 < len(hi) else len(rows)
    data = [r for r in rows[start + 1:end] if len(r) == len(h) and r[0] != "Address"]
    si, ci, ie = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    f = lambda x: float(x or 0)  # noqa: E731
    tot = sum(f(r[ci]) for r in data)
    print(rows[start - 1][:2], "samples", tot, "instructions", len(data))
    agg = collections.Counter({h[c]: sum(f(r[c]) for r in data) for c in stall_cols})
    for name, v in agg.most_common(9):
        print(f"  {name:26s} {v / tot * 100:5.1f}%")
    byop, cnt = collections.Counter(), collections.Counter()
    for r in data:
        t = r[si].split()
        op = t[1] if t[0].startswith("@") else t[0]
        byop[op] += f(r[ci])
        cnt[op] += f(r[ie])
    print("  -- by opcode")
    for name, v in byop.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 18):
        print(f"  {name:28s} samples {v / tot * 100:5.1f}%   warp-instr executed {cnt[name]:.3g}")
    print("  -- hottest instructions")
    for r in sorted(data, key=lambda r: -f(r[ci]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 16]:
        st = sorted([(f(r[c]), h[c]) for c in stall_cols], reverse=True)[:2]
        print(f"  {f(r[ci]) / tot * 100:5.1f}% {r[si][:64]:64s} {st[0][1]}:{st[0][0]:.0f} {st[1][1]}:{st[1][0]:.0f}")

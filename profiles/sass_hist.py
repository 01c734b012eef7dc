"""Opcode histogram + hottest stall lines from `ncu -i rep --page source --csv`:
   ncu -i x.ncu-rep --page source --csv > /tmp/src.csv; python profiles/sass_hist.py /tmp/src.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
H = rows[1]
si, ci, st = H.index("Source"), H.index("Instructions Executed"), H.index("Warp Stall Sampling (All Samples)")
agg, tot_st = collections.Counter(), 0
lines = []
for r in rows[2:]:
    try:
        n = int(r[ci])
    except (ValueError, IndexError):
        continue
    toks = r[si].split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "")
    agg[op.split(".")[0]] += n
    lines.append((int(r[st] or 0), n, r[si].strip()))
    tot_st += int(r[st] or 0)
tot = sum(agg.values())
print("warp instructions:", tot)
for k, v in agg.most_common(22):
    print(f"  {k:12s} {v:12d} {100 * v / tot:5.1f}%")
print("top stall lines (samples, executed, sass):")
for s, n, src in sorted(lines, reverse=True)[:18]:
    print(f"  {100 * s / max(tot_st, 1):5.1f}% {n:10d}  {src[:100]}")

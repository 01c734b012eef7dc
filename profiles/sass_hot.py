"""Per-region stall samples of a kernel from `ncu -i rep --page source --csv --print-source sass`:
python profiles/sass_hot.py <csv> [n_regions]  -- splits the SASS into equal-count address regions and prints samples,
executed instructions and the top opcodes / stall reasons per region, plus the hottest single instructions."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr_i]
R = [r for r in rows[hdr_i + 1:] if len(r) == len(H)]
ci = {h: i for i, h in enumerate(H)}
S, X = ci["# Samples"], ci["Instructions Executed"]
stalls = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[S] or 0) for r in R)
totx = sum(int(r[X] or 0) for r in R)
print(f"instructions {len(R)}  samples {tot}  executed warp-instructions {totx}")
nreg = int(sys.argv[2]) if len(sys.argv) > 2 else 12
per = (len(R) + nreg - 1) // nreg
for k in range(nreg):
    seg = R[k * per:(k + 1) * per]
    if not seg: break
    s = sum(int(r[S] or 0) for r in seg); x = sum(int(r[X] or 0) for r in seg)
    ops = collections.Counter()
    st = collections.Counter()
    for r in seg:
        op = r[ci["Source"]].split()[0] if r[ci["Source"]].split() else "?"
        if op.startswith("@"): op = r[ci["Source"]].split()[1]
        ops[op.split(".")[0]] += int(r[X] or 0)
        for h in stalls:
            st[h[6:]] += int(r[ci[h]] or 0)
    print(f"region {k:2d} [{k*per:5d}..): samples {100*s/tot:5.1f}%  exec {100*x/totx:5.1f}%  top ops {ops.most_common(4)}  stalls {st.most_common(3)}")
print("hottest instructions:")
for r in sorted(R, key=lambda r: -int(r[S] or 0))[:14]:
    print(f"  {R.index(r):5d} {int(r[S]):6d} ({100*int(r[S])/tot:4.1f}%)  {r[ci['Source']][:90]}")

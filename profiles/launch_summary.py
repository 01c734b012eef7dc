"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel:
   python profiles/launch_summary.py gpurun_out/launches.csv [top]"""
import collections
import csv
import sys


def main(path, top=16):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[ui], 1.0)
        agg[r[ki][:84]][0] += 1
        agg[r[ki][:84]][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'total ms':>10} {'calls':>6} {'avg us':>10} {'share':>6}  kernel")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f"{t / 1e6:10.3f} {n:6d} {t / n / 1e3:10.2f} {100 * t / tot:5.1f}%  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 16)

"""DRAM bytes per launch of one kernel from an `ncu --set full` report -> small JSON that bench.py reads for
`roofline.traffic` (so the number in the bench line names the capture it came from instead of being a constant).

    python profiles/extract_traffic.py <report.ncu-rep> <kernel-substring> <out.json> "<command that produced it>"
"""
import csv
import json
import subprocess
import sys


def main(rep, kernel, out, command):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    H, U = rows[0], rows[1]
    hit = [r for r in rows[2:] if kernel in r[H.index("Kernel Name")]]
    assert hit, f"no launch of {kernel} in {rep}"
    r = hit[0]

    def val(name):
        i = H.index(name)
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[U[i]]
        return float(r[i]) * scale
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    json.dump({"kernel": r[H.index("Kernel Name")][:120], "dram_bytes_read": rd, "dram_bytes_write": wr,
               "dram_bytes": rd + wr, "launches_in_report": len(hit), "report": rep, "command": command,
               "extracted_at_commit": head,
               "duration_ms_under_ncu": float(r[H.index("gpu__time_duration.sum")]) *
               {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[U[H.index("gpu__time_duration.sum")]]},
              open(out, "w"), indent=1)
    print(open(out).read())


if __name__ == "__main__":
    main(*sys.argv[1:5])

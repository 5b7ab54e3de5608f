#!/usr/bin/env python
"""Writes profiles/ncu_traffic.json -- the measured DRAM traffic bench.py reports in `roofline.traffic` -- from committed
`ncu --set full` captures (no GPU needed):

    python tools/ncu_traffic.py oned=gpurun_out/r2f_svat.ncu-rep:16777216 sas=gpurun_out/r2g_sas.ncu-rep:131072x1500

`name=report:cells` takes dram__bytes_read.sum + dram__bytes_write.sum of the LAST captured launch divided by the cells of that
launch (`cells x bins` for the SAS kernel: bytes per cell and age bin)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def dram_bytes(report):
    out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, last = rows[0], rows[1], rows[-1]
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(k)
        tot += float(last[i]) * UNIT[units[i]]
    return tot, last[hdr.index("Kernel Name")], float(last[hdr.index("gpu__time_duration.sum")]), units[hdr.index("gpu__time_duration.sum")]


def main():
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    rec = json.load(open(path)) if os.path.exists(path) else {}
    for arg in sys.argv[1:]:
        name, rest = arg.split("=")
        report, size = rest.rsplit(":", 1)
        b, kernel, dur, dur_unit = dram_bytes(report)
        entry = {"kernel": kernel.split("(")[0].strip(), "capture": os.path.basename(report), "dram_bytes_per_launch": b,
                 "duration": f"{dur} {dur_unit} (under ncu: cold caches, serialised)"}
        if "x" in size:
            cells, bins = (int(v) for v in size.split("x"))
            entry.update(cells=cells, bins=bins, dram_bytes_per_bin_day=b / cells / bins)
        else:
            entry.update(cells=int(size), dram_bytes_per_cell=b / int(size))
        entry["source"] = f"profiles/ncu_traffic.json <- ncu --set full capture {os.path.basename(report)} (dram__bytes_read.sum + dram__bytes_write.sum of one launch / its cells)"
        rec[name] = entry
    json.dump(rec, open(path, "w"), indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()

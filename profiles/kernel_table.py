#!/usr/bin/env python
"""Summarise the ncu launch logs of profiles/kernel_table.sh: per kernel launches, time, DRAM bytes, achieved GB/s and
the fraction of the measured HBM peak (MEASURED_PEAKS.json).  python profiles/kernel_table.py gpurun_out/kt_*.csv"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(paths):
    peak = 6551.4
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    for path in paths:
        rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
        hdr = rows[0]
        ix = {h: i for i, h in enumerate(hdr)}
        agg = {}
        for r in rows[1:]:
            name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("pgr::", "")
            a = agg.setdefault(name, {"n": set(), "ns": 0.0, "rd": 0.0, "wr": 0.0})
            a["n"].add(r[ix["ID"]])
            v = float(r[ix["Metric Value"]].replace(",", ""))
            unit = r[ix["Metric Unit"]]
            m = r[ix["Metric Name"]]
            if m == "gpu__time_duration.sum":
                a["ns"] += v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
            else:
                b = v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
                a["rd" if "read" in m else "wr"] += b
        print("== %s   (HBM peak %.0f GB/s)" % (os.path.basename(path), peak))
        print("%-46s %7s %10s %10s %10s %9s %7s" % ("kernel", "launch", "ms total", "read MB", "write MB", "GB/s", "% peak"))
        tot = sum(a["ns"] for a in agg.values())
        for name, a in sorted(agg.items(), key=lambda t: -t[1]["ns"]):
            gbs = (a["rd"] + a["wr"]) / max(a["ns"], 1.0)
            print("%-46s %7d %10.3f %10.1f %10.1f %9.1f %6.1f%%   (%4.1f %% of the kernel time)" % (name[:46], len(a["n"]), a["ns"] / 1e6, a["rd"] / 1e6, a["wr"] / 1e6, gbs,
                                                                                     100 * gbs / peak, 100 * a["ns"] / tot))
        print()


if __name__ == "__main__":
    main(sys.argv[1:])

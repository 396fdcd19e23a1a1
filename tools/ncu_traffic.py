#!/usr/bin/env python
"""DRAM traffic of one kernel launch from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,
gpu__time_duration.sum --csv` log -> the small JSON bench.py reads for roofline.traffic.
usage: ncu_traffic.py log.csv workload-label exp_mode [ntracks-of-the-slice] > profiles/rNN_K1_dram_traffic.json
With a fourth argument the capture was taken on a slice of the workload's 2D tracks (same slabs, fewer tracks): bench.py
scales the traffic by tracks (scale_to_full)."""
import csv
import json
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1], errors="replace")))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[hdr]
    ki, mi, ui, vi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6,
             "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "nsecond": 1e-9, "s": 1, "second": 1}
    out = {}
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        out.setdefault(r[ki], {})[r[mi]] = float(r[vi].replace(",", "")) * scale.get(r[ui], 1)
    name, m = next(iter(out.items()))
    extra = {"ntracks": int(sys.argv[4]), "scale_to_full": True} if len(sys.argv) > 4 else {}
    print(json.dumps({"kernel": name, "workload": sys.argv[2], "exp": sys.argv[3], **extra,
                      "dram_bytes_read": m["dram__bytes_read.sum"], "dram_bytes_write": m["dram__bytes_write.sum"],
                      "traffic": m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"],
                      "gpu_time_s_under_ncu": m.get("gpu__time_duration.sum"),
                      "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, one "
                                + ("launch on a slice of the workload's 2D tracks" if extra else "full-size launch of bench.py's workload")},
                     indent=1))


if __name__ == "__main__":
    main()

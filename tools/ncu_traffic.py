#!/usr/bin/env python
"""Writes profiles/traffic_ncu.json: DRAM bytes per k_intersect_closest launch
(dram__bytes_read.sum + dram__bytes_write.sum, averaged over the captured launches) from
`ncu --set full` reports, with the provenance of each capture.  bench.py prints these as
roofline.traffic (+ traffic_source): a profiler number from a separate run of the same
build, not a measurement of the bench run itself.
usage: python tools/ncu_traffic.py <tag> workload=report.ncu-rep[:note] ..."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def launches(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        rec = {"kernel": r[hdr.index("Kernel Name")]}
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
                    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
                    "smsp__thread_inst_executed_per_inst_executed.ratio",
                    "sm__warps_active.avg.pct_of_peak_sustained_active",
                    "smsp__issue_active.avg.pct_of_peak_sustained_active",
                    "smsp__sass_inst_executed_op_local_ld.sum", "sass__inst_executed_local_loads"):
            if key in hdr:
                i = hdr.index(key)
                v = float(r[i].replace(",", ""))
                if "bytes" in key:
                    v *= UNIT.get(units[i], 1.0)
                elif key == "gpu__time_duration.sum":
                    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[i], 1.0)
                rec[key] = v
        out.append(rec)
    return out


def main():
    tag = sys.argv[1]
    path = os.path.join(ROOT, "profiles", "traffic_ncu.json")
    table = json.load(open(path)) if os.path.exists(path) else {}
    for arg in sys.argv[2:]:
        workload, rest = arg.split("=", 1)
        rep, _, note = rest.partition(":")
        ls = launches(rep)
        n = len(ls)
        total = sum(l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"] for l in ls)
        table[workload] = {
            "dram_bytes_per_launch": total / n, "launches_captured": n,
            "per_launch": [{"dram_bytes": l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"],
                            "ms": l.get("gpu__time_duration.sum"),
                            "l2_hit_pct": l.get("lts__t_sector_hit_rate.pct"),
                            "l1_hit_pct": l.get("l1tex__t_sector_hit_rate.pct"),
                            "lanes_per_inst": l.get(
                                "smsp__thread_inst_executed_per_inst_executed.ratio")}
                           for l in ls],
            "source": "ncu --set full --clock-control none, report %s (%s)"
                      % (os.path.basename(rep), tag),
            "note": note or "same build as the bench line of this tag; the capture renders one "
                            "batch of the workload, all of its closest-hit launches averaged",
        }
        print(workload, json.dumps(table[workload])[:400])
    with open(path, "w") as f:
        json.dump(table, f, indent=1)


if __name__ == "__main__":
    main()

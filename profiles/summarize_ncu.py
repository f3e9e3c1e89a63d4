#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep + launches.csv into the small text summaries kept
under profiles/ (the .ncu-rep files themselves are scratch).
usage: python profiles/summarize_ncu.py <tag>   # e.g. r01a"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_bytes.sum", "l1tex__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
    "local_load_bytes", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
]


def launches(tag):
    path = os.path.join(OUT, "launches.csv")
    if not os.path.exists(path):
        return
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(ROOT, "profiles", "%s_launch_list.txt" % tag), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised:\n"
                "# compare SHARES, not absolutes)\n")
        for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("%-44s launches=%4d total_us=%10.1f share=%.3f avg_us=%.1f\n"
                    % (n[:44], a[0], a[1] / 1e3, a[1] / tot, a[1] / a[0] / 1e3))


def report(tag, rep):
    path = os.path.join(OUT, rep)
    if not os.path.exists(path):
        return
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    name = rep.replace(".ncu-rep", "")
    with open(os.path.join(ROOT, "profiles", "%s_%s.txt" % (tag, name)), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on, %s\n" % rep)
        ki = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
        for w in hdr:
            if w in WANT or w.startswith("smsp__average_warps_issue_stalled") or w.startswith("sass__inst_executed_local"):
                i = hdr.index(w)
                f.write("%-78s %s\n" % (w, [r[i] for r in rows[1:]]))
        if ki is not None:
            f.write("kernels: %s\n" % [r[ki][:60] for r in rows[2:]])
    if name == "prof_closest":
        # bench.py reports this as roofline.traffic (DRAM bytes per launch of the dominant kernel)
        import json

        def col(metric):
            i = hdr.index(metric)
            unit = rows[1][i]
            mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
            return [float(r[i].replace(",", "")) * mul for r in rows[2:]]
        rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
        per = [a + b for a, b in zip(rd, wr)]
        json.dump({"kernel": "k_intersect_closest", "launches_captured": len(per),
                   "dram_bytes_per_launch": sum(per) / len(per), "per_launch": per,
                   "workload": "terrain_1002k 1920x1080, 16 spp batch (33.2 M camera rays, then the "
                               "bounce rays), ncu --set full --clock-control none",
                   "source": "profiles/%s_prof_closest.txt" % tag},
                  open(os.path.join(ROOT, "profiles", "traffic_intersect_closest.json"), "w"), indent=1)


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "rXX"
    launches(tag)
    for rep in sorted(os.listdir(OUT)):
        if rep.endswith(".ncu-rep"):
            report(tag, rep)

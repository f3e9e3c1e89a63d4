#!/usr/bin/env python
"""Per-source-line cost of one kernel from an ncu report captured with --import-source on
(-lineinfo build): warp instructions executed and stall samples per CUDA source line,
summed per file and listed for the hottest lines.
usage: python tools/srclines.py <report.ncu-rep> [kernel-index] [top-n]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
k = -1
cur_file = None
hdr = None
per_line = collections.OrderedDict()
per_file = collections.Counter()
per_file_s = collections.Counter()
kname = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        if r[1] != kname:
            kname = r[1]
            k += 1
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if k != kidx or hdr is None or len(r) != len(hdr):
        continue
    if r[0] == "":
        continue  # SASS row
    iS, iE = hdr.index("# Samples"), hdr.index("Instructions Executed")
    try:
        s, e = int(r[iS]), int(r[iE])
    except ValueError:
        continue
    key = (cur_file, int(r[0]), r[1].strip()[:90])
    a = per_line.setdefault(key, [0, 0])
    a[0] += s
    a[1] += e
    per_file[cur_file] += e
    per_file_s[cur_file] += s
tot_e = sum(per_file.values()) or 1
tot_s = sum(per_file_s.values()) or 1
print("kernel", kidx, kname, "warp-instr", tot_e, "samples", tot_s)
for f, e in per_file.most_common():
    print("  %-24s instr %5.1f%%  samples %5.1f%%" % (f, 100.0 * e / tot_e, 100.0 * per_file_s[f] / tot_s))
print("hottest lines by samples:")
for (f, ln, src), (s, e) in sorted(per_line.items(), key=lambda x: -x[1][0])[:topn]:
    print("  %5.2f%% smp %5.2f%% ins  %s:%d  %s" % (100.0 * s / tot_s, 100.0 * e / tot_e, f, ln, src))

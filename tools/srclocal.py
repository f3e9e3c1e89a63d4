#!/usr/bin/env python
"""Which source lines execute local-memory (LDL/STL) and shared (LDS/STS) instructions:
from an ncu report with --import-source on.  usage: srclocal.py report [kernel-index]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
k = -1; cur_file = None; hdr = None; kname = None; cur_line = None
agg = collections.defaultdict(lambda: collections.Counter())
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        if r[1] != kname: kname = r[1]; k += 1
        continue
    if r[0] == "Line No": hdr = r; continue
    if k != kidx or hdr is None or len(r) != len(hdr): continue
    iE = hdr.index("Instructions Executed")
    if r[0] != "":
        cur_line = (cur_file, int(r[0]), r[1].strip()[:80]); continue
    sass = r[3].strip()
    op = sass.split()[1] if sass.startswith("@") else sass.split()[0]
    for pre in ("LDL", "STL", "LDS", "STS", "LDG", "STG", "LD.", "ST."):
        if op.startswith(pre):
            try: agg[pre][cur_line] += int(r[iE])
            except ValueError: pass
for pre in ("LDL", "STL"):
    tot = sum(agg[pre].values())
    print(pre, "total warp-instr", tot)
    for line, n in agg[pre].most_common(18):
        print("   %5.1f%%  %s:%d  %s" % (100.0 * n / max(tot, 1), line[0], line[1], line[2]))

#!/usr/bin/env python
"""Per-region hot spots of one kernel from an ncu report (--import-source on):
program-order chunks with sample share, issued-instruction share and lane use."""
import csv, subprocess, sys, io
rep = sys.argv[1]; chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None; data = []; k = 0
for r in rows:
    if r and r[0] == 'Kernel Name':
        k += 1; continue
    if r and r[0] == 'Address':
        hdr = r; continue
    if k == 1 and hdr and len(r) == len(hdr): data.append(r)
ia = hdr.index('Source'); iss = hdr.index('# Samples'); ie = hdr.index('Instructions Executed'); it = hdr.index('Avg. Threads Executed')
tot = sum(int(r[iss]) for r in data); totinst = sum(int(r[ie]) for r in data)
print('instr rows', len(data), 'samples', tot, 'warp-instr', totinst)
for c in range(0, len(data), chunk):
    seg = data[c:c + chunk]
    s = sum(int(r[iss]) for r in seg); e = sum(int(r[ie]) for r in seg)
    thr = sum(float(r[it]) * int(r[ie]) for r in seg) / max(e, 1)
    ops = [r[ia].split()[0] if not r[ia].strip().startswith('@') else r[ia].split()[1] for r in seg]
    mem = [o for o in ops if o.startswith(('LDG', 'LDL', 'STL', 'STG', 'LDS', 'LDGSTS', 'ATOM', 'BAR', 'CALL', 'VOTE', 'WARPSYNC'))]
    print('%4d-%4d samples %5.1f%% instr %5.1f%% avgthr %4.1f  %s' % (c, c + chunk, 100 * s / tot, 100 * e / totinst, thr, ' '.join(mem)[:100]))

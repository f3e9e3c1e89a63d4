import sys
sys.path.insert(0, '.')
import numpy as np
from oracle import cycles_ref as cr
from raytracingproject_b200 import scenes
from raytracingproject_b200.device import B200Device
desc = scenes.instanced(width=3840, height=2160)
rs = cr.build_scene(desc, kernel=0)
dev = B200Device(0)
dev.upload_scene(rs.device_arrays())
rays, _ = rs.camera_rays(1, 0, 0, 3840, 1092)
ref = rs.intersect(rays); got = dev.trace_batch(rays)
same = (ref["prim"] == got["prim"]) & (ref["object"] == got["object"])
rel = np.abs(ref["t"] - got["t"]) / np.maximum(np.abs(ref["t"]), 1e-30)
hard = ~same & ~((ref["prim"] >= 0) & (got["prim"] >= 0) & (rel < 1e-5))
idx = np.nonzero(hard)[0]
print('hard', idx)
for i in idx:
    print('ray', rays[i]); print(' ref', ref[i]); print(' got', got[i], 'rel dt', rel[i])
    # re-trace with tmax just beyond each candidate to see whether the other side can see it
    for cand in (ref[i], got[i]):
        r2 = rays[i:i+1].copy(); r2['t'] = cand['t'] * (1 + 1e-4)
        print('  limit t to', r2['t'][0], '-> ref', rs.intersect(r2)[0], ' gpu', dev.trace_batch(r2)[0])

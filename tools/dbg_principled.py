import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
np.set_printoptions(precision=6, suppress=True, linewidth=200)
from oracle import cycles_ref as cr
from raytracingproject_b200 import scenes
from raytracingproject_b200.device import B200Device
dev = B200Device(0)
W, H = 128, 72
d = scenes.cornell(W, H, materials="glass", max_bounce=3)
rs = cr.build_scene(d)
dev.upload_scene(rs.device_arrays())
names = "P.x P.y P.z t D.x D.y D.z hit_t prim obj sP.x sP.y sP.z N.x N.y N.z flag ncl c0type c0sw c1type c1sw label pdf wi.x wi.y wi.z thr.x thr.y thr.z u v".split()
for (x, y) in [(69, 0), (67, 0)]:
    # slot of pixel (x,y) for sample 0 in the 8x4 tiling
    t = (y // 4) * (W // 8) + (x // 8); l = (y % 4) * 8 + (x % 8); slot = t * 32 + l
    dev.set_option("debug_slot", slot)
    got_img = dev.render(W, H, rs.pass_stride, 0, 1)
    g = dev.debug_read(); r = rs.path_dump(0, x, y)
    print('pixel', x, y, 'slot', slot, 'gpu px', got_img[y, x, :3])
    for b in range(4):
        print(' bounce', b + 1)
        for k in range(32):
            if abs(g[b, k] - r[b, k]) > 1e-4 * max(1, abs(r[b, k])):
                print('    %-6s ref %14.7g  gpu %14.7g' % (names[k], r[b, k], g[b, k]))
        print('    ref:', r[b, [8, 16, 17, 18, 19, 20, 21, 22, 23]], 'gpu:', g[b, [8, 16, 17, 18, 19, 20, 21, 22, 23]])

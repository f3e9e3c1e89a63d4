"""Where do mesh-light renders differ from the reference?  1 spp, max_bounce 1..N,
per-pixel diff statistics and path dumps of the worst pixels (run on the GPU box)."""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
np.set_printoptions(precision=7, suppress=True, linewidth=200)
from oracle import cycles_ref as cr
from raytracingproject_b200 import scenes
from raytracingproject_b200.device import B200Device
dev = B200Device(0)
W, H = 128, 72
light = sys.argv[1] if len(sys.argv) > 1 else "mesh"
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 1
d = scenes.cornell(W, H, materials="diffuse", max_bounce=mb, light=light)
rs = cr.build_scene(d)
dev.upload_scene(rs.device_arrays())
names = "P.x P.y P.z t D.x D.y D.z hit_t prim obj sP.x sP.y sP.z N.x N.y N.z flag ncl c0type c0sw c1type c1sw label pdf wi.x wi.y wi.z thr.x thr.y thr.z u v".split()
for s in range(3):
    ref, _ = rs.render(s, 1, tile_size=0)
    got = dev.render(W, H, rs.pass_stride, s, 1)
    diff = np.abs(ref[..., :3] - got[..., :3]).max(axis=-1)
    rel = diff / np.maximum(np.abs(ref[..., :3]).max(axis=-1), 1e-6)
    print('sample', s, 'pixels differing', int((diff > 0).sum()), 'of', W * H, ' rel>1e-5:', int((rel > 1e-5).sum()),
          ' rel>1e-3:', int((rel > 1e-3).sum()), 'max abs', diff.max())
    if s == 0:
        ys, xs = np.where(rel > 1e-5)
        order = np.argsort(-rel[ys, xs])[:4]
        for i in order:
            x, y = int(xs[i]), int(ys[i])
            t = (y // 4) * (W // 8) + (x // 8); l = (y % 4) * 8 + (x % 8); slot = t * 32 + l
            dev.set_option("debug_slot", slot)
            g_img = dev.render(W, H, rs.pass_stride, 0, 1)
            g = dev.debug_read(); r = rs.path_dump(0, x, y)
            print('pixel', x, y, 'ref', ref[y, x, :3], 'gpu', got[y, x, :3])
            for b in range(mb + 2):
                bad = [k for k in range(32) if abs(g[b, k] - r[b, k]) > 1e-6 * max(1, abs(r[b, k]))]
                print('  bounce', b + 1, 'prim ref/gpu', r[b, 8], g[b, 8], 'diff fields:',
                      [(names[k], float(r[b, k]), float(g[b, k])) for k in bad][:6])

import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from oracle import cycles_ref as cr
from raytracingproject_b200 import scenes
from raytracingproject_b200.device import B200Device
dev = B200Device(0)
W, H = 128, 72
base = scenes.cornell(W, H, materials="procedural", max_bounce=0)
def run(label, xml):
    d = scenes.SceneDesc(base.name, xml, W, H, meshes=base.meshes, objects=base.objects, spp=1)
    rs = cr.build_scene(d)
    dev.upload_scene(rs.device_arrays())
    want, _ = rs.render(0, 1, tile_size=0)
    got = dev.render(W, H, rs.pass_stride, 0, 1)
    a, b = want[..., :3], got[..., :3]
    bad = np.argwhere(np.any(a != b, axis=-1))
    print('%-14s rmse %.3e  max %.3e  identical %.4f' % (label, np.sqrt(np.mean((a - b) ** 2)), np.abs(a - b).max(), np.all(want == got, axis=-1).mean()),
          'first bad', (tuple(bad[0]), a[tuple(bad[0])], b[tuple(bad[0])]) if len(bad) else None)
    rs.close()
x = base.xml
for node, sock in [("m1", "value"), ("m2", "value"), ("m3", "value"), ("vm", "vector"), ("vm", "value"), ("vl", "value"), ("m4", "value"),
                   ("cl", "result"), ("cmb", "vector"), ("mx", "color"), ("lw", "facing"), ("mx2", "color"), ("mx3", "color"),
                   ("bc", "color"), ("gm", "color"), ("inv", "color"), ("fr", "fac"), ("m5", "value")]:
    y = x.replace('  <connect from="mc closure" to="output surface"/>',
                  '  <emission name="dbg_e" strength="1"/>\n  <connect from="%s %s" to="dbg_e color"/>\n'
                  '  <connect from="dbg_e emission" to="output surface"/>' % (node, sock))
    run(node + '.' + sock, y)

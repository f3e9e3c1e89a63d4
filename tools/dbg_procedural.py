import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from oracle import cycles_ref as cr
from raytracingproject_b200 import scenes
from raytracingproject_b200.device import B200Device
dev = B200Device(0)
W, H = 128, 72
base = scenes.cornell(W, H, materials="procedural", max_bounce=int(sys.argv[1]) if len(sys.argv) > 1 else 0)
def run(label, xml):
    d = scenes.SceneDesc(base.name, xml, W, H, meshes=base.meshes, objects=base.objects, spp=4)
    rs = cr.build_scene(d)
    dev.upload_scene(rs.device_arrays())
    want, _ = rs.render(0, 4, tile_size=0)
    got = dev.render(W, H, rs.pass_stride, 0, 4)
    a, b = want[..., :3] / 4, got[..., :3] / 4
    print('%-28s rmse %.3e  max %.3e  identical %.3f' % (label, np.sqrt(np.mean((a - b) ** 2)), np.abs(a - b).max(), np.all(want == got, axis=-1).mean()))
    rs.close()
x = base.xml
run('full', x)
run('no roughness link', x.replace('  <connect from="m5 value" to="gl roughness"/>\n', ''))
run('no gl color link', x.replace('  <connect from="cl result" to="gl color"/>\n', ''))
run('neither', x.replace('  <connect from="cl result" to="gl color"/>\n', '').replace('  <connect from="m5 value" to="gl roughness"/>\n', ''))
run('diffuse only', x.replace('  <connect from="mc closure" to="output surface"/>', '  <connect from="d bsdf" to="output surface"/>'))
run('glossy only', x.replace('  <connect from="mc closure" to="output surface"/>', '  <connect from="gl bsdf" to="output surface"/>'))
run('const fac', x.replace('  <connect from="fr fac" to="mc fac"/>\n', ''))

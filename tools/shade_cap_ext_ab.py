"""Register cap of the FULL shading kernels (k_shade_surface<true, true>): scenes that run
them (PMJ pattern, world AO) with the library named by B200_CYCLES_LIB."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raytracingproject_b200 import scenes
from raytracingproject_b200.device import B200Device
from oracle import cycles_ref as ref

W, H, SPP = 1920, 1080, 64
MS = "Multiscatter GGX"
cube = scenes.default_cube(W, H, material="principled", distribution=MS)
cube.xml = cube.xml.replace('sampling_pattern="sobol"', 'sampling_pattern="pmj"')
assert 'sampling_pattern="pmj"' in cube.xml
cases = {
    "cube_principled pmj": cube,
    "cornell_principled pmj": scenes.cornell(W, H, materials="principled", distribution=MS,
                                             pattern="pmj"),
    "cornell_principled ao": scenes.cornell(W, H, materials="principled", distribution=MS,
                                            ao=(0.3, 5.0)),
    "cornell_textured": scenes.cornell(W, H, materials="textured3"),
}
dev = B200Device(0)
if os.environ.get("SHADE_WIDE"):
    dev.set_option("shade_wide", int(os.environ["SHADE_WIDE"]))
for name, desc in cases.items():
    rs = ref.build_scene(desc)
    dev.upload_scene(rs.device_arrays(), rs.textures())
    best = 1e30
    for it in range(3):
        dev.render(W, H, rs.pass_stride, 0, SPP)
        d = dev.stats()
        best = min(best, d["device_ms"])
    rays = d["primary_rays"] + d["bounce_rays"] + d["shadow_rays"]
    print("%-28s device_ms %.1f  closest %.1f shade %.1f shadow %.1f  Mrays/s %.0f  ext=%d" % (
        name, best, d["closest_ms"], d["shade_ms"], d["shadow_ms"], rays / best / 1e3,
        d["svm_extended"]) + " dense=%d" % d.get("shade_wide", -9))
    rs.close()

"""Which scenes gain from a 3-blocks/SM register cap of the lean multiscatter shading
kernel?  Renders a few scenes with the library named by B200_CYCLES_LIB, prints device ms."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raytracingproject_b200 import scenes
from raytracingproject_b200.device import B200Device
from oracle import cycles_ref as ref

W, H, SPP = 1920, 1080, 64
MS = os.environ.get("DIST", "Multiscatter GGX")
cases = {
    "cube_principled": scenes.default_cube(W, H, material="principled", distribution=MS),
    "cube_principled shallow (4 bounces)": scenes.default_cube(W, H, material="principled",
                                                               distribution=MS, max_bounce=4),
    "cornell_principled(metal+glass)": scenes.cornell(W, H, materials="principled",
                                                      distribution=MS),
    "cornell_metal(no glass)": scenes.cornell(W, H, materials="metal", distribution=MS),
    "cornell_glass(no metal)": scenes.cornell(W, H, materials="glass", distribution=MS),
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
        s = dev.stats()
        d = s
        best = min(best, d["device_ms"])
    rays = d["primary_rays"] + d["bounce_rays"] + d["shadow_rays"]
    print("%-34s device_ms %.1f  closest %.1f shade %.1f shadow %.1f  Mrays/s %.0f  ext=%d dense=%d" % (
        name, best, d["closest_ms"], d["shade_ms"], d["shadow_ms"], rays / best / 1e3,
        d["svm_extended"], d.get("shade_wide", -9)))
    rs.close()

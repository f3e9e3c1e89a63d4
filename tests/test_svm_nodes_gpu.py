"""Value nodes of the SVM interpreter (svm_nodes.cuh) against the reference kernels.

node_chart: one emissive quad per operator group, the emitted colour is the node
output at the hit point; camera rays only, so both sides evaluate the nodes on
bit-identical inputs.  +, -, *, /, sqrt, floor ... must agree exactly; libm
transcendentals (sin, pow, exp ...) differ by a few ulp between glibc and CUDA."""
import numpy as np
import pytest

from raytracingproject_b200 import scenes
from test_render_gpu import image_gates

pytestmark = pytest.mark.gpu


def test_value_node_chart_matches_reference(ref, device):
    desc = scenes.node_chart()
    rs = ref.build_scene(desc)
    try:
        device.upload_scene(rs.device_arrays())
        want, _ = rs.render(0, 1, tile_size=0)
        got = device.render(desc.width, desc.height, rs.pass_stride, 0, 1)
    finally:
        rs.close()
    lit = want[..., :3].max(axis=-1) != 0
    assert lit.mean() > 0.5
    same = np.all(want == got, axis=-1)
    print("node chart: bit-identical pixels %.4f" % same.mean())
    assert same.mean() > 0.9
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=1e-6)


def test_procedural_material_image(ref, device):
    """Cornell box whose two boxes carry graphs of value nodes feeding a diffuse +
    glossy-GGX mix (8 bounces): the usual image gates."""
    desc = scenes.cornell(256, 144, materials="procedural")
    rs = ref.build_scene(desc)
    try:
        device.upload_scene(rs.device_arrays())
        want, _ = rs.render(0, 16, tile_size=64)
        got = device.render(desc.width, desc.height, rs.pass_stride, 0, 16)
        image_gates(want, got, 16, "cornell procedural")
    finally:
        rs.close()


def test_unsupported_node_is_refused(ref, device):
    """A node outside the supported subset (Bump: it needs ray differentials) is refused when its program
    is bound - never skipped or approximated."""
    from raytracingproject_b200.device import DeviceError
    desc = scenes.cornell(64, 36, materials="diffuse")
    desc.xml = desc.xml.replace(
        '  <diffuse_bsdf name="d" color="0.73 0.73 0.73"/>\n',
        '  <diffuse_bsdf name="d"/>\n  <geometry name="g"/>\n'
        '  <vector_math name="l" type="length"/>\n'
        '  <connect from="g position" to="l vector1"/>\n'
        '  <math name="t" type="multiply_add" value2="100" value3="450"/>\n'
        '  <connect from="l value" to="t value1"/>\n'
        '  <bump name="m" strength="0.6"/>\n'
        '  <connect from="t value" to="m height"/>\n'
        '  <connect from="m normal" to="d normal"/>\n', 1)
    assert "<bump" in desc.xml
    rs = ref.build_scene(desc)
    try:
        arrays = rs.device_arrays()
        with pytest.raises(DeviceError) as e:      # refused when the program is bound
            device.upload_scene(arrays)
            device.render(desc.width, desc.height, rs.pass_stride, 0, 1)
        assert "SVM node opcode" in str(e.value)
    finally:
        rs.close()

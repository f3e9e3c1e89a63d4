"""Edge cases of DeviceTask::RENDER through the C ABI: ragged frames, degenerate
tiles, empty sample ranges, a film that already holds samples, and the refusal of
out-of-scope scenes (no silent approximation)."""
import os
import re

import numpy as np
import pytest

from raytracingproject_b200 import scenes
from raytracingproject_b200.device import DeviceError, DeviceMemory
from test_render_gpu import image_gates

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kd_offsets():
    text = open(os.path.join(ROOT, "include", "cycles_abi.h")).read()
    return {k: int(v) for k, v in re.findall(r"#define (KD_\w+)[ \t]+(\d+)", text)}


def test_ragged_frame_matches_reference(ref, device):
    """101 x 37: neither a multiple of the 8x4 pixel tile nor of the warp size (the
    row-major fallback of batch_pixel and partly filled warps everywhere)."""
    desc = scenes.cornell(width=101, height=37, spp=8, materials="diffuse")
    rs = ref.build_scene(desc)
    try:
        device.upload_scene(rs.device_arrays())
        ref_img, _ = rs.render(0, desc.spp, tile_size=64)
        got = device.render(desc.width, desc.height, rs.pass_stride, 0, desc.spp)
        image_gates(ref_img, got, desc.spp, "ragged cornell")

        # the same frame as odd tiles, one of them a single pixel
        w, h, ps = desc.width, desc.height, rs.pass_stride
        film = DeviceMemory("RenderBuffers", np.zeros((h, w, ps), np.float32))
        device.mem_zero(film)
        cuts_x, cuts_y = [0, 1, 34, w], [0, 5, h]
        for y0, y1 in zip(cuts_y[:-1], cuts_y[1:]):
            for x0, x1 in zip(cuts_x[:-1], cuts_x[1:]):
                device.render_tile(film.device_pointer, x0, y0, x1 - x0, y1 - y0, 0, desc.spp, 0, w)
        device.mem_copy_from(film)
        device.mem_free(film)
        assert np.array_equal(film.host, got)
    finally:
        rs.close()


def test_empty_work_leaves_the_film_alone(ref, device):
    desc = scenes.default_cube(width=64, height=32, spp=2)
    rs = ref.build_scene(desc)
    try:
        device.upload_scene(rs.device_arrays())
        w, h, ps = desc.width, desc.height, rs.pass_stride
        marker = np.full((h, w, ps), 3.25, np.float32)
        film = DeviceMemory("RenderBuffers", marker.copy())
        device.mem_alloc(film)
        device.mem_copy_to(film)
        device.render_tile(film.device_pointer, 0, 0, w, h, 0, 0, 0, w)   # no samples
        device.render_tile(film.device_pointer, 0, 0, 0, h, 0, 4, 0, w)   # no columns
        device.render_tile(film.device_pointer, 0, 0, w, 0, 0, 4, 0, w)   # no rows
        device.mem_copy_from(film)
        assert np.array_equal(film.host, marker)

        # RENDER accumulates into what the film already holds (kernel_write_pass_float4 is +=)
        device.render_tile(film.device_pointer, 0, 0, w, h, 0, 2, 0, w)
        device.mem_copy_from(film)
        fresh = device.render(w, h, ps, 0, 2)
        np.testing.assert_allclose(film.host[..., :4], fresh[..., :4] + 3.25, rtol=1e-6)
        device.mem_free(film)
    finally:
        rs.close()


@pytest.mark.parametrize("field,value,needle", [
    ("KD_INT_USE_VOLUMES", 1, "volumes"),
    ("KD_INT_BRANCHED", 1, "branched"),
    ("KD_BVH_HAVE_CURVES", 1, "curves"),
    ("KD_FILM_PASS_DENOISING_DATA", 4, "denoising data passes do not fit"),
    ("KD_FILM_PASS_DENOISING_CLEAN", 4, "clean pass needs"),
    ("KD_FILM_CRYPTOMATTE_PASSES", 1, "cryptomatte"),
    ("KD_BG_NUM_PORTALS", 1, "portals"),
    ("KD_INT_MAX_CLOSURES", 33, "closures per shader"),
])
def test_out_of_scope_scenes_are_refused(ref, device, field, value, needle):
    """KernelData asking for a feature outside SURVEY.md 8 is refused with
    B200_ERR_UNSUPPORTED when the scene is prepared - never rendered approximately."""
    desc = scenes.default_cube(width=32, height=16, spp=1)
    rs = ref.build_scene(desc)
    try:
        arrays = dict(rs.device_arrays())
        data = arrays["__data"][0].copy()
        data[kd_offsets()[field]:kd_offsets()[field] + 4] = np.frombuffer(
            np.int32(value).tobytes(), np.uint8)
        arrays["__data"] = (data, 1)
        device.upload_scene(arrays)
        with pytest.raises(DeviceError) as e:
            device.render(desc.width, desc.height, rs.pass_stride, 0, 1)
        assert needle in str(e.value)
        # the device recovers once a supported scene is bound again
        device.upload_scene(rs.device_arrays())
        device.render(desc.width, desc.height, rs.pass_stride, 0, 1)
    finally:
        rs.close()

"""FILM_CONVERT (kernel_film.h:90-130) and the in-process film reduce against the
reference: the SAME float film goes through the reference CPU kernels and through
b200_film_convert.

 - half output: integer bit manipulation of one float product -> bit-exact;
 - byte output: sRGB uses powf, CUDA and glibc differ by <= 1-2 ulp, so a byte may
   flip at a rounding boundary: |diff| <= 1 and fewer than 0.1 % of the bytes."""
import numpy as np
import pytest

from raytracingproject_b200 import scenes
from raytracingproject_b200.device import B200Device, B200HostDevice, DeviceMemory

pytestmark = pytest.mark.gpu

CASES = {
    "cornell": lambda: scenes.cornell(width=160, height=120, spp=8, materials="diffuse"),
    "cube": lambda: scenes.default_cube(width=128, height=72, spp=4),
}


def _scene_on_device(ref, device, desc):
    rs = ref.build_scene(desc)
    device.upload_scene(rs.device_arrays())
    device.build_bvh()
    return rs


@pytest.mark.parametrize("name", sorted(CASES))
def test_film_convert_matches_reference(ref, device, name):
    desc = CASES[name]()
    rs = _scene_on_device(ref, device, desc)
    w, h, n = desc.width, desc.height, desc.spp
    film, _ = rs.render(0, n, tile_size=0)               # reference film (float sums)
    want_b = rs.film_convert(n)                          # reference CPU kernels
    want_h = rs.film_convert(n, half_float=True)

    mem = DeviceMemory("RenderBuffers", np.ascontiguousarray(film))
    device.mem_alloc(mem)
    device.mem_copy_to(mem)
    try:
        got_b = device.film_convert(mem, w, h, n)
        got_h = device.film_convert(mem, w, h, n, half_float=True)
    finally:
        device.mem_free(mem)
    rs.close()

    assert got_h.dtype == np.uint16 and np.array_equal(got_h, want_h), "half pixels differ"
    d = np.abs(got_b.astype(np.int16) - want_b.astype(np.int16))
    assert d.max() <= 1
    assert (d != 0).mean() < 1e-3
    assert want_b[..., :3].max() > 0  # not a black frame


def test_film_convert_through_device_task(ref, device):
    """The reference's DeviceTask::FILM_CONVERT on the C++ B200Device equals the C-ABI
    call on the same film."""
    desc = CASES["cornell"]()
    host = B200HostDevice(0)
    rs = ref.build_scene(desc, external_device=host.ptr)
    film, _ = rs.render(0, desc.spp, tile_size=0)
    via_task_b = rs.film_convert(desc.spp)
    via_task_h = rs.film_convert(desc.spp, half_float=True)
    rs.close()
    host.close()

    rs2 = _scene_on_device(ref, device, desc)
    mem = DeviceMemory("RenderBuffers", np.ascontiguousarray(film))
    device.mem_alloc(mem)
    device.mem_copy_to(mem)
    try:
        assert np.array_equal(device.film_convert(mem, desc.width, desc.height, desc.spp), via_task_b)
        assert np.array_equal(
            device.film_convert(mem, desc.width, desc.height, desc.spp, half_float=True), via_task_h)
    finally:
        device.mem_free(mem)
    rs2.close()


def test_film_reduce_two_contexts(ref, device):
    """b200_film_reduce (replaces MultiDevice::mem_copy_from slicing,
    device_multi.cpp:374-393): two sample ranges rendered by two contexts sum to the
    film one context renders for the whole range.  Uses two GPUs when the box has them,
    else two contexts on the same GPU (the peer copy degenerates to a device copy)."""
    import torch
    desc = CASES["cornell"]()
    rs = ref.build_scene(desc)
    arrays = rs.device_arrays()
    w, h, ps, n = desc.width, desc.height, rs.pass_stride, desc.spp
    second = B200Device(1 if torch.cuda.device_count() > 1 else 0)
    try:
        for d in (device, second):
            d.upload_scene(arrays)
            d.build_bvh()
        whole = device.render(w, h, ps, 0, n).copy()
        films = []
        for d, (s0, cnt) in zip((device, second), ((0, n // 2), (n // 2, n - n // 2))):
            m = DeviceMemory("RenderBuffers", np.zeros((h, w, ps), np.float32))
            d.mem_zero(m)
            d.render_tile(m.device_pointer, 0, 0, w, h, s0, cnt, 0, w)
            films.append(m)
        B200Device.film_reduce([device, second], films, h * w * ps)
        device.mem_copy_from(films[0])
        # per-pixel sums are in sample order in both cases; the split adds (a) + (b)
        # instead of one running sum, so agreement is to rounding, not bitwise
        np.testing.assert_allclose(films[0].host, whole, rtol=2e-6, atol=1e-6)
        device.mem_free(films[0])
        second.mem_free(films[1])
    finally:
        second.close()
        rs.close()

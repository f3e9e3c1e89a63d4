"""The drop-in path end to end: the reference's own Scene::device_update and
DeviceTask::RENDER (acquire_tile / release_tile callbacks) drive the C++
`B200Device : ccl::Device`; the film read back through RenderBuffers must equal
the film rendered through the Python mirror of the same C ABI, and match the
reference CPU device."""
import numpy as np
import pytest

from scene_cases import small_cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["cornell", "instanced"])
def test_reference_scene_drives_b200_device(ref, device, name):
    from raytracingproject_b200.device import B200HostDevice
    desc = small_cases()[name]
    spp = 8
    cpu = ref.build_scene(desc)
    ref_img, _ = cpu.render(0, spp, tile_size=64)
    device.upload_scene(cpu.device_arrays())
    py_img = device.render(desc.width, desc.height, cpu.pass_stride, 0, spp).copy()
    cpu.close()

    host = B200HostDevice(0)
    try:
        rs = ref.build_scene(desc, external_device=host.ptr)
        try:
            full, _ = rs.render(0, spp, tile_size=0)        # one full-frame tile
            tiled, _ = rs.render(0, spp, tile_size=64)      # Session-style 64x64 tiles
            assert host.error_message() == ""
            assert np.array_equal(full, py_img), "C++ shim and Python mirror disagree"
            assert np.array_equal(tiled, py_img), "tiled RENDER differs from full frame"
            rmse = np.sqrt(np.mean((full[..., :3] / spp - ref_img[..., :3] / spp) ** 2))
            print(name, "shim vs CPU rmse", rmse, host.stats())
            assert rmse <= 1e-3
        finally:
            rs.close()
    finally:
        host.close()

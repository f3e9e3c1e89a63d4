"""The drop-in path end to end: the reference's own Scene::device_update and
DeviceTask::RENDER (acquire_tile / release_tile callbacks) drive the C++
`B200Device : ccl::Device`; the film read back through RenderBuffers must equal
the film rendered through the Python mirror of the same C ABI, and match the
reference CPU device."""
import os

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
            # the device asked the host for ITS layout: the reference's BVH::create built a
            # `BVH8 : BVH`, no packed BVH2 reached the device and nothing was built there
            info = host.bvh_info()
            packed, seconds, err = host.host_bvh8_report()
            print(name, "bvh", info, "host pack s", seconds)
            assert info["host_packed"] == 1 and info["build_ms"] == 0.0 and err == ""
            assert info["num_nodes"] == packed["num_nodes"] > 0
        finally:
            rs.close()
    finally:
        host.close()

    # the other route (a host that only knows BVH2): same film
    os.environ["B200_HOST_BVH"] = "bvh2"
    try:
        host = B200HostDevice(0)
        try:
            rs = ref.build_scene(desc, external_device=host.ptr)
            try:
                full2, _ = rs.render(0, spp, tile_size=0)
                info = host.bvh_info()
                assert info["host_packed"] == 0 and info["build_ms"] > 0.0
                assert np.array_equal(full2, py_img)
            finally:
                rs.close()
        finally:
            host.close()
    finally:
        del os.environ["B200_HOST_BVH"]


def test_device_comes_out_of_the_reference_registry(ref, device):
    """`cycles --device B200`, as far as it can be driven here: the reference's own
    Device::type_from_string / available_types / available_devices / Device::create
    (device/device.cpp:367-550, with the "B200" rows of the registration patch -
    oracle/device_registry_hook.sed) hand out the B200Device; a reference Scene rendered on
    it gives the film of the Python mirror."""
    from raytracingproject_b200.device import RegisteredDevice
    desc = small_cases()["cornell"]
    spp = 8
    cpu = ref.build_scene(desc)
    device.upload_scene(cpu.device_arrays())
    py_img = device.render(desc.width, desc.height, cpu.pass_stride, 0, spp).copy()
    cpu.close()

    reg = RegisteredDevice("B200", index=0)
    try:
        assert reg.type_name() == "B200" and reg.type_available()
        listed = reg.available()
        print("Device::available_devices(DEVICE_MASK_B200):", listed)
        assert listed and listed[0][0].startswith("B200_") and "B200" in listed[0][1]
        rs = ref.build_scene(desc, external_device=reg.ptr)
        try:
            img, _ = rs.render(0, spp, tile_size=64)
            st = reg.stats()
            assert st["kernel_launches"] > 0 and st["primary_rays"] > 0
            assert np.array_equal(img, py_img)
        finally:
            rs.close()
    finally:
        reg.close()


def test_registry_makes_one_device_of_a_b200_list(ref, device):
    """Device::get_multi_device of a B200-only list + Device::create (device.cpp:367-375,
    583-655, patched rows): not the generic MultiDevice (tiles fanned out, films sliced
    through the host) but ONE B200MultiDevice - samples split over the GPUs, films summed
    on the device.  Two GPUs when the box has them, else two contexts on one."""
    from raytracingproject_b200.device import RegisteredDevice
    desc = small_cases()["cornell"]
    spp = 8
    cpu = ref.build_scene(desc)
    device.upload_scene(cpu.device_arrays())
    py_img = device.render(desc.width, desc.height, cpu.pass_stride, 0, spp).copy()
    cpu.close()
    reg = RegisteredDevice("B200", index=0, count=2)
    try:
        rs = ref.build_scene(desc, external_device=reg.ptr)
        try:
            img, _ = rs.render(0, spp, tile_size=0)
            st = reg.stats()           # only a B200MultiDevice / B200Device answers
            assert st["primary_rays"] == desc.width * desc.height * spp
            # two partial films summed: equal up to the order of the float additions
            a, b = img[..., :4] / spp, py_img[..., :4] / spp
            assert np.abs(a - b).max() <= 1e-5 * max(1.0, np.abs(b).max())
        finally:
            rs.close()
    finally:
        reg.close()


def test_multi_device_in_one_process(ref):
    """B200MultiDevice: the reference Scene + DeviceTask drive several contexts through
    ONE ccl::Device; every GPU renders its share of the samples, the films are summed on
    the device (b200_film_reduce).  Equals the single-device film up to the order of the
    float additions; FILM_CONVERT of the summed film works.  Uses two GPUs when the box
    has them, else two contexts on one GPU."""
    import numpy as np
    import torch
    from raytracingproject_b200 import scenes
    from raytracingproject_b200.device import B200HostDevice
    desc = scenes.cornell(160, 96, spp=9, materials="diffuse")   # 9 samples: uneven split
    single = B200HostDevice(0)
    rs = ref.build_scene(desc, external_device=single.ptr)
    want, _ = rs.render(0, desc.spp, tile_size=0)
    rs.close()
    single.close()

    ordinals = [0, 1] if torch.cuda.device_count() > 1 else [0, 0]
    multi = B200HostDevice(ordinals)
    rs = ref.build_scene(desc, external_device=multi.ptr)
    got, _ = rs.render(0, desc.spp, tile_size=0)
    st = multi.stats()
    # a second task accumulates on top (the other GPUs' films were cleared after the sum)
    got2, _ = rs.render(desc.spp, 3, tile_size=64, accumulate=True)
    rgba = rs.film_convert(desc.spp + 3)
    rs.close()
    multi.close()

    np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-6)
    assert st["primary_rays"] == desc.width * desc.height * desc.spp
    assert (got2[..., 3] == desc.spp + 3).all()      # alpha counts every sample once
    assert rgba[..., :3].max() > 0


def test_two_contexts_with_different_scenes_on_one_gpu(ref):
    """The kernels read the scene from one __constant__ block per GPU.  Two contexts on
    the same GPU holding DIFFERENT scenes, rendering in turn (and from two threads at
    once): each frame must be the one its own scene gives alone."""
    import threading
    from raytracingproject_b200 import scenes
    from raytracingproject_b200.device import B200Device
    descs = [scenes.cornell(96, 64, spp=4, materials="diffuse"),
             scenes.default_cube(96, 64, spp=4, material="principled")]
    devs = [B200Device(0), B200Device(0)]
    rss = [ref.build_scene(d) for d in descs]
    try:
        alone = []
        for dev, rs, d in zip(devs, rss, descs):
            dev.upload_scene(rs.device_arrays())
            alone.append(dev.render(d.width, d.height, rs.pass_stride, 0, 4).copy())
        assert not np.array_equal(alone[0], alone[1])
        for _ in range(2):  # in turn: the other context used the GPU last each time
            for k in (0, 1):
                got = devs[k].render(descs[k].width, descs[k].height, rss[k].pass_stride, 0, 4)
                assert np.array_equal(got, alone[k]), k
        out = [None, None]

        def work(k):
            for _ in range(3):
                out[k] = devs[k].render(descs[k].width, descs[k].height, rss[k].pass_stride,
                                        0, 4).copy()
        threads = [threading.Thread(target=work, args=(k,)) for k in (0, 1)]
        [t.start() for t in threads]
        [t.join() for t in threads]
        assert np.array_equal(out[0], alone[0]) and np.array_equal(out[1], alone[1])
    finally:
        for rs in rss:
            rs.close()
        for dev in devs:
            dev.close()


def test_session_style_tile_buffers_are_freed_on_the_worker_thread(ref):
    """Background renders give every tile its own RenderBuffers and Session::release_tile
    deletes it on the device's worker thread (session.cpp:449-460, 517-518): mem_free runs
    INSIDE the running task and must not wait for the task pool.  The stitched tiles equal
    the full-frame render bit for bit; the same flow on the two-context multi device."""
    from raytracingproject_b200 import scenes
    from raytracingproject_b200.device import B200HostDevice
    desc = scenes.cornell(160, 96, spp=4, materials="diffuse")
    for ordinals in (0, [0, 0]):
        host = B200HostDevice(ordinals)
        try:
            rs = ref.build_scene(desc, external_device=host.ptr)
            try:
                full, _ = rs.render(0, desc.spp, tile_size=0)
                tiles, done = rs.render_tile_buffers(0, desc.spp, tile_size=64)
                assert host.error_message() == ""
                assert done == 3 * 2
                if ordinals == 0:
                    assert np.array_equal(tiles, full)
                else:
                    np.testing.assert_allclose(tiles, full, rtol=2e-6, atol=1e-6)
            finally:
                rs.close()
        finally:
            host.close()


def test_task_get_cancel_stops_the_render(ref):
    """task.get_cancel() (Session::cancel / progress.set_cancel) ends the tile loop: after
    two released tiles no further tile is completed, and the device stays usable."""
    from raytracingproject_b200 import scenes
    from raytracingproject_b200.device import B200HostDevice
    desc = scenes.cornell(192, 128, spp=2, materials="diffuse")
    host = B200HostDevice(0)
    try:
        rs = ref.build_scene(desc, external_device=host.ptr)
        try:
            _, done = rs.render_tile_buffers(0, desc.spp, tile_size=64, cancel_after=2)
            assert host.error_message() == ""
            assert done == 2
            film, done = rs.render_tile_buffers(0, desc.spp, tile_size=64)
            assert done == 6 and film[..., 3].min() == desc.spp
        finally:
            rs.close()
    finally:
        host.close()


def test_holdout_and_shadow_catcher_objects_are_refused(ref, device):
    """Per-object holdout masks / shadow catchers are not implemented: binding an
    __object_flag array that carries them is refused, not rendered as a plain surface."""
    from raytracingproject_b200.device import DeviceError, DeviceMemory, MEM_GLOBAL
    desc = small_cases()["cornell"]
    rs = ref.build_scene(desc)
    try:
        arrays = rs.device_arrays()
    finally:
        rs.close()
    device.upload_scene(arrays)
    for bit, what in ((0x1, "holdout"), (0x80, "shadow catcher")):
        flags = arrays["__object_flag"][0].copy().view(np.uint32)
        flags[1] |= bit
        with pytest.raises(DeviceError, match=what):
            device.mem_copy_to(DeviceMemory("__object_flag", flags.view(np.uint8), MEM_GLOBAL))
    # the unmodified array binds again and the device still renders
    device._error = ""
    device.upload_scene(arrays)
    assert device.render(desc.width, desc.height, 4, 0, 1)[..., 3].min() == 1.0

"""GPU image parity of the whole wavefront (b200_render through the C ABI) against
the reference CPU kernel (generic scalar variant, oracle/_ref) on identical
scenes, seeds and Sobol samples.  Gates (BASELINE.json): per-pixel RMSE <= 1e-3
and mean luminance within 0.1 %.  Both sides use the same sample set, so the
gates hold at any spp, not only at 1024."""
import numpy as np
import pytest

from scene_cases import (adaptive_cases, ao_cases, camera_cases, closure_cases, denoising_cases,
                         image_cases, light_cases,
                         pass_cases, principled_cases, sampling_cases, small_cases,
                         texture_cases, world_light_cases)

pytestmark = pytest.mark.gpu

SPP = 16
W_SMALL, H_SMALL = 256, 144


def luminance(img):
    return 0.2126 * img[..., 0] + 0.7152 * img[..., 1] + 0.0722 * img[..., 2]


def image_gates(ref_img, got, spp, label):
    a = ref_img[..., :4].astype(np.float64) / spp
    b = got[..., :4].astype(np.float64) / spp
    rmse = float(np.sqrt(np.mean((a[..., :3] - b[..., :3]) ** 2)))
    la, lb = luminance(a).mean(), luminance(b).mean()
    rel = abs(la - lb) / max(la, 1e-12)
    exact = float(np.mean(np.all(ref_img == got, axis=-1)))
    print("%s: rmse=%.3e mean_lum ref=%.6f got=%.6f rel=%.3e bit-identical pixels=%.4f "
          "max|d|=%.3e alpha max|d|=%.3e" % (label, rmse, la, lb, rel, exact,
          np.abs(a[..., :3] - b[..., :3]).max(), np.abs(a[..., 3] - b[..., 3]).max()))
    assert rmse <= 1e-3, label
    assert rel <= 1e-3, label
    assert np.abs(a[..., 3] - b[..., 3]).max() <= 1e-6
    return rmse, rel


@pytest.mark.parametrize("name", ["cube", "cornell", "terrain", "instanced"])
def test_image_matches_reference(ref, device, name):
    desc = small_cases()[name]
    rs = ref.build_scene(desc)
    try:
        device.upload_scene(rs.device_arrays())
        ref_img, _ = rs.render(0, SPP, tile_size=64)
        got = device.render(desc.width, desc.height, rs.pass_stride, 0, SPP)
        print(name, device.stats())
        image_gates(ref_img, got, SPP, name)
    finally:
        rs.close()


@pytest.mark.parametrize("name", ["cube_principled", "cornell_principled"])
def test_principled_image_matches_reference(ref, device, name):
    desc = principled_cases()[name]
    rs = ref.build_scene(desc)
    try:
        device.upload_scene(rs.device_arrays())
        ref_img, _ = rs.render(0, SPP, tile_size=64)
        got = device.render(desc.width, desc.height, rs.pass_stride, 0, SPP)
        print(name, device.stats())
        # no texture / attribute node, sheen a constant zero: the lean shading kernels ran
        assert device.stats()["svm_extended"] == 0
        image_gates(ref_img, got, SPP, name)
    finally:
        rs.close()


@pytest.mark.parametrize("name", ["cube_principled_multiscatter",
                                  "cornell_principled_multiscatter"])
def test_multiscatter_ggx_matches_reference(ref, device, name):
    """The Principled BSDF's DEFAULT distribution (Multiscatter GGX): stochastic lobes
    evaluated by a random walk over the microsurface with the shading point's LCG.  The
    device follows the reference's walk number for number, so even at 16 spp the images
    agree to the usual gates."""
    desc = principled_cases()[name]
    assert 'distribution="Multiscatter GGX"' in desc.xml
    rs = ref.build_scene(desc)
    try:
        device.upload_scene(rs.device_arrays())
        ref_img, _ = rs.render(0, SPP, tile_size=64)
        got = device.render(desc.width, desc.height, rs.pass_stride, 0, SPP)
        print(name, device.stats())
        image_gates(ref_img, got, SPP, name)
    finally:
        rs.close()


@pytest.mark.parametrize("name", ["cube_spot", "cube_mixed_lights", "cube_light_falloff", "cornell_mesh_light",
                                  "cornell_mesh_light_instanced"])
def test_lamp_types_match_reference(ref, device, name):
    desc = light_cases()[name]
    rs = ref.build_scene(desc)
    try:
        device.upload_scene(rs.device_arrays())
        ref_img, _ = rs.render(0, SPP, tile_size=64)
        got = device.render(desc.width, desc.height, rs.pass_stride, 0, SPP)
        assert ref_img[..., :3].max() > 0.0
        image_gates(ref_img, got, SPP, name)
    finally:
        rs.close()


@pytest.mark.parametrize("name", ["cornell_closures", "cornell_closures2", "cornell_closures_multi",
                                  "cornell_transparent_opaque_shadow", "cornell_transparent",
                                  "cornell_transparent_panes",
                                  "cornell_transparent_panes_limit"])
def test_closure_nodes_match_reference(ref, device, name):
    desc = closure_cases()[name]
    rs = ref.build_scene(desc)
    try:
        device.upload_scene(rs.device_arrays())
        ref_img, _ = rs.render(0, SPP, tile_size=64)
        got = device.render(desc.width, desc.height, rs.pass_stride, 0, SPP)
        st = device.stats()
        print(name, st)
        image_gates(ref_img, got, SPP, name)
        if "transparent" in name and "opaque_shadow" not in name:
            # the stepping loop of transparent shadows is queued blind (ts_rounds): the host
            # stops the stream once per call, like every other scene - and the same film
            # comes out when it asks after every step instead
            assert st["host_syncs"] == 1
            device.set_option("sync_iterations", 1)
            try:
                synced = device.render(desc.width, desc.height, rs.pass_stride, 0, SPP)
                assert device.stats()["host_syncs"] > 1
            finally:
                device.set_option("sync_iterations", 0)
            assert np.array_equal(got, synced)
    finally:
        rs.close()


@pytest.mark.parametrize("name", ["cube_dof_disk", "cube_dof_blades", "cube_ortho",
                                  "cube_ortho_dof", "cornell_equirect", "cornell_equirect_dof",
                                  "cornell_fisheye", "cornell_fisheye_equisolid",
                                  "cornell_mirrorball"])
def test_camera_models_match_reference(ref, device, name):
    desc = camera_cases()[name]
    rs = ref.build_scene(desc)
    try:
        device.upload_scene(rs.device_arrays())
        ref_img, _ = rs.render(0, SPP, tile_size=64)
        got = device.render(desc.width, desc.height, rs.pass_stride, 0, SPP)
        assert ref_img[..., :3].max() > 0.0
        image_gates(ref_img, got, SPP, name)
    finally:
        rs.close()


@pytest.mark.parametrize("name", ["cornell_textured", "cornell_textured2", "cornell_textured3", "cornell_textured4",
                                  "cornell_textured_mesh_light", "cornell_textured_ortho"])
def test_texture_nodes_match_reference(ref, device, name):
    desc = texture_cases()[name]
    rs = ref.build_scene(desc)
    try:
        device.upload_scene(rs.device_arrays())
        ref_img, _ = rs.render(0, SPP, tile_size=64)
        got = device.render(desc.width, desc.height, rs.pass_stride, 0, SPP)
        assert ref_img[..., :3].max() > 0.0
        assert device.stats()["svm_extended"] == 1  # the full-interpreter kernels ran
        image_gates(ref_img, got, SPP, name)
    finally:
        rs.close()


@pytest.mark.parametrize("name", ["cornell_image", "cornell_image2", "cube_env_equirect",
                                  "cube_env_mirrorball"])
def test_image_textures_match_reference(ref, device, name):
    """Image / Environment Texture nodes: the images the reference's ImageManager loaded
    are bound per slot (tex_alloc -> b200_texture_set) and sampled with the CPU device's
    arithmetic.  Through the Python mirror, and through the C++ shim with the reference's
    own ImageManager handing the device_texture to Device::mem_copy_to."""
    from raytracingproject_b200.device import B200HostDevice
    desc = image_cases()[name]
    rs = ref.build_scene(desc)
    try:
        textures = rs.textures()
        assert textures
        device.upload_scene(rs.device_arrays(), textures)
        ref_img, _ = rs.render(0, SPP, tile_size=64)
        got = device.render(desc.width, desc.height, rs.pass_stride, 0, SPP).copy()
        assert ref_img[..., :3].max() > 0.0
        assert device.stats()["svm_extended"] == 1
        image_gates(ref_img, got, SPP, name)
    finally:
        rs.close()
    host = B200HostDevice(0)
    try:
        rs = ref.build_scene(desc, external_device=host.ptr)
        try:
            shim, _ = rs.render(0, SPP, tile_size=0)
            assert host.error_message() == ""
            assert np.array_equal(shim, got), ("C++ shim and Python mirror disagree",
                                               float(np.abs(shim - got).max()),
                                               int((shim != got).any(axis=-1).sum()))
        finally:
            rs.close()
    finally:
        host.close()


@pytest.mark.parametrize("name", ["passes_cornell_principled", "passes_cornell_image",
                                  "passes_cube_env_transparent_film",
                                  "passes_cornell_mesh_light", "passes_data_only",
                                  "passes_ao", "passes_ao_data_only",
                                  "light_invisible_to_glossy_rays",
                                  "passes_transparent_shadows",
                                  "clamp_after_transparent_shadows"])
def test_render_passes_match_reference(ref, device, name):
    """Light and data passes (kernel_passes.h, kernel_accumulate.h with use_light_pass):
    every pass of the film against the reference CPU kernel - the colour passes to the
    image gates, the id passes exactly - plus the combined pass, which with light passes
    is the SUM of the per-class passes on both sides."""
    from raytracingproject_b200 import scenes
    desc = pass_cases()[name]
    rs = ref.build_scene(desc)
    try:
        textures = rs.textures()
        device.upload_scene(rs.device_arrays(), textures)
        ref_img, _ = rs.render(0, SPP, tile_size=64)
        got = device.render(desc.width, desc.height, rs.pass_stride, 0, SPP)
        assert got.shape == ref_img.shape and ref_img.shape[-1] == rs.pass_stride
        off, _ = rs.pass_offset(1)   # PASS_COMBINED
        image_gates(ref_img[..., off:off + 4], got[..., off:off + 4], SPP, name + " combined")
        names = {v: k for k, v in scenes.PASS.items()}
        for pass_type in desc.passes:
            off, comps = rs.pass_offset(pass_type)
            n = {1: 1, 4: 3}[comps] if names[pass_type] != "shadow" else 4
            a = ref_img[..., off:off + n].astype(np.float64) / SPP
            b = got[..., off:off + n].astype(np.float64) / SPP
            rmse = float(np.sqrt(np.mean((a - b) ** 2)))
            scale = max(float(np.abs(a).mean()), 1e-9)
            rel = abs(float(a.mean()) - float(b.mean())) / scale
            print("%s %-22s rmse=%.3e mean ref=%.6f got=%.6f rel=%.2e max|d|=%.2e" % (
                name, names[pass_type], rmse, a.mean(), b.mean(), rel, np.abs(a - b).max()))
            must_have = {"diffuse_direct", "diffuse_indirect", "shadow", "depth", "normal"}
            if "transparent" not in name:
                must_have.add("glossy_indirect")
            if "mesh_light" in name:
                must_have.add("emission")       # lamps are not visible to the camera
            if "env" in name:
                must_have.add("background")     # the Cornell box is closed, its world black
            if name == "passes_ao":
                must_have.add("ao")
            if name == "passes_cornell_principled":
                must_have |= {"transmission_indirect", "transmission_color", "glossy_color",
                              "diffuse_color", "mist", "uv"}
            assert np.abs(a).max() > 0.0 or names[pass_type] not in must_have, \
                names[pass_type] + ": the reference pass is empty, the case tests nothing"
            if names[pass_type] in ("object_id", "material_id"):
                assert np.array_equal(a, b)
            else:
                assert rmse <= 1e-3 * max(1.0, scale) and rel <= 1e-3, names[pass_type]
        # the 4th float of a 3-component pass is never touched
        for pass_type in desc.passes:
            off, comps = rs.pass_offset(pass_type)
            if comps == 4 and names[pass_type] != "shadow":
                assert not got[..., off + 3].any()
        got = got.copy()
    finally:
        rs.close()
    if name == "passes_cornell_mesh_light":
        # the same film through the C++ shim: the reference's Film / RenderBuffers decide
        # the layout, the device reads it from KernelData
        from raytracingproject_b200.device import B200HostDevice
        host = B200HostDevice(0)
        try:
            rs = ref.build_scene(desc, external_device=host.ptr)
            try:
                shim, _ = rs.render(0, SPP, tile_size=64)
                assert host.error_message() == ""
                assert np.array_equal(shim, got)
            finally:
                rs.close()
        finally:
            host.close()


@pytest.mark.parametrize("name", ["cube_world_light", "cube_world_light_principled",
                                  "cube_world_light_mirrorball", "cube_world_light_passes"])
def test_world_light_matches_reference(ref, device, name):
    """Background MIS (kernel_light_background.h): the world is in the light distribution,
    sampled by its importance map, and BSDF-sampled background hits are MIS-weighted.
    Python mirror: the map comes from the reference's own host update.  C++ shim: the
    reference's LightManager runs DeviceTask::SHADER on THIS device
    (b200_shader_eval_background) and builds the map from its answer - the map must agree
    with the CPU device's, the film with the gates."""
    from raytracingproject_b200.device import B200HostDevice
    desc = world_light_cases()[name]
    rs = ref.build_scene(desc)
    try:
        arrays = rs.device_arrays()
        assert "__light_background_marginal_cdf" in arrays
        cpu_marg = arrays["__light_background_marginal_cdf"][0].view(np.float32).copy()
        cpu_cond = arrays["__light_background_conditional_cdf"][0].view(np.float32).copy()
        device.upload_scene(arrays, rs.textures())
        ref_img, _ = rs.render(0, SPP, tile_size=64)
        got = device.render(desc.width, desc.height, rs.pass_stride, 0, SPP).copy()
        off, _ = rs.pass_offset(1)
        image_gates(ref_img[..., off:off + 4], got[..., off:off + 4], SPP, name)
        if desc.passes:
            d = np.abs(ref_img - got).max() / SPP
            print(name, "all passes max|d|", d)
            assert d < 1e-3 * max(1.0, float(np.abs(ref_img).max()) / SPP)
    finally:
        rs.close()
    host = B200HostDevice(0)
    try:
        rs = ref.build_scene(desc, external_device=host.ptr)
        try:
            marg = rs.global_array("__light_background_marginal_cdf")[0].view(np.float32)
            cond = rs.global_array("__light_background_conditional_cdf")[0].view(np.float32)
            assert marg.shape == cpu_marg.shape and cond.shape == cpu_cond.shape
            print(name, "importance map: max|d| marginal %.3e conditional %.3e" % (
                np.abs(marg - cpu_marg).max(), np.abs(cond - cpu_cond).max()))
            np.testing.assert_allclose(marg, cpu_marg, rtol=1e-4, atol=1e-6)
            np.testing.assert_allclose(cond, cpu_cond, rtol=1e-4, atol=1e-6)
            shim, _ = rs.render(0, SPP, tile_size=64)
            assert host.error_message() == ""
            image_gates(ref_img[..., off:off + 4], shim[..., off:off + 4], SPP, name + " (shim)")
        finally:
            rs.close()
    finally:
        host.close()


@pytest.mark.parametrize("name", ["adaptive_cornell", "adaptive_cube"])
def test_adaptive_sampling_matches_reference(ref, device, name):
    """Adaptive sampling (kernel_adaptive_sampling.h; device_cpu.cpp:838-945): pixels stop
    at the filter points where the reference CPU device stops them - the sample-count pass
    says which - and the film is rescaled to a uniform sample count at the end.  Through
    the Python mirror and the C++ shim (whose KernelData says adaptive_stop_per_sample = 0:
    the convergence test runs as a kernel at the filter points, like the CUDA device)."""
    from raytracingproject_b200 import scenes
    from raytracingproject_b200.device import B200HostDevice
    desc = adaptive_cases()[name]
    spp = desc.spp
    rs = ref.build_scene(desc)
    try:
        device.upload_scene(rs.device_arrays(), rs.textures())
        ref_img, _ = rs.render(0, spp, tile_size=0)
        got = device.render(desc.width, desc.height, rs.pass_stride, 0, spp).copy()
        c_off, _ = rs.pass_offset(scenes.PASS["sample_count"])
        a_off, _ = rs.pass_offset(scenes.PASS["adaptive_aux_buffer"])
    finally:
        rs.close()

    def check(img, label):
        ref_cnt, cnt = ref_img[..., c_off], img[..., c_off]
        same = float(np.mean(ref_cnt == cnt))
        print(label, "sample count: ref mean %.2f got %.2f, identical pixels %.4f, histogram %s"
              % (ref_cnt.mean(), cnt.mean(), same,
                 dict(zip(*[a.tolist() for a in np.unique(cnt, return_counts=True)]))))
        assert ref_cnt.min() < spp and ref_cnt.max() == spp   # the case does stop pixels
        assert same >= 0.995
        image_gates(ref_img[..., :4], img[..., :4], spp, label)
        d = np.abs(ref_img[..., a_off:a_off + 3] - img[..., a_off:a_off + 3])[ref_cnt == cnt]
        assert d.max() / spp < 2e-3

    check(got, name)
    host = B200HostDevice(0)
    try:
        rs = ref.build_scene(desc, external_device=host.ptr)
        try:
            shim, _ = rs.render(0, spp, tile_size=0)
            assert host.error_message() == ""
            check(shim, name + " (shim)")
        finally:
            rs.close()
    finally:
        host.close()


def test_program_with_image_nodes_needs_bound_images(ref, device):
    """A compiled program that names image slots while none is bound is refused, not
    rendered with missing textures."""
    from raytracingproject_b200.device import DeviceError
    desc = image_cases()["cornell_image"]
    rs = ref.build_scene(desc)
    try:
        device.upload_scene(rs.device_arrays())        # no textures
        with pytest.raises(DeviceError, match="image"):
            device.render(desc.width, desc.height, rs.pass_stride, 0, 1)
    finally:
        rs.close()


def test_sheen_alone_selects_the_full_interpreter(ref, device):
    """A Principled BSDF with sheen and no texture node: the host's scan of the program
    (constant inputs of the closure node) must route it to the full interpreter."""
    from raytracingproject_b200 import scenes
    desc = scenes.cornell(W_SMALL, H_SMALL, materials="principled")
    assert 'name="p" distribution="GGX"' in desc.xml
    desc.xml = desc.xml.replace('<principled_bsdf name="p" ', '<principled_bsdf name="p" sheen="0.7" '
                                'sheen_tint="0.3" ')
    rs = ref.build_scene(desc)
    try:
        device.upload_scene(rs.device_arrays())
        ref_img, _ = rs.render(0, SPP, tile_size=64)
        got = device.render(desc.width, desc.height, rs.pass_stride, 0, SPP)
        assert device.stats()["svm_extended"] == 1
        image_gates(ref_img, got, SPP, "principled with sheen")
    finally:
        rs.close()


@pytest.mark.parametrize("name", ["cornell_ao", "cornell_ao_principled", "cornell_ao_textured"])
def test_ambient_occlusion_matches_reference(ref, device, name):
    desc = ao_cases()[name]
    rs = ref.build_scene(desc)
    try:
        device.upload_scene(rs.device_arrays())
        ref_img, _ = rs.render(0, SPP, tile_size=64)
        got = device.render(desc.width, desc.height, rs.pass_stride, 0, SPP)
        st = device.stats()
        # every surviving surface hit adds an AO ray to the light ray
        assert st["shadow_rays"] > st["bounce_rays"]
        image_gates(ref_img, got, SPP, name)
    finally:
        rs.close()


def test_shadow_terminator_offset_matches_reference(ref, device):
    """Object::shadow_terminator_offset on every object: the host sees it in __objects
    and selects the full kernels, whose bsdf_eval / bsdf_sample apply shift_cos_in."""
    from raytracingproject_b200 import scenes
    desc = scenes.cornell(W_SMALL, H_SMALL, materials="principled")
    desc.terminator_offset = 0.6
    rs = ref.build_scene(desc)
    try:
        device.upload_scene(rs.device_arrays())
        ref_img, _ = rs.render(0, SPP, tile_size=64)
        got = device.render(desc.width, desc.height, rs.pass_stride, 0, SPP)
        assert device.stats()["svm_extended"] == 1
        image_gates(ref_img, got, SPP, "terminator offset")
        plain = scenes.cornell(W_SMALL, H_SMALL, materials="principled")
        rs2 = ref.build_scene(plain)
        try:
            other, _ = rs2.render(0, SPP, tile_size=64)
        finally:
            rs2.close()
        assert not np.array_equal(other, ref_img)  # the offset does change the picture
    finally:
        rs.close()


def test_bent_normal_redoes_the_batch_with_the_full_kernels(ref, device):
    """Only lean nodes, but a BSDF normal that is not the shading normal: the host's scan
    cannot know, the lean kernels notice on the device, nothing of that batch reaches
    the film, the batch is traced again with the full kernels (bump shadowing term of
    bsdf_eval / bsdf_sample) - and the frame matches the reference."""
    from raytracingproject_b200 import scenes
    desc = scenes.cornell(W_SMALL, H_SMALL, materials="diffuse")
    desc.xml = desc.xml.replace(
        '  <diffuse_bsdf name="d" color="0.73 0.73 0.73"/>\n',
        '  <diffuse_bsdf name="d" color="0.73 0.73 0.73"/>\n  <geometry name="g"/>\n'
        '  <vector_math name="bn" type="add" vector2="0.3 0.2 0.1"/>\n'
        '  <connect from="g normal" to="bn vector1"/>\n'
        '  <vector_math name="bnn" type="normalize"/>\n'
        '  <connect from="bn vector" to="bnn vector1"/>\n'
        '  <connect from="bnn vector" to="d normal"/>\n', 1)
    assert "bnn vector" in desc.xml
    rs = ref.build_scene(desc)
    try:
        device.upload_scene(rs.device_arrays())
        ref_img, _ = rs.render(0, SPP, tile_size=64)
        got = device.render(desc.width, desc.height, rs.pass_stride, 0, SPP)
        assert device.stats()["svm_extended"] == 1
        image_gates(ref_img, got, SPP, "bent normal")
        again = device.render(desc.width, desc.height, rs.pass_stride, 0, SPP)
        assert np.array_equal(again, got)  # the context stays on the full kernels
    finally:
        rs.close()


def test_window_coordinates_need_a_perspective_camera(ref, device):
    """NODE_TEXCO_WINDOW reads the ray origin under an orthographic camera and the
    panorama projection under a panoramic one; both are refused, not approximated."""
    from raytracingproject_b200.device import DeviceError
    ortho = scenes_cornell_ortho_textured2()  # the "glass" shader reads window coords
    rs = ref.build_scene(ortho)
    try:
        device.upload_scene(rs.device_arrays())
        with pytest.raises(DeviceError) as e:
            device.render(ortho.width, ortho.height, rs.pass_stride, 0, 1)
        assert "window texture coordinates" in str(e.value)
    finally:
        rs.close()


def scenes_cornell_ortho_textured2():
    from raytracingproject_b200 import scenes
    return scenes.cornell(64, 48, spp=1, materials="textured2", cam_type="orthograph")


@pytest.mark.parametrize("name", ["cornell_cmj16", "cornell_cmj12", "cornell_pmj"])
def test_cmj_sampling_matches_reference(ref, device, name):
    desc = sampling_cases()[name]
    rs = ref.build_scene(desc)
    try:
        device.upload_scene(rs.device_arrays())
        ref_img, _ = rs.render(0, desc.spp, tile_size=64)
        got = device.render(desc.width, desc.height, rs.pass_stride, 0, desc.spp)
        image_gates(ref_img, got, desc.spp, name)
    finally:
        rs.close()


def test_sample_ranges_add_up(ref, device):
    """Rendering [0,8) then [8,16) into the same film equals [0,16) (the property
    the multi-GPU sample split relies on)."""
    from raytracingproject_b200.device import DeviceMemory
    desc = small_cases()["cornell"]
    rs = ref.build_scene(desc)
    try:
        device.upload_scene(rs.device_arrays())
        w, h, ps = desc.width, desc.height, rs.pass_stride
        whole = device.render(w, h, ps, 0, 16).copy()
        film = DeviceMemory("RenderBuffers", np.zeros((h, w, ps), np.float32))
        device.mem_zero(film)
        device.render_tile(film.device_pointer, 0, 0, w, h, 0, 8, 0, w)
        device.render_tile(film.device_pointer, 0, 0, w, h, 8, 8, 0, w)
        device.mem_copy_from(film)
        device.mem_free(film)
        assert np.array_equal(whole, film.host)
    finally:
        rs.close()


def test_tiles_and_small_pool(ref, device):
    """Tiled rendering with a tiny path pool (forces row bands and several batches)
    gives the same film as one full-frame launch."""
    from raytracingproject_b200.device import DeviceMemory
    desc = small_cases()["cube"]
    rs = ref.build_scene(desc)
    try:
        device.upload_scene(rs.device_arrays())
        w, h, ps = desc.width, desc.height, rs.pass_stride
        whole = device.render(w, h, ps, 0, 4).copy()
        device.set_option("batch_paths", 5000)
        film = DeviceMemory("RenderBuffers", np.zeros((h, w, ps), np.float32))
        device.mem_zero(film)
        for ty in range(0, h, 64):
            for tx in range(0, w, 96):
                device.render_tile(film.device_pointer, tx, ty, min(96, w - tx), min(64, h - ty),
                                   0, 4, 0, w)
        device.mem_copy_from(film)
        device.mem_free(film)
        device.set_option("batch_paths", 0)
        assert np.array_equal(whole, film.host)
    finally:
        rs.close()


DENOISING_FEATURES = (("normal", 0, 3), ("normal_var", 3, 3), ("albedo", 6, 3),
                      ("albedo_var", 9, 3), ("depth", 12, 1), ("depth_var", 13, 1),
                      ("shadow_a", 14, 3), ("shadow_b", 17, 3), ("color", 20, 3),
                      ("color_var", 23, 3))


@pytest.mark.parametrize("name", ["denoising_data", "denoising_clean", "denoising_ao",
                                  "denoising_transparent_shadows", "denoising_cube_env"])
def test_denoising_data_passes_match_reference(ref, device, name):
    """The denoising data passes behind the regular ones (kernel_passes.h:21-122, 355-389;
    DENOISING_PASS_* offsets of kernel_types.h:414-437) against the reference CPU kernel,
    feature by feature, plus the combined pass and the clean pass."""
    desc = denoising_cases()[name]
    rs = ref.build_scene(desc)
    try:
        device.upload_scene(rs.device_arrays(), rs.textures())
        ref_img, _ = rs.render(0, SPP, tile_size=64)
        got = device.render(desc.width, desc.height, rs.pass_stride, 0, SPP)
        assert got.shape == ref_img.shape
        off, _ = rs.pass_offset(1)
        image_gates(ref_img[..., off:off + 4], got[..., off:off + 4], SPP, name + " combined")
        dn, clean = rs.denoising_offset()
        assert dn > 0 and dn + 26 <= rs.pass_stride
        features = list(DENOISING_FEATURES)
        if desc.denoising[0]:
            assert clean > 0
            features.append(("clean", clean - dn, 3))
        for label, o, n in features:
            a = ref_img[..., dn + o:dn + o + n].astype(np.float64) / SPP
            b = got[..., dn + o:dn + o + n].astype(np.float64) / SPP
            rmse = float(np.sqrt(np.mean((a - b) ** 2)))
            scale = max(float(np.abs(a).mean()), 1e-9)
            rel = abs(float(a.mean()) - float(b.mean())) / scale
            print("%s %-12s rmse=%.3e mean ref=%.6f got=%.6f rel=%.2e max|d|=%.2e" % (
                name, label, rmse, a.mean(), b.mean(), rel, np.abs(a - b).max()))
            assert np.abs(a).max() > 0.0, label + ": the reference feature is empty"
            assert rmse <= 1e-3 * max(1.0, scale) and rel <= 1e-3, label
        # nothing is written past the passes the film holds
        used = dn + 26 + (3 if desc.denoising[0] else 0)
        assert not got[..., used:].any() and not ref_img[..., used:].any()
    finally:
        rs.close()


@pytest.mark.parametrize("name", ["cube_principled_multiscatter",
                                  "cornell_principled_multiscatter"])
def test_shading_block_shapes_render_the_same_film(ref, device, name):
    """The lean multiscatter shading kernel exists in three block shapes (k_shade_surface
    WIDE: two blocks of 256 threads per SM, one of 512, one of 1024); the device times them
    on the first batches of a scene and keeps the fastest.
    Same code, same arithmetic: whichever runs, the film is bit-identical - forced either
    way, and while the probe alternates between them batch by batch."""
    from scene_cases import principled_cases
    desc = principled_cases()[name]
    rs = ref.build_scene(desc)
    try:
        device.upload_scene(rs.device_arrays())
        films = {}
        for mode in (0, 1, 2, -1):
            device.set_option("shade_wide", mode)
            device.set_option("batch_paths", 0 if mode >= 0 else 1 << 16)
            films[mode] = device.render(desc.width, desc.height, rs.pass_stride, 0, SPP)
            st = device.stats()
            assert st["svm_extended"] == 0
            assert st["shade_wide"] == (mode if mode >= 0 else st["shade_wide"])
            if mode < 0:
                assert st["batches"] >= 8   # the probe saw every shape
        assert np.array_equal(films[0], films[1])
        assert np.array_equal(films[0], films[2])
        assert np.array_equal(films[0], films[-1])
        ref_img, _ = rs.render(0, SPP, tile_size=64)
        image_gates(ref_img, films[1], SPP, name + " wide")
    finally:
        device.set_option("shade_wide", -1)
        device.set_option("batch_paths", 0)
        rs.close()


def test_full_kernel_block_shapes_render_the_same_film(ref, device):
    """The same for the full shading kernel (world AO routes the scene there)."""
    desc = ao_cases()["cornell_ao_principled"]
    rs = ref.build_scene(desc)
    try:
        device.upload_scene(rs.device_arrays())
        films = {}
        for mode in (0, 1, 2):
            device.set_option("shade_wide", mode)
            films[mode] = device.render(desc.width, desc.height, rs.pass_stride, 0, SPP)
            st = device.stats()
            assert st["svm_extended"] == 1 and st["shade_wide"] == mode
        # the AO and the light ray of a path are added with atomics, in either order
        for mode in (1, 2):
            a, b = films[0][..., :3] / SPP, films[mode][..., :3] / SPP
            assert np.abs(a - b).max() <= 1e-5 * max(1.0, np.abs(a).max())
    finally:
        device.set_option("shade_wide", -1)
        rs.close()

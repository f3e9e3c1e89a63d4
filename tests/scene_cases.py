"""Small, seconds-sized scene cases shared by the parity tests."""
from raytracingproject_b200 import scenes

W, H = 256, 144


def small_cases():
    return {
        "cube": scenes.default_cube(W, H, material="diffuse"),
        "cornell": scenes.cornell(W, H, materials="diffuse"),
        "terrain": scenes.terrain(W, H, n=96),
        "instanced": scenes.instanced(W, H, grid=8, subdiv=3),
    }


def principled_cases():
    """Config 1 / config 3 materials: Principled BSDF (GGX distribution), metallic and
    glass variants, 12 / 8 bounces."""
    return {
        "cube_principled": scenes.default_cube(W, H, material="principled"),
        "cornell_principled": scenes.cornell(W, H, materials="principled"),
        # the node's own default distribution: random-walk lobes (microfacet_multi.cuh)
        "cube_principled_multiscatter": scenes.default_cube(
            W, H, material="principled", distribution="Multiscatter GGX"),
        "cornell_principled_multiscatter": scenes.cornell(
            W, H, materials="principled", distribution="Multiscatter GGX"),
    }


def sampling_cases():
    """Correlated multi-jittered sampling instead of Sobol (kernel_jitter.h); aa_samples
    16 = a 4 x 4 grid, 12 = a 3 x 4 grid with the non-power-of-two permutation."""
    return {
        "cornell_cmj16": scenes.cornell(W, H, spp=16, materials="diffuse", pattern="cmj"),
        "cornell_cmj12": scenes.cornell(W, H, spp=12, materials="principled", pattern="cmj"),
        # progressive multi-jitter: the host's 48 x 4096-point table, xor-scrambled
        "cornell_pmj": scenes.cornell(W, H, spp=16, materials="principled", pattern="pmj"),
    }


def camera_cases():
    """Camera models beyond the pinhole: thin lens with a disk and with a rotated
    six-blade anamorphic aperture (kernel_camera.h:21-40), orthographic with and
    without a lens."""
    cube = lambda **kw: scenes.default_cube(W, H, material="diffuse", **kw)
    # panoramic projections look down the camera's +X axis (kernel_projection.h): a pose
    # inside the box, just behind its open front, with X pointing at the back wall
    import numpy as np
    inside = np.array([[0, 0, 1, 0.0], [1, 0, 0, -0.95], [0, 1, 0, 1.0]], np.float32)
    box = lambda **kw: scenes.cornell(W, H, materials="diffuse", cam_type="panorama",
                                      cam_pose=inside, **kw)
    return {
        "cube_dof_disk": cube(cam_extra='aperturesize="0.35" focaldistance="9.5"'),
        "cube_dof_blades": cube(cam_extra='aperturesize="0.3" focaldistance="11" blades="6" '
                                'bladesrotation="0.4" aperture_ratio="1.6"'),
        "cube_ortho": cube(cam_type="orthograph"),
        "cube_ortho_dof": cube(cam_type="orthograph",
                               cam_extra='aperturesize="0.15" focaldistance="10"'),
        # panoramic cameras from inside the Cornell box; the fisheyes leave the frame
        # corners outside the lens (camera rays with t = 0)
        "cornell_equirect": box(cam_extra='panorama_type="equirectangular"'),
        "cornell_equirect_dof": box(cam_extra='panorama_type="equirectangular" '
                                    'aperturesize="0.03" focaldistance="1.5"'),
        "cornell_fisheye": box(cam_extra='panorama_type="fisheye_equidistant" '
                               'fisheye_fov="3.4"'),
        "cornell_fisheye_equisolid": box(cam_extra='panorama_type="fisheye_equisolid" '
                                         'fisheye_lens="10.5" fisheye_fov="3.14159" '
                                         'sensorwidth="36" sensorheight="20.25"'),
        "cornell_mirrorball": box(cam_extra='panorama_type="mirrorball"'),
    }


def closure_cases():
    """BSDF nodes beyond Diffuse / Principled / Glossy-GGX: Glass and Refraction (GGX and
    sharp), sharp Glossy, Translucent, Oren-Nayar - 8 bounces in a Cornell box."""
    return {
        "cornell_closures": scenes.cornell(W, H, materials="closures"),
        "cornell_closures2": scenes.cornell(W, H, materials="closures2"),
        # the Multiscatter GGX option of the Glossy and the Glass node
        "cornell_closures_multi": scenes.cornell(W, H, materials="closures_multi"),
        # Transparent BSDF: straight-through bounces, alpha, terminate-after-transparent
        "cornell_transparent_opaque_shadow": scenes.cornell(
            W, H, materials="transparent_opaque_shadow"),
        # ... and transparent shadows (kernel_shadow.h stepped loop)
        "cornell_transparent": scenes.cornell(W, H, materials="transparent"),
        # five stacked transparent sheets under the light; with a limit of 4 transparent
        # bounces the shadow rays through all of them count as blocked
        "cornell_transparent_panes": scenes.cornell(W, H, materials="transparent", panes=5),
        "cornell_transparent_panes_limit": scenes.cornell(W, H, materials="transparent",
                                                          panes=5, transparent_max=4),
    }


def texture_cases():
    """Texture coordinates, mesh attributes (generated, UV), mapping, the procedural
    textures (noise 1D-4D, wave, magic, checker, brick, gradient) and the Principled
    features built on them: anisotropy with the generated-coordinates tangent, sheen,
    clearcoat, and a mix of two Principled BSDFs (16 closures per shader)."""
    return {
        "cornell_textured": scenes.cornell(W, H, materials="textured"),
        "cornell_textured2": scenes.cornell(W, H, materials="textured2"),
        # HSV / map range / vector rotate + transform / object info / camera / white noise
        "cornell_textured3": scenes.cornell(W, H, materials="textured3"),
        # Voronoi (five features, 1D-4D, four metrics) and Musgrave (five types)
        "cornell_textured4": scenes.cornell(W, H, materials="textured4"),
        # the same programs under a lamp-less mesh light (emissive-triangle MIS evaluates
        # the surface shader with PATH_RAY_EMISSION) and through an orthographic camera
        "cornell_textured_mesh_light": scenes.cornell(W, H, materials="textured", light="mesh"),
        "cornell_textured_ortho": scenes.cornell(
            W, H, materials="textured", cam_type="orthograph"),
    }


def image_cases():
    """Image textures (svm_image.cuh): byte / ushort / half / float images with one and
    four channels, closest / linear / cubic lookups, repeat / extend / clip, flat / box /
    sphere / tube projections, sRGB decompression, alpha handling, UDIM tiles, a missing
    image - and the Environment Texture node lighting the startup scene."""
    return {
        "cornell_image": scenes.cornell(W, H, materials="image"),
        "cornell_image2": scenes.cornell(W, H, materials="image2"),
        "cube_env_equirect": scenes.default_cube(W, H, world="env_equirect"),
        "cube_env_mirrorball": scenes.default_cube(W, H, world="env_mirrorball",
                                                   material="diffuse"),
    }


def adaptive_cases():
    """Adaptive sampling (adaptive.cuh): PMJ pattern, the aux-buffer and sample-count passes,
    a threshold loose enough that flat regions stop at the first filter point and tight
    enough that edges sample on - the sample-count pass shows who stopped when."""
    cases = {}
    for name, d in (("adaptive_cornell", scenes.cornell(W, H, spp=32, materials="diffuse",
                                                        pattern="pmj")),
                    ("adaptive_cube", scenes.default_cube(W, H, spp=32, material="diffuse"))):
        if "pmj" not in d.xml:
            d = _replace(d, 'sampling_pattern="sobol"', 'sampling_pattern="pmj"')
        d = _replace(d, 'filter_glossy="0"', 'filter_glossy="0" adaptive_threshold="0.02" '
                     'adaptive_min_samples="8"')
        d.passes = [scenes.PASS["adaptive_aux_buffer"], scenes.PASS["sample_count"]]
        d.name += "_adaptive"
        cases[name] = d
    return cases


def world_light_cases():
    """The world as a light (background MIS): an environment texture lights the startup
    scene and is importance-sampled through the luminance map the host builds from a
    DeviceTask::SHADER evaluation of the world shader; diffuse and Principled cube, and the
    passes on top (the world's light arrives in the direct passes, is_lamp = false)."""
    cases = {
        "cube_world_light": scenes.default_cube(W, H, world="env_equirect", world_light=64,
                                                material="diffuse"),
        "cube_world_light_principled": scenes.default_cube(W, H, world="env_equirect",
                                                           world_light=128),
        "cube_world_light_mirrorball": scenes.default_cube(W, H, world="env_mirrorball",
                                                           world_light=64),
    }
    d = scenes.default_cube(W, H, world="env_equirect", world_light=64)
    d.passes = [scenes.PASS[k] for k in ("diffuse_direct", "glossy_direct", "shadow",
                                         "background", "diffuse_indirect")]
    d.name += "_passes"
    cases["cube_world_light_passes"] = d
    return cases


ALL_PASSES = ("depth", "normal", "uv", "object_id", "material_id", "mist", "emission",
              "background", "shadow", "diffuse_direct", "diffuse_indirect", "diffuse_color",
              "glossy_direct", "glossy_indirect", "glossy_color", "transmission_direct",
              "transmission_indirect", "transmission_color")


def pass_cases():
    """Render passes next to the combined one (passes.cuh): every light and data pass in
    scope on the Cornell box with the Principled metal / glass boxes and on the textured
    variant; the startup scene under an environment texture with a transparent film (the
    background pass keeps the colour behind it); a mesh-light scene (emission seen
    directly, after one bounce and later); only data passes (no light-pass machinery);
    and a lamp that glossy rays do not see - its ray-visibility flags alone switch the
    per-class split on and remove that class of its light."""
    with_mist = lambda d: _replace(d, 'exposure="1" ', 'exposure="1" mist_start="2.5" '
                                   'mist_depth="3.0" mist_falloff="2.0" ')
    cases = {}
    d = with_mist(scenes.cornell(W, H, materials="principled"))
    # rough glass instead of the sharp one: light sampling then reaches the transmission class
    d = _replace(d, 'metallic="0" roughness="0" ', 'metallic="0" roughness="0.2" ')
    d.passes = [scenes.PASS[k] for k in ALL_PASSES]
    cases["passes_cornell_principled"] = d
    d = scenes.cornell(W, H, materials="image")
    d.passes = [scenes.PASS[k] for k in ALL_PASSES if k != "mist"]
    cases["passes_cornell_image"] = d
    d = _replace(scenes.default_cube(W, H, world="env_equirect"), "<background>",
                 '<background transparent="true">')
    d.passes = [scenes.PASS[k] for k in ("background", "emission", "diffuse_direct",
                                         "glossy_direct", "normal", "depth", "mist")]
    cases["passes_cube_env_transparent_film"] = d
    d = scenes.cornell(W, H, materials="principled", light="mesh")
    d.passes = [scenes.PASS[k] for k in ("emission", "diffuse_direct", "diffuse_indirect",
                                         "glossy_direct", "glossy_indirect", "shadow")]
    cases["passes_cornell_mesh_light"] = d
    d = scenes.cornell(W, H, materials="principled")
    d.passes = [scenes.PASS[k] for k in ("depth", "normal", "uv", "object_id", "material_id")]
    cases["passes_data_only"] = d
    d = _replace(scenes.cornell(W, H, materials="principled"), 'use_mis="true"',
                 'use_mis="true" use_glossy="false"')
    cases["light_invisible_to_glossy_rays"] = d
    # transparent shadows: the light crosses stacked transparent sheets on its way, every
    # class of it attenuated; the shadow pass holds the attenuation colour.  With a direct
    # clamp low enough to bite: the clamp applies to the ATTENUATED light
    d = _replace(scenes.cornell(W, H, materials="transparent", panes=3),
                 'sample_clamp_direct="0"', 'sample_clamp_direct="0.02"')
    d.passes = [scenes.PASS[k] for k in ("diffuse_direct", "diffuse_indirect", "glossy_direct",
                                         "shadow", "diffuse_color", "normal", "depth")]
    d.name += "_passes"
    cases["passes_transparent_shadows"] = d
    d = _replace(scenes.cornell(W, H, materials="transparent", panes=3),
                 'sample_clamp_direct="0"', 'sample_clamp_direct="0.02"')
    d.name += "_clamped"
    cases["clamp_after_transparent_shadows"] = d
    # world AO with passes (path_radiance_accum_ao): the AO pass itself, AO light in the
    # direct diffuse pass at the first surface and in the indirect light afterwards; and
    # with data passes alone, where it is one more colour in the combined pass
    d = scenes.cornell(W, H, materials="principled", ao=(0.3, 5.0))
    d.passes = [scenes.PASS[k] for k in ("ao", "diffuse_direct", "diffuse_indirect",
                                         "glossy_direct", "glossy_indirect", "shadow",
                                         "diffuse_color", "normal", "depth")]
    d.name += "_passes"
    cases["passes_ao"] = d
    d = scenes.cornell(W, H, materials="diffuse", ao=(0.6, 0.8))
    d.passes = [scenes.PASS[k] for k in ("normal", "depth", "object_id")]
    d.name += "_data_passes"
    cases["passes_ao_data_only"] = d
    return cases


def denoising_cases():
    """Denoising data passes (kernel_passes.h:21-122, 355-389): normal / albedo / depth
    features of the first non-specular surface (through the glass box, off the metal one),
    the shadowing buffers (light that arrived over light that could have), colour with
    variance; alone, next to light passes with a clean pass that takes three components out
    of the noisy colour, and with world AO (whose rays count for the shadowing too)."""
    cases = {}
    d = scenes.cornell(W, H, materials="principled")
    d.denoising = (False, 0)
    d.name += "_denoising"
    cases["denoising_data"] = d
    d = scenes.cornell(W, H, materials="principled")
    d.passes = [scenes.PASS[k] for k in ("diffuse_direct", "diffuse_indirect", "glossy_indirect",
                                         "normal")]
    d.denoising = (True, 1 | 8 | 16)  # diffuse direct, glossy indirect, transmission direct
    d.name += "_denoising_clean"
    cases["denoising_clean"] = d
    d = scenes.cornell(W, H, materials="principled", ao=(0.3, 5.0))
    d.denoising = (False, 0)
    d.name += "_denoising_ao"
    cases["denoising_ao"] = d
    d = scenes.cornell(W, H, materials="transparent", panes=3)
    d.denoising = (False, 0)
    d.name += "_denoising"
    cases["denoising_transparent_shadows"] = d
    d = scenes.default_cube(W, H, world="env_equirect")
    d.denoising = (False, 0)
    d.name += "_denoising"
    cases["denoising_cube_env"] = d
    return cases


def _replace(desc, old, new):
    assert old in desc.xml, old
    desc.xml = desc.xml.replace(old, new)
    return desc


def ao_cases():
    """World ambient occlusion (kernel_path_ao): a second shadow ray per path and bounce,
    cosine-sampled around the averaged diffuse normal, short and long reach."""
    return {
        "cornell_ao": scenes.cornell(W, H, materials="diffuse", ao=(0.6, 0.8)),
        "cornell_ao_principled": scenes.cornell(W, H, materials="principled", ao=(0.3, 5.0)),
        "cornell_ao_textured": scenes.cornell(W, H, materials="textured3", ao=(0.5, 0.4)),
    }


def light_cases():
    """Lamp types of kernel_light.h beyond the configs' point / sun / area: a spot with
    a smooth edge, and three lamps of different types in one light distribution."""
    return {
        "cube_spot": scenes.default_cube(W, H, material="principled", lights="spot"),
        "cube_mixed_lights": scenes.default_cube(W, H, material="diffuse", lights="mixed"),
        "cube_light_falloff": scenes.default_cube(W, H, material="diffuse", lights="falloff"),
        # emissive triangles in the light distribution (kernel_light.h:302-581)
        "cornell_mesh_light": scenes.cornell(W, H, materials="diffuse", light="mesh"),
        "cornell_mesh_light_instanced": scenes.cornell(W, H, materials="principled",
                                                       light="mesh_instanced"),
    }

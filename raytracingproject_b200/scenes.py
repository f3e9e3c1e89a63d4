"""Procedural scene descriptions for the BASELINE.json configs.

A SceneDesc is neutral data: a Cycles-XML header (camera, integrator, film,
background, shaders, lights and small meshes - the syntax read by the
reference's intern/cycles/app/cycles_xml.cpp:614-660) plus numpy triangle
meshes and object instances that are too big for XML.  It contains no
renderer code; the reference's own host code (Scene / BVHBuild / SVM compiler)
flattens it into the device arrays that both the CPU oracle and the B200 device
consume.  All generators are seeded and deterministic.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np


@dataclass
class MeshDesc:
    P: np.ndarray  # (nv, 3) float32
    tris: np.ndarray  # (nt, 3) int32
    shader: str
    smooth: bool = False


@dataclass
class SceneDesc:
    name: str
    xml: str
    width: int
    height: int
    meshes: List[MeshDesc] = field(default_factory=list)
    # (mesh index, 3x4 row-major object-to-world transform)
    objects: List[Tuple[int, np.ndarray]] = field(default_factory=list)
    spp: int = 64
    notes: str = ""
    # image textures: file name (as the XML names it, relative to the scene file) -> pixels,
    # (h, w) or (h, w, c) of uint8 / uint16 / float16 / float32, row 0 = bottom row
    images: Dict[str, np.ndarray] = field(default_factory=dict)
    # render passes next to the combined one: PassType values (kernel_types.h:353-402),
    # see PASS below
    passes: List[int] = field(default_factory=list)
    # UDIM tile numbers of Image Texture nodes: (shader name, node name, [tiles])
    image_tiles: List[Tuple[str, str, List[int]]] = field(default_factory=list)
    # denoising data passes behind the others (film.cpp:604-620): None, or
    # (clean pass?, DenoiseFlag bits of the components the clean pass takes)
    denoising: Optional[Tuple[bool, int]] = None

    @property
    def num_triangles(self):
        return int(sum(len(m.tris) for m in self.meshes))

    @property
    def num_instanced_triangles(self):
        return int(sum(len(self.meshes[m].tris) for m, _ in self.objects))


# kernel_types.h PassType (checked against include/cycles_abi.h by tests/test_scenes_cpu.py)
PASS = {"depth": 2, "normal": 3, "uv": 4, "object_id": 5, "material_id": 6, "mist": 32,
        "emission": 33, "background": 34, "ao": 35, "shadow": 36, "diffuse_direct": 38,
        "diffuse_indirect": 39, "diffuse_color": 40, "glossy_direct": 41, "glossy_indirect": 42,
        "glossy_color": 43, "transmission_direct": 44, "transmission_indirect": 45,
        "transmission_color": 46, "adaptive_aux_buffer": 13, "sample_count": 14}


def _f(v):
    return " ".join("%.9g" % float(x) for x in np.asarray(v, dtype=np.float64).ravel())


def _matrix_attr(m34):
    """cycles_xml.cpp:550-559 reads 16 floats and transposes: column-major."""
    m = np.eye(4)
    m[:3, :] = np.asarray(m34, dtype=np.float64).reshape(3, 4)
    return _f(m.T)


def look_at(eye, target, up=(0.0, 0.0, 1.0)):
    """Cycles camera space: +Z forward, +X right, +Y up (camera.cpp / Blender's
    -Z camera flipped by scale(1,1,-1) in blender_camera.cpp)."""
    eye = np.asarray(eye, dtype=np.float64)
    fwd = np.asarray(target, dtype=np.float64) - eye
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, np.asarray(up, dtype=np.float64))
    right /= np.linalg.norm(right)
    upv = np.cross(right, fwd)
    m = np.zeros((3, 4))
    m[:, 0], m[:, 1], m[:, 2], m[:, 3] = right, upv, fwd, eye
    return m


def euler_xyz_camera(loc, rot):
    """Blender camera object (XYZ euler, looks down -Z) -> Cycles matrix."""
    rx, ry, rz = rot
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    R = Rz @ Ry @ Rx
    m = np.zeros((3, 4))
    m[:, :3] = R @ np.diag([1.0, 1.0, -1.0])
    m[:, 3] = loc
    return m


def _header(width, height, cam_m34, fov, integrator, film="", nearclip=0.1, farclip=1000.0,
            cam_type="perspective", cam_extra=""):
    """cam_type: "perspective" or "orthograph" (cycles_xml's spelling); cam_extra: more
    camera attributes, e.g. 'aperturesize="0.2" focaldistance="9" blades="6"'."""
    return (
        '<camera width="%d" height="%d"/>\n'
        '<transform matrix="%s">\n'
        '  <camera type="%s" fov="%s" nearclip="%s" farclip="%s" shuttertime="-1" %s/>\n'
        "</transform>\n"
        "<integrator %s/>\n"
        '<film filter_type="blackman_harris" filter_width="1.5" exposure="1" %s/>\n'
        % (width, height, _matrix_attr(cam_m34), cam_type, _f(fov), _f(nearclip), _f(farclip),
           cam_extra, integrator, film)
    )


def _background(color, strength=1.0, attrs=""):
    return (
        "<background%s>\n" % attrs +
        '  <background name="bg" color="%s" strength="%s"/>\n'
        '  <connect from="bg background" to="output surface"/>\n'
        "</background>\n" % (_f(color), _f(strength))
    )


def _diffuse_shader(name, color):
    return (
        '<shader name="%s">\n'
        '  <diffuse_bsdf name="d" color="%s"/>\n'
        '  <connect from="d bsdf" to="output surface"/>\n'
        "</shader>\n" % (name, _f(color))
    )


def _emission_shader(name, color=(1, 1, 1), strength=1.0):
    return (
        '<shader name="%s">\n'
        '  <emission name="e" color="%s" strength="%s"/>\n'
        '  <connect from="e emission" to="output surface"/>\n'
        "</shader>\n" % (name, _f(color), _f(strength))
    )


def _principled_shader(name, base_color, metallic=0.0, roughness=0.5, specular=0.5,
                       transmission=0.0, ior=1.45, distribution="GGX"):
    return (
        '<shader name="%s">\n'
        '  <principled_bsdf name="p" distribution="%s" base_color="%s" metallic="%s" '
        'roughness="%s" specular="%s" transmission="%s" ior="%s"/>\n'
        '  <connect from="p bsdf" to="output surface"/>\n'
        "</shader>\n"
        % (name, distribution, _f(base_color), _f(metallic), _f(roughness), _f(specular),
           _f(transmission), _f(ior))
    )


def _procedural_shader(name, variant=0):
    """A material built from value nodes only (no textures): position stripes through
    separate / math, facing and fresnel weights, colour mix / brightness / gamma / invert,
    vector math + clamp, a diffuse + glossy-GGX mix.  `variant` changes the blend modes
    and math operators so that two shaders cover most of svm_nodes.cuh."""
    blend_a, blend_b, op_a, op_b, vop = [
        ("mix", "multiply", "sine", "multiply_add", "cross_product"),
        ("screen", "overlay", "pingpong", "smoothmin", "reflect"),
    ][variant]
    return (
        '<shader name="%s">\n'
        '  <geometry name="g"/>\n'
        '  <separate_xyz name="sep"/>\n'
        '  <connect from="g position" to="sep vector"/>\n'
        '  <math name="m1" type="multiply" value2="5.5"/>\n'
        '  <connect from="sep x" to="m1 value1"/>\n'
        '  <math name="m2" type="%s" value2="0.7"/>\n'
        '  <connect from="m1 value" to="m2 value1"/>\n'
        '  <math name="m3" type="%s" value2="0.45" value3="0.5"/>\n'
        '  <connect from="m2 value" to="m3 value1"/>\n'
        '  <vector_math name="vd" type="dot_product" vector2="0.3 0.5 0.8"/>\n'
        '  <connect from="g normal" to="vd vector1"/>\n'
        '  <vector_math name="vm" type="%s" vector2="0.3 0.5 0.8"/>\n'
        '  <connect from="g normal" to="vm vector1"/>\n'
        '  <vector_math name="vl" type="length"/>\n'
        '  <connect from="vm vector" to="vl vector1"/>\n'
        '  <math name="m4" type="add"/>\n'
        '  <connect from="vd value" to="m4 value1"/>\n'
        '  <connect from="vl value" to="m4 value2"/>\n'
        '  <clamp name="cl" type="minmax" min="0.1" max="0.9"/>\n'
        '  <connect from="m4 value" to="cl value"/>\n'
        '  <combine_xyz name="cmb" z="0.25"/>\n'
        '  <connect from="cl result" to="cmb x"/>\n'
        '  <connect from="m3 value" to="cmb y"/>\n'
        '  <mix name="mx" type="%s" color1="0.8 0.25 0.1" color2="0.1 0.35 0.8"/>\n'
        '  <connect from="m3 value" to="mx fac"/>\n'
        '  <layer_weight name="lw" blend="0.3"/>\n'
        '  <mix name="mx2" type="%s" fac="0.6"/>\n'
        '  <connect from="mx color" to="mx2 color1"/>\n'
        '  <connect from="cmb vector" to="mx2 color2"/>\n'
        '  <mix name="mx3" type="lighten" color2="0.3 0.3 0.3"/>\n'
        '  <connect from="lw facing" to="mx3 fac"/>\n'
        '  <connect from="mx2 color" to="mx3 color1"/>\n'
        '  <brightness_contrast name="bc" bright="0.04" contrast="0.15"/>\n'
        '  <connect from="mx3 color" to="bc color"/>\n'
        '  <light_path name="lp"/>\n'
        '  <math name="m6" type="multiply_add" value2="0.03" value3="0.01"/>\n'
        '  <connect from="lp ray_depth" to="m6 value1"/>\n'
        '  <connect from="m6 value" to="bc bright"/>\n'
        '  <gamma name="gm" gamma="1.3"/>\n'
        '  <connect from="bc color" to="gm color"/>\n'
        '  <invert name="inv" fac="0.1"/>\n'
        '  <connect from="gm color" to="inv color"/>\n'
        '  <diffuse_bsdf name="d"/>\n'
        '  <connect from="inv color" to="d color"/>\n'
        '  <glossy_bsdf name="gl" distribution="GGX" roughness="0.3" color="0.9 0.9 0.9"/>\n'
        '  <connect from="cl result" to="gl color"/>\n'
        '  <math name="m5" type="multiply_add" value2="0.3" value3="0.15"/>\n'
        '  <connect from="inv color" to="m5 value1"/>\n'
        '  <connect from="m5 value" to="gl roughness"/>\n'
        '  <fresnel name="fr" IOR="1.6"/>\n'
        '  <mix_closure name="mc"/>\n'
        '  <connect from="fr fac" to="mc fac"/>\n'
        '  <connect from="d bsdf" to="mc closure1"/>\n'
        '  <connect from="gl bsdf" to="mc closure2"/>\n'
        '  <connect from="mc closure" to="output surface"/>\n'
        "</shader>\n" % (name, op_a, op_b, vop, blend_a, blend_b)
    )


def _node_shader(name, body, out, attrs=""):
    return ('<shader name="%s"%s>\n%s  <connect from="%s" to="output surface"/>\n</shader>\n'
            % (name, attrs, body, out))


def _closure_shaders(variant):
    """Materials made of the BSDF nodes beyond Diffuse / Principled / Glossy-GGX: Glass
    (GGX and sharp), Refraction (GGX and sharp), sharp Glossy, Translucent, Oren-Nayar."""
    white = _node_shader("white", '  <diffuse_bsdf name="d" color="0.73 0.73 0.73" '
                         'roughness="0.6"/>\n', "d bsdf")
    red = _node_shader(
        "red", '  <diffuse_bsdf name="d" color="0.65 0.05 0.05" roughness="0.3"/>\n'
        '  <glossy_bsdf name="g" distribution="sharp" color="0.9 0.9 0.9"/>\n'
        '  <mix_closure name="m" fac="0.15"/>\n'
        '  <connect from="d bsdf" to="m closure1"/>\n  <connect from="g bsdf" to="m closure2"/>\n',
        "m closure")
    green = _node_shader(
        "green", '  <diffuse_bsdf name="d" color="0.12 0.45 0.15"/>\n'
        '  <translucent_bsdf name="t" color="0.3 0.6 0.3"/>\n'
        '  <mix_closure name="m" fac="0.4"/>\n'
        '  <connect from="d bsdf" to="m closure1"/>\n  <connect from="t bsdf" to="m closure2"/>\n',
        "m closure")
    if variant in (2, 3):
        # Transparent BSDF: one box half transparent over diffuse, one a tinted pure
        # transparent shell.  Variant 2 switches the shaders' transparent shadows off
        # (opaque shadow rays), variant 3 leaves them on (integrator.transparent_shadows).
        attrs = ' use_transparent_shadow="false"' if variant == 2 else ""
        metal = _node_shader(
            "metal", '  <diffuse_bsdf name="d" color="0.8 0.5 0.2"/>\n'
            '  <transparent_bsdf name="t" color="1 1 1"/>\n'
            '  <mix_closure name="m" fac="0.5"/>\n'
            '  <connect from="d bsdf" to="m closure1"/>\n'
            '  <connect from="t bsdf" to="m closure2"/>\n', "m closure", attrs)
        glass = _node_shader("glass", '  <transparent_bsdf name="t" color="0.75 0.9 1"/>\n',
                             "t bsdf", attrs)
    elif variant == 4:
        # the multi-scatter options of the Glossy and Glass nodes (random-walk lobes)
        metal = _node_shader("metal", '  <glossy_bsdf name="g" distribution="Multiscatter GGX" '
                             'roughness="0.45" color="0.9 0.7 0.3"/>\n', "g bsdf")
        glass = _node_shader("glass", '  <glass_bsdf name="g" distribution="Multiscatter GGX" '
                             'roughness="0.3" IOR="1.45" color="0.9 0.97 1"/>\n', "g bsdf")
    elif variant == 0:
        metal = _node_shader("metal", '  <glass_bsdf name="g" distribution="GGX" roughness="0.15" '
                             'IOR="1.45" color="0.95 0.97 1"/>\n', "g bsdf")
        glass = _node_shader("glass", '  <glass_bsdf name="g" distribution="sharp" IOR="1.5" '
                             'color="1 0.95 0.9"/>\n', "g bsdf")
    else:
        metal = _node_shader("metal", '  <refraction_bsdf name="g" distribution="GGX" '
                             'roughness="0.2" IOR="1.33" color="0.9 0.95 1"/>\n', "g bsdf")
        glass = _node_shader("glass", '  <refraction_bsdf name="g" distribution="sharp" IOR="1.2" '
                             'color="1 1 1"/>\n', "g bsdf")
    return white + red + green + metal + glass


def _textured_shaders(variant):
    """Materials driven by texture coordinates, mesh attributes and the procedural
    textures (svm_tex.cuh), plus the Principled features that need them: anisotropy
    with a tangent from the generated coordinates, sheen, clearcoat, and a mix of two
    Principled BSDFs (16 closures).  Two variants cover the texture types, noise
    dimensions, mapping types and coordinate sources between them."""
    c = lambda a, b: '  <connect from="%s" to="%s"/>\n' % (a, b)
    if variant == 0:
        white = _node_shader(
            "white", '  <texture_coordinate name="tc"/>\n'
            '  <checker_texture name="t" scale="3.0" color1="0.73 0.73 0.73" '
            'color2="0.3 0.32 0.4"/>\n' + c("tc generated", "t vector") +
            '  <diffuse_bsdf name="d"/>\n' + c("t color", "d color"), "d bsdf")
        red = _node_shader(
            "red", '  <brick_texture name="t" scale="2.5" color1="0.65 0.05 0.05" '
            'color2="0.4 0.2 0.05" mortar="0.6 0.6 0.6" mortar_size="0.03" mortar_smooth="0.4" '
            'bias="0.1" brick_width="0.6" row_height="0.3" offset="0.5" squash="0.8" '
            'squash_frequency="3" tex_mapping.rotation="0 1.5707963 0"/>\n'
            '  <diffuse_bsdf name="d"/>\n' + c("t color", "d color"), "d bsdf")
        green = _node_shader(
            "green", '  <texture_coordinate name="tc"/>\n'
            '  <mapping name="mp" type="point" location="0.2 0.1 0" rotation="0.3 0.2 0.5" '
            'scale="1.5 1 2"/>\n' + c("tc object", "mp vector") +
            '  <wave_texture name="t" type="rings" rings_direction="spherical" profile="sine" '
            'scale="0.4" distortion="2.5" detail="2.5" detail_scale="1.5" '
            'detail_roughness="0.6" phase="0.3"/>\n' + c("mp vector", "t vector") +
            '  <mix name="mx" type="mix" color1="0.12 0.45 0.15" color2="0.6 0.7 0.2"/>\n' +
            c("t fac", "mx fac") + '  <diffuse_bsdf name="d"/>\n' + c("mx color", "d color"),
            "d bsdf")
        metal = _node_shader(
            "metal", '  <noise_texture name="t" dimensions="4D" scale="3.0" w="0.7" detail="3.0" '
            'roughness="0.55" distortion="0.6"/>\n'
            '  <principled_bsdf name="p" distribution="GGX" metallic="1.0" roughness="0.35" '
            'anisotropic="0.7" anisotropic_rotation="0.15" specular="0.5"/>\n' +
            c("t color", "p base_color"), "p bsdf")
        glass = _node_shader(
            "glass", '  <magic_texture name="t" depth="4" scale="2.0" distortion="1.5"/>\n'
            '  <gradient_texture name="gr" type="spherical"/>\n'
            '  <principled_bsdf name="p1" distribution="GGX" roughness="0.6" sheen="1.0" '
            'sheen_tint="0.5" specular="0.3"/>\n' + c("t color", "p1 base_color") +
            '  <principled_bsdf name="p2" distribution="GGX" base_color="0.1 0.2 0.7" '
            'roughness="0.4" clearcoat="1.0" clearcoat_roughness="0.05" metallic="0.3"/>\n'
            '  <mix_closure name="m"/>\n' + c("gr fac", "m fac") +
            c("p1 bsdf", "m closure1") + c("p2 bsdf", "m closure2"), "m closure")
    elif variant == 3:
        # Voronoi (every feature, 1D-4D, all metrics) and Musgrave (all five types)
        white = _node_shader(
            "white", '  <texture_coordinate name="tc"/>\n'
            '  <voronoi_texture name="t" dimensions="3D" feature="f1" metric="euclidean" '
            'scale="3.0" randomness="0.9"/>\n' + c("tc generated", "t vector") +
            '  <math name="tk" type="multiply_add" value2="9000" value3="1000"/>\n' +
            c("t distance", "tk value1") + '  <blackbody name="bb"/>\n' +
            c("tk value", "bb temperature") +
            '  <mix name="mx0" type="multiply" fac="0.3"/>\n' + c("t color", "mx0 color1") +
            c("bb color", "mx0 color2") +
            '  <mix name="mx" type="mix" color1="0.73 0.73 0.73" fac="0.5"/>\n' +
            c("mx0 color", "mx color2") + '  <diffuse_bsdf name="d"/>\n' + c("mx color", "d color"),
            "d bsdf")
        red = _node_shader(
            "red", '  <texture_coordinate name="tc"/>\n'
            '  <voronoi_texture name="t" dimensions="2D" feature="smooth_f1" metric="manhattan" '
            'scale="4.0" smoothness="0.6"/>\n' + c("tc object", "t vector") +
            '  <voronoi_texture name="t1" dimensions="1D" feature="f2" scale="6.0"/>\n'
            '  <separate_xyz name="sep"/>\n' + c("tc object", "sep vector") + c("sep z", "t1 w") +
            '  <mix name="mx" type="mix" color1="0.65 0.05 0.05" color2="0.9 0.7 0.2"/>\n' +
            c("t distance", "mx fac") +
            '  <mix name="mx2" type="multiply" fac="0.7"/>\n' + c("mx color", "mx2 color1") +
            c("t1 color", "mx2 color2") + '  <diffuse_bsdf name="d"/>\n' + c("mx2 color", "d color"),
            "d bsdf")
        green = _node_shader(
            "green", '  <texture_coordinate name="tc"/>\n'
            '  <musgrave_texture name="t" dimensions="3D" type="fBM" scale="2.5" detail="3.5" '
            'dimension="1.2" lacunarity="2.1"/>\n' + c("tc object", "t vector") +
            '  <musgrave_texture name="t2" dimensions="2D" type="multifractal" scale="3.0" '
            'detail="2.0" dimension="0.8" lacunarity="1.9"/>\n' + c("tc object", "t2 vector") +
            '  <math name="m" type="multiply_add" value2="0.3" value3="0.5"/>\n' +
            c("t fac", "m value1") + '  <math name="m2" type="multiply" value2="0.4"/>\n' +
            c("t2 fac", "m2 value1") + '  <combine_xyz name="cmb" x="0.12"/>\n' +
            c("m value", "cmb y") + c("m2 value", "cmb z") +
            '  <math name="nm" type="multiply_add" value2="150" value3="520"/>\n' +
            c("t fac", "nm value1") + '  <wavelength name="wl"/>\n' + c("nm value", "wl wavelength") +
            '  <mix name="mxw" type="add" fac="0.3"/>\n' + c("cmb vector", "mxw color1") +
            c("wl color", "mxw color2") + '  <diffuse_bsdf name="d"/>\n' +
            c("mxw color", "d color"), "d bsdf")
        metal = _node_shader(
            "metal", '  <texture_coordinate name="tc"/>\n'
            '  <voronoi_texture name="t" dimensions="3D" feature="distance_to_edge" scale="3.0"/>\n' +
            c("tc object", "t vector") +
            '  <voronoi_texture name="t4" dimensions="4D" feature="n_sphere_radius" scale="2.5" '
            'w="0.4"/>\n' + c("tc object", "t4 vector") +
            '  <voronoi_texture name="t3" dimensions="3D" feature="f2" metric="chebychev" '
            'scale="2.0"/>\n' + c("tc object", "t3 vector") +
            '  <combine_xyz name="cmb"/>\n' + c("t distance", "cmb x") + c("t4 radius", "cmb y") +
            c("t3 distance", "cmb z") +
            '  <anisotropic_bsdf name="g" distribution="GGX" roughness="0.3" anisotropy="0.6" '
            'rotation="0.2"/>\n' + c("cmb vector", "g color"), "g bsdf")
        glass = _node_shader(
            "glass", '  <texture_coordinate name="tc"/>\n'
            '  <musgrave_texture name="t" dimensions="4D" type="ridged_multifractal" scale="2.0" '
            'detail="3.0" dimension="1.0" lacunarity="2.0" offset="1.0" gain="2.0" w="0.3"/>\n' +
            c("tc generated", "t vector") +
            '  <musgrave_texture name="t2" dimensions="2D" type="hetero_terrain" scale="2.0" '
            'detail="2.5" dimension="1.0" lacunarity="2.0" offset="0.5"/>\n' +
            c("tc generated", "t2 vector") +
            '  <musgrave_texture name="t3" dimensions="1D" type="hybrid_multifractal" scale="3.0" '
            'detail="2.0" dimension="1.0" lacunarity="2.0" offset="0.6" gain="1.5"/>\n'
            '  <separate_xyz name="sep"/>\n' + c("tc generated", "sep vector") + c("sep x", "t3 w") +
            '  <voronoi_texture name="v" dimensions="4D" feature="f1" metric="minkowski" '
            'exponent="1.5" scale="2.0" w="0.2"/>\n' + c("tc generated", "v vector") +
            '  <combine_xyz name="cmb"/>\n' + c("t fac", "cmb x") + c("t2 fac", "cmb y") +
            c("t3 fac", "cmb z") +
            '  <mix name="mx" type="mix" fac="0.5"/>\n' + c("cmb vector", "mx color1") +
            c("v position", "mx color2") +
            '  <vector_math name="ab" type="absolute"/>\n' + c("mx color", "ab vector1") +
            '  <vector_math name="fr" type="fraction"/>\n' + c("ab vector", "fr vector1") +
            '  <principled_bsdf name="p" distribution="GGX" roughness="0.5"/>\n' +
            c("fr vector", "p base_color"), "p bsdf")
    elif variant == 2:
        # colour / range / vector / info nodes: HSV, separate + combine HSV, the HSV and
        # dodge / burn blend modes, map range (all four kinds), normal, vector rotate,
        # vector transform, object info, camera data, white noise
        white = _node_shader(
            "white", '  <texture_coordinate name="tc"/>\n'
            '  <checker_texture name="t" scale="2.0" color1="0.7 0.3 0.2" color2="0.2 0.5 0.7"/>\n' +
            c("tc generated", "t vector") + '  <geometry name="g"/>\n'
            '  <separate_xyz name="sep"/>\n' + c("g position", "sep vector") +
            '  <map_range name="mr" type="smoothstep" from_min="-1" from_max="1" to_min="0.2" '
            'to_max="0.8"/>\n' + c("sep x", "mr value") +
            '  <hsv name="h" saturation="0.8" value="0.9" fac="0.9"/>\n' + c("t color", "h color") +
            c("mr result", "h hue") +
            # a linked BSDF normal (bent shading normal): the bump shadowing term of
            # bsdf_eval / bsdf_sample
            '  <vector_math name="bn" type="add" vector2="0.25 0.15 0.0"/>\n' +
            c("g normal", "bn vector1") + '  <vector_math name="bnn" type="normalize"/>\n' +
            c("bn vector", "bnn vector1") +
            '  <diffuse_bsdf name="d"/>\n' + c("h color", "d color") + c("bnn vector", "d normal"),
            "d bsdf")
        chain = ""
        prev = "t color"
        for i, mode in enumerate(("hue", "saturation", "value", "color", "dodge", "burn")):
            col2 = ("0.2 0.6 0.9", "0.5 0.5 0.2", "0.3 0.3 0.3", "0.9 0.4 0.1", "0.3 0.2 0.1",
                    "0.8 0.9 0.7")[i]
            chain += '  <mix name="m%d" type="%s" fac="0.6" color2="%s"/>\n' % (i, mode, col2)
            chain += c(prev, "m%d color1" % i)
            prev = "m%d color" % i
        red = _node_shader(
            "red", '  <wave_texture name="t" type="bands" bands_direction="z" profile="sine" '
            'scale="0.6"/>\n  <mix name="base" type="mix" color1="0.65 0.05 0.05" '
            'color2="0.1 0.3 0.6"/>\n' + c("t fac", "base fac") +
            chain.replace('"t color"', '"base color"') +
            '  <diffuse_bsdf name="d"/>\n' + c(prev, "d color"), "d bsdf")
        green = _node_shader(
            "green", '  <texture_coordinate name="tc"/>\n'
            '  <vector_math name="sc" type="scale" scale="3.0"/>\n' + c("tc object", "sc vector1") +
            '  <vector_math name="fl" type="floor"/>\n' + c("sc vector", "fl vector1") +
            '  <white_noise_texture name="t" dimensions="3D"/>\n' + c("fl vector", "t vector") +
            '  <separate_hsv name="sh"/>\n' + c("t color", "sh color") +
            '  <math name="ms" type="multiply" value2="0.5"/>\n' + c("sh s", "ms value1") +
            '  <white_noise_texture name="t4" dimensions="4D" w="0.37"/>\n' +
            c("fl vector", "t4 vector") +
            '  <map_range name="mr" type="stepped" from_min="0" from_max="1" to_min="0.3" '
            'to_max="0.9" steps="3"/>\n' + c("t4 value", "mr value") +
            '  <combine_hsv name="ch"/>\n' + c("sh h", "ch h") + c("ms value", "ch s") +
            c("mr result", "ch v") + '  <diffuse_bsdf name="d"/>\n' + c("ch color", "d color"),
            "d bsdf")
        metal = _node_shader(
            "metal", '  <object_info name="oi"/>\n  <camera_info name="ci"/>\n'
            '  <map_range name="mr" type="smootherstep" from_min="2.5" from_max="4.5" '
            'to_min="0.1" to_max="0.9"/>\n' + c("ci view_z_depth", "mr value") +
            '  <map_range name="mr2" type="linear" from_min="3" from_max="5"/>\n' +
            c("ci view_distance", "mr2 value") +
            '  <geometry name="g"/>\n'
            '  <vector_rotate name="vr" type="euler_xyz" rotation="0.4 0.2 0.7" '
            'center="0.1 0 0.2"/>\n' + c("g normal", "vr vector") +
            '  <vector_transform name="vt" type="normal" convert_from="world" '
            'convert_to="camera"/>\n' + c("vr vector", "vt vector") +
            '  <vector_math name="ab" type="absolute"/>\n' + c("vt vector", "ab vector1") +
            '  <combine_xyz name="cmb"/>\n' + c("mr result", "cmb x") + c("oi random", "cmb y") +
            c("mr2 result", "cmb z") +
            '  <mix name="mx" type="mix" fac="0.5"/>\n' + c("ab vector", "mx color1") +
            c("cmb vector", "mx color2") +
            '  <vector_math name="ad" type="add"/>\n' + c("mx color", "ad vector1") +
            c("oi location", "ad vector2") +
            '  <vector_math name="fr" type="fraction"/>\n' + c("ad vector", "fr vector1") +
            '  <tangent name="tg" direction_type="radial" axis="y"/>\n'
            '  <anisotropic_bsdf name="gl" distribution="GGX" roughness="0.3" anisotropy="-0.5"/>\n' +
            c("fr vector", "gl color") + c("tg tangent", "gl tangent"), "gl bsdf")
        glass = _node_shader(
            "glass", '  <geometry name="g"/>\n'
            '  <normal name="nn" direction="0.3 -0.5 0.8"/>\n' + c("g normal", "nn normal") +
            '  <camera_info name="ci"/>\n'
            '  <vector_rotate name="vr" type="axis" axis="0.2 0.5 1.0" invert="true"/>\n' +
            c("g position", "vr vector") + c("ci view_distance", "vr angle") +
            '  <vector_rotate name="vr2" type="y_axis" angle="0.8"/>\n' + c("vr vector", "vr2 vector") +
            '  <vector_transform name="vt" type="point" convert_from="object" '
            'convert_to="camera"/>\n' + c("vr2 vector", "vt vector") +
            '  <vector_transform name="vt2" type="vector" convert_from="camera" '
            'convert_to="object"/>\n' + c("ci view_vector", "vt2 vector") +
            '  <vector_math name="ad" type="add"/>\n' + c("vt vector", "ad vector1") +
            c("vt2 vector", "ad vector2") +
            '  <vector_math name="fr" type="fraction"/>\n' + c("ad vector", "fr vector1") +
            '  <mix name="mx" type="mix" color2="0.9 0.9 0.9"/>\n' + c("fr vector", "mx color1") +
            c("nn dot", "mx fac") +
            '  <normal_map name="nmap" space="object" strength="0.7"/>\n' +
            c("fr vector", "nmap color") +
            '  <principled_bsdf name="p" distribution="GGX" roughness="0.4" sheen="0.5"/>\n' +
            c("mx color", "p base_color") + c("nmap normal", "p normal"), "p bsdf")
    else:
        white = _node_shader(
            "white", '  <texture_coordinate name="tc"/>\n'
            '  <noise_texture name="t" dimensions="2D" scale="4.0" detail="2.0" '
            'roughness="0.5"/>\n' + c("tc uv", "t vector") +
            '  <mix name="mx" type="mix" color1="0.73 0.73 0.73" color2="0.35 0.3 0.25"/>\n' +
            c("t fac", "mx fac") + '  <diffuse_bsdf name="d"/>\n' + c("mx color", "d color"),
            "d bsdf")
        red = _node_shader(
            "red", '  <texture_coordinate name="tc"/>\n'
            '  <wave_texture name="t" type="bands" bands_direction="diagonal" profile="saw" '
            'scale="1.2" tex_mapping.scale="1 2 1" tex_mapping.use_minmax="true" '
            'tex_mapping.min="-0.8 -0.8 -0.8" tex_mapping.max="0.8 0.8 0.8"/>\n' +
            c("tc normal", "t vector") +
            '  <gradient_texture name="gr" type="radial"/>\n' + c("tc camera", "gr vector") +
            '  <mix name="mx" type="mix" color1="0.65 0.05 0.05" color2="0.7 0.5 0.1"/>\n' +
            c("t fac", "mx fac") +
            '  <mix name="mx2" type="multiply" color2="0.8 0.8 0.8" fac="0.5"/>\n' +
            c("mx color", "mx2 color1") + c("gr color", "mx2 color2") +
            '  <diffuse_bsdf name="d"/>\n' + c("mx2 color", "d color"), "d bsdf")
        green = _node_shader(
            "green", '  <attribute name="at" attribute="UVMap"/>\n'
            '  <mapping name="mp" type="texture" location="0.1 0.2 0" rotation="0 0 0.4" '
            'scale="0.5 0.25 1"/>\n' + c("at vector", "mp vector") +
            '  <checker_texture name="t" scale="2.0" color1="0.12 0.45 0.15" '
            'color2="0.5 0.6 0.1"/>\n' + c("mp vector", "t vector") +
            '  <noise_texture name="n1" dimensions="1D" scale="3.0" detail="1.5" '
            'distortion="0.3"/>\n' + c("at fac", "n1 w") +
            '  <mix name="mx" type="mix" color2="0.9 0.9 0.9"/>\n' + c("t color", "mx color1") +
            c("n1 fac", "mx fac") + '  <diffuse_bsdf name="d"/>\n' + c("mx color", "d color"),
            "d bsdf")
        metal = _node_shader(
            "metal", '  <texture_coordinate name="tc"/>\n'
            '  <mapping name="mp" type="normal" rotation="0.2 0.1 0" scale="1 2 1"/>\n' +
            c("tc reflection", "mp vector") +
            '  <wave_texture name="t" type="rings" rings_direction="z" profile="tri" '
            'scale="0.5"/>\n' + c("mp vector", "t vector") +
            '  <noise_texture name="n3" dimensions="3D" scale="2.0" detail="2.5" '
            'roughness="0.7" distortion="0.4"/>\n'
            '  <mix name="mx" type="mix" fac="0.5"/>\n' + c("t color", "mx color1") +
            c("n3 color", "mx color2") +
            '  <glossy_bsdf name="g" distribution="GGX" roughness="0.25"/>\n' +
            c("mx color", "g color"), "g bsdf")
        glass = _node_shader(
            "glass", '  <texture_coordinate name="tc"/>\n'
            '  <mapping name="mp" type="vector" rotation="0 0.5 0" scale="8 8 8"/>\n' +
            c("tc window", "mp vector") +
            '  <gradient_texture name="g1" type="easing"/>\n' + c("mp vector", "g1 vector") +
            '  <gradient_texture name="g2" type="quadratic_sphere"/>\n'
            '  <gradient_texture name="g3" type="diagonal"/>\n' + c("tc object", "g3 vector") +
            '  <combine_xyz name="cmb"/>\n' + c("g1 fac", "cmb x") + c("g2 fac", "cmb y") +
            c("g3 fac", "cmb z") +
            '  <magic_texture name="t" depth="9" scale="1.5" distortion="0.8"/>\n' +
            c("tc generated", "t vector") +
            '  <mix name="mx" type="add" fac="0.4"/>\n' + c("cmb vector", "mx color1") +
            c("t color", "mx color2") +
            '  <principled_bsdf name="p" distribution="GGX" roughness="0.5" sheen="0.8" '
            'clearcoat="0.5" anisotropic="0.4"/>\n' + c("mx color", "p base_color"), "p bsdf")
    return white + red + green + metal + glass



# ---------------------------------------------------------------- image textures
_B2IM_KINDS = {np.dtype(np.uint8): 0, np.dtype(np.float32): 1, np.dtype(np.float16): 2,
               np.dtype(np.uint16): 3}


def write_b2im(path, pixels):
    """The raw image container of this repo's scenes: "B2IM", uint32 width, height,
    channels, kind (0 uint8, 1 float32, 2 float16, 3 uint16), then the rows bottom-up.
    (There is no OpenImageIO in the image; the test harness's stand-in for the reference's
    image loader reads this, oracle/ref_stubs.cpp.)"""
    a = np.ascontiguousarray(pixels)
    if a.ndim == 2:
        a = a[:, :, None]
    h, w, c = a.shape
    with open(path, "wb") as f:
        f.write(b"B2IM")
        f.write(np.array([w, h, c, _B2IM_KINDS[a.dtype]], np.uint32).tobytes())
        f.write(a.tobytes())


def test_image(kind, width, height, channels, seed):
    """A small deterministic test image: smooth colour ramps, a coarse checker and noise,
    alpha (when there is a fourth channel) with fully transparent, partial and opaque
    regions.  `kind`: "u8" | "u16" | "f16" | "f32" (float images reach above 1)."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:height, 0:width]
    u, v = (x + 0.5) / width, (y + 0.5) / height
    chan = [0.5 + 0.5 * np.sin(6.0 * u + 2.0 * v + seed), v * (0.3 + 0.7 * u),
            0.25 + 0.75 * (((x // 3) + (y // 2)) % 2)]
    img = np.stack(chan[:max(1, min(channels, 3))], axis=-1)
    img = np.clip(img + rng.uniform(-0.08, 0.08, img.shape), 0.0, 1.0)
    if channels == 2 or channels == 4:
        alpha = np.clip(1.6 * np.abs(np.sin(3.0 * u * np.pi) * np.cos(2.0 * v * np.pi)) - 0.15,
                        0.0, 1.0)
        img = np.concatenate([img[..., :channels - 1], alpha[..., None]], axis=-1)
    if kind == "u8":
        return np.round(img * 255.0).astype(np.uint8)
    if kind == "u16":
        return np.round(img * 65535.0).astype(np.uint16)
    if kind == "f16":
        return (img * 1.5).astype(np.float16)
    return (img * 2.0).astype(np.float32)


def _image_shaders(variant):
    """Cornell materials that read image textures (svm_image.cuh): every pixel format,
    interpolation, extension and projection between the two variants, alpha handling,
    UDIM tiles.  Returns (xml, {file name: pixels})."""
    c = lambda a, b: '  <connect from="%s" to="%s"/>\n' % (a, b)
    images = {}
    if variant == 0:
        images["rgba8.b2im"] = test_image("u8", 37, 23, 4, 1)
        images["rgb32f.b2im"] = test_image("f32", 32, 32, 3, 2)
        images["gray8.b2im"] = test_image("u8", 16, 12, 1, 3)
        images["rgba16f.b2im"] = test_image("f16", 24, 20, 4, 4)
        images["tile.1001.b2im"] = test_image("u8", 16, 16, 3, 5)
        images["tile.1002.b2im"] = test_image("u8", 20, 12, 4, 6)
        white = _node_shader(
            "white", '  <texture_coordinate name="tc"/>\n'
            '  <image_texture name="t" filename="rgba8.b2im" colorspace="sRGB" '
            'interpolation="linear" extension="periodic" tex_mapping.scale="1.7 2.3 1"/>\n' +
            c("tc generated", "t vector") +
            '  <diffuse_bsdf name="d"/>\n' + c("t color", "d color"), "d bsdf")
        red = _node_shader(
            "red", '  <texture_coordinate name="tc"/>\n'
            '  <image_texture name="t" filename="rgb32f.b2im" interpolation="cubic" '
            'extension="clamp" projection="box" projection_blend="0.35"/>\n' +
            c("tc object", "t vector") +
            '  <mix name="mx" type="multiply" fac="1.0" color2="0.65 0.2 0.2"/>\n' +
            c("t color", "mx color1") +
            '  <diffuse_bsdf name="d"/>\n' + c("mx color", "d color"), "d bsdf")
        green = _node_shader(
            "green", '  <texture_coordinate name="tc"/>\n'
            '  <image_texture name="t" filename="gray8.b2im" colorspace="Raw" '
            'interpolation="closest" extension="black" '
            'tex_mapping.scale="0.8 0.7 1" tex_mapping.location="-0.1 -0.2 0"/>\n' +
            c("tc generated", "t vector") +
            '  <mix name="mx" type="mix" color1="0.12 0.45 0.15" color2="0.8 0.8 0.3"/>\n' +
            c("t color", "mx fac") +
            '  <diffuse_bsdf name="d"/>\n' + c("mx color", "d color"), "d bsdf")
        metal = _node_shader(
            "metal", '  <texture_coordinate name="tc"/>\n'
            '  <image_texture name="t" filename="rgba16f.b2im" interpolation="linear" '
            'extension="periodic" projection="sphere"/>\n' + c("tc generated", "t vector") +
            '  <principled_bsdf name="p" distribution="GGX" metallic="1.0" specular="0.5"/>\n' +
            c("t color", "p base_color") + c("t alpha", "p roughness"), "p bsdf")
        glass = _node_shader(
            "glass", '  <texture_coordinate name="tc"/>\n'
            '  <image_texture name="t" filename="tile.&lt;UDIM&gt;.b2im" '
            'colorspace="sRGB" interpolation="linear" extension="clamp" '
            'tex_mapping.scale="2.5 1.2 1" tex_mapping.location="-0.1 0 0"/>\n' +
            c("tc generated", "t vector") +
            '  <diffuse_bsdf name="d"/>\n' + c("t color", "d color"), "d bsdf")
    else:
        images["rgba8.b2im"] = test_image("u8", 29, 31, 4, 11)
        images["rgba16.b2im"] = test_image("u16", 18, 22, 4, 12)
        images["gray16.b2im"] = test_image("u16", 16, 16, 1, 13)
        images["gray32f.b2im"] = test_image("f32", 12, 20, 1, 14)
        images["gray16f.b2im"] = test_image("f16", 14, 14, 1, 15)
        images["rgb32f.b2im"] = test_image("f32", 20, 16, 3, 16)
        white = _node_shader(
            "white", '  <texture_coordinate name="tc"/>\n'
            '  <image_texture name="t" filename="rgba8.b2im" colorspace="sRGB" '
            'alpha_type="channel_packed" interpolation="smart" extension="black" '
            'tex_mapping.scale="1.3 1.3 1" tex_mapping.location="-0.15 -0.1 0"/>\n' +
            c("tc generated", "t vector") +
            '  <mix name="mx" type="mix" color1="0.73 0.73 0.73"/>\n' + c("t color", "mx color2") +
            c("t alpha", "mx fac") +
            '  <diffuse_bsdf name="d"/>\n' + c("mx color", "d color"), "d bsdf")
        red = _node_shader(
            "red", '  <texture_coordinate name="tc"/>\n'
            '  <image_texture name="t" filename="rgba16.b2im" colorspace="Raw" '
            'interpolation="linear" extension="clamp" projection="tube"/>\n' +
            c("tc generated", "t vector") +
            '  <image_texture name="t2" filename="gray16.b2im" colorspace="Raw" '
            'interpolation="cubic" extension="periodic" tex_mapping.scale="3 3 1"/>\n' +
            c("tc generated", "t2 vector") +
            '  <mix name="mx" type="multiply" fac="1.0"/>\n' + c("t color", "mx color1") +
            c("t2 color", "mx color2") +
            '  <diffuse_bsdf name="d"/>\n' + c("mx color", "d color"), "d bsdf")
        green = _node_shader(
            "green", '  <texture_coordinate name="tc"/>\n'
            '  <image_texture name="t" filename="gray32f.b2im" interpolation="linear" '
            'extension="black" projection="box" projection_blend="0.0"/>\n' +
            c("tc object", "t vector") +
            '  <image_texture name="t2" filename="gray16f.b2im" interpolation="closest" '
            'extension="periodic" tex_mapping.scale="4 4 1"/>\n' + c("tc generated", "t2 vector") +
            '  <mix name="mx" type="mix" color1="0.12 0.45 0.15" color2="0.1 0.2 0.6"/>\n' +
            c("t color", "mx fac") +
            '  <mix name="mx2" type="multiply" fac="0.7"/>\n' + c("mx color", "mx2 color1") +
            c("t2 color", "mx2 color2") +
            '  <diffuse_bsdf name="d"/>\n' + c("mx2 color", "d color"), "d bsdf")
        metal = _node_shader(
            "metal", '  <texture_coordinate name="tc"/>\n'
            '  <image_texture name="t" filename="rgb32f.b2im" colorspace="sRGB" '
            'interpolation="cubic" extension="black" projection="box" projection_blend="1.0"/>\n' +
            c("tc object", "t vector") +
            '  <principled_bsdf name="p" distribution="GGX" metallic="0.6" roughness="0.3" '
            'specular="0.5"/>\n' + c("t color", "p base_color"), "p bsdf")
        glass = _node_shader(
            "glass", '  <texture_coordinate name="tc"/>\n'
            '  <image_texture name="t" filename="rgba8.b2im" colorspace="sRGB" '
            'alpha_type="ignore" interpolation="closest" extension="clamp"/>\n' +
            c("tc uv", "t vector") +
            '  <image_texture name="missing" filename="does_not_exist.b2im"/>\n' +
            c("tc uv", "missing vector") +
            '  <mix name="mx" type="mix" fac="0.25"/>\n' + c("t color", "mx color1") +
            c("missing color", "mx color2") +
            '  <diffuse_bsdf name="d"/>\n' + c("mx color", "d color"), "d bsdf")
    tiles = [("glass", "t", [1001, 1002])] if variant == 0 else []
    return white + red + green + metal + glass, images, tiles


def _integrator(max_bounce, diffuse=None, glossy=None, transmission=None, transparent=8,
                clamp_indirect=0.0, seed=0, light_threshold=0.01, caustics=True,
                pattern="sobol", aa_samples=0):
    diffuse = max_bounce if diffuse is None else diffuse
    glossy = max_bounce if glossy is None else glossy
    transmission = max_bounce if transmission is None else transmission
    return (
        'method="path" sampling_pattern="%s" aa_samples="%d" seed="%d" min_bounce="0" max_bounce="%d" '
        'max_diffuse_bounce="%d" max_glossy_bounce="%d" max_transmission_bounce="%d" '
        'transparent_min_bounce="0" transparent_max_bounce="%d" sample_clamp_direct="0" '
        'sample_clamp_indirect="%s" light_sampling_threshold="%s" caustics_reflective="%s" '
        'caustics_refractive="%s" filter_glossy="0"'
        % (pattern, aa_samples, seed, max_bounce, diffuse, glossy, transmission, transparent,
           _f(clamp_indirect),
           _f(light_threshold), "true" if caustics else "false", "true" if caustics else "false")
    )


def _state(shader, body):
    return '<state shader="%s">\n%s</state>\n' % (shader, body)


CUBE_VERTS = np.array(
    [[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1],
     [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=np.float32)
# outward-facing quads (counter-clockwise seen from outside)
CUBE_QUADS = np.array(
    [[0, 3, 2, 1], [4, 5, 6, 7], [0, 1, 5, 4], [1, 2, 6, 5], [2, 3, 7, 6], [3, 0, 4, 7]],
    dtype=np.int32)


def quads_to_tris(quads):
    q = np.asarray(quads, dtype=np.int32)
    return np.concatenate([q[:, [0, 1, 2]], q[:, [0, 2, 3]]], axis=1).reshape(-1, 3)


def box_mesh(lo, hi):
    lo, hi = np.asarray(lo, np.float32), np.asarray(hi, np.float32)
    P = lo + (CUBE_VERTS * 0.5 + 0.5) * (hi - lo)
    return P.astype(np.float32), quads_to_tris(CUBE_QUADS)


# ---------------------------------------------------------------- config 1
def default_cube(width=1920, height=1080, spp=64, material="principled", max_bounce=12,
                 lights="point", cam_type="perspective", cam_extra="", distribution="GGX",
                 world="grey", world_light=0):
    """BASELINE config 1 - Blender's startup scene, values extracted from
    release/datafiles/startup.blend.  `lights`: "point" (the
    startup scene), "falloff" (its lamp shader goes through a Light Falloff node), "spot"
    (the same lamp as a 50 degree spot aimed at the cube, soft
    edge) or "mixed" (point + round area + sun: three entries in the light
    distribution) - variants used by the parity tests only.  `world`: "grey" (the
    startup scene), "env_equirect" / "env_mirrorball" (an Environment Texture node lights
    the scene: a float image looked up by the ray direction).  `world_light` > 0 adds a
    background light of that map resolution: the world enters the light distribution and
    is importance-sampled by its luminance map (background MIS, kernel_light_background.h),
    which the host builds from one DeviceTask::SHADER evaluation of the world shader."""
    cam = euler_xyz_camera((7.358891, -6.925791, 4.958309), (1.109319, 0.0, 0.814928))
    fov = 2.0 * np.arctan(0.5 * 36.0 / 50.0 / (width / height))
    xml = "<cycles>\n"
    xml += _header(
        width, height, cam, fov,
        _integrator(max_bounce, diffuse=min(4, max_bounce), glossy=min(4, max_bounce),
                    transmission=max_bounce, clamp_indirect=10.0),
        nearclip=0.1, farclip=100.0, cam_type=cam_type, cam_extra=cam_extra)
    images = {}
    if world == "grey":
        xml += _background((0.05087609, 0.05087609, 0.05087609))
    else:
        images["env.b2im"] = test_image("f32", 48, 24, 3, 21)
        xml += ("<background>\n"
                '  <environment_texture name="env" filename="env.b2im" interpolation="%s" '
                'projection="%s"/>\n'
                '  <background name="bg" strength="0.6"/>\n'
                '  <connect from="env color" to="bg color"/>\n'
                '  <connect from="bg background" to="output surface"/>\n'
                "</background>\n"
                % (("linear", "equirectangular") if world == "env_equirect" else
                   ("cubic", "mirror_ball")))
    if material == "principled":
        xml += _principled_shader("cube", (0.8, 0.8, 0.8), 0.0, 0.5, 0.5,
                                  distribution=distribution)
    else:
        xml += _diffuse_shader("cube", (0.8, 0.8, 0.8))
    if lights == "falloff":
        # lamp shader with a Light Falloff node: evaluated per light sample, not constant
        xml += _node_shader(
            "lamp", '  <light_falloff name="lf" strength="1.4" smooth="2.5"/>\n'
            '  <emission name="e" color="1 0.95 0.9"/>\n'
            '  <connect from="lf linear" to="e strength"/>\n', "e emission")
    else:
        xml += _emission_shader("lamp", (1, 1, 1), 1.0)
    co = np.array([4.076245, 1.005454, 5.903862])
    aim = -co / np.linalg.norm(co)
    point = ('<light type="point" co="%s" size="0.1" strength="1000 1000 1000" '
             'use_mis="true"/>\n' % " ".join(_f(c) for c in co))
    if lights in ("point", "falloff"):
        body = point
    elif lights == "spot":
        body = ('<light type="spot" co="%s" dir="%s" spot_angle="%s" spot_smooth="0.25" size="0.1" '
                'strength="3000 3000 3000" use_mis="true"/>\n'
                % (" ".join(_f(c) for c in co), " ".join(_f(c) for c in aim),
                   _f(np.radians(50.0))))
    elif lights == "mixed":
        body = point
        body += ('<light type="area" co="-3 -2 4" dir="0.5570860 0.3713907 -0.7427814" '
                 'axisu="0.5547002 -0.8320503 0" axisv="0.6180207 0.4120138 0.6695225" '
                 'sizeu="1.5" sizev="1" size="1" strength="400 380 350" use_mis="true"/>\n')
        body += ('<light type="distant" dir="0.3 0.2 -0.9327379" angle="0.02" '
                 'strength="1.5 1.5 1.6" use_mis="true"/>\n')
    else:
        raise ValueError(lights)
    xml += _state("lamp", body)
    if world_light:
        xml += _state("default_background",
                      '<light type="background" map_resolution="%d" use_mis="true" '
                      'strength="1 1 1"/>\n' % int(world_light))
    xml += "</cycles>\n"
    P, tris = box_mesh((-1, -1, -1), (1, 1, 1))
    multi = "_multiscatter" if (material == "principled" and distribution != "GGX") else ""
    return SceneDesc(
        "default_cube_" + material + multi + ("" if lights == "point" else "_" + lights) +
        ("" if world == "grey" else "_" + world) + ("_mis" if world_light else ""), xml,
        width, height,
        meshes=[MeshDesc(P, tris, "cube")], objects=[(0, np.eye(4, dtype=np.float32)[:3])],
        spp=spp, notes="config 1", images=images)


# ---------------------------------------------------------------- config 2
def value_noise_height(n, seed=1234, octaves=5, amplitude=1.5, base_cells=4):
    """Sum of `octaves` octaves of bilinear-smoothstep value noise on an
    (n+1)x(n+1) lattice over [0,1]^2."""
    rng = np.random.default_rng(seed)
    u = np.linspace(0.0, 1.0, n + 1)
    X, Y = np.meshgrid(u, u, indexing="xy")
    H = np.zeros_like(X)
    amp, cells, total = 1.0, base_cells, 0.0
    for _ in range(octaves):
        g = rng.random((cells + 2, cells + 2))
        fx, fy = X * cells, Y * cells
        ix, iy = np.minimum(fx.astype(np.int64), cells), np.minimum(fy.astype(np.int64), cells)
        tx, ty = fx - ix, fy - iy
        tx, ty = tx * tx * (3 - 2 * tx), ty * ty * (3 - 2 * ty)
        v = (g[iy, ix] * (1 - tx) + g[iy, ix + 1] * tx) * (1 - ty) + (
            g[iy + 1, ix] * (1 - tx) + g[iy + 1, ix + 1] * tx) * ty
        H += amp * (v - 0.5)
        total += amp
        amp *= 0.5
        cells *= 2
    return (H / total) * 2.0 * amplitude


def grid_mesh(n, extent=10.0, height=None):
    u = np.linspace(-extent, extent, n + 1, dtype=np.float64)
    X, Y = np.meshgrid(u, u, indexing="xy")
    Z = np.zeros_like(X) if height is None else height
    P = np.stack([X, Y, Z], axis=-1).reshape(-1, 3).astype(np.float32)
    i = np.arange(n, dtype=np.int64)
    I, J = np.meshgrid(i, i, indexing="xy")
    v00 = (J * (n + 1) + I).ravel()
    v10, v01, v11 = v00 + 1, v00 + (n + 1), v00 + (n + 2)
    tris = np.stack(
        [np.stack([v00, v10, v11], -1), np.stack([v00, v11, v01], -1)], axis=1).reshape(-1, 3)
    return P, tris.astype(np.int32)


def terrain(width=1920, height=1080, spp=256, n=708, max_bounce=0):
    """BASELINE config 2 - 708x708 grid (1 002 528 tris) displaced by 5 octaves of
    value noise; diffuse only; one sun; max_bounce 0 (traversal-bound)."""
    H = value_noise_height(n)
    P, tris = grid_mesh(n, 10.0, H)
    elev = np.deg2rad(35.0)
    dist = 24.0
    cam = look_at((0.0, -dist * np.cos(elev), dist * np.sin(elev)), (0.0, 0.0, 0.0))
    xml = "<cycles>\n"
    xml += _header(width, height, cam, 0.62, _integrator(max_bounce), nearclip=0.1, farclip=1000.0)
    xml += _background((0.3, 0.4, 0.6), 0.25)
    xml += _diffuse_shader("ground", (0.8, 0.8, 0.8))
    xml += _emission_shader("sun", (1.0, 0.95, 0.9), 1.0)
    sun_dir = np.array([0.35, 0.45, -0.82])
    sun_dir /= np.linalg.norm(sun_dir)
    xml += _state(
        "sun",
        '<light type="distant" dir="%s" angle="%s" strength="3 3 3" use_mis="true"/>\n'
        % (_f(sun_dir), _f(np.deg2rad(0.5))))
    xml += "</cycles>\n"
    return SceneDesc(
        "terrain_%dk" % (len(tris) // 1000), xml, width, height,
        meshes=[MeshDesc(P, tris, "ground")], objects=[(0, np.eye(4, dtype=np.float32)[:3])],
        spp=spp, notes="config 2")


# ---------------------------------------------------------------- config 3
def cornell(width=1920, height=1080, spp=512, max_bounce=8, distribution="GGX",
            materials="principled", light="area", panes=0, transparent_max=8, pattern="sobol",
            cam_type="perspective", cam_extra="", cam_pose=None, ao=None):
    """BASELINE config 3 - Cornell box, ceiling area light, one metallic and one
    glass Principled box (materials="diffuse" gives the all-diffuse variant).
    light="mesh" replaces the lamp by an emissive quad (a mesh light: its two triangles
    enter the light distribution), light="mesh_instanced" by two instances of one small
    emissive quad with different rotations and non-uniform scales (mesh lights whose
    object transform is not applied) - parity-test variants.  `panes` adds that many
    horizontal sheets of the "glass" material under the light (stacked transparent
    surfaces for the shadow rays), `transparent_max` is the transparent bounce limit."""
    xml = "<cycles>\n"
    cam = look_at((0.0, -3.6, 1.0), (0.0, 0.0, 1.0)) if cam_pose is None else cam_pose
    fov = 2.0 * np.arctan(0.5 * 36.0 / 50.0 / (width / height)) * 1.6
    xml += _header(width, height, cam, fov,
                   _integrator(max_bounce, clamp_indirect=10.0, transparent=transparent_max,
                               pattern=pattern, aa_samples=spp),
                   nearclip=0.01, farclip=100.0, cam_type=cam_type, cam_extra=cam_extra)
    # ao = (factor, distance): world ambient occlusion (kernel_path_ao)
    xml += _background((0, 0, 0), 0.0, "" if ao is None else
                       ' use_ao="true" ao_factor="%s" ao_distance="%s"' % (_f(ao[0]), _f(ao[1])))
    closure_variants = {"closures": 0, "closures2": 1, "transparent_opaque_shadow": 2,
                        "transparent": 3, "closures_multi": 4, "textured": 10, "textured2": 11,
                        "textured3": 12, "textured4": 13, "image": 20, "image2": 21}
    images, image_tiles = {}, []
    if materials in ("image", "image2"):
        shaders, images, image_tiles = _image_shaders(closure_variants[materials] - 20)
        xml += shaders
    elif materials in ("textured", "textured2", "textured3", "textured4"):
        xml += _textured_shaders(closure_variants[materials] - 10)
    elif materials in closure_variants:
        xml += _closure_shaders(closure_variants[materials])
    else:
        xml += _diffuse_shader("white", (0.73, 0.73, 0.73))
        xml += _diffuse_shader("red", (0.65, 0.05, 0.05))
        xml += _diffuse_shader("green", (0.12, 0.45, 0.15))
    if materials in closure_variants:
        pass
    elif materials == "procedural":
        xml += _procedural_shader("metal", 0)
    elif materials in ("principled", "metal"):
        xml += _principled_shader("metal", (0.9, 0.85, 0.7), 1.0, 0.2, 0.5, 0.0, 1.45, distribution)
    else:
        xml += _diffuse_shader("metal", (0.9, 0.85, 0.7))
    if materials in closure_variants:
        pass
    elif materials == "procedural":
        xml += _procedural_shader("glass", 1)
    elif materials in ("principled", "glass"):
        xml += _principled_shader("glass", (1, 1, 1), 0.0, 0.0, 0.5, 1.0, 1.45, distribution)
    else:
        xml += _diffuse_shader("glass", (0.6, 0.7, 0.9))
    if light == "area":
        xml += _emission_shader("lamp", (1.0, 0.9, 0.7), 1.0)
        xml += _state(
            "lamp",
            '<light type="area" co="0 0 1.98" dir="0 0 -1" axisu="1 0 0" axisv="0 1 0" '
            'sizeu="0.5" sizev="0.5" size="1" strength="12 12 12" use_mis="true"/>\n')
    else:
        xml += _emission_shader("lamp", (1.0, 0.9, 0.7), 18.0)
    xml += "</cycles>\n"

    def quad(a, b, c, d):
        return np.array([a, b, c, d], np.float32), np.array([[0, 1, 2], [0, 2, 3]], np.int32)

    meshes, objects = [], []

    def add(P, tris, shader, tfm=None):
        meshes.append(MeshDesc(np.asarray(P, np.float32), np.asarray(tris, np.int32), shader))
        objects.append((len(meshes) - 1, np.eye(4, dtype=np.float32)[:3] if tfm is None else tfm))

    add(*quad((-1, -1, 0), (1, -1, 0), (1, 1, 0), (-1, 1, 0)), "white")  # floor
    add(*quad((-1, -1, 2), (-1, 1, 2), (1, 1, 2), (1, -1, 2)), "white")  # ceiling
    add(*quad((-1, 1, 0), (1, 1, 0), (1, 1, 2), (-1, 1, 2)), "white")  # back
    add(*quad((-1, -1, 0), (-1, 1, 0), (-1, 1, 2), (-1, -1, 2)), "red")  # left
    add(*quad((1, -1, 0), (1, -1, 2), (1, 1, 2), (1, 1, 0)), "green")  # right

    def rot_z(deg, t):
        a = np.deg2rad(deg)
        m = np.eye(4, dtype=np.float32)[:3].copy()
        m[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]
        m[:, 3] = t
        return m

    # the boxes float 2 mm above the floor: coplanar overlapping faces would make the
    # closest hit an exact tie whose winner depends on the BVH traversal order
    Pb, Tb = box_mesh((-0.3, -0.3, 0.002), (0.3, 0.3, 1.2))
    add(Pb, Tb, "metal", rot_z(20.0, (-0.35, 0.3, 0.0)))
    Ps, Ts = box_mesh((-0.3, -0.3, 0.002), (0.3, 0.3, 0.6))
    add(Ps, Ts, "glass", rot_z(-18.0, (0.35, -0.3, 0.0)))
    for k in range(panes):
        z = 1.25 + 0.12 * k
        add(*quad((-0.6, -0.7, z), (0.1 + 0.1 * k, -0.7, z), (0.1 + 0.1 * k, 0.7, z),
                  (-0.6, 0.7, z)), "glass")
    if light == "mesh":
        add(*quad((-0.25, -0.25, 1.98), (-0.25, 0.25, 1.98), (0.25, 0.25, 1.98),
                  (0.25, -0.25, 1.98)), "lamp")
    elif light == "mesh_instanced":
        Pq, Tq = quad((-0.5, -0.5, 0), (-0.5, 0.5, 0), (0.5, 0.5, 0), (0.5, -0.5, 0))
        meshes.append(MeshDesc(Pq, Tq, "lamp"))
        mi = len(meshes) - 1
        for deg, sx, sy, t in ((25.0, 0.45, 0.25, (-0.4, 0.2, 1.97)),
                               (-40.0, 0.2, 0.5, (0.45, -0.1, 1.95))):
            m = rot_z(deg, t)
            m[:, 0] *= sx
            m[:, 1] *= sy
            objects.append((mi, m))
    elif light != "area":
        raise ValueError(light)
    multi = "_multiscatter" if (materials in ("principled", "metal", "glass") and
                                distribution != "GGX") else ""
    return SceneDesc("cornell_" + materials + multi + ("" if light == "area" else "_" + light),
                     xml, width, height, meshes=meshes, objects=objects, spp=spp, notes="config 3",
                     images=images, image_tiles=image_tiles)


# ---------------------------------------------------------------- config 4
def icosphere(subdiv):
    t = (1.0 + 5.0 ** 0.5) / 2.0
    V = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t),
         (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    F = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4),
         (11, 10, 2), (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8),
         (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    V = np.array(V, dtype=np.float64)
    V /= np.linalg.norm(V, axis=1, keepdims=True)
    F = np.array(F, dtype=np.int64)
    for _ in range(subdiv):
        e = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]], axis=0)
        es = np.sort(e, axis=1)
        uniq, inv = np.unique(es, axis=0, return_inverse=True)
        mid = V[uniq[:, 0]] + V[uniq[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        base = len(V)
        V = np.concatenate([V, mid], axis=0)
        nF = len(F)
        inv = inv.reshape(-1)
        m01, m12, m20 = base + inv[:nF], base + inv[nF:2 * nF], base + inv[2 * nF:]
        F = np.concatenate([
            np.stack([F[:, 0], m01, m20], 1), np.stack([F[:, 1], m12, m01], 1),
            np.stack([F[:, 2], m20, m12], 1), np.stack([m01, m12, m20], 1)], axis=0)
    return V, F


def geodesic_sphere(freq):
    """Icosahedron with every face cut into freq x freq triangles, pushed out onto the unit
    sphere: 20 * freq^2 triangles (freq 71 -> 100 820), vertices shared along the edges."""
    V0, F0 = icosphere(0)
    verts, index, faces = [], {}, []

    def vid(p):
        key = tuple(np.round(p, 9))
        if key not in index:
            index[key] = len(verts)
            verts.append(p)
        return index[key]

    n = int(freq)
    for a, b, c in F0:
        A, B, C = V0[a], V0[b], V0[c]
        grid = {}
        for i in range(n + 1):
            for j in range(n + 1 - i):
                p = (A * (n - i - j) + B * i + C * j) / n
                grid[(i, j)] = vid(p / np.linalg.norm(p))
        for i in range(n):
            for j in range(n - i):
                faces.append((grid[(i, j)], grid[(i + 1, j)], grid[(i, j + 1)]))
                if i + j < n - 1:
                    faces.append((grid[(i + 1, j)], grid[(i + 1, j + 1)], grid[(i, j + 1)]))
    return np.array(verts, dtype=np.float64), np.array(faces, dtype=np.int64)


def rock_mesh(subdiv=6, seed=42, amplitude=0.18, freq=0):
    """A sphere with lumpy radial noise: an icosphere (20*4^subdiv tris; subdiv 6 ->
    81 920) or, with freq > 0, a geodesic sphere of 20*freq^2 triangles."""
    V, F = geodesic_sphere(freq) if freq > 0 else icosphere(subdiv)
    rng = np.random.default_rng(seed)
    r = np.ones(len(V))
    for k in range(1, 5):
        for _ in range(6):
            d = rng.normal(size=3)
            d /= np.linalg.norm(d)
            ph = rng.uniform(0, 2 * np.pi)
            r += (amplitude / (k * 6)) * np.sin(k * 3.0 * (V @ d) + ph)
    return (V * r[:, None]).astype(np.float32), F.astype(np.int32)


def instanced(width=3840, height=2160, spp=256, grid=100, subdiv=6, max_bounce=2, seed=7,
              freq=None):
    """BASELINE config 4 - grid x grid instances of one lumpy sphere BLAS (two-level BVH,
    transform_applied=false) over a ground plane.  At the configuration's size (grid 100)
    the BLAS is a geodesic sphere of 20 * 71^2 = 100 820 triangles: 10 000 instances of
    a 100k-triangle mesh, ~1 G effective triangles; the small test variants keep the
    icosphere (20 * 4^subdiv)."""
    if freq is None:
        freq = 71 if grid >= 100 else 0
    P, T = rock_mesh(subdiv, freq=freq)
    rng = np.random.default_rng(seed)
    meshes = [MeshDesc(P, T, "rock")]
    objects = []
    spacing = 3.0
    half = 0.5 * spacing * (grid - 1)
    for j in range(grid):
        for i in range(grid):
            s = rng.uniform(0.6, 1.2)
            ax = rng.normal(size=3)
            ax /= np.linalg.norm(ax)
            ang = rng.uniform(0, 2 * np.pi)
            K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
            R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)
            m = np.zeros((3, 4), dtype=np.float32)
            m[:, :3] = (R * s).astype(np.float32)
            m[:, 3] = (i * spacing - half + rng.uniform(-0.6, 0.6),
                       j * spacing - half + rng.uniform(-0.6, 0.6), 1.25 * s)
            objects.append((0, m))
    g = half + 10.0
    Pg = np.array([(-g, -g, 0), (g, -g, 0), (g, g, 0), (-g, g, 0)], np.float32)
    meshes.append(MeshDesc(Pg, np.array([[0, 1, 2], [0, 2, 3]], np.int32), "ground"))
    objects.append((1, np.eye(4, dtype=np.float32)[:3]))

    cam = look_at((0.0, -half * 0.55, half * 0.22), (0.0, half * 0.1, 0.0))
    xml = "<cycles>\n"
    xml += _header(width, height, cam, 0.55, _integrator(max_bounce), nearclip=0.1, farclip=5000.0)
    xml += _background((0.35, 0.45, 0.65), 0.3)
    xml += _diffuse_shader("rock", (0.7, 0.6, 0.5))
    xml += _diffuse_shader("ground", (0.5, 0.5, 0.5))
    xml += _emission_shader("sun", (1.0, 0.95, 0.9), 1.0)
    sun_dir = np.array([0.4, 0.3, -0.85])
    sun_dir /= np.linalg.norm(sun_dir)
    xml += _state(
        "sun",
        '<light type="distant" dir="%s" angle="%s" strength="3 3 3" use_mis="true"/>\n'
        % (_f(sun_dir), _f(np.deg2rad(0.5))))
    xml += "</cycles>\n"
    return SceneDesc("instanced_%dx%d" % (grid, grid), xml, width, height, meshes=meshes,
                     objects=objects, spp=spp, notes="config 4")


CONFIGS = {
    "cube": default_cube,
    "terrain": terrain,
    "cornell": cornell,
    "instanced": instanced,
}


# ------------------------------------------------- value-node chart (tests)
MATH_OPS = ["add", "subtract", "multiply", "divide", "multiply_add", "sine", "cosine", "tangent",
            "sinh", "cosh", "tanh", "arcsine", "arccosine", "arctangent", "power", "logarithm",
            "minimum", "maximum", "round", "less_than", "greater_than", "modulo", "absolute",
            "arctan2", "floor", "ceil", "fraction", "trunc", "snap", "wrap", "pingpong", "sqrt",
            "inversesqrt", "sign", "exponent", "radians", "degrees", "smoothmin", "smoothmax",
            "compare"]
VECTOR_OPS = ["add", "subtract", "multiply", "divide", "cross_product", "project", "reflect",
              "dot_product", "distance", "length", "scale", "normalize", "snap", "floor", "ceil",
              "modulo", "wrap", "fraction", "absolute", "minimum", "maximum", "sine", "cosine",
              "tangent"]
MIX_OPS = ["mix", "add", "multiply", "screen", "overlay", "subtract", "divide", "difference",
           "darken", "lighten", "soft_light", "linear_light"]


def node_chart(width=256, height=144, spp=1):
    """One emissive quad per group of value nodes; the emitted colour IS the node output,
    evaluated at the hit position, so one frame compares every operator of svm_nodes.cuh
    with the reference (camera rays only: the inputs are bit-identical on both sides)."""
    head = ('  <geometry name="g"/>\n  <separate_xyz name="sep"/>\n'
            '  <connect from="g position" to="sep vector"/>\n'
            '  <math name="a" type="multiply_add" value2="1.7" value3="0.13"/>\n'
            '  <connect from="sep x" to="a value1"/>\n'
            '  <math name="b" type="multiply_add" value2="-0.9" value3="0.41"/>\n'
            '  <connect from="sep y" to="b value1"/>\n')
    tail = ('  <emission name="e" strength="1"/>\n  <connect from="out vector" to="e color"/>\n'
            '  <connect from="e emission" to="output surface"/>\n')
    shaders = []

    def scalar_group(kind_ops):
        body = head + '  <combine_xyz name="out"/>\n'
        for ch, op in zip("xyz", kind_ops):
            body += ('  <math name="op_%s" type="%s" value3="0.35"/>\n'
                     '  <connect from="a value" to="op_%s value1"/>\n'
                     '  <connect from="b value" to="op_%s value2"/>\n'
                     '  <connect from="op_%s value" to="out %s"/>\n' % (ch, op, ch, ch, ch, ch))
        return body + tail

    for i in range(0, len(MATH_OPS), 3):
        shaders.append(scalar_group(MATH_OPS[i:i + 3]))

    for op in VECTOR_OPS:
        body = head + ('  <combine_xyz name="va" z="0.6"/>\n'
                       '  <connect from="a value" to="va x"/>\n'
                       '  <connect from="b value" to="va y"/>\n'
                       '  <vector_math name="vop" type="%s" vector2="0.45 -0.8 0.3" '
                       'vector3="-0.2 0.1 0.05" scale="1.5"/>\n'
                       '  <connect from="va vector" to="vop vector1"/>\n' % op)
        if op in ("dot_product", "distance", "length"):
            body += ('  <combine_xyz name="out" y="0.2" z="0.1"/>\n'
                     '  <connect from="vop value" to="out x"/>\n')
        else:
            body += ('  <vector_math name="out" type="add" vector2="0 0 0"/>\n'
                     '  <connect from="vop vector" to="out vector1"/>\n')
        shaders.append(body + tail)

    for op in MIX_OPS:
        body = head + ('  <combine_xyz name="c1" z="0.3"/>\n'
                       '  <connect from="a value" to="c1 x"/>\n'
                       '  <connect from="b value" to="c1 y"/>\n'
                       '  <math name="f" type="fraction"/>\n'
                       '  <connect from="a value" to="f value1"/>\n'
                       '  <mix name="mop" type="%s" color2="0.25 0.6 0.9"/>\n'
                       '  <connect from="f value" to="mop fac"/>\n'
                       '  <connect from="c1 vector" to="mop color1"/>\n'
                       '  <vector_math name="out" type="add" vector2="0 0 0"/>\n'
                       '  <connect from="mop color" to="out vector1"/>\n' % op)
        shaders.append(body + tail)

    # a few single-purpose ones: clamp (both types), gamma, brightness, invert, fresnel,
    # layer weight (both outputs), float -> colour and colour -> float conversion
    extra = head + (
        '  <clamp name="cl1" type="minmax" min="0.2" max="0.7"/>\n'
        '  <connect from="a value" to="cl1 value"/>\n'
        '  <clamp name="cl2" type="range" min="0.8" max="0.1"/>\n'
        '  <connect from="b value" to="cl2 value"/>\n'
        '  <fresnel name="fr" IOR="1.3"/>\n'
        '  <combine_xyz name="out"/>\n'
        '  <connect from="cl1 result" to="out x"/>\n'
        '  <connect from="cl2 result" to="out y"/>\n'
        '  <connect from="fr fac" to="out z"/>\n')
    shaders.append(extra + tail)
    extra2 = head + (
        '  <combine_xyz name="c1" z="0.3"/>\n'
        '  <connect from="a value" to="c1 x"/>\n'
        '  <connect from="b value" to="c1 y"/>\n'
        '  <gamma name="gm" gamma="2.2"/>\n'
        '  <connect from="c1 vector" to="gm color"/>\n'
        '  <brightness_contrast name="bc" bright="0.1" contrast="0.3"/>\n'
        '  <connect from="gm color" to="bc color"/>\n'
        '  <invert name="out2" fac="0.7"/>\n'
        '  <connect from="bc color" to="out2 color"/>\n'
        '  <vector_math name="out" type="add" vector2="0 0 0"/>\n'
        '  <connect from="out2 color" to="out vector1"/>\n')
    shaders.append(extra2 + tail)
    extra3 = head + (
        '  <layer_weight name="lw" blend="0.35"/>\n'
        '  <combine_xyz name="c1" z="0.3"/>\n'
        '  <connect from="a value" to="c1 x"/>\n'
        '  <connect from="b value" to="c1 y"/>\n'
        '  <math name="gray" type="multiply" value2="1.0"/>\n'
        '  <connect from="c1 vector" to="gray value1"/>\n'     # colour -> float
        '  <combine_xyz name="out"/>\n'
        '  <connect from="lw fresnel" to="out x"/>\n'
        '  <connect from="lw facing" to="out y"/>\n'
        '  <connect from="gray value" to="out z"/>\n')
    shaders.append(extra3 + tail)
    extra4 = head + (
        '  <emission name="e" strength="1"/>\n'
        '  <connect from="a value" to="e color"/>\n'           # float -> colour
        '  <connect from="e emission" to="output surface"/>\n')
    shaders.append(extra4)

    # light path (camera-ray flags, ray length, depth), ColorRamp, RGB / vector curves
    rng = np.random.default_rng(5)
    ramp = np.sort(rng.random((8, 3)), axis=0)
    ramp_attr = " ".join("%.6g" % v for v in ramp.ravel())
    alpha_attr = " ".join("%.6g" % v for v in np.linspace(0.2, 1.0, 8))
    curve = np.clip(np.linspace(0, 1, 16)[:, None] ** np.array([0.5, 1.0, 2.0]), 0, 1)
    curve_attr = " ".join("%.6g" % v for v in curve.ravel())
    extra5 = head + (
        '  <light_path name="lp"/>\n'
        '  <math name="rl" type="multiply" value2="0.08"/>\n'
        '  <connect from="lp ray_length" to="rl value1"/>\n'
        '  <math name="dp" type="add" value2="0.25"/>\n'
        '  <connect from="lp ray_depth" to="dp value1"/>\n'
        '  <combine_xyz name="out"/>\n'
        '  <connect from="lp is_camera_ray" to="out x"/>\n'
        '  <connect from="rl value" to="out y"/>\n'
        '  <connect from="dp value" to="out z"/>\n')
    shaders.append(extra5 + tail)
    for interp in ("true", "false"):
        extra6 = head + (
            '  <rgb_ramp name="rr" ramp="%s" ramp_alpha="%s" interpolate="%s"/>\n'
            '  <connect from="a value" to="rr fac"/>\n'
            '  <mix name="mo" type="multiply" fac="1"/>\n'
            '  <connect from="rr color" to="mo color1"/>\n'
            '  <connect from="rr alpha" to="mo color2"/>\n'
            '  <vector_math name="out" type="add" vector2="0 0 0"/>\n'
            '  <connect from="mo color" to="out vector1"/>\n' % (ramp_attr, alpha_attr, interp))
        shaders.append(extra6 + tail)
    for kind in ("rgb_curves", "vector_curves"):
        sock = "value"  # cycles_xml matches the socket's internal name
        extra7 = head + (
            '  <combine_xyz name="c1" z="0.3"/>\n'
            '  <connect from="a value" to="c1 x"/>\n'
            '  <connect from="b value" to="c1 y"/>\n'
            '  <%s name="cv" curves="%s" min_x="-0.25" max_x="1.25" fac="0.8"/>\n'
            '  <connect from="c1 vector" to="cv %s"/>\n'
            '  <vector_math name="out" type="add" vector2="0 0 0"/>\n'
            '  <connect from="cv %s" to="out vector1"/>\n' % (kind, curve_attr, sock, sock))
        shaders.append(extra7 + tail)

    n = len(shaders)
    cols = 12
    rows = (n + cols - 1) // cols
    cam = look_at((0.0, 0.0, 10.0), (0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0))
    fov = 2.0 * np.arctan((rows * 0.5 + 0.2) / 10.0)
    xml = "<cycles>\n"
    xml += _header(width, height, cam, fov, _integrator(0), nearclip=0.1, farclip=100.0)
    xml += _background((0, 0, 0), 0.0)
    for i, body in enumerate(shaders):
        xml += '<shader name="s%d">\n%s</shader>\n' % (i, body)
    xml += "</cycles>\n"
    meshes, objects = [], []
    for i in range(n):
        cx = (i % cols) - (cols - 1) / 2.0
        cy = (rows - 1) / 2.0 - (i // cols)
        P = np.array([(cx - 0.46, cy - 0.46, 0), (cx + 0.46, cy - 0.46, 0),
                      (cx + 0.46, cy + 0.46, 0), (cx - 0.46, cy + 0.46, 0)], np.float32)
        meshes.append(MeshDesc(P, np.array([[0, 1, 2], [0, 2, 3]], np.int32), "s%d" % i))
        objects.append((i, np.eye(4, dtype=np.float32)[:3]))
    return SceneDesc("node_chart", xml, width, height, meshes=meshes, objects=objects, spp=spp,
                     notes="svm value nodes")

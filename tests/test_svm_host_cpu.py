"""Node-by-node parity of the device-side SVM texture / attribute / mapping nodes
(csrc/svm_tex.cuh) with the reference's svm_node_* functions, on the CPU.

tests/host_check/svm_tex_host.cpp compiles THE SAME svm_tex.cuh the CUDA kernels include
for the host (g++, CUDA built-ins mapped to their host meaning); oracle/ref_probe.cpp
runs the reference's own function for the same opcode.  Both get identical node words,
stacks and shading points; the stacks they leave must agree.  Two sources of programs:

 - every texture-family node of the real compiled programs of the "textured" Cornell
   scenes (real attribute maps, objects, camera);
 - random encodings of each opcode (random stack slots / defaults / enum values).

This checks decoding and arithmetic of the nodes before any GPU time is spent; the GPU
render-parity tests (test_render_gpu.py, texture_cases) check them in the kernels."""
import ctypes as C
import os
import re
import subprocess
import zlib

import numpy as np
import pytest

from raytracingproject_b200 import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"


def abi():
    text = open(os.path.join(ROOT, "include", "cycles_abi.h")).read()
    return {k: int(v) for k, v in re.findall(r"#define CY_(NODE_\w+)[ \t]+(\d+)", text)}


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    out = tmp_path_factory.mktemp("host_check") / "libsvm_tex_host.so"
    src = os.path.join(ROOT, "tests", "host_check", "svm_tex_host.cpp")
    subprocess.run(["g++", "-std=c++14", "-O2", "-fPIC", "-shared", "-w", "-ffp-contract=off",
                    "-fvisibility=hidden", "-Wl,-Bsymbolic",  # g_scene must not interpose
                    "-I" + CUDA_INC, src, "-o", str(out)], check=True)
    L = C.CDLL(str(out))
    L.host_svm_node.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    L.host_svm_node.restype = C.c_int
    L.host_svm_bind.argtypes = [C.c_void_p]
    return L


class Bound:
    """The scene arrays svm_tex.cuh reads, bound to the host build; keeps them alive."""
    FIELDS = ["svm_nodes", "objects", "tri_vindex", "lights", "shaders", "attributes_map",
              "attributes_float", "attributes_float2", "attributes_float3",
              "attributes_uchar4", "kernel_data"]

    def __init__(self, L, arrays, nodes, textures=()):
        self.keep = {"svm_nodes": nodes}
        for f in self.FIELDS[1:]:
            name = "__data" if f == "kernel_data" else "__" + f
            if name in arrays:
                self.keep[f] = np.ascontiguousarray(arrays[name][0])
        # image slots: the reference's TextureInfo records, `data` = where the host copy of
        # the pixels lives (what B200Device.tex_alloc does with a device address)
        from raytracingproject_b200.device import SIZEOF_TEXTURE_INFO
        n_tex = 1 + max([slot for slot, _, _ in textures], default=-1)
        table = np.zeros((n_tex, SIZEOF_TEXTURE_INFO), np.uint8)
        self.pixels = []
        for slot, info, pix in textures:
            pix = np.ascontiguousarray(pix)
            self.pixels.append(pix)
            table[slot] = info
            table[slot, :8] = np.array([pix.ctypes.data], np.uint64).view(np.uint8)
        self.keep["texture_info"] = table
        ptrs = (C.c_void_p * (len(self.FIELDS) + 2))(
            *([self.keep[f].ctypes.data if f in self.keep else None for f in self.FIELDS] +
              [table.ctypes.data if n_tex else None, n_tex]))
        L.host_svm_bind(ptrs)


def shading_points(arrays, n, rng):
    from oracle.cycles_ref import SHADING_POINT_DTYPE
    prim_object = arrays["__prim_object"][0].view(np.int32)
    prim_index = arrays["__prim_index"][0].view(np.int32)
    ok = np.nonzero((prim_object >= 0) & (prim_index >= 0))[0]
    pts = np.zeros(n, SHADING_POINT_DTYPE)
    pick = rng.choice(ok, n)
    pts["object"], pts["prim"], pts["lamp"] = prim_object[pick], prim_index[pick], -1
    for k in ("P", "dPdu"):
        pts[k] = rng.uniform(-1.5, 1.5, (n, 3))
    for k in ("N", "I"):
        v = rng.normal(size=(n, 3))
        pts[k] = v / np.linalg.norm(v, axis=1, keepdims=True)
    u = rng.uniform(0, 1, n)
    pts["u"], pts["v"] = u, rng.uniform(0, 1, n) * (1 - u)
    # a few points that are not on a surface: background, and a lamp's emission shader
    pts["object"][: n // 16] = -1
    pts["prim"][: n // 16] = -1
    if "__lights" in arrays:
        pts["lamp"][: n // 32] = 0
    pts["shader"] = rng.integers(0, arrays["__shaders"][0].size // 32, n)  # SIZEOF_KERNEL_SHADER
    pts["backfacing"] = rng.integers(0, 2, n)
    return pts


def run_both(L, rs, nodes, offset, stack0, pt):
    s_ref, s_dev = stack0.copy(), stack0.copy()
    n_ref = rs.svm_node(nodes, offset, s_ref, pt)
    n_dev = L.host_svm_node(offset, s_dev.ctypes.data, pt.ctypes.data)
    return n_ref, n_dev, s_ref, s_dev


PRINCIPLED_ID = int(re.search(r"#define CY_CLOSURE_BSDF_PRINCIPLED_ID[ \t]+(\d+)",
                              open(os.path.join(ROOT, "include", "cycles_abi.h")).read()).group(1))


def texture_ops():
    a = abi()
    return {a[k]: k for k in ("NODE_ATTR", "NODE_TEX_COORD", "NODE_MAPPING",
                              "NODE_TEXTURE_MAPPING", "NODE_MIN_MAX", "NODE_TEX_NOISE",
                              "NODE_TEX_CHECKER", "NODE_TEX_GRADIENT", "NODE_TEX_WAVE",
                              "NODE_TEX_MAGIC", "NODE_TEX_BRICK", "NODE_GEOMETRY",
                              "NODE_HSV", "NODE_SEPARATE_HSV", "NODE_COMBINE_HSV",
                              "NODE_MAP_RANGE", "NODE_NORMAL", "NODE_VECTOR_ROTATE",
                              "NODE_VECTOR_TRANSFORM", "NODE_OBJECT_INFO", "NODE_CAMERA",
                              "NODE_TEX_WHITE_NOISE", "NODE_MIX", "NODE_MATH",
                              "NODE_TEX_VORONOI", "NODE_TEX_MUSGRAVE", "NODE_BLACKBODY",
                              "NODE_WAVELENGTH", "NODE_TANGENT", "NODE_NORMAL_MAP",
                              "NODE_VECTOR_MATH", "NODE_CONVERT", "NODE_INVERT", "NODE_GAMMA",
                              "NODE_BRIGHTCONTRAST", "NODE_CLAMP", "NODE_FRESNEL",
                              "NODE_LAYER_WEIGHT", "NODE_RGB_RAMP", "NODE_RGB_CURVES",
                              "NODE_VECTOR_CURVES", "NODE_TEX_IMAGE", "NODE_TEX_IMAGE_BOX",
                              "NODE_TEX_ENVIRONMENT")}


def compare(name, s_ref, s_dev, tol):
    bad = ~np.isclose(s_ref, s_dev, rtol=tol, atol=tol, equal_nan=True)
    assert not bad.any(), (name, np.nonzero(bad)[0][:8], s_ref[bad][:8], s_dev[bad][:8])


@pytest.mark.parametrize("materials", ["textured", "textured2", "textured3", "textured4",
                                       "procedural",
                                       "node_chart", "image", "image2", "env_equirect",
                                       "env_mirrorball"])
def test_compiled_texture_nodes_match_reference(ref, host_lib, materials):
    if materials == "node_chart":
        desc = scenes.node_chart()
    elif materials.startswith("env_"):
        desc = scenes.default_cube(64, 48, spp=1, world=materials)
    else:
        desc = scenes.cornell(64, 48, spp=1, materials=materials)
    rs = ref.build_scene(desc)
    try:
        arrays = rs.device_arrays()
        nodes = np.zeros((arrays["__svm_nodes"][0].size // 16 + 8, 4), np.uint32)
        real = arrays["__svm_nodes"][0].view(np.uint32).reshape(-1, 4)
        nodes[: len(real)] = real
        bound = Bound(host_lib, arrays, nodes, rs.textures())
        ops = texture_ops()
        a = abi()
        rng = np.random.default_rng(7)
        pts = shading_points(arrays, 64, rng)
        seen = {}
        # walk the program instruction by instruction: nodes the probes dispatch report
        # their own length, the rest (closures, jumps, constants) have a fixed one
        one = {a[k] for k in ("NODE_END", "NODE_SHADER_JUMP", "NODE_CLOSURE_EMISSION",
                              "NODE_CLOSURE_BACKGROUND", "NODE_CLOSURE_SET_WEIGHT",
                              "NODE_CLOSURE_WEIGHT", "NODE_EMISSION_WEIGHT", "NODE_MIX_CLOSURE",
                              "NODE_JUMP_IF_ZERO", "NODE_JUMP_IF_ONE", "NODE_GEOMETRY",
                              "NODE_VALUE_F", "NODE_LIGHT_PATH", "NODE_LIGHT_FALLOFF",
                              "NODE_SEPARATE_VECTOR", "NODE_COMBINE_VECTOR")}
        off = 0
        while off < len(real):
            op = int(nodes[off, 0])
            probed = op in ops and not (op == a["NODE_GEOMETRY"] and nodes[off, 1] != 2)
            if not probed:
                if op in one:
                    off += 1
                elif op == a["NODE_VALUE_V"]:
                    off += 2
                elif op == a["NODE_CLOSURE_BSDF"]:
                    off += 6 if (nodes[off, 1] & 0xff) == PRINCIPLED_ID else 2
                else:
                    raise AssertionError("walk: opcode %d at %d" % (op, off))
                continue
            nxt = None
            if os.environ.get("SVM_HOST_TRACE"):
                print("node", ops[op], off, nodes[off].tolist(), flush=True)
            surface_only = ops[op] in ("NODE_TANGENT", "NODE_NORMAL_MAP")
            for i in range(len(pts)):
                if surface_only and pts["object"][i] < 0:
                    continue  # off a surface the reference transforms by an unset matrix
                stack0 = rng.uniform(-2.0, 2.0, 264).astype(np.float32)
                if i % 4 == 0 and ops[op] in ("NODE_TEX_IMAGE", "NODE_TEX_IMAGE_BOX"):
                    # lookups near and across the image borders and the texel centres
                    stack0 = np.round(stack0 * 8.0) / 16.0 + np.float32(rng.choice(
                        [0.0, 1e-7, -1e-7, 0.5 / 16, 1.0]))
                n_ref, n_dev, s_ref, s_dev = run_both(host_lib, rs, nodes, off, stack0,
                                                      pts[i:i + 1])
                assert n_ref == n_dev and n_ref > off, (ops[op], off, n_ref, n_dev)
                compare((ops[op], off, i), s_ref, s_dev, 3e-5)
                if not np.array_equal(s_ref, stack0):
                    seen[ops[op]] = seen.get(ops[op], 0) + 1
                nxt = n_ref
            off = nxt
        # every node family the scenes were written to contain did run and wrote output
        common = {"NODE_ATTR", "NODE_TEX_NOISE", "NODE_TEX_CHECKER", "NODE_TEX_WAVE",
                  "NODE_TEX_MAGIC", "NODE_MAPPING", "NODE_TEX_GRADIENT", "NODE_TEXTURE_MAPPING"}
        want = {"textured": common | {"NODE_TEX_BRICK", "NODE_GEOMETRY"},
                "textured2": common | {"NODE_TEX_COORD", "NODE_MIN_MAX"},
                "textured3": {"NODE_HSV", "NODE_SEPARATE_HSV", "NODE_COMBINE_HSV",
                              "NODE_MAP_RANGE", "NODE_NORMAL", "NODE_VECTOR_ROTATE",
                              "NODE_VECTOR_TRANSFORM", "NODE_OBJECT_INFO", "NODE_CAMERA",
                              "NODE_TEX_WHITE_NOISE", "NODE_MIX", "NODE_MATH",
                              "NODE_VECTOR_MATH"},
                "textured4": {"NODE_TEX_VORONOI", "NODE_TEX_MUSGRAVE", "NODE_TEX_COORD"},
                "procedural": {"NODE_MATH", "NODE_VECTOR_MATH", "NODE_MIX", "NODE_CLAMP",
                               "NODE_GAMMA", "NODE_INVERT", "NODE_BRIGHTCONTRAST",
                               "NODE_FRESNEL", "NODE_LAYER_WEIGHT"},
                "node_chart": {"NODE_MATH", "NODE_VECTOR_MATH", "NODE_MIX"},
                "image": {"NODE_TEX_IMAGE", "NODE_TEX_IMAGE_BOX", "NODE_TEX_COORD"},
                "image2": {"NODE_TEX_IMAGE", "NODE_TEX_IMAGE_BOX"},
                "env_equirect": {"NODE_TEX_ENVIRONMENT"},
                "env_mirrorball": {"NODE_TEX_ENVIRONMENT"}}[materials]
        assert want <= set(seen), sorted(want - set(seen))
        del bound
    finally:
        rs.close()


O = "off"
# Encodings of the value nodes (svm/svm_*.h): what the y, z, w words of the instruction
# and of its data nodes hold - a stack offset, four packed offsets, a float, an enum range
# (int), or a tuple of packed bytes.
VALUE_NODE_SPECS = {
    "NODE_MATH": {"yzw": (41, "packed", O)},
    "NODE_VECTOR_MATH": {"yzw": (25, "packed", "packed"), "extra": [("off", O, O, O)]},
    "NODE_MIX": {"yzw": (O, O, O), "extra": [(O, 19, O, O)]},
    "NODE_CONVERT": {"yzw": (12, O, O)},
    "NODE_INVERT": {"yzw": (O, O, O)},
    "NODE_GAMMA": {"yzw": (O, O, O)},
    "NODE_BRIGHTCONTRAST": {"yzw": (O, O, "packed")},
    "NODE_CLAMP": {"yzw": (O, (O, O, 2), O), "extra": [("float", "float", "float", "float")]},
    "NODE_FRESNEL": {"yzw": (O, "float", "packed")},
    "NODE_LAYER_WEIGHT": {"yzw": (O, "float", (2, O, O))},
    "NODE_HSV": {"yzw": ("packed", "packed", O)},
    "NODE_SEPARATE_HSV": {"yzw": (O, O, O), "extra": [(O, O, O, O)]},
    "NODE_COMBINE_HSV": {"yzw": (O, O, O), "extra": [(O, O, O, O)]},
    "NODE_MAP_RANGE": {"yzw": (O, "packed", (4, O, O)),
                       "extra": [("float",) * 4, ("float",) * 4]},
    "NODE_NORMAL": {"yzw": (O, O, O), "extra": [("float",) * 4]},
    "NODE_VECTOR_ROTATE": {"yzw": ((5, O, O, 2), "packed", O)},
    "NODE_VECTOR_TRANSFORM": {"yzw": ((3, 3, 3), "packed", O)},
    "NODE_OBJECT_INFO": {"yzw": (5, O, O)},
    "NODE_CAMERA": {"yzw": (O, O, O)},
    "NODE_TEX_WHITE_NOISE": {"yzw": ("dims", "packed", "packed")},
    "NODE_TANGENT": {"yzw": ((O, 2, 3), "attr", "attr")},
    "NODE_NORMAL_MAP": {"yzw": ((O, O, O, 5), "attr", "attr")},
    "NODE_BLACKBODY": {"yzw": (O, O, O)},
    "NODE_WAVELENGTH": {"yzw": (O, O, O)},
    "NODE_TEX_MUSGRAVE": {"yzw": ((5, "dims", O, O), "packed", "packed"),
                          "extra": [("float",) * 4, ("float",) * 4]},
    "NODE_TEX_VORONOI": {"yzw": ("dims", 5, 4),
                         "extra": [("packed", "packed", "packed", "float"), ("float",) * 4]},
}


def random_program(op_name, rng, a):
    """One random encoding of `op_name` followed by its data words."""
    text = open(os.path.join(ROOT, "include", "cycles_abi.h")).read()
    a_std = {k: int(v) for k, v in re.findall(r"#define CY_ATTR_STD_(\w+)[ \t]+(\d+)", text)}
    slot = lambda: int(rng.choice([255, int(rng.integers(0, 60)) * 4]))  # invalid -> default
    used = lambda: int(rng.integers(0, 60)) * 4
    pack = lambda *b: int(sum(int(x) << (8 * i) for i, x in enumerate(b)))
    fbits = lambda lo, hi: int(np.float32(rng.uniform(lo, hi)).view(np.uint32))
    prog = np.zeros((8, 4), np.uint32)
    prog[0, 0] = a[op_name]
    if op_name == "NODE_MAPPING":
        prog[0, 1:] = [rng.integers(0, 4), pack(used(), used(), used(), used()), used()]
    elif op_name == "NODE_TEXTURE_MAPPING":
        prog[0, 1:3] = [used(), used()]
        prog[1:4] = rng.uniform(-2, 2, (3, 4)).astype(np.float32).view(np.uint32)
    elif op_name == "NODE_MIN_MAX":
        prog[0, 1:3] = [used(), used()]
        prog[1] = rng.uniform(-2, 0, 4).astype(np.float32).view(np.uint32)
        prog[2] = rng.uniform(0, 2, 4).astype(np.float32).view(np.uint32)
    elif op_name == "NODE_TEX_NOISE":
        prog[0, 1:] = [rng.integers(1, 5), pack(used(), slot(), slot(), slot()),
                       pack(slot(), slot(), slot(), slot())]
        prog[1] = [fbits(-2, 2), fbits(0.5, 4), fbits(0, 5), fbits(0, 1)]
        prog[2, 0] = fbits(0, 1.5) if rng.random() < 0.7 else 0
    elif op_name == "NODE_TEX_CHECKER":
        prog[0, 1:] = [pack(used(), used(), used(), slot()), pack(slot(), slot()), fbits(0.5, 6)]
    elif op_name == "NODE_TEX_GRADIENT":
        prog[0, 1] = pack(rng.integers(0, 7), used(), slot(), slot())
    elif op_name == "NODE_TEX_WAVE":
        prog[0, 1:] = [pack(rng.integers(0, 2), rng.integers(0, 4), rng.integers(0, 4),
                            rng.integers(0, 3)),
                       pack(used(), slot(), slot()), pack(slot(), slot(), slot(), slot())]
        prog[1] = [pack(slot(), slot()), fbits(0.2, 3), fbits(0, 2) if rng.random() < 0.7 else 0,
                   fbits(0, 4)]
        prog[2, :3] = [fbits(0.5, 2), fbits(0, 1), fbits(-3, 3)]
    elif op_name == "NODE_TEX_MAGIC":
        prog[0, 1:3] = [pack(rng.integers(0, 11), slot(), slot()), pack(used(), slot(), slot())]
        prog[1, :2] = [fbits(0.5, 5), fbits(0, 2) if rng.random() < 0.8 else 0]
    elif op_name == "NODE_TEX_BRICK":
        prog[0, 1:] = [pack(used(), used(), used(), used()), pack(slot(), slot(), slot(), slot()),
                       pack(slot(), slot(), slot(), slot())]
        prog[1] = [pack(rng.integers(0, 4), rng.integers(0, 4)), fbits(1, 6), fbits(0, 0.1),
                   fbits(-0.5, 0.5)]
        prog[2] = [fbits(0.2, 1), fbits(0.1, 0.5), fbits(0, 1), fbits(0.5, 1.5)]
        prog[3, 0] = fbits(0, 1) if rng.random() < 0.6 else 0
    elif op_name in VALUE_NODE_SPECS:
        spec = VALUE_NODE_SPECS[op_name]
        word = {"off": used, "packed": lambda: pack(used(), used(), used(), used()),
                "float": lambda: fbits(0.2, 2.5), "dims": lambda: int(rng.integers(1, 5)),
                # generated, UV, vertex normal, or an id no mesh has
                "attr": lambda: int(rng.choice([a_std["GENERATED"], a_std["UV"],
                                                a_std["VERTEX_NORMAL"], 4000]))}
        for col, kind in zip((1, 2, 3), spec["yzw"]):
            if isinstance(kind, int):
                prog[0, col] = rng.integers(0, kind)
            elif isinstance(kind, tuple):  # packed bytes: ints are enum ranges
                prog[0, col] = pack(*[rng.integers(0, k) if isinstance(k, int) else
                                      (int(rng.integers(1, 5)) if k == "dims" else used())
                                      for k in kind])
            else:
                prog[0, col] = word[kind]()
        if op_name == "NODE_VECTOR_MATH":
            # only the output the operator defines is given a slot: the reference leaves
            # the other one uninitialised (svm_math_util.h svm_vector_math)
            scalar = int(prog[0, 1]) in (7, 8, 9)  # dot product, distance, length
            prog[0, 3] = pack(used(), 255) if scalar else pack(255, used())
        if op_name == "NODE_TEX_MUSGRAVE":
            # parameters from the node's defaults (sane ranges): a negative lacunarity or
            # dimension from a random stack slot turns the fractal sums into an amplifier
            # of the 1e-7 differences between the SSE and the scalar Perlin noise
            prog[0, 2] = pack(255, 255, 255, 255)
            prog[0, 3] = pack(255, 255, used())
        for row, kinds in enumerate(spec.get("extra", []), start=1):
            for col, kind in enumerate(kinds):
                prog[row, col] = rng.integers(0, kind) if isinstance(kind, int) else word[kind]()
    if op_name == "NODE_TEX_VORONOI":
        # outputs a feature does not define stay unassigned (the reference leaves its
        # position locals uninitialised for distance-to-edge / n-sphere radius)
        feature, dims = int(prog[0, 2]), int(prog[0, 1])
        cells = feature in (0, 1, 2)
        prog[1, 0] = pack(used(), slot(), slot(), slot())
        prog[1, 1] = pack(slot(), slot(), used() if feature != 4 else 255,
                          used() if cells else 255)
        prog[1, 2] = pack(used() if cells and dims != 1 else 255,
                          used() if cells and dims in (1, 4) else 255,
                          used() if feature == 4 else 255)
        prog[1, 3] = fbits(-2, 2)
        prog[2] = [fbits(0.5, 4), fbits(0.1, 1.5), fbits(0.5, 3), fbits(0, 1.2)]
    elif op_name == "NODE_TEX_COORD":
        kind = int(rng.choice([0, 1, 1, 2, 3, 4]))
        prog[0, 1:] = [kind, used(), int(kind == 1 and rng.random() < 0.5)]
        prog[1:4] = rng.uniform(-2, 2, (3, 4)).astype(np.float32).view(np.uint32)
    return prog


@pytest.mark.parametrize("op_name", ["NODE_MAPPING", "NODE_TEXTURE_MAPPING", "NODE_MIN_MAX",
                                     "NODE_TEX_NOISE", "NODE_TEX_CHECKER", "NODE_TEX_GRADIENT",
                                     "NODE_TEX_WAVE", "NODE_TEX_MAGIC", "NODE_TEX_BRICK",
                                     "NODE_TEX_COORD"] + sorted(VALUE_NODE_SPECS))
def test_random_node_encodings_match_reference(ref, host_lib, op_name):
    desc = scenes.cornell(64, 48, spp=1, materials="textured2")
    rs = ref.build_scene(desc)
    try:
        arrays = rs.device_arrays()
        a = abi()
        rng = np.random.default_rng(zlib.crc32(op_name.encode()))
        pts = shading_points(arrays, 64, rng)
        if op_name in ("NODE_TANGENT", "NODE_NORMAL_MAP"):
            # surface points only: off a surface the reference transforms by an unset matrix
            pts["object"], pts["prim"] = np.maximum(pts["object"], 0), np.maximum(pts["prim"], 0)
            pts["lamp"] = -1
        wrote = 0
        for trial in range(300):
            prog = random_program(op_name, rng, a)
            bound = Bound(host_lib, arrays, prog)
            stack0 = rng.uniform(-2.0, 2.0, 264).astype(np.float32)
            if op_name == "NODE_BLACKBODY":  # temperatures across all six bands
                stack0 = rng.uniform(500.0, 14000.0, 264).astype(np.float32)
            if op_name == "NODE_WAVELENGTH":  # nanometres, a little beyond the table
                stack0 = rng.uniform(350.0, 810.0, 264).astype(np.float32)
            pt = pts[trial % len(pts):trial % len(pts) + 1]
            n_ref, n_dev, s_ref, s_dev = run_both(host_lib, rs, prog, 0, stack0, pt)
            assert n_ref == n_dev and n_ref > 0, (op_name, trial, n_ref, n_dev)
            compare((op_name, trial, prog[:4].tolist()), s_ref, s_dev, 3e-5)
            wrote += int(not np.array_equal(s_ref, stack0))
            del bound
        assert wrote > 150
    finally:
        rs.close()


def test_scope_check_without_a_device(ref):
    """b200_validate_svm: programs of the supported scenes pass, a Bump node or a
    truncated program is refused with a reason - on the host, no GPU involved."""
    from raytracingproject_b200.device import validate_svm
    for materials in ("principled", "closures", "procedural", "textured", "textured2",
                      "textured3", "transparent"):
        rs = ref.build_scene(scenes.cornell(64, 48, spp=1, materials=materials))
        try:
            assert validate_svm(rs.device_arrays()["__svm_nodes"][0]) is None, materials
        finally:
            rs.close()
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
    rs = ref.build_scene(desc)
    try:
        svm = rs.device_arrays()["__svm_nodes"][0]
        assert "SVM node opcode" in validate_svm(svm)
        assert validate_svm(svm[:-8]) is not None          # not a whole number of nodes
    finally:
        rs.close()


CLOSURE_SCENES = ["principled", "closures", "closures2", "transparent", "textured", "textured3",
                  "textured4", "procedural", "principled+terminator_offset",
                  "principled+multiscatter", "closures_multi"]


@pytest.mark.parametrize("materials", CLOSURE_SCENES)
def test_closure_setup_eval_sample_match_reference(ref, host_lib, materials):
    """Every NODE_CLOSURE_BSDF of the compiled programs: closure setup (svm_closure.cuh),
    then bsdf_eval for a random direction and bsdf_sample for random numbers on each
    closure it made (bsdf.cuh, bsdf_principled.cuh), against the reference's
    svm_node_closure_bsdf / bsdf_eval / bsdf_sample - device source built for the host.
    Parameters come from a stack of plausible values (0.05 .. 0.95)."""
    host_lib.host_svm_closure.restype = C.c_int
    host_lib.host_svm_closure.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint,
                                          C.c_void_p, C.c_float, C.c_float, C.c_void_p]
    text = open(os.path.join(ROOT, "include", "cycles_abi.h")).read()
    ray_diffuse = int(re.search(r"#define CY_PATH_RAY_DIFFUSE[ \t]+(\w+)", text).group(1)
                      .rstrip("u"), 0)
    desc = scenes.cornell(64, 48, spp=1, materials=materials.split("+")[0],
                          distribution="Multiscatter GGX" if materials.endswith("+multiscatter")
                          else "GGX")
    if materials.endswith("+terminator_offset"):
        desc.terminator_offset = 0.6  # every object: shift_cos_in in bsdf_eval / bsdf_sample
    rs = ref.build_scene(desc)
    try:
        arrays = rs.device_arrays()
        real = arrays["__svm_nodes"][0].view(np.uint32).reshape(-1, 4)
        nodes = np.zeros((len(real) + 8, 4), np.uint32)
        nodes[: len(real)] = real
        bound = Bound(host_lib, arrays, nodes)
        a = abi()
        rng = np.random.default_rng(11)
        pts = shading_points(arrays, 32, rng)
        pts["object"], pts["lamp"] = np.maximum(pts["object"], 0), -1
        pts["prim"] = np.maximum(pts["prim"], 0)
        closures_seen, types_seen = 0, set()
        for off in range(len(real)):
            # closure instructions are recognisable: opcode + a closure type in the low byte
            if int(nodes[off, 0]) != a["NODE_CLOSURE_BSDF"] or off < 2:
                continue
            if int(nodes[off - 1, 0]) == a["NODE_VALUE_V"]:
                continue  # the float3 payload of a value node, not an instruction
            for i in range(len(pts)):
                pt = pts[i:i + 1].copy()
                n = pt["N"][0]
                wo = rng.normal(size=3)
                wo = wo / np.linalg.norm(wo)
                pt["I"][0] = wo if wo @ n > 0 else -wo  # viewer above the surface
                wi = rng.normal(size=3).astype(np.float32)
                wi /= np.linalg.norm(wi)
                stack0 = rng.uniform(0.05, 0.95, 264).astype(np.float32)
                if i % 8 == 5:
                    # the zero-parameter branches: Lambert instead of Oren-Nayar, sharp
                    # instead of rough, no sheen / clearcoat / transmission
                    stack0[:] = 0.0
                    stack0[rng.integers(0, 264, 40)] = rng.uniform(0.2, 0.9, 40)
                if (int(nodes[off, 1]) & 0xff) == PRINCIPLED_ID:
                    # subsurface stays zero: anything else is a BSSRDF, refused on the host
                    ss_slot = (int(nodes[off, 1]) >> 16) & 0xff
                    if ss_slot != 255:
                        stack0[ss_slot] = 0.0
                weight = rng.uniform(0.1, 1.0, 3).astype(np.float32)
                flag = ray_diffuse if i % 4 == 3 else 0
                ru, rv = float(rng.uniform(0.01, 0.99)), float(rng.uniform(0.01, 0.99))
                s_ref, s_dev = stack0.copy(), stack0.copy()
                n_ref, o_ref = rs.svm_closure(nodes, off, s_ref, pt, weight, flag, wi, ru, rv)
                o_dev = np.zeros_like(o_ref)
                n_dev = host_lib.host_svm_closure(off, s_dev.ctypes.data, pt.ctypes.data,
                                                  weight.ctypes.data, flag, wi.ctypes.data,
                                                  ru, rv, o_dev.ctypes.data)
                assert n_ref == n_dev, (off, n_ref, n_dev)
                assert o_ref[0] == o_dev[0], (off, "closure count", o_ref[0], o_dev[0])
                k = int(o_ref[0])
                r = o_ref[1:1 + 20 * k].reshape(k, 20)
                d = o_dev[1:1 + 20 * k].reshape(k, 20)
                assert np.array_equal(r[:, 0], d[:, 0]), (off, "closure types", r[:, 0], d[:, 0])
                assert np.array_equal(r[:, 9], d[:, 9]), (off, "sample labels", r[:, 9], d[:, 9])
                bad = ~np.isclose(r, d, rtol=2e-4, atol=2e-5, equal_nan=True)
                assert not bad.any(), (off, i, np.argwhere(bad)[:6].tolist(), r[bad][:6], d[bad][:6], r[0, :17].tolist(), d[0, :17].tolist(), pt.tolist(), wi.tolist())
                closures_seen += k
                types_seen |= set(int(t) for t in r[:, 0])
        assert closures_seen > 0
        print(materials, "closures compared:", closures_seen, "types:", sorted(types_seen))
        del bound
    finally:
        rs.close()

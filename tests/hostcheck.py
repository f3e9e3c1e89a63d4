"""ctypes wrapper over tests/_build/libbvh8_hostcheck.so (test infrastructure)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB = os.path.join(_HERE, "_build", "libbvh8_hostcheck.so")

HIT5 = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("prim", "<i4"), ("object", "<i4")])


def build():
    srcs = [os.path.join(_HERE, "bvh8_hostcheck.cpp"),
            os.path.join(_ROOT, "raytracingproject_b200", "csrc", "bvh8_build.cpp")]
    deps = srcs + [os.path.join(_ROOT, "raytracingproject_b200", "csrc", "bvh8_build.h"),
                   os.path.join(_ROOT, "raytracingproject_b200", "csrc", "bvh8.h")]
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared",
                           "-o", LIB] + srcs)
    return LIB


class HostBVH8:
    def __init__(self, arrays):
        """arrays: {name: (uint8 bytes, elem_size)} as RefScene.device_arrays()."""
        L = C.CDLL(build())
        L.hostcheck_build.restype = C.c_void_p
        L.hostcheck_build.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                      C.c_void_p, C.c_size_t, C.c_int32, C.c_char_p, C.c_size_t]
        L.hostcheck_free.argtypes = [C.c_void_p]
        L.hostcheck_invariants.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.hostcheck_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.hostcheck_info.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        self._L = L
        self._keep = {k: np.ascontiguousarray(v[0]) for k, v in arrays.items()}
        g = lambda n: self._keep[n] if n in self._keep else np.zeros(0, np.uint8)
        p = lambda a: a.ctypes.data if a.size else None
        import re
        abi = open(os.path.join(_ROOT, "include", "cycles_abi.h")).read()
        root_off = int(re.search(r"#define KD_BVH_ROOT\s+(\d+)", abi).group(1))
        ko = int(re.search(r"#define SIZEOF_KERNEL_OBJECT\s+(\d+)", abi).group(1))
        root = int(g("__data")[root_off:root_off + 4].view(np.int32)[0])
        self.num_prims = g("__prim_tri_index").size // 4
        err = C.create_string_buffer(256)
        self._h = L.hostcheck_build(
            p(g("__bvh_nodes")), g("__bvh_nodes").size // 16, p(g("__bvh_leaf_nodes")),
            g("__bvh_leaf_nodes").size // 16, p(g("__prim_tri_verts")), p(g("__prim_tri_index")),
            p(g("__prim_visibility")), p(g("__prim_object")), self.num_prims,
            p(g("__object_node")), p(g("__objects")), g("__objects").size // ko, root, err, 256)
        if not self._h:
            raise RuntimeError("bvh8 build: " + err.value.decode())

    def info(self):
        n, r, t, i = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        d, s = C.c_uint32(), C.c_float()
        self._L.hostcheck_info(self._h, C.byref(n), C.byref(r), C.byref(t), C.byref(i),
                               C.byref(d), C.byref(s))
        return dict(nodes=n.value, records=r.value, triangles=t.value, instances=i.value,
                    depth=d.value, sah=s.value)

    def invariants(self):
        cnt = np.zeros(max(self.num_prims, 1), np.uint32)
        bad = self._L.hostcheck_invariants(self._h, cnt.ctypes.data, self.num_prims)
        return bad, cnt[:self.num_prims]

    def intersect(self, rays):
        rays = np.ascontiguousarray(rays)
        hits = np.zeros(len(rays), HIT5)
        self._L.hostcheck_intersect(self._h, rays.ctypes.data, len(rays), hits.ctypes.data)
        return hits

    def __del__(self):
        try:
            self._L.hostcheck_free(self._h)
        except Exception:
            pass

"""ctypes binding of the plain-C restatement (oracle/cycles_port.c).
TEST INFRASTRUCTURE ONLY - see the header of cycles_port.c."""
import ctypes as C
import os
import re
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_port", "libcycles_port.so")


class PortScene(C.Structure):
    _fields_ = [("bvh_nodes", C.c_void_p), ("bvh_leaf_nodes", C.c_void_p),
                ("prim_tri_verts", C.c_void_p), ("prim_tri_index", C.c_void_p),
                ("prim_visibility", C.c_void_p), ("prim_object", C.c_void_p),
                ("object_node", C.c_void_p), ("objects", C.c_void_p),
                ("object_stride", C.c_uint32), ("object_itfm_offset", C.c_uint32),
                ("root", C.c_int32)]


def build():
    src = os.path.join(_HERE, "cycles_port.c")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-f", "Makefile.port"])
    return LIB


def _abi(name):
    text = open(os.path.join(os.path.dirname(_HERE), "include", "cycles_abi.h")).read()
    return int(re.search(r"#define %s\s+(\d+)" % name, text).group(1))


class PortOracle:
    def __init__(self, arrays):
        """arrays: {kernel_textures name: uint8 ndarray} + "__data"."""
        self._L = C.CDLL(build())
        self._keep = {k: np.ascontiguousarray(v) for k, v in arrays.items()}
        g = lambda n: self._keep[n].ctypes.data if n in self._keep and self._keep[n].size else None
        kd = self._keep["__data"]
        off = _abi("KD_BVH_ROOT")
        self.scene = PortScene(g("__bvh_nodes"), g("__bvh_leaf_nodes"), g("__prim_tri_verts"),
                               g("__prim_tri_index"), g("__prim_visibility"), g("__prim_object"),
                               g("__object_node"), g("__objects"), _abi("SIZEOF_KERNEL_OBJECT"),
                               _abi("KO_ITFM"), int(kd[off:off + 4].view(np.int32)[0]))

    def intersect(self, rays):
        from oracle.cycles_ref import HIT_DTYPE, RAY_DTYPE
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(len(rays), dtype=HIT_DTYPE)
        self._L.port_scene_intersect(C.byref(self.scene), C.c_void_p(rays.ctypes.data),
                                     C.c_void_p(hits.ctypes.data), C.c_uint64(len(rays)))
        return hits

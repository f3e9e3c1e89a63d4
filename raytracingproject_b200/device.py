"""Python host-side mirror of the reference Device interface for the B200
path-tracing device (intern/cycles/device/device.h:288-500), over the C ABI of
include/b200_cycles.h (raytracingproject_b200/libb200cycles.so).

Method names and argument meaning follow the reference: mem_alloc / mem_copy_to /
mem_copy_from / mem_zero / mem_free act on a `DeviceMemory` handle the way
Device::mem_* act on a ccl::device_memory (device/device_memory.h:198-320);
const_copy_to("__data", ...) is Device::const_copy_to; task RENDER is
`render_tile`.  Errors are latched like Device::set_error (device.h:341-348)
AND raised, because Python callers do not poll have_error().

There is no CPU fallback: constructing a device without the CUDA library or
without a B200 raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200_CYCLES_LIB: developer override to A/B-test differently compiled builds (tools/variants.sh)
LIB_PATH = os.environ.get("B200_CYCLES_LIB") or os.path.join(_HERE, "libb200cycles.so")

RAY_DTYPE = np.dtype(
    [("P", "<f4", 3), ("t", "<f4"), ("D", "<f4", 3), ("visibility", "<u4")], align=False)
HIT_DTYPE = np.dtype(
    [("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("prim", "<i4"), ("object", "<i4"), ("type", "<i4")],
    align=False)

# device_memory.h:35-42
MEM_READ_ONLY, MEM_READ_WRITE, MEM_DEVICE_ONLY, MEM_GLOBAL, MEM_TEXTURE, MEM_PIXELS = range(6)
SIZEOF_TEXTURE_INFO = 96  # util/util_texture.h TextureInfo (include/cycles_abi.h; checked in tests)

EXPORTS = [
    "b200_abi_version", "b200_device_count", "b200_device_name", "b200_create", "b200_destroy",
    "b200_last_error", "b200_alloc", "b200_free", "b200_h2d", "b200_d2h", "b200_zero",
    "b200_mem_used", "b200_bind_global", "b200_set_kernel_data", "b200_build_bvh", "b200_render",
    "b200_trace_batch", "b200_film_convert", "b200_film_reduce", "b200_get_stats",
    "b200_synchronize", "b200_set_option", "b200_set_stream", "b200_debug_read",
    "b200_validate_svm", "b200_device_pci_id", "b200_set_cancel_callback", "b200_film_allreduce",
    "b200_texture_set", "b200_texture_clear", "b200_bvh8_pack", "b200_bvh8_free",
    "b200_shader_eval_background",
]

BVH_LAYOUT_BVH2, BVH_LAYOUT_BVH8 = 1 << 0, 1 << 3  # kernel_types.h BVHLayout + INTEGRATION.md


class WorkTile(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("w", C.c_int32), ("h", C.c_int32),
                ("start_sample", C.c_int32), ("num_samples", C.c_int32),
                ("offset", C.c_int32), ("stride", C.c_int32), ("buffer", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [("primary_rays", C.c_uint64), ("bounce_rays", C.c_uint64),
                ("shadow_rays", C.c_uint64),
                ("closest_nodes", C.c_uint64), ("closest_tris", C.c_uint64),
                ("closest_instances", C.c_uint64),
                ("shadow_nodes", C.c_uint64), ("shadow_tris", C.c_uint64),
                ("shadow_instances", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("closest_launches", C.c_uint64),
                ("shadow_launches", C.c_uint64),
                ("device_ms", C.c_double), ("closest_ms", C.c_double), ("shadow_ms", C.c_double),
                ("svm_extended", C.c_uint64),
                ("shade_ms", C.c_double), ("batches", C.c_uint64), ("iterations", C.c_uint64),
                ("host_syncs", C.c_uint64), ("host_waits", C.c_uint64),
                ("shade_wide", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class BVHInfo(C.Structure):
    _fields_ = [("num_nodes", C.c_uint64), ("num_tri_records", C.c_uint64),
                ("num_triangles", C.c_uint64), ("num_instances", C.c_uint64),
                ("node_bytes", C.c_uint64), ("tri_bytes", C.c_uint64),
                ("build_ms", C.c_double), ("sah_cost", C.c_float), ("max_depth", C.c_uint32),
                ("host_packed", C.c_uint32), ("pad", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class PackedBVH2(C.Structure):
    """b200_packed_bvh2: the reference's PackedBVH arrays as BVH2::pack_nodes leaves them."""
    _fields_ = [("nodes", C.c_void_p), ("num_nodes_f4", C.c_size_t),
                ("leaf_nodes", C.c_void_p), ("num_leaf_nodes_f4", C.c_size_t),
                ("prim_tri_verts", C.c_void_p), ("prim_tri_index", C.c_void_p),
                ("prim_visibility", C.c_void_p), ("prim_object", C.c_void_p),
                ("num_prims", C.c_size_t), ("object_node", C.c_void_p),
                ("object_tfm", C.c_void_p), ("num_objects", C.c_size_t), ("root", C.c_int)]


class PackedBVH8(C.Structure):
    _fields_ = [("nodes", C.c_void_p), ("node_bytes", C.c_size_t),
                ("records", C.c_void_p), ("record_bytes", C.c_size_t),
                ("object_node", C.POINTER(C.c_int)), ("root", C.c_uint32),
                ("info", BVHInfo)]



_lib = None


def load_library():
    """dlopen the C-ABI library and declare the prototypes.  Raises if missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "%s is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C raytracingproject_b200/csrc).  There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, u64, sz = C.c_void_p, C.c_uint64, C.c_size_t
    L.b200_create.restype = vp
    L.b200_create.argtypes = [C.c_int, C.c_char_p, sz]
    L.b200_destroy.argtypes = [vp]
    L.b200_last_error.restype = C.c_char_p
    L.b200_last_error.argtypes = [vp]
    L.b200_device_name.argtypes = [C.c_int, C.c_char_p, sz, C.POINTER(C.c_int),
                                   C.POINTER(C.c_int), C.POINTER(u64), C.POINTER(C.c_int)]
    L.b200_alloc.argtypes = [vp, sz, C.POINTER(u64)]
    L.b200_free.argtypes = [vp, u64]
    L.b200_h2d.argtypes = [vp, u64, vp, sz, sz]
    L.b200_d2h.argtypes = [vp, u64, vp, sz, sz]
    L.b200_zero.argtypes = [vp, u64, sz, sz]
    L.b200_mem_used.restype = sz
    L.b200_mem_used.argtypes = [vp]
    L.b200_bind_global.argtypes = [vp, C.c_char_p, u64, vp, sz]
    L.b200_set_kernel_data.argtypes = [vp, vp, sz]
    L.b200_build_bvh.argtypes = [vp, C.POINTER(BVHInfo)]
    L.b200_render.argtypes = [vp, C.POINTER(WorkTile), vp]
    L.b200_trace_batch.argtypes = [vp, u64, u64, u64, C.c_int]
    L.b200_film_convert.argtypes = [vp, u64, u64, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_int]
    L.b200_film_reduce.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(u64), sz]
    L.b200_film_allreduce.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(u64), sz]
    L.b200_device_pci_id.argtypes = [C.c_int, C.c_char_p, sz]
    L.b200_set_cancel_callback.argtypes = [vp, vp, vp]
    L.b200_texture_set.argtypes = [vp, C.c_int, vp, sz, u64]
    L.b200_texture_clear.argtypes = [vp, C.c_int]
    L.b200_bvh8_pack.argtypes = [C.POINTER(PackedBVH2), C.POINTER(PackedBVH8), C.c_char_p, sz]
    L.b200_bvh8_free.argtypes = [C.POINTER(PackedBVH8)]
    L.b200_bvh8_free.restype = None
    L.b200_shader_eval_background.argtypes = [vp, u64, u64, C.c_int, C.c_int]
    L.b200_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.b200_synchronize.argtypes = [vp]
    L.b200_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    L.b200_set_stream.argtypes = [vp, u64]
    L.b200_debug_read.argtypes = [vp, vp, sz]
    L.b200_validate_svm.argtypes = [vp, sz, C.c_char_p, sz]
    _lib = L
    return L


def pack_bvh8(arrays, object_tfm=None):
    """Host-only b200_bvh8_pack over {kernel array name: (bytes, elem size)} holding the
    reference's packed BVH2 (RefScene.device_arrays()): what `BVH8::pack_nodes` does on
    the host.  Returns (nodes bytes, records bytes, object_node int32[], root, info)."""
    import re
    L = load_library()
    g = lambda n: np.ascontiguousarray(arrays[n][0]) if n in arrays else np.zeros(0, np.uint8)
    keep = {n: g(n) for n in ("__bvh_nodes", "__bvh_leaf_nodes", "__prim_tri_verts",
                              "__prim_tri_index", "__prim_visibility", "__prim_object",
                              "__object_node", "__objects", "__data")}
    p = lambda n: keep[n].ctypes.data if keep[n].size else None
    abi = open(os.path.join(os.path.dirname(_HERE), "include", "cycles_abi.h")).read()
    num = lambda k: int(re.search(r"#define %s\s+(\d+)" % k, abi).group(1))
    n_obj = keep["__objects"].size // num("SIZEOF_KERNEL_OBJECT")
    if object_tfm is None and n_obj == 0:
        object_tfm = np.zeros(0, np.float32)
    if object_tfm is None:  # KernelObject::tfm, the first 48 bytes of each record
        rec = keep["__objects"].reshape(n_obj, -1)
        object_tfm = np.ascontiguousarray(rec[:, num("KO_TFM"):num("KO_TFM") + 48]).view(np.float32)
    object_tfm = np.ascontiguousarray(object_tfm, np.float32)
    src = PackedBVH2(p("__bvh_nodes"), keep["__bvh_nodes"].size // 16,
                     p("__bvh_leaf_nodes"), keep["__bvh_leaf_nodes"].size // 16,
                     p("__prim_tri_verts"), p("__prim_tri_index"), p("__prim_visibility"),
                     p("__prim_object"), keep["__prim_tri_index"].size // 4,
                     p("__object_node"), object_tfm.ctypes.data if object_tfm.size else None,
                     min(n_obj, keep["__object_node"].size // 4),
                     int(keep["__data"][num("KD_BVH_ROOT"):num("KD_BVH_ROOT") + 4].view(np.int32)[0]))
    out = PackedBVH8()
    err = C.create_string_buffer(512)
    rc = L.b200_bvh8_pack(C.byref(src), C.byref(out), err, len(err))
    if rc != 0:
        raise DeviceError("b200_bvh8_pack: " + err.value.decode())
    try:
        nodes = np.ctypeslib.as_array((C.c_uint8 * out.node_bytes).from_address(out.nodes)).copy()
        recs = np.ctypeslib.as_array(
            (C.c_uint8 * out.record_bytes).from_address(out.records)).copy() \
            if out.record_bytes else np.zeros(0, np.uint8)
        onode = np.array([out.object_node[i] for i in range(src.num_objects)], np.int32)
        return nodes, recs, onode, int(out.root), out.info.as_dict()
    finally:
        L.b200_bvh8_free(C.byref(out))


class DeviceError(RuntimeError):
    pass


def validate_svm(svm_nodes):
    """b200_validate_svm: None when the compiled SVM program (uint4 array) lies inside
    the supported subset, else the reason it would be refused.  Needs no GPU."""
    import numpy as np
    nodes = np.ascontiguousarray(svm_nodes).view(np.uint8).reshape(-1)
    err = C.create_string_buffer(512)
    rc = load_library().b200_validate_svm(nodes.ctypes.data, nodes.size, err, len(err))
    return None if rc == 0 else err.value.decode()


class DeviceMemory:
    """Counterpart of ccl::device_memory: a host array + its device allocation."""

    def __init__(self, name, host=None, mem_type=MEM_READ_WRITE):
        self.name = name
        self.type = mem_type
        self.host = host  # numpy array or None
        self.device_pointer = 0
        self.device_size = 0

    @property
    def memory_size(self):
        return 0 if self.host is None else self.host.nbytes


class B200Device:
    """Device subclass equivalent (device.h:288).  One instance = one GPU context."""

    def __init__(self, ordinal=0):
        self._L = load_library()
        err = C.create_string_buffer(512)
        self._ctx = self._L.b200_create(int(ordinal), err, len(err))
        self._error = ""
        if not self._ctx:
            raise DeviceError("b200_create: " + err.value.decode())
        self.ordinal = ordinal
        self._globals = {}
        self._textures = {}

    # -- error latch (Device::set_error / have_error / error_message) --
    def have_error(self):
        return bool(self._error)

    def error_message(self):
        return self._error

    def _check(self, rc, what):
        if rc != 0:
            msg = "%s: %s (code %d)" % (what, self._L.b200_last_error(self._ctx).decode(), rc)
            if not self._error:
                self._error = msg
            raise DeviceError(msg)

    def close(self):
        if getattr(self, "_ctx", None):
            self._L.b200_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- memory (device.h:484-488) --
    def mem_alloc(self, mem):
        if mem.device_pointer:
            return
        p = C.c_uint64()
        self._check(self._L.b200_alloc(self._ctx, max(mem.memory_size, 16), C.byref(p)),
                    "mem_alloc(%s)" % mem.name)
        mem.device_pointer = p.value
        mem.device_size = mem.memory_size

    def mem_copy_to(self, mem):
        """(Re)allocate, upload and - for MEM_GLOBAL - bind by name, as
        CUDADevice::mem_copy_to does (device_cuda_impl.cpp:1020-1096)."""
        if mem.device_pointer and mem.device_size != mem.memory_size:
            self.mem_free(mem)
        self.mem_alloc(mem)
        host = np.ascontiguousarray(mem.host)
        if host.nbytes:
            self._check(self._L.b200_h2d(self._ctx, mem.device_pointer, host.ctypes.data, 0,
                                         host.nbytes), "mem_copy_to(%s)" % mem.name)
        if mem.type == MEM_GLOBAL:
            self._check(self._L.b200_bind_global(self._ctx, mem.name.encode(), mem.device_pointer,
                                                 host.ctypes.data if host.nbytes else None,
                                                 host.nbytes), "bind(%s)" % mem.name)
            self._globals[mem.name] = mem

    def mem_copy_from(self, mem, y=0, w=None, h=None, elem=None):
        """Rows [y, y+h) of width w elements of `elem` bytes (device.h:485)."""
        host = mem.host
        if w is None:
            offset, size = 0, host.nbytes
        else:
            offset, size = elem * y * w, elem * w * h
        flat = host.reshape(-1).view(np.uint8)
        self._check(self._L.b200_d2h(self._ctx, mem.device_pointer,
                                     flat[offset:].ctypes.data, offset, size),
                    "mem_copy_from(%s)" % mem.name)

    def mem_zero(self, mem):
        self.mem_alloc(mem)
        self._check(self._L.b200_zero(self._ctx, mem.device_pointer, 0, mem.memory_size),
                    "mem_zero(%s)" % mem.name)
        if mem.host is not None:
            mem.host[...] = 0

    def mem_free(self, mem):
        if mem.device_pointer:
            self._check(self._L.b200_free(self._ctx, mem.device_pointer), "mem_free")
            mem.device_pointer = 0
            mem.device_size = 0
            self._globals.pop(mem.name, None)

    def mem_used(self):
        return self._L.b200_mem_used(self._ctx)

    def const_copy_to(self, name, host_bytes):
        """Device::const_copy_to - only ever "__data" (render/scene.cpp:307)."""
        if name != "__data":
            raise DeviceError("const_copy_to: unknown constant " + name)
        buf = np.ascontiguousarray(host_bytes).view(np.uint8)
        self._check(self._L.b200_set_kernel_data(self._ctx, buf.ctypes.data, buf.nbytes),
                    "const_copy_to(__data)")

    # -- image textures (CUDADevice::tex_alloc / tex_free) --
    def tex_alloc(self, slot, texture_info, pixels):
        """One image slot: `texture_info` = the reference's TextureInfo record (bytes),
        `pixels` = the host pixel array (any dtype; uploaded as it is)."""
        mem = DeviceMemory("tex_%d" % slot, np.ascontiguousarray(pixels).view(np.uint8),
                           MEM_TEXTURE)
        old = self._textures.pop(slot, None)
        if old is not None:
            self.mem_free(old)
        self.mem_alloc(mem)
        host = np.ascontiguousarray(mem.host)
        self._check(self._L.b200_h2d(self._ctx, mem.device_pointer, host.ctypes.data, 0,
                                     host.nbytes), "tex_alloc(%d)" % slot)
        info = np.ascontiguousarray(texture_info).view(np.uint8)
        self._check(self._L.b200_texture_set(self._ctx, int(slot), info.ctypes.data, info.nbytes,
                                             mem.device_pointer), "tex_alloc(%d)" % slot)
        self._textures[slot] = mem

    def tex_free(self, slot):
        mem = self._textures.pop(slot, None)
        if mem is not None:
            self._check(self._L.b200_texture_clear(self._ctx, int(slot)), "tex_free")
            self.mem_free(mem)

    # -- convenience: upload everything Scene::device_update would --
    def upload_scene(self, arrays, textures=None):
        """arrays: {kernel_textures name: (uint8 bytes, elem_size)} + "__data";
        textures: [(slot, TextureInfo bytes, pixels)] - the ImageManager's images."""
        for slot in list(self._textures):
            self.tex_free(slot)
        for slot, info, pixels in (textures or []):
            self.tex_alloc(slot, info, pixels)
        for name, (data, _es) in arrays.items():
            if name == "__data":
                continue
            mem = DeviceMemory(name, np.ascontiguousarray(data), MEM_GLOBAL)
            self.mem_copy_to(mem)
        self.const_copy_to("__data", arrays["__data"][0])

    def build_bvh(self):
        info = BVHInfo()
        self._check(self._L.b200_build_bvh(self._ctx, C.byref(info)), "build_bvh")
        return info.as_dict()

    def set_option(self, name, value):
        self._check(self._L.b200_set_option(self._ctx, name.encode(), int(value)), "set_option")

    def debug_read(self):
        """(16, 32) float32 records of the path selected with option debug_slot."""
        out = np.zeros((16, 32), np.float32)
        self._check(self._L.b200_debug_read(self._ctx, out.ctypes.data, out.size), "debug_read")
        return out

    def set_stream(self, cuda_stream):
        """Issue all work on an existing CUDA stream (0 = the private stream)."""
        self._check(self._L.b200_set_stream(self._ctx, int(cuda_stream)), "set_stream")

    def synchronize(self):
        self._check(self._L.b200_synchronize(self._ctx), "synchronize")

    def stats(self):
        s = Stats()
        self._check(self._L.b200_get_stats(self._ctx, C.byref(s)), "get_stats")
        return s.as_dict()

    # -- parity / bench hook --
    def trace_batch(self, rays, any_hit=False):
        """scene_intersect on a host ray batch; returns the HIT_DTYPE array."""
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        rmem = DeviceMemory("rays", rays.view(np.uint8), MEM_READ_ONLY)
        hits = np.zeros(len(rays), dtype=HIT_DTYPE)
        hmem = DeviceMemory("hits", hits.view(np.uint8), MEM_READ_WRITE)
        self.mem_copy_to(rmem)
        self.mem_alloc(hmem)
        try:
            self.trace_device(rmem.device_pointer, hmem.device_pointer, len(rays), any_hit)
            self.mem_copy_from(hmem)
        finally:
            self.mem_free(rmem)
            self.mem_free(hmem)
        return hits

    def trace_device(self, rays_dptr, hits_dptr, n, any_hit=False):
        self._check(self._L.b200_trace_batch(self._ctx, rays_dptr, hits_dptr, int(n),
                                             int(bool(any_hit))), "trace_batch")

    # -- DeviceTask::RENDER for one tile --
    def render_tile(self, film_dptr, x, y, w, h, start_sample, num_samples, offset, stride):
        wt = WorkTile(x, y, w, h, start_sample, num_samples, offset, stride, film_dptr)
        self._check(self._L.b200_render(self._ctx, C.byref(wt), None), "render")

    # -- DeviceTask::FILM_CONVERT --
    def film_convert(self, film, width, height, num_samples, half_float=False):
        """kernel_film_convert_to_byte / _to_half_float over a full frame: `film` is a
        DeviceMemory holding (h, w, pass_stride) float sums on the device; returns
        (h, w, 4) uint8 display bytes or (h, w, 4) uint16 half bit patterns.
        sample_scale = 1 / num_samples (device_cpu.cpp:1324)."""
        out = DeviceMemory("display_rgba",
                           np.zeros((height, width, 4), np.uint16 if half_float else np.uint8))
        self.mem_zero(out)
        try:
            self._check(self._L.b200_film_convert(
                self._ctx, film.device_pointer, out.device_pointer, int(bool(half_float)),
                C.c_float(1.0 / float(num_samples)), 0, 0, width, height, 0, width),
                "film_convert")
            self.mem_copy_from(out)
        finally:
            self.mem_free(out)
        return out.host

    @staticmethod
    def film_reduce(devices, films, n_floats):
        """In-process multi-GPU film sum into films[0] (b200_film_reduce): devices[i] owns
        the DeviceMemory films[i]."""
        n = len(devices)
        ctxs = (C.c_void_p * n)(*[d._ctx for d in devices])
        ptrs = (C.c_uint64 * n)(*[f.device_pointer for f in films])
        rc = devices[0]._L.b200_film_reduce(ctxs, n, ptrs, int(n_floats))
        devices[0]._check(rc, "film_reduce")

    @staticmethod
    def film_allreduce(devices, films, n_floats):
        """The same sum as one NCCL all-reduce over NVLink (b200_film_allreduce): one
        device per GPU, every film ends up holding the sum."""
        n = len(devices)
        ctxs = (C.c_void_p * n)(*[d._ctx for d in devices])
        ptrs = (C.c_uint64 * n)(*[f.device_pointer for f in films])
        rc = devices[0]._L.b200_film_allreduce(ctxs, n, ptrs, int(n_floats))
        devices[0]._check(rc, "film_allreduce")

    def render(self, width, height, pass_stride, start_sample, num_samples, film=None):
        """Full-frame RENDER into a fresh (or given) RenderBuffers-like film and
        read it back: returns (h, w, pass_stride) float32 sums."""
        if film is None:
            film = DeviceMemory("RenderBuffers",
                                np.zeros((height, width, pass_stride), np.float32))
            self.mem_zero(film)
            own = True
        else:
            own = False
        try:
            self.render_tile(film.device_pointer, 0, 0, width, height, start_sample, num_samples,
                             0, width)
            self.mem_copy_from(film)
        finally:
            if own:
                self.mem_free(film)
        return film.host


SHIM_PATH = os.path.join(_HERE, "libcycles_device_b200.so")


class RegisteredDevice:
    """A ccl::Device made by the REFERENCE's own registry - `Device::type_from_string`,
    `Device::available_devices`, `Device::create` (device/device.cpp:367-550) - the way
    `cycles --device B200` or Blender's preferences would make it, once the registration
    patch of INTEGRATION.md section 2 is in.  `count` > 1 asks for
    `Device::get_multi_device` of the first `count` devices of the type (the reference's
    MultiDevice over B200 sub-devices).  `.ptr` is the ccl::Device*."""

    def __init__(self, type_name="B200", index=0, count=1):
        load_library()
        if not os.path.exists(SHIM_PATH):
            raise DeviceError("%s is missing" % SHIM_PATH)
        S = C.CDLL(SHIM_PATH, mode=C.RTLD_GLOBAL)  # loading it registers the device type
        S.ref_host_device_type_from_string.argtypes = [C.c_char_p]
        S.ref_host_device_type_name.argtypes = [C.c_int, C.c_char_p, C.c_int]
        S.ref_host_device_type_available.argtypes = [C.c_int]
        S.ref_host_available_devices.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_int,
                                                 C.c_char_p, C.c_int]
        S.ref_host_device_create.restype = C.c_void_p
        S.ref_host_device_create.argtypes = [C.c_int, C.c_int, C.c_int]
        S.ref_host_device_free.argtypes = [C.c_void_p]
        S.b200_registered_device_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        self._S = S
        self.type = S.ref_host_device_type_from_string(type_name.encode())
        if self.type == 0:
            raise DeviceError("the reference's device registry does not know %r" % type_name)
        self._ptr = None
        if count > 0:
            self._ptr = S.ref_host_device_create(self.type, int(index), int(count))
            if not self._ptr:
                raise DeviceError("Device::create(%s #%d) returned NULL" % (type_name, index))

    def type_name(self):
        buf = C.create_string_buffer(64)
        self._S.ref_host_device_type_name(self.type, buf, len(buf))
        return buf.value.decode()

    def type_available(self):
        """Device::available_types() lists the type."""
        return bool(self._S.ref_host_device_type_available(self.type))

    def available(self):
        """[(id, description)] of Device::available_devices(mask of the type)."""
        out = []
        n = self._S.ref_host_available_devices(self.type, -1, None, 0, None, 0)
        for i in range(n):
            a, b = C.create_string_buffer(256), C.create_string_buffer(256)
            self._S.ref_host_available_devices(self.type, i, a, len(a), b, len(b))
            out.append((a.value.decode(), b.value.decode()))
        return out

    @property
    def ptr(self):
        return self._ptr

    def stats(self):
        s = Stats()
        if self._S.b200_registered_device_stats(self._ptr, C.byref(s)) != 0:
            raise DeviceError("the registry's device is not a B200Device")
        return s.as_dict()

    def close(self):
        if getattr(self, "_ptr", None):
            self._S.ref_host_device_free(self._ptr)
            self._ptr = None


class B200HostDevice:
    """The C++ `B200Device : ccl::Device` (csrc/device_b200.cpp) as a handle whose
    `.ptr` is a ccl::Device* that a reference Scene / DeviceTask can drive - the real
    drop-in path.  Needs libcycles_device_b200.so (built where the reference headers
    are available) and the host application library it links to."""

    def __init__(self, ordinal=0):
        """ordinal: one GPU (B200Device), or a list of GPUs for the in-process
        B200MultiDevice (sample split + device-side film sum); a list may repeat an
        ordinal, which puts several contexts on one GPU."""
        load_library()
        if not os.path.exists(SHIM_PATH):
            raise DeviceError("%s is missing (make -C raytracingproject_b200/csrc -f "
                              "Makefile.device, needs /root/reference)" % SHIM_PATH)
        S = C.CDLL(SHIM_PATH, mode=C.RTLD_GLOBAL)
        S.b200_host_device_create.restype = C.c_void_p
        S.b200_host_device_create.argtypes = [C.c_int, C.c_char_p, C.c_size_t]
        S.b200_host_device_ptr.restype = C.c_void_p
        S.b200_host_device_ptr.argtypes = [C.c_void_p]
        S.b200_host_device_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        S.b200_host_device_error.restype = C.c_char_p
        S.b200_host_device_error.argtypes = [C.c_void_p]
        S.b200_host_device_destroy.argtypes = [C.c_void_p]
        S.b200_host_device_bvh_info.argtypes = [C.c_void_p, C.POINTER(BVHInfo)]
        S.b200_host_bvh8_report.argtypes = [C.POINTER(BVHInfo), C.POINTER(C.c_double),
                                            C.c_char_p, C.c_size_t]
        self._S = S
        S.b200_host_multi_device_create.restype = C.c_void_p
        S.b200_host_multi_device_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_char_p,
                                                    C.c_size_t]
        err = C.create_string_buffer(512)
        if isinstance(ordinal, (list, tuple)):
            ords = (C.c_int * len(ordinal))(*[int(o) for o in ordinal])
            self._h = S.b200_host_multi_device_create(ords, len(ordinal), err, len(err))
        else:
            self._h = S.b200_host_device_create(int(ordinal), err, len(err))
        if not self._h:
            raise DeviceError("B200Device: " + err.value.decode())

    @property
    def ptr(self):
        return self._S.b200_host_device_ptr(self._h)

    def stats(self):
        s = Stats()
        self._S.b200_host_device_stats(self._h, C.byref(s))
        return s.as_dict()

    def error_message(self):
        return self._S.b200_host_device_error(self._h).decode()

    def bvh_info(self):
        """The BVH the device traverses for the scene bound last; host_packed = 1 when the
        host's `BVH8 : BVH` delivered it (BVH_LAYOUT_BVH8), 0 when the device derived it
        from packed BVH2 arrays."""
        info = BVHInfo()
        rc = self._S.b200_host_device_bvh_info(self._h, C.byref(info))
        if rc != 0:
            raise DeviceError("bvh_info: " + self.error_message())
        return info.as_dict()

    def host_bvh8_report(self):
        """(info, seconds, error) of the last top-level BVH8::pack_nodes on the host."""
        info, sec, err = BVHInfo(), C.c_double(), C.create_string_buffer(512)
        self._S.b200_host_bvh8_report(C.byref(info), C.byref(sec), err, len(err))
        return info.as_dict(), sec.value, err.value.decode()

    def close(self):
        if getattr(self, "_h", None):
            self._S.b200_host_device_destroy(self._h)
            self._h = None

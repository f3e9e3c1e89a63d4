"""ctypes binding over oracle/_ref/libcycles_ref.so - the REFERENCE's own CPU
Cycles (BVH2 path) compiled from /root/reference by oracle/Makefile.

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs - never by the product
path (raytracingproject_b200/), which must fail loudly without its CUDA library.

The library is the oracle ("kind": "reference"): scene flattening
(render/scene.cpp:193-321), BVH2 build (bvh/bvh_build.cpp:370), and the CPU
kernels (kernel/kernels/cpu/kernel.cpp generic = parity variant,
kernel_avx2.cpp = speed variant) are the reference's code, unmodified.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libcycles_ref.so")

RAY_DTYPE = np.dtype(
    [("P", "<f4", 3), ("t", "<f4"), ("D", "<f4", 3), ("visibility", "<u4")], align=False
)
HIT_DTYPE = np.dtype(
    [("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("prim", "<i4"), ("object", "<i4"), ("type", "<i4")],
    align=False,
)
assert RAY_DTYPE.itemsize == 32 and HIT_DTYPE.itemsize == 24

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(
                "%s missing: run `make -C oracle -j8` where /root/reference exists" % LIB_PATH
            )
        L = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
        L.ref_scene_new.restype = C.c_void_p
        L.ref_scene_new.argtypes = [C.c_char_p, C.c_int, C.c_void_p]
        L.ref_scene_free.argtypes = [C.c_void_p]
        L.ref_scene_error.restype = C.c_char_p
        L.ref_scene_error.argtypes = [C.c_void_p]
        L.ref_scene_add_mesh.argtypes = [
            C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_char_p, C.c_int]
        L.ref_scene_add_object.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_scene_update.argtypes = [C.c_void_p, C.c_int, C.c_int]
        for f in ("ref_scene_width", "ref_scene_height", "ref_scene_pass_stride"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.ref_scene_global.argtypes = [
            C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64),
            C.POINTER(C.c_uint32)]
        L.ref_global_name.restype = C.c_char_p
        L.ref_global_name.argtypes = [C.c_int]
        L.ref_scene_data.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        L.ref_scene_num_textures.argtypes = [C.c_void_p]
        L.ref_scene_texture.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64,
                                        C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        L.ref_render.argtypes = [
            C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_double)]
        L.ref_film_convert.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.ref_render_tile_buffers.argtypes = [
            C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        L.ref_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.ref_camera_rays.argtypes = [
            C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_shadow_rays.argtypes = [
            C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.ref_init.argtypes = [C.c_int]
        L.ref_path_dump.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.ref_svm_closure.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_uint, C.c_void_p, C.c_float, C.c_float,
                                      C.c_void_p, C.c_void_p]
        L.ref_svm_node.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p]
        L.ref_count_rays.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.ref_num_threads.restype = C.c_int
        _lib = L
    return _lib


def global_names():
    L = lib()
    out, i = [], 0
    while True:
        n = L.ref_global_name(i)
        if n is None:
            return out
        out.append(n.decode())
        i += 1


SHADING_POINT_DTYPE = np.dtype([("P", "<f4", 3), ("N", "<f4", 3), ("I", "<f4", 3),
                                ("dPdu", "<f4", 3), ("u", "<f4"), ("v", "<f4"),
                                ("object", "<i4"), ("prim", "<i4"), ("lamp", "<i4"),
                                ("shader", "<i4"), ("backfacing", "<i4")])


class RefScene:
    """One reference Scene bound to the reference CPUDevice (or to an external
    ccl::Device* - used to drive the B200 device shim through the same Scene)."""

    GENERIC, AVX2 = 0, 1

    def __init__(self, xml_path, kernel=GENERIC, external_device=None, threads=0):
        L = lib()
        L.ref_init(int(threads))
        self._L = L
        self._h = L.ref_scene_new(os.fsencode(xml_path), int(kernel), external_device)
        if not self._h:
            raise RuntimeError("ref_scene_new failed")
        self.kernel = kernel

    def close(self):
        if self._h:
            self._L.ref_scene_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("reference: " + self._L.ref_scene_error(self._h).decode())

    def add_mesh(self, P, tris, shader, smooth=False):
        P = np.ascontiguousarray(P, dtype=np.float32).reshape(-1, 3)
        tris = np.ascontiguousarray(tris, dtype=np.int32).reshape(-1, 3)
        m = self._L.ref_scene_add_mesh(
            self._h, P.ctypes.data, len(P), tris.ctypes.data, len(tris), shader.encode(),
            int(smooth))
        if m < 0:
            self._check(1)
        return m

    def add_object(self, mesh, tfm=None):
        if tfm is None:
            tfm = np.eye(4, dtype=np.float32)[:3]
        tfm = np.ascontiguousarray(tfm, dtype=np.float32).reshape(3, 4)
        o = self._L.ref_scene_add_object(self._h, int(mesh), tfm.ctypes.data)
        if o < 0:
            raise RuntimeError("bad mesh handle")
        return o

    def update(self, width=0, height=0):
        self._check(self._L.ref_scene_update(self._h, int(width), int(height)))

    @property
    def width(self):
        return self._L.ref_scene_width(self._h)

    @property
    def height(self):
        return self._L.ref_scene_height(self._h)

    @property
    def pass_stride(self):
        return self._L.ref_scene_pass_stride(self._h)

    def kernel_data(self):
        p, n = C.c_void_p(), C.c_uint64()
        self._L.ref_scene_data(self._h, C.byref(p), C.byref(n))
        return np.ctypeslib.as_array((C.c_uint8 * n.value).from_address(p.value)).copy()

    def global_array(self, name):
        """Raw bytes of one kernel_textures.h array + its element size."""
        p, n, es = C.c_void_p(), C.c_uint64(), C.c_uint32()
        rc = self._L.ref_scene_global(self._h, name.encode(), C.byref(p), C.byref(n), C.byref(es))
        if rc != 0:
            raise KeyError(name)
        nbytes = n.value * es.value
        if nbytes == 0 or not p.value:
            return np.zeros(0, dtype=np.uint8), es.value
        return (
            np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(p.value)).copy(),
            es.value,
        )

    def device_arrays(self):
        """{name: (bytes, elem_size)} for every non-empty kernel array + '__data'."""
        out = {}
        for name in global_names():
            try:
                a, es = self.global_array(name)
            except KeyError:  # device-owned arrays (__texture_info) are not in DeviceScene
                continue
            if a.size:
                out[name] = (a, es)
        out["__data"] = (self.kernel_data(), 1)
        return out

    def pack_bvh(self, layout):
        """The scene's top-level BVH packed in `layout` by the reference's own
        BVH::create / BVH::build (no device): (nodes bytes, leaf_nodes bytes, object_node
        int32[], root)."""
        L = self._L
        L.ref_scene_pack_bvh.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 7
        n, nb, l, lb = C.c_void_p(), C.c_uint64(), C.c_void_p(), C.c_uint64()
        o, no, root = C.c_void_p(), C.c_uint64(), C.c_int()
        self._check(L.ref_scene_pack_bvh(self._h, int(layout), C.byref(n), C.byref(nb),
                                         C.byref(l), C.byref(lb), C.byref(o), C.byref(no),
                                         C.byref(root)))
        grab = lambda p, nbytes: (np.ctypeslib.as_array(
            (C.c_uint8 * nbytes).from_address(p.value)).copy() if nbytes and p.value
            else np.zeros(0, np.uint8))
        return (grab(n, nb.value), grab(l, lb.value),
                grab(o, no.value * 4).view(np.int32), root.value)

    def pass_offset(self, pass_type):
        """(float offset inside a film pixel, components) of a render pass, or None."""
        self._L.ref_scene_pass_offset.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        n = C.c_int()
        off = self._L.ref_scene_pass_offset(self._h, int(pass_type), C.byref(n))
        return None if off < 0 else (off, n.value)

    def denoising_offset(self):
        """(float offset of the denoising data passes, of the clean pass) in a film pixel,
        0 = absent (KernelFilm::pass_denoising_data / pass_denoising_clean)."""
        self._L.ref_scene_denoising_offset.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        clean = C.c_int()
        return self._L.ref_scene_denoising_offset(self._h, C.byref(clean)), clean.value

    def textures(self):
        """[(slot, TextureInfo bytes, pixel bytes)] - the images the reference's
        ImageManager loaded into its device (CPUDevice::tex_alloc), for
        B200Device.upload_scene(arrays, textures)."""
        from raytracingproject_b200.device import SIZEOF_TEXTURE_INFO
        out = []
        for slot in range(self._L.ref_scene_num_textures(self._h)):
            info = np.zeros(SIZEOF_TEXTURE_INFO, np.uint8)
            p, n = C.c_void_p(), C.c_uint64()
            if self._L.ref_scene_texture(self._h, slot, info.ctypes.data, info.nbytes,
                                         C.byref(p), C.byref(n)) != 0:
                raise RuntimeError("ref_scene_texture(%d)" % slot)
            if n.value and p.value:
                pix = np.ctypeslib.as_array((C.c_uint8 * n.value).from_address(p.value)).copy()
                out.append((slot, info, pix))
        return out

    def render(self, start_sample, num_samples, tile_size=64, accumulate=False):
        """Film sums, shape (h, w, pass_stride) float32, and the wall seconds."""
        w, h, ps = self.width, self.height, self.pass_stride
        out = np.empty((h, w, ps), dtype=np.float32)
        sec = C.c_double()
        self._check(self._L.ref_render(
            self._h, int(start_sample), int(num_samples), int(tile_size), int(accumulate),
            out.ctypes.data, C.byref(sec)))
        return out, sec.value

    def render_tile_buffers(self, start_sample, num_samples, tile_size=64, cancel_after=-1):
        """The background shape of Session's tile flow: a RenderBuffers per tile, deleted
        in release_tile on the device's worker thread; task.get_cancel() turns true after
        `cancel_after` released tiles.  Returns (film, tiles completed)."""
        w, h, ps = self.width, self.height, self.pass_stride
        out = np.empty((h, w, ps), dtype=np.float32)
        done = C.c_int()
        self._check(self._L.ref_render_tile_buffers(
            self._h, int(start_sample), int(num_samples), int(tile_size), int(cancel_after),
            out.ctypes.data, C.byref(done)))
        return out, done.value

    def film_convert(self, num_samples, half_float=False):
        """DeviceTask::FILM_CONVERT of the last rendered film on this scene's device:
        (h, w, 4) uint8 display bytes, or (h, w, 4) uint16 half bit patterns."""
        w, h = self.width, self.height
        out = np.empty((h, w, 4), dtype=np.uint16 if half_float else np.uint8)
        self._check(self._L.ref_film_convert(self._h, int(num_samples), int(bool(half_float)),
                                             out.ctypes.data))
        return out

    def path_dump(self, sample, x, y):
        """(16, 32) per-bounce debug records of one reference path (ref_probe_path_dump)."""
        out = np.zeros((16, 32), np.float32)
        self._check(self._L.ref_path_dump(self._h, sample, x, y, out.ctypes.data))
        return out

    def svm_node(self, nodes, offset, stack, point):
        """Runs the node at `offset` of the uint4 program `nodes` through the reference's
        svm_node_* function on one shading point (SHADING_POINT_DTYPE record); `stack`
        (float32, >= 260) is updated in place.  Returns the offset after the node, -1
        for an opcode the probe does not dispatch."""
        nxt = C.c_int(-1)
        self._check(self._L.ref_svm_node(self._h, nodes.ctypes.data, int(offset),
                                         stack.ctypes.data, point.ctypes.data, C.byref(nxt)))
        return nxt.value

    def svm_closure(self, nodes, offset, stack, point, closure_weight, path_flag, omega_in,
                    randu, randv):
        """NODE_CLOSURE_BSDF at `offset` through the reference, then bsdf_eval(omega_in) and
        bsdf_sample(randu, randv) on every closure: (next offset, float32[1 + 20 * 32])."""
        out = np.zeros(1 + 20 * 32, np.float32)
        nxt = C.c_int(-1)
        w = np.ascontiguousarray(closure_weight, np.float32)
        wi = np.ascontiguousarray(omega_in, np.float32)
        self._check(self._L.ref_svm_closure(self._h, nodes.ctypes.data, int(offset),
                                            stack.ctypes.data, point.ctypes.data, w.ctypes.data,
                                            int(path_flag), wi.ctypes.data, float(randu),
                                            float(randv), out.ctypes.data, C.byref(nxt)))
        return nxt.value, out

    def count_rays(self, start_sample, num_samples):
        """(camera, bounce, shadow) rays the reference traces for these samples."""
        counts = np.zeros(3, dtype=np.uint64)
        self._check(self._L.ref_count_rays(self._h, int(start_sample), int(num_samples),
                                           counts.ctypes.data))
        return tuple(int(c) for c in counts)

    def num_threads(self):
        return int(self._L.ref_num_threads())

    def intersect(self, rays):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.empty(len(rays), dtype=HIT_DTYPE)
        self._check(self._L.ref_intersect(self._h, rays.ctypes.data, hits.ctypes.data, len(rays)))
        return hits

    def camera_rays(self, sample, x0, y0, w, h):
        rays = np.empty(w * h, dtype=RAY_DTYPE)
        hashes = np.empty(w * h, dtype=np.uint32)
        self._check(self._L.ref_camera_rays(
            self._h, sample, x0, y0, w, h, rays.ctypes.data, hashes.ctypes.data))
        return rays, hashes

    def shadow_rays(self, sample, x0, y0, w, h):
        rays = np.empty(w * h, dtype=RAY_DTYPE)
        self._check(self._L.ref_shadow_rays(self._h, sample, x0, y0, w, h, rays.ctypes.data))
        return rays


def build_scene(desc, kernel=RefScene.GENERIC, external_device=None, threads=0, tmpdir=None):
    """Flatten a raytracingproject_b200.scenes.SceneDesc through the reference's
    own host code; returns the updated RefScene."""
    import tempfile

    d = tmpdir or tempfile.mkdtemp(prefix="cyref_")
    path = os.path.join(d, desc.name + ".xml")
    with open(path, "w") as f:
        f.write(desc.xml)
    if getattr(desc, "images", None):
        from raytracingproject_b200.scenes import write_b2im
        for fname, pixels in desc.images.items():
            write_b2im(os.path.join(d, fname), pixels)
    rs = RefScene(path, kernel=kernel, external_device=external_device, threads=threads)
    for pass_type in getattr(desc, "passes", None) or []:
        rs._L.ref_scene_add_pass.argtypes = [C.c_void_p, C.c_int]
        rs._L.ref_scene_add_pass(rs._h, int(pass_type))
    dn = getattr(desc, "denoising", None)  # (clean pass?, DenoiseFlag bits)
    if dn is not None:
        rs._L.ref_scene_set_denoising.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        rs._L.ref_scene_set_denoising(rs._h, 1, int(bool(dn[0])), int(dn[1]))
    handles = [rs.add_mesh(m.P, m.tris, m.shader, m.smooth) for m in desc.meshes]
    for mi, tfm in desc.objects:
        o = rs.add_object(handles[mi], tfm)
        off = getattr(desc, "terminator_offset", 0.0)
        if off:
            rs._L.ref_scene_set_terminator_offset.argtypes = [C.c_void_p, C.c_int, C.c_float]
            rs._L.ref_scene_set_terminator_offset(rs._h, o, C.c_float(off))
    for shader, node, tiles in getattr(desc, "image_tiles", None) or []:
        arr = (C.c_int * len(tiles))(*tiles)
        rs._L.ref_scene_set_image_tiles.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p,
                                                    C.POINTER(C.c_int), C.c_int]
        rs._check(rs._L.ref_scene_set_image_tiles(rs._h, shader.encode(), node.encode(), arr,
                                                  len(tiles)))
    rs.update(desc.width, desc.height)
    return rs

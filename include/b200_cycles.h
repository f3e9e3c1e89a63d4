/* b200_cycles.h - C ABI of the B200-native Cycles path-tracing device.
 *
 * This is the drop-in boundary: everything CUDA-specific lives behind these
 * entry points (libb200cycles.so); the host side is either the C++ Device
 * subclass (raytracingproject_b200/csrc/device_b200.cpp, a ccl::Device exactly
 * like the reference's CUDADevice) or the Python mirror
 * (raytracingproject_b200/device.py).  Plain pointers and sizes only; no
 * exceptions cross the boundary; every function returns 0 on success and a
 * non-zero B200_ERR_* code on failure, with the text in b200_last_error().
 *
 * Each entry point cites the reference interface it stands in for.  Reference
 * paths are relative to /root/reference/blender/intern/cycles/.
 */
#ifndef B200_CYCLES_H
#define B200_CYCLES_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_ABI_VERSION 1

enum {
  B200_OK = 0,
  B200_ERR_CUDA = 1,        /* a CUDA runtime call failed */
  B200_ERR_INVALID = 2,     /* bad argument / unknown name / size mismatch */
  B200_ERR_UNSUPPORTED = 3, /* scene uses a feature outside the hot-path scope */
  B200_ERR_NOT_READY = 4,   /* render before the scene was bound */
  B200_ERR_OOM = 5,
  B200_ERR_CANCELLED = 6
};

typedef struct b200_ctx b200_ctx;

/* Ray / hit records of the batch-trace hook.  Same layout as the oracle's
 * RefProbeRay / RefProbeHit so one dumped batch feeds both sides.
 * b200_hit mirrors struct Intersection (kernel/kernel_types.h:672-686). */
typedef struct b200_ray {
  float P[3];
  float t; /* max distance; 0 = inactive ray */
  float D[3];
  uint32_t visibility; /* PATH_RAY_* mask tested against __prim_visibility */
} b200_ray;

typedef struct b200_hit {
  float t, u, v;
  int32_t prim;   /* index into the reference's packed prim arrays; -1 = miss */
  int32_t object; /* instance object id, -1 (OBJECT_NONE) for static geometry */
  int32_t type;   /* PRIMITIVE_* */
} b200_hit;

/* Fields of WorkTile (kernel/kernel_types.h:1690-1700) / RenderTile
 * (render/buffers.h:135-160) that RENDER needs. */
typedef struct b200_work_tile {
  int32_t x, y, w, h;
  int32_t start_sample, num_samples;
  int32_t offset, stride;
  uint64_t buffer; /* device pointer of the float film (RenderBuffers::buffer) */
} b200_work_tile;

/* Counters of the last b200_render / b200_trace_batch (device-side counts, not
 * estimates): rays actually traversed and the BVH work they did. */
typedef struct b200_stats {
  uint64_t primary_rays;  /* scene_intersect calls for camera rays */
  uint64_t bounce_rays;   /* scene_intersect calls for bounce rays */
  uint64_t shadow_rays;   /* shadow_blocked traversals */
  /* BVH work, counted only when option "count_traversal" is on (else 0): the
   * equivalent of the reference's __KERNEL_DEBUG__ passes (bvh_types.h:46-70) */
  uint64_t closest_nodes, closest_tris, closest_instances; /* intersect_closest */
  uint64_t shadow_nodes, shadow_tris, shadow_instances;    /* intersect_shadow */
  uint64_t kernel_launches;  /* our kernels launched inside the call */
  uint64_t closest_launches; /* intersect_closest launches among them */
  uint64_t shadow_launches;  /* intersect_shadow launches among them */
  double device_ms;          /* CUDA-event time of the call's device work */
  double closest_ms;         /* CUDA-event time inside intersect_closest launches */
  double shadow_ms;          /* CUDA-event time inside intersect_shadow launches */
  uint64_t svm_extended;     /* 1: shading ran the full SVM interpreter kernels (the bound
                              * program holds texture / attribute / colour nodes or sheen) */
  /* wavefront control (b200_render): the bounce loop is queued ahead of the host, which
   * reads each iteration's counters one iteration late - see DESIGN.md section 5 */
  double shade_ms;           /* CUDA-event time of sort + shade_background + shade_surface */
  uint64_t batches;          /* wavefront batches (pixels x samples that fit the path pool) */
  uint64_t iterations;       /* bounce iterations that had paths to work on */
  uint64_t host_syncs;       /* stream synchronisations: the device drains, then waits for
                              * the host (one per call; one per step of a transparent shadow) */
  uint64_t host_waits;       /* waits on an iteration's counters while later work is queued */
  int64_t shade_wide;        /* block shape of the lean multiscatter / full shading kernel after
                                this call: -1 still probing, 0 = two blocks of 256 threads per
                                SM, 1 = one block of 512, 2 = one block of 1024 */
} b200_stats;

/* BVH8 build report (host builder). */
typedef struct b200_bvh_info {
  uint64_t num_nodes, num_tri_records, num_triangles, num_instances;
  uint64_t node_bytes, tri_bytes;
  double build_ms;
  float sah_cost;
  uint32_t max_depth;
  uint32_t host_packed; /* 1: the host bound BVH8 arrays (layout below), nothing was built here */
  uint32_t pad;
} b200_bvh_info;

/* The device's own host BVH layout - what `BVH_LAYOUT_BVH8` is in the reference's
 * BVHLayout enum once the patch of INTEGRATION.md section 2 is applied
 * (kernel/kernel_types.h:1396-1406; BVH_LAYOUT_BVH2/EMBREE/OPTIX are bits 0..2).
 * With KernelData.bvh.bvh_layout == B200_BVH_LAYOUT_BVH8 the host's `BVH8 : BVH`
 * (csrc/bvh8_host.cpp, standing where bvh/bvh2.cpp stands) has packed
 *   __bvh_nodes       the 80-byte BVH8 nodes (five uint4 each, csrc/bvh8.h)
 *   __bvh_leaf_nodes  the 48-byte leaf records (three float4 each)
 *   KernelData.bvh.root = the BVH8 root, __object_node = BVH8 root of each object's BLAS
 * and the device traverses them as bound: no BVH2 reaches the device and nothing is built
 * at bind time.  With BVH_LAYOUT_BVH2 (an unpatched host) the device derives the same
 * BVH8 from the packed binary tree itself (b200_build_bvh). */
#define B200_BVH_LAYOUT_BVH2 (1u << 0)
#define B200_BVH_LAYOUT_BVH8 (1u << 3)

/* Host-only: the reference's packed binary BVH (PackedBVH, bvh/bvh.h:38-77, as
 * BVH2::pack_nodes leaves it) -> BVH8 nodes + leaf records.  No device, no context: this
 * is what `BVH8::pack_nodes` calls on the host.  `object_tfm` = 12 floats (3x4, row-major,
 * Object::tfm) per object.  The output arrays are owned by the library until
 * b200_bvh8_free.  Returns B200_OK, or B200_ERR_UNSUPPORTED with the reason in `err`
 * (curve / motion / unaligned nodes, a tree deeper than the traversal stack). */
typedef struct b200_packed_bvh2 {
  const void *nodes;           /* PackedBVH::nodes, int4 units */
  size_t num_nodes_f4;
  const void *leaf_nodes;      /* PackedBVH::leaf_nodes */
  size_t num_leaf_nodes_f4;
  const void *prim_tri_verts;  /* float4 */
  const void *prim_tri_index;  /* uint */
  const void *prim_visibility; /* uint */
  const void *prim_object;     /* int */
  size_t num_prims;
  const void *object_node;     /* int per object: root of its BLAS in `nodes` */
  const float *object_tfm;
  size_t num_objects;
  int root;                    /* PackedBVH::root_index */
} b200_packed_bvh2;

typedef struct b200_packed_bvh8 {
  void *nodes;          /* 80 bytes each */
  size_t node_bytes;
  void *records;        /* 48 bytes each */
  size_t record_bytes;
  int *object_node;     /* num_objects entries: BVH8 root of the object's BLAS, -1 if none */
  uint32_t root;
  b200_bvh_info info;
} b200_packed_bvh8;

int b200_bvh8_pack(const b200_packed_bvh2 *in, b200_packed_bvh8 *out, char *err, size_t errlen);
void b200_bvh8_free(b200_packed_bvh8 *out);

int b200_abi_version(void);

/* Device enumeration - device_cuda_info() / device_cuda_init()
 * (device/device_intern.h:34-35,47; device/device_cuda.cpp:100-190). */
int b200_device_count(void);
int b200_device_name(int ordinal, char *name, size_t len, int *sm_major, int *sm_minor,
                     uint64_t *total_mem, int *num_sms);
/* "dddd:bb:dd" PCI location of the device - the stable part of DeviceInfo::id, as
 * device_cuda_info() builds it (device/device_cuda.cpp:144-152). */
int b200_device_pci_id(int ordinal, char *buf, size_t len);

/* Context - CUDADevice ctor/dtor (device/cuda/device_cuda_impl.cpp:199-262).
 * Fails (returns NULL, message in err) unless the device is sm_100. */
b200_ctx *b200_create(int cuda_ordinal, char *err, size_t errlen);
void b200_destroy(b200_ctx *ctx);
const char *b200_last_error(b200_ctx *ctx);

/* Memory - Device::mem_alloc / mem_copy_to / mem_copy_from / mem_zero / mem_free
 * (device/device.h:484-488; CUDADevice::generic_alloc .. mem_free,
 * device_cuda_impl.cpp:806-1086). */
int b200_alloc(b200_ctx *ctx, size_t bytes, uint64_t *dptr);
int b200_free(b200_ctx *ctx, uint64_t dptr);
int b200_h2d(b200_ctx *ctx, uint64_t dptr, const void *host, size_t offset, size_t bytes);
int b200_d2h(b200_ctx *ctx, uint64_t dptr, void *host, size_t offset, size_t bytes);
int b200_zero(b200_ctx *ctx, uint64_t dptr, size_t offset, size_t bytes);
size_t b200_mem_used(b200_ctx *ctx);

/* MEM_GLOBAL binding by kernel_textures.h name ("__bvh_nodes", "__svm_nodes"..) -
 * the const_copy_to(mem.name, &device_pointer) self-call of CUDADevice::mem_copy_to
 * (device_cuda_impl.cpp:1088-1096) / kernel_global_memory_copy
 * (kernel/kernels/cpu/kernel.cpp:77-92).  `host` may be NULL; when given, the
 * host copy is kept for the BVH8 build (the BVH arrays) and validation
 * (__svm_nodes is scanned for opcodes outside the supported subset ->
 * B200_ERR_UNSUPPORTED). */
int b200_bind_global(b200_ctx *ctx, const char *name, uint64_t dptr, const void *host,
                     size_t bytes);

/* Image textures - CUDADevice::tex_alloc / tex_free (device_cuda_impl.cpp:1105-1304) for
 * the device_texture the ImageManager hands over per image slot (render/image.cpp:
 * 700-790).  The pixels are an ordinary allocation of this context (b200_alloc + b200_h2d);
 * `texture_info` is the reference's TextureInfo record for the slot (util/util_texture.h:
 * 93-107, SIZEOF_TEXTURE_INFO bytes) - its `data` member is replaced by `pixels`.  The
 * kernels sample with the CPU device's arithmetic (kernel_cpu_image.h), not a texture
 * unit.  2D images of every pixel format; 3D (volume) images are refused. */
int b200_texture_set(b200_ctx *ctx, int slot, const void *texture_info, size_t bytes,
                     uint64_t pixels);
int b200_texture_clear(b200_ctx *ctx, int slot);

/* DeviceTask::SHADER with SHADER_EVAL_BACKGROUND - CUDADevice::shader
 * (device_cuda_impl.cpp:2019-2093) for the one use the render path has of it: the light
 * manager evaluates the world shader over an equirectangular grid to build the importance
 * map of the background light (shade_background_pixels, render/light.cpp:38-102).
 * `input` = uint4 per point, (u, v) as float bits; `output` = float4 per point, the colour
 * is ADDED; points [shader_x, shader_x + shader_w).  Synchronous. */
int b200_shader_eval_background(b200_ctx *ctx, uint64_t input, uint64_t output, int shader_x,
                                int shader_w);

/* Host-only scope check of a compiled SVM program (the `__svm_nodes` array built by
 * SVMShaderManager::device_update_shader, render/svm.cpp:70-133): the same walk
 * b200_bind_global runs, without a device or a context.  Returns B200_OK, or
 * B200_ERR_UNSUPPORTED with the first opcode / closure / option outside the supported
 * subset described in `err`.  Lets an integrator decide up front whether a scene can
 * go to this device (the reference has no such query: its devices take every node). */
int b200_validate_svm(const void *svm_nodes, size_t bytes, char *err, size_t errlen);

/* Device::const_copy_to("__data", &KernelData, sizeof) (render/scene.cpp:307). */
int b200_set_kernel_data(b200_ctx *ctx, const void *kernel_data, size_t bytes);

/* Builds the compressed BVH8 from the bound BVH2 arrays and uploads it.  Called
 * implicitly by the first render/trace after a (re)bind; exposed so the build
 * time can be reported - stands where BVH::build + pack_nodes + copy_to_device
 * stand (bvh/bvh.cpp:128-178, render/geometry.cpp:1093-1098). */
int b200_build_bvh(b200_ctx *ctx, b200_bvh_info *info);

/* DeviceTask::RENDER for one tile - CUDADevice::render
 * (device_cuda_impl.cpp:1853-1952): runs the whole wavefront (init_from_camera,
 * intersect_closest, shade_*, intersect_shadow, film write) for samples
 * [start_sample, start_sample+num_samples) of the tile and accumulates into the
 * film at tile.buffer.  `cancel` (may be NULL) is polled between batches
 * (task.get_cancel(), device_cuda_impl.cpp:1908). */
int b200_render(b200_ctx *ctx, const b200_work_tile *tile, volatile const int *cancel);

/* task.get_cancel() (device/device_task.h:157-163; polled by CUDADevice::render every
 * sample step, device_cuda_impl.cpp:1939): a host predicate b200_render asks between
 * wavefront batches, next to the `cancel` flag.  NULL removes it. */
typedef int (*b200_cancel_fn)(void *user);
int b200_set_cancel_callback(b200_ctx *ctx, b200_cancel_fn fn, void *user);

/* Parity / benchmark hook: scene_intersect (kernel/bvh/bvh.h:154-237) on a batch
 * of rays resident in device memory.  any_hit != 0 gives the shadow-ray
 * early-out (bvh_traversal.h:144-147): only `prim >= 0` is meaningful then. */
int b200_trace_batch(b200_ctx *ctx, uint64_t rays, uint64_t hits, uint64_t n, int any_hit);

/* DeviceTask::FILM_CONVERT - kernel_film_convert_to_byte / _half_float
 * (kernel/kernel_film.h:90-130; CUDADevice::film_convert
 * device_cuda_impl.cpp:1954-2017). */
int b200_film_convert(b200_ctx *ctx, uint64_t film, uint64_t rgba, int half_float,
                      float sample_scale, int x, int y, int w, int h, int offset, int stride);

/* Multi-GPU film reduction inside ONE process (device/device_multi.cpp:374-393
 * replaced by a device-side sum): films[i] lives on ctxs[i]; after the call
 * films[0] holds the element-wise sum.  Peer copies over NVLink + one add
 * kernel per peer.  The one-process-per-GPU path uses NCCL all-reduce through
 * torch.distributed on the same device pointer instead (bench.py). */
int b200_film_reduce(b200_ctx **ctxs, int n, const uint64_t *films, size_t n_floats);

/* The same sum as ONE NCCL all-reduce over NVLink / NVSwitch for contexts on n DISTINCT
 * GPUs of this process (ncclCommInitAll once per set of GPUs, ncclAllReduce on each
 * context's stream inside a group): afterwards EVERY films[i] holds the sum.  This is
 * what the in-process multi device uses where the reference's MultiDevice copies tile
 * slices through the host (device/device_multi.cpp:374-393).  NCCL is resolved at first
 * use (libnccl.so.2); B200_ERR_UNSUPPORTED when it is missing or a GPU repeats. */
int b200_film_allreduce(b200_ctx **ctxs, int n, const uint64_t *films, size_t n_floats);

int b200_get_stats(b200_ctx *ctx, b200_stats *out);
int b200_synchronize(b200_ctx *ctx);

/* Debugging aid: with option "debug_slot" = s (>= 0) the path in pool slot s records
 * 32 floats per bounce in shade_surface; this reads up to 16 bounces back. */
int b200_debug_read(b200_ctx *ctx, float *out, size_t n_floats);

/* Run all work of this context on an existing CUDA stream (cudaStream_t handle,
 * e.g. torch.cuda.Stream.cuda_stream) so callers can bracket it with their own
 * events; 0 restores the context's private stream. */
int b200_set_stream(b200_ctx *ctx, uint64_t cuda_stream);

/* Tunables (0 keeps the default): "batch_paths" paths per wavefront batch,
 * "count_traversal" 1 = count BVH nodes / triangles per ray (slower),
 * "shade_wide" block shape of the lean multiscatter / the full shading kernel: -1 (default) =
 * time them on the first batches of a scene and keep the fastest, 0 = two blocks of 256
 * threads per SM, 1 = one block of 512, 2 = one block of 1024 (same arithmetic, same film
 * whichever runs). */
int b200_set_option(b200_ctx *ctx, const char *name, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* B200_CYCLES_H */

/* b200_cycles.cu - libb200cycles.so: the C ABI of include/b200_cycles.h and every
 * CUDA kernel behind it (one translation unit; compiled for sm_100a only with
 * -fmad=false, see Makefile).  Host code here is the thin runtime a Device
 * needs: context, memory, by-name array binding, the BVH8 build + upload and
 * the wavefront launch loop.  No CPU fallback: without a B200 b200_create fails.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h> /* types only: the library is resolved at first use (nccl_api) */

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>

#include "b200_internal.h"
#include "device_scene.cuh"
#include "traverse.cuh"

/* one __constant__ DeviceScene per GPU: which context's block it holds, and a lock per GPU
 * for the calls that launch kernels reading it (see DeviceUse) */
#define MAX_GPUS 16
static std::mutex g_owner_mutex;
static std::mutex g_device_mutex[MAX_GPUS];
static b200_ctx *g_constant_owner[MAX_GPUS] = {};


/* ------------------------------------------------------------------ utils */

#define CUDA_TRY(ctx, call) \
  do { \
    cudaError_t err__ = (call); \
    if (err__ != cudaSuccess) { \
      set_error(ctx, std::string(#call) + ": " + cudaGetErrorString(err__)); \
      return B200_ERR_CUDA; \
    } \
  } while (0)

static void set_error(b200_ctx *ctx, const std::string &msg)
{
  if (ctx)
    ctx->error = msg;
}

static int fail(b200_ctx *ctx, int code, const std::string &msg)
{
  set_error(ctx, msg);
  return code;
}

struct DeviceGuard {
  int prev = 0;
  explicit DeviceGuard(int ordinal)
  {
    cudaGetDevice(&prev);
    if (prev != ordinal)
      cudaSetDevice(ordinal);
  }
  ~DeviceGuard()
  {
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != prev)
      cudaSetDevice(prev);
  }
};

/* device counter slots (ctx->d_counters) */
enum {
  CNT_WORK = 0,     /* persistent-kernel work cursor */
  CNT_NODES = 1,    /* traversal statistics (counter build) */
  CNT_TRIS = 2,
  CNT_INSTANCES = 3,
  CNT_ACTIVE = 4,   /* wavefront queue sizes */
  CNT_NEXT = 5,
  CNT_SHADOW = 6,
  CNT_PRIMARY = 7,
  CNT_BOUNCE = 8,
  CNT_SHADOW_TOTAL = 9,
  CNT_WORK2 = 10,
  CNT_NODES_HI = 11,
  CNT_TRIS_HI = 12,
  CNT_NUM = 64
};

/* ------------------------------------------------------------ trace batch */

/* Job of the batch hook: 32-byte ray records in (two 16-byte halves), 24-byte hit
 * records out. */
struct BatchJob {
  static constexpr bool QUEUE_RAYS = false; /* the caller's 32-byte ray records, interleaved */
  const b200_ray *rays;
  b200_hit *hits;
  __device__ __forceinline__ const float4 *ray_P(unsigned int qi) const
  {
    return (const float4 *)(rays + qi);
  }
  __device__ __forceinline__ const float4 *ray_D(unsigned int qi) const
  {
    return (const float4 *)(rays + qi) + 1;
  }
  __device__ __forceinline__ void store(unsigned int qi, const TraceHit &h, bool found)
  {
    b200_hit out;
    out.t = h.t;
    out.u = h.u;
    out.v = h.v;
    out.prim = h.prim;
    out.object = h.object;
    out.type = found ? (int)CY_PRIMITIVE_TRIANGLE : 0;
    hits[qi] = out;
  }
};

/* Persistent warps with dynamic lane refill (trace_persistent, traverse.cuh): the
 * grid is sized to the machine (a multiple of the SM count), not to the batch. */
template<bool ANY_HIT, bool COUNT>
__global__ void __launch_bounds__(TRACE_BLOCK)
    k_trace_batch(const b200_ray *__restrict__ rays, b200_hit *__restrict__ hits, uint64_t n,
                  unsigned int *counters, int refill_threshold)
{
  const unsigned lane = threadIdx.x & 31u;
  TraceCounters cnt;
  cnt.nodes = cnt.tris = cnt.instances = 0;
  BatchJob job;
  job.rays = rays;
  job.hits = hits;
  trace_persistent<ANY_HIT, COUNT>(job, (unsigned int)n, &counters[CNT_WORK], refill_threshold,
                                   cnt);
  if (COUNT) {
    for (int o = 16; o > 0; o >>= 1) {
      cnt.nodes += __shfl_xor_sync(0xffffffffu, cnt.nodes, o);
      cnt.tris += __shfl_xor_sync(0xffffffffu, cnt.tris, o);
      cnt.instances += __shfl_xor_sync(0xffffffffu, cnt.instances, o);
    }
    if (lane == 0) {
      atomicAdd((unsigned long long *)&counters[16], (unsigned long long)cnt.nodes);
      atomicAdd((unsigned long long *)&counters[18], (unsigned long long)cnt.tris);
      atomicAdd((unsigned long long *)&counters[20], (unsigned long long)cnt.instances);
    }
  }
}

/* --------------------------------------------------------------- C ABI */

extern "C" {

int b200_abi_version(void)
{
  return B200_ABI_VERSION;
}

int b200_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess)
    return 0;
  return n;
}

int b200_device_name(int ordinal, char *name, size_t len, int *sm_major, int *sm_minor,
                     uint64_t *total_mem, int *num_sms)
{
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, ordinal) != cudaSuccess)
    return B200_ERR_CUDA;
  if (name && len) {
    strncpy(name, prop.name, len - 1);
    name[len - 1] = 0;
  }
  if (sm_major)
    *sm_major = prop.major;
  if (sm_minor)
    *sm_minor = prop.minor;
  if (total_mem)
    *total_mem = prop.totalGlobalMem;
  if (num_sms)
    *num_sms = prop.multiProcessorCount;
  return B200_OK;
}

int b200_device_pci_id(int ordinal, char *buf, size_t len)
{
  if (!buf || len < 13)
    return B200_ERR_INVALID;
  /* cudaDeviceGetPCIBusId: "dddd:bb:dd.f" */
  if (cudaDeviceGetPCIBusId(buf, (int)len, ordinal) != cudaSuccess)
    return B200_ERR_CUDA;
  char *dot = strchr(buf, '.');
  if (dot)
    *dot = 0;
  return B200_OK;
}

b200_ctx *b200_create(int cuda_ordinal, char *err, size_t errlen)
{
  auto report = [&](const std::string &m) {
    if (err && errlen) {
      strncpy(err, m.c_str(), errlen - 1);
      err[errlen - 1] = 0;
    }
  };
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    report(std::string("no CUDA device: ") + cudaGetErrorString(e));
    return nullptr;
  }
  if (cuda_ordinal < 0 || cuda_ordinal >= n) {
    report("CUDA ordinal out of range");
    return nullptr;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cuda_ordinal) != cudaSuccess) {
    report("cudaGetDeviceProperties failed");
    return nullptr;
  }
  if (prop.major != 10) {
    report(std::string("device ") + prop.name +
           " is not sm_100: this library carries sm_100a code only");
    return nullptr;
  }
  b200_ctx *ctx = new b200_ctx();
  ctx->ordinal = cuda_ordinal;
  ctx->num_sms = prop.multiProcessorCount;
  DeviceGuard guard(cuda_ordinal);
  bool ok = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess;
  ctx->stream = ctx->own_stream;
  ok = ok && cudaEventCreate(&ctx->ev0) == cudaSuccess && cudaEventCreate(&ctx->ev1) == cudaSuccess;
  ok = ok && cudaEventCreate(&ctx->ev2) == cudaSuccess && cudaEventCreate(&ctx->ev3) == cudaSuccess;
  ok = ok && cudaEventCreate(&ctx->ev4) == cudaSuccess && cudaEventCreate(&ctx->ev5) == cudaSuccess;
  ok = ok && cudaEventCreate(&ctx->ev6) == cudaSuccess;
  ok = ok && cudaMalloc(&ctx->d_counters, CNT_NUM * sizeof(unsigned int)) == cudaSuccess;
  ok = ok && cudaMallocHost(&ctx->h_counters, CNT_NUM * sizeof(unsigned int)) == cudaSuccess;
  if (!ok) {
    report(std::string("context setup failed: ") + cudaGetErrorString(cudaGetLastError()));
    delete ctx;
    return nullptr;
  }
  cudaMemset(ctx->d_counters, 0, CNT_NUM * sizeof(unsigned int));
  return ctx;
}

static void free_pool(b200_ctx *ctx);

void b200_destroy(b200_ctx *ctx)
{
  if (!ctx)
    return;
  DeviceGuard guard(ctx->ordinal);
  cudaStreamSynchronize(ctx->stream);
  {
    std::lock_guard<std::mutex> lock(g_owner_mutex);
    if (g_constant_owner[ctx->ordinal & (MAX_GPUS - 1)] == ctx)
      g_constant_owner[ctx->ordinal & (MAX_GPUS - 1)] = nullptr;
  }
  free_pool(ctx);
  for (auto &a : ctx->allocs)
    cudaFree((void *)a.first);
  if (ctx->d_nodes)
    cudaFree(ctx->d_nodes);
  if (ctx->d_records)
    cudaFree(ctx->d_records);
  if (ctx->d_counters)
    cudaFree(ctx->d_counters);
  if (ctx->d_debug)
    cudaFree(ctx->d_debug);
  if (ctx->reduce_tmp)
    cudaFree(ctx->reduce_tmp);
  if (ctx->d_texture_info)
    cudaFree(ctx->d_texture_info);
  if (ctx->h_counters)
    cudaFreeHost(ctx->h_counters);
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  cudaEventDestroy(ctx->ev2);
  cudaEventDestroy(ctx->ev3);
  cudaEventDestroy(ctx->ev4);
  cudaEventDestroy(ctx->ev5);
  cudaEventDestroy(ctx->ev6);
  cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

const char *b200_last_error(b200_ctx *ctx)
{
  return ctx ? ctx->error.c_str() : "null context";
}

int b200_alloc(b200_ctx *ctx, size_t bytes, uint64_t *dptr)
{
  if (!ctx || !dptr)
    return B200_ERR_INVALID;
  DeviceGuard guard(ctx->ordinal);
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
  if (e == cudaErrorMemoryAllocation) {
    cudaGetLastError();
    return fail(ctx, B200_ERR_OOM, "out of device memory");
  }
  if (e != cudaSuccess)
    return fail(ctx, B200_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  std::lock_guard<std::mutex> lock(ctx->mutex);
  ctx->allocs[(uint64_t)p] = bytes;
  ctx->mem_used += bytes;
  *dptr = (uint64_t)p;
  return B200_OK;
}

int b200_free(b200_ctx *ctx, uint64_t dptr)
{
  if (!ctx)
    return B200_ERR_INVALID;
  if (!dptr)
    return B200_OK;
  DeviceGuard guard(ctx->ordinal);
  {
    std::lock_guard<std::mutex> lock(ctx->mutex);
    auto it = ctx->allocs.find(dptr);
    if (it == ctx->allocs.end())
      return fail(ctx, B200_ERR_INVALID, "b200_free: unknown pointer");
    ctx->mem_used -= it->second;
    ctx->allocs.erase(it);
    /* an image slot whose pixels these were goes back to "no image" */
    for (size_t t = 0; t + SIZEOF_TEXTURE_INFO <= ctx->texture_info.size();
         t += SIZEOF_TEXTURE_INFO) {
      uint64_t data = 0;
      memcpy(&data, ctx->texture_info.data() + t + TI_DATA, 8);
      if (data == dptr) {
        memset(ctx->texture_info.data() + t, 0, SIZEOF_TEXTURE_INFO);
        ctx->scene_dirty = true;
      }
    }
    /* drop any by-name binding of this buffer */
    for (auto g = ctx->globals.begin(); g != ctx->globals.end();) {
      if (g->second.dptr == dptr) {
        g = ctx->globals.erase(g);
        ctx->scene_dirty = true;
      }
      else
        ++g;
    }
  }
  cudaStreamSynchronize(ctx->stream);
  CUDA_TRY(ctx, cudaFree((void *)dptr));
  return B200_OK;
}

int b200_h2d(b200_ctx *ctx, uint64_t dptr, const void *host, size_t offset, size_t bytes)
{
  if (!ctx || !dptr || (!host && bytes))
    return B200_ERR_INVALID;
  DeviceGuard guard(ctx->ordinal);
  CUDA_TRY(ctx, cudaMemcpyAsync((char *)dptr + offset, host, bytes, cudaMemcpyHostToDevice,
                                ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}

int b200_d2h(b200_ctx *ctx, uint64_t dptr, void *host, size_t offset, size_t bytes)
{
  if (!ctx || !dptr || (!host && bytes))
    return B200_ERR_INVALID;
  DeviceGuard guard(ctx->ordinal);
  CUDA_TRY(ctx, cudaMemcpyAsync(host, (const char *)dptr + offset, bytes, cudaMemcpyDeviceToHost,
                                ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}

int b200_zero(b200_ctx *ctx, uint64_t dptr, size_t offset, size_t bytes)
{
  if (!ctx || !dptr)
    return B200_ERR_INVALID;
  DeviceGuard guard(ctx->ordinal);
  CUDA_TRY(ctx, cudaMemsetAsync((char *)dptr + offset, 0, bytes, ctx->stream));
  return B200_OK;
}

size_t b200_mem_used(b200_ctx *ctx)
{
  return ctx ? ctx->mem_used + ctx->pool_bytes : 0;
}

int b200_synchronize(b200_ctx *ctx)
{
  if (!ctx)
    return B200_ERR_INVALID;
  DeviceGuard guard(ctx->ordinal);
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}

/* SVM opcodes / closures the shading kernels implement; anything else in a
 * bound __svm_nodes stream is refused up front instead of rendering wrongly. */
static bool svm_validate(const uint32_t *nodes, size_t n_nodes, std::string &why,
                         uint32_t *features, int *max_image_slot);

int b200_bind_global(b200_ctx *ctx, const char *name, uint64_t dptr, const void *host,
                     size_t bytes)
{
  if (!ctx || !name)
    return B200_ERR_INVALID;
  static const char *known[] = {
#define KNOWN(n) #n,
      KNOWN(__bvh_nodes) KNOWN(__bvh_leaf_nodes) KNOWN(__prim_tri_verts) KNOWN(__prim_tri_index)
          KNOWN(__prim_type) KNOWN(__prim_visibility) KNOWN(__prim_index) KNOWN(__prim_object)
              KNOWN(__object_node) KNOWN(__prim_time) KNOWN(__objects) KNOWN(__object_motion_pass)
                  KNOWN(__object_motion) KNOWN(__object_flag) KNOWN(__object_volume_step)
                      KNOWN(__camera_motion) KNOWN(__tri_shader) KNOWN(__tri_vnormal)
                          KNOWN(__tri_vindex) KNOWN(__tri_patch) KNOWN(__tri_patch_uv)
                              KNOWN(__curves) KNOWN(__curve_keys) KNOWN(__patches)
                                  KNOWN(__attributes_map) KNOWN(__attributes_float)
                                      KNOWN(__attributes_float2) KNOWN(__attributes_float3)
                                          KNOWN(__attributes_uchar4) KNOWN(__light_distribution)
                                              KNOWN(__lights)
                                                  KNOWN(__light_background_marginal_cdf)
                                                      KNOWN(__light_background_conditional_cdf)
                                                          KNOWN(__particles) KNOWN(__svm_nodes)
                                                              KNOWN(__shaders)
                                                                  KNOWN(__lookup_table)
                                                                      KNOWN(__sample_pattern_lut)
                                                                          KNOWN(__texture_info)
                                                                              KNOWN(__ies)
#undef KNOWN
  };
  bool found = false;
  for (const char *k : known)
    if (strcmp(k, name) == 0)
      found = true;
  if (!found)
    return fail(ctx, B200_ERR_INVALID, std::string("unknown kernel array ") + name);

  static const char *keep_host[] = {"__bvh_nodes",     "__bvh_leaf_nodes", "__prim_tri_verts",
                                    "__prim_tri_index", "__prim_type",      "__prim_visibility",
                                    "__prim_object",    "__object_node",    "__objects",
                                    "__svm_nodes",      "__lights",         "__curves",
                                    "__tri_patch",      "__attributes_map"};
  static const char *bvh_inputs[] = {"__bvh_nodes",       "__bvh_leaf_nodes",  "__prim_tri_verts",
                                     "__prim_tri_index",  "__prim_visibility", "__prim_object",
                                     "__object_node",     "__objects"};
  for (const char *k : bvh_inputs)
    if (strcmp(k, name) == 0)
      ctx->bvh_dirty = true;
  HostArray ha;
  ha.dptr = dptr;
  ha.bytes = bytes;
  bool keep = false;
  for (const char *k : keep_host)
    if (strcmp(k, name) == 0)
      keep = true;
  if (keep && bytes) {
    ha.host.resize(bytes);
    if (host) {
      memcpy(ha.host.data(), host, bytes);
    }
    else {
      int rc = b200_d2h(ctx, dptr, ha.host.data(), 0, bytes);
      if (rc)
        return rc;
    }
  }
  if (strcmp(name, "__svm_nodes") == 0) {
    ctx->force_svm_ext = false;
    /* new shader mix: probe the block shapes of the shading kernels again */
    ctx->shade_probe[0] = ctx->shade_probe[1] = b200_ctx::ShadeProbe();
    if (!bytes)
      ctx->svm_features = 0;
  }
  if (strcmp(name, "__svm_nodes") == 0 && bytes) {
    std::string why;
    ctx->svm_max_image_slot = -1;
    if (!svm_validate((const uint32_t *)ha.host.data(), bytes / 16, why, &ctx->svm_features,
                      &ctx->svm_max_image_slot))
      return fail(ctx, B200_ERR_UNSUPPORTED, why);
  }
  if (strcmp(name, "__objects") == 0) {
    ctx->has_terminator_offset = false;
    for (size_t o = 0; o < bytes / SIZEOF_KERNEL_OBJECT; o++) {
      float f;
      memcpy(&f, ha.host.data() + o * SIZEOF_KERNEL_OBJECT + KO_SHADOW_TERMINATOR_OFFSET, 4);
      if (f > 1.0f)
        ctx->has_terminator_offset = true;
    }
  }
  if (strcmp(name, "__object_flag") == 0 && bytes) {
    /* per-object holdout masks and shadow catchers change alpha and colour of every
     * path that meets the object (kernel_path.h:254-321): not implemented, so refused
     * rather than rendered as an ordinary surface */
    std::vector<uint32_t> flags(bytes / 4);
    if (host)
      memcpy(flags.data(), host, flags.size() * 4);
    else if (int rc = b200_d2h(ctx, dptr, flags.data(), 0, flags.size() * 4))
      return rc;
    for (uint32_t f : flags) {
      if (f & CY_SD_OBJECT_HOLDOUT_MASK)
        return fail(ctx, B200_ERR_UNSUPPORTED,
                    "objects used as holdout masks are outside the hot-path scope");
      if (f & CY_SD_OBJECT_SHADOW_CATCHER)
        return fail(ctx, B200_ERR_UNSUPPORTED,
                    "shadow catcher objects are outside the hot-path scope");
    }
  }
  if (strcmp(name, "__attributes_map") == 0) {
    ctx->has_generated_attr = false;
    const uint32_t *map = (const uint32_t *)ha.host.data(); /* uint4 entries, x = id */
    for (size_t i = 0; i < bytes / 16; i++)
      if (map[4 * i] == CY_ATTR_STD_GENERATED && map[4 * i + 1] != CY_ATTR_ELEMENT_NONE)
        ctx->has_generated_attr = true;
    ha.host = std::vector<uint8_t>();
  }
  if (strcmp(name, "__tri_patch") == 0) {
    ctx->has_subd_patches = false;
    const uint32_t *patch = (const uint32_t *)ha.host.data();
    for (size_t i = 0; i < bytes / 4; i++)
      if (patch[i] != ~0u)
        ctx->has_subd_patches = true;
    ha.host = std::vector<uint8_t>(); /* only the flag is kept */
  }
  if ((strcmp(name, "__curves") == 0 || strcmp(name, "__curve_keys") == 0) && bytes)
    return fail(ctx, B200_ERR_UNSUPPORTED, "hair curves are outside the hot-path scope");
  std::lock_guard<std::mutex> lock(ctx->mutex);
  ctx->globals[name] = std::move(ha);
  ctx->scene_dirty = true;
  return B200_OK;
}

int b200_texture_set(b200_ctx *ctx, int slot, const void *texture_info, size_t bytes,
                     uint64_t pixels)
{
  if (!ctx || slot < 0 || !texture_info)
    return B200_ERR_INVALID;
  if (bytes != SIZEOF_TEXTURE_INFO)
    return fail(ctx, B200_ERR_INVALID, "TextureInfo size mismatch (ABI)");
  uint32_t depth = 0, type = 0;
  memcpy(&depth, (const uint8_t *)texture_info + TI_DEPTH, 4);
  memcpy(&type, (const uint8_t *)texture_info + TI_DATA_TYPE, 4);
  if (depth > 1)
    return fail(ctx, B200_ERR_UNSUPPORTED,
                "3D (volume) image textures are outside the hot-path scope");
  if (type > CY_IMAGE_DATA_TYPE_USHORT)
    return fail(ctx, B200_ERR_UNSUPPORTED, "unknown image data type");
  std::lock_guard<std::mutex> lock(ctx->mutex);
  const size_t need = ((size_t)slot + 1) * SIZEOF_TEXTURE_INFO;
  if (ctx->texture_info.size() < need)
    ctx->texture_info.resize(need, 0); /* unset slots: data == NULL reads as black */
  uint8_t *rec = ctx->texture_info.data() + (size_t)slot * SIZEOF_TEXTURE_INFO;
  memcpy(rec, texture_info, SIZEOF_TEXTURE_INFO);
  memcpy(rec + TI_DATA, &pixels, 8);
  ctx->scene_dirty = true;
  return B200_OK;
}

int b200_texture_clear(b200_ctx *ctx, int slot)
{
  if (!ctx || slot < 0)
    return B200_ERR_INVALID;
  std::lock_guard<std::mutex> lock(ctx->mutex);
  if (((size_t)slot + 1) * SIZEOF_TEXTURE_INFO <= ctx->texture_info.size()) {
    memset(ctx->texture_info.data() + (size_t)slot * SIZEOF_TEXTURE_INFO, 0, SIZEOF_TEXTURE_INFO);
    ctx->scene_dirty = true;
  }
  return B200_OK;
}

int b200_validate_svm(const void *svm_nodes, size_t bytes, char *err, size_t errlen)
{
  std::string why;
  uint32_t features = 0;
  if (!svm_nodes || bytes % 16 != 0)
    why = "an SVM program is an array of 16-byte nodes";
  else if (svm_validate((const uint32_t *)svm_nodes, bytes / 16, why, &features, nullptr))
    return B200_OK;
  if (err && errlen) {
    strncpy(err, why.c_str(), errlen - 1);
    err[errlen - 1] = 0;
  }
  return (svm_nodes && bytes % 16 == 0) ? B200_ERR_UNSUPPORTED : B200_ERR_INVALID;
}

int b200_bvh8_pack(const b200_packed_bvh2 *in, b200_packed_bvh8 *out, char *err, size_t errlen)
{
  if (!in || !out)
    return B200_ERR_INVALID;
  memset(out, 0, sizeof(*out));
  b200::BVH2Input bi;
  memset(&bi, 0, sizeof(bi));
  bi.nodes = (const float *)in->nodes;
  bi.num_nodes_f4 = in->num_nodes_f4;
  bi.leaf_nodes = (const float *)in->leaf_nodes;
  bi.num_leaf_nodes_f4 = in->num_leaf_nodes_f4;
  bi.prim_tri_verts = (const float *)in->prim_tri_verts;
  bi.prim_tri_index = (const uint32_t *)in->prim_tri_index;
  bi.prim_visibility = (const uint32_t *)in->prim_visibility;
  bi.prim_object = (const uint32_t *)in->prim_object;
  bi.num_prims = in->num_prims;
  bi.object_node = (const int32_t *)in->object_node;
  bi.objects = (const uint8_t *)in->object_tfm;
  bi.object_stride = 12 * sizeof(float);
  bi.object_tfm_offset = 0;
  bi.num_objects = in->num_objects;
  bi.root = in->root;
  bi.node_unaligned_flag = CY_PATH_RAY_NODE_UNALIGNED;
  bi.primitive_all = CY_PRIMITIVE_ALL;
  bi.primitive_triangle = CY_PRIMITIVE_TRIANGLE;
  bi.tighten_instances = true;
  b200::BVH8Output bo;
  std::string why;
  if (!bi.leaf_nodes || !b200::build_bvh8(bi, bo, why)) {
    if (why.empty())
      why = "no packed BVH2 to convert";
    if (err && errlen) {
      strncpy(err, why.c_str(), errlen - 1);
      err[errlen - 1] = 0;
    }
    return B200_ERR_UNSUPPORTED;
  }
  out->node_bytes = bo.nodes.size() * sizeof(BVH8Node);
  out->record_bytes = bo.records.size() * sizeof(float);
  out->nodes = malloc(std::max<size_t>(out->node_bytes, 16));
  out->records = malloc(std::max<size_t>(out->record_bytes, 16));
  out->object_node = (int *)malloc(std::max<size_t>(bo.object_root8.size() * sizeof(int), 16));
  if (!out->nodes || !out->records || !out->object_node) {
    b200_bvh8_free(out);
    return B200_ERR_OOM;
  }
  memcpy(out->nodes, bo.nodes.data(), out->node_bytes);
  memcpy(out->records, bo.records.data(), out->record_bytes);
  memcpy(out->object_node, bo.object_root8.data(), bo.object_root8.size() * sizeof(int));
  out->root = bo.root;
  out->info.num_nodes = bo.nodes.size();
  out->info.num_tri_records = bo.records.size() / 12;
  out->info.num_triangles = bo.num_triangles;
  out->info.num_instances = bo.num_instances;
  out->info.node_bytes = out->node_bytes;
  out->info.tri_bytes = out->record_bytes;
  out->info.build_ms = bo.build_ms;
  out->info.sah_cost = bo.sah_cost;
  out->info.max_depth = bo.max_depth;
  out->info.host_packed = 1;
  return B200_OK;
}

void b200_bvh8_free(b200_packed_bvh8 *out)
{
  if (!out)
    return;
  free(out->nodes);
  free(out->records);
  free(out->object_node);
  out->nodes = out->records = nullptr;
  out->object_node = nullptr;
}

int b200_set_kernel_data(b200_ctx *ctx, const void *kernel_data, size_t bytes)
{
  if (!ctx || !kernel_data)
    return B200_ERR_INVALID;
  if (bytes != SIZEOF_KERNEL_DATA)
    return fail(ctx, B200_ERR_INVALID, "KernelData size mismatch (ABI)");
  ctx->kernel_data.assign((const uint8_t *)kernel_data, (const uint8_t *)kernel_data + bytes);
  ctx->have_data = true;
  ctx->scene_dirty = true;
  return B200_OK;
}

} /* extern "C" */

/* -------------------------------------------------- scene (re)build + upload */

template<typename T> static T kd_host(const b200_ctx *ctx, int off)
{
  T v;
  memcpy(&v, ctx->kernel_data.data() + off, sizeof(T));
  return v;
}

static const HostArray *find_global(b200_ctx *ctx, const char *name)
{
  auto it = ctx->globals.find(name);
  return (it == ctx->globals.end()) ? nullptr : &it->second;
}

static int check_scope(b200_ctx *ctx);

/* for_shader_task: DeviceTask::SHADER runs in the middle of Scene::device_update (the light
 * manager evaluates the world shader before the film / integrator are final), so the
 * scope check waits for the first render */
static int prepare_scene(b200_ctx *ctx, bool for_shader_task = false)
{
  if (!ctx->scene_dirty)
    return B200_OK;
  if (!ctx->have_data)
    return fail(ctx, B200_ERR_NOT_READY, "KernelData (\"__data\") was never uploaded");
  const HostArray *nodes = find_global(ctx, "__bvh_nodes");
  const HostArray *leaves = find_global(ctx, "__bvh_leaf_nodes");
  const HostArray *verts = find_global(ctx, "__prim_tri_verts");
  const HostArray *tri_index = find_global(ctx, "__prim_tri_index");
  const HostArray *vis = find_global(ctx, "__prim_visibility");
  const HostArray *pobj = find_global(ctx, "__prim_object");
  const HostArray *onode = find_global(ctx, "__object_node");
  const HostArray *objects = find_global(ctx, "__objects");
  if (!leaves || !verts || !tri_index || !vis || !pobj || !objects)
    return fail(ctx, B200_ERR_NOT_READY, "BVH arrays are not bound (scene without geometry?)");
  int rc = for_shader_task ? B200_OK : check_scope(ctx);
  if (rc)
    return rc;

  DeviceGuard guard(ctx->ordinal);
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  const uint4 *dev_nodes = nullptr;
  const float4 *dev_records = nullptr;
  uint32_t bvh_root = 0;
  const uint32_t layout = (uint32_t)kd_host<int>(ctx, KD_BVH_LAYOUT);
  if (layout != B200_BVH_LAYOUT_BVH8 && !ctx->bvh_dirty && ctx->d_nodes && ctx->d_records) {
    /* only KernelData or non-BVH arrays changed: the BVH8 derived last time still holds */
    dev_nodes = (const uint4 *)ctx->d_nodes;
    dev_records = (const float4 *)ctx->d_records;
    bvh_root = ctx->bvh_root8;
  }
  else if (layout == B200_BVH_LAYOUT_BVH8) {
    /* the host's BVH8 class packed the device layout: traverse the arrays as bound */
    if (!nodes || nodes->bytes % sizeof(BVH8Node) != 0 || leaves->bytes % 48 != 0)
      return fail(ctx, B200_ERR_INVALID, "BVH_LAYOUT_BVH8: __bvh_nodes / __bvh_leaf_nodes are "
                                         "not whole BVH8 nodes / leaf records");
    dev_nodes = (const uint4 *)nodes->dptr;
    dev_records = (const float4 *)leaves->dptr;
    bvh_root = (uint32_t)kd_host<int>(ctx, KD_BVH_ROOT);
    if ((size_t)bvh_root >= nodes->bytes / sizeof(BVH8Node))
      return fail(ctx, B200_ERR_INVALID, "BVH_LAYOUT_BVH8: root outside the node array");
    memset(&ctx->bvh_info, 0, sizeof(ctx->bvh_info));
    ctx->bvh_info.num_nodes = nodes->bytes / sizeof(BVH8Node);
    ctx->bvh_info.num_tri_records = leaves->bytes / 48;
    ctx->bvh_info.node_bytes = nodes->bytes;
    ctx->bvh_info.tri_bytes = leaves->bytes;
    ctx->bvh_info.host_packed = 1;
  }
  else {
    if (ctx->d_nodes)
      cudaFree(ctx->d_nodes);
    if (ctx->d_records)
      cudaFree(ctx->d_records);
    ctx->d_nodes = ctx->d_records = nullptr;
    b200::BVH2Input in;
    memset(&in, 0, sizeof(in));
    in.nodes = nodes ? (const float *)nodes->host.data() : nullptr;
    in.num_nodes_f4 = nodes ? nodes->bytes / 16 : 0;
    in.leaf_nodes = (const float *)leaves->host.data();
    in.num_leaf_nodes_f4 = leaves->bytes / 16;
    in.prim_tri_verts = (const float *)verts->host.data();
    in.prim_tri_index = (const uint32_t *)tri_index->host.data();
    in.prim_visibility = (const uint32_t *)vis->host.data();
    in.prim_object = (const uint32_t *)pobj->host.data();
    in.num_prims = tri_index->bytes / 4;
    in.object_node = onode ? (const int32_t *)onode->host.data() : nullptr;
    in.objects = objects->host.data();
    in.object_stride = SIZEOF_KERNEL_OBJECT;
    in.object_tfm_offset = KO_TFM;
    in.num_objects = objects->bytes / SIZEOF_KERNEL_OBJECT;
    in.root = kd_host<int>(ctx, KD_BVH_ROOT);
    in.node_unaligned_flag = CY_PATH_RAY_NODE_UNALIGNED;
    in.primitive_all = CY_PRIMITIVE_ALL;
    in.primitive_triangle = CY_PRIMITIVE_TRIANGLE;
    in.tighten_instances = ctx->opt_loose_instances == 0;
    in.instance_detail_boxes = (int)ctx->opt_instance_detail_boxes;

    b200::BVH8Output out;
    std::string err;
    if (!b200::build_bvh8(in, out, err))
      return fail(ctx, B200_ERR_UNSUPPORTED, "BVH8 build: " + err);

    const size_t node_bytes = out.nodes.size() * sizeof(BVH8Node);
    const size_t rec_bytes = out.records.size() * sizeof(float);
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_nodes, std::max<size_t>(node_bytes, 16)));
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_records, std::max<size_t>(rec_bytes, 16)));
    CUDA_TRY(ctx, cudaMemcpy(ctx->d_nodes, out.nodes.data(), node_bytes, cudaMemcpyHostToDevice));
    CUDA_TRY(ctx,
             cudaMemcpy(ctx->d_records, out.records.data(), rec_bytes, cudaMemcpyHostToDevice));
    dev_nodes = (const uint4 *)ctx->d_nodes;
    dev_records = (const float4 *)ctx->d_records;
    bvh_root = out.root;
    ctx->bvh_root8 = out.root;
    ctx->bvh_dirty = false;

    memset(&ctx->bvh_info, 0, sizeof(ctx->bvh_info));
    ctx->bvh_info.num_nodes = out.nodes.size();
    ctx->bvh_info.num_tri_records = out.records.size() / 12;
    ctx->bvh_info.num_triangles = out.num_triangles;
    ctx->bvh_info.num_instances = out.num_instances;
    ctx->bvh_info.node_bytes = node_bytes;
    ctx->bvh_info.tri_bytes = rec_bytes;
    ctx->bvh_info.build_ms = out.build_ms;
    ctx->bvh_info.sah_cost = out.sah_cost;
    ctx->bvh_info.max_depth = out.max_depth;
  }

  DeviceScene ds;
  memset(&ds, 0, sizeof(ds));
  ds.nodes = dev_nodes;
  ds.records = dev_records;
  ds.bvh_root = bvh_root;
  ctx->dev_nodes = (const void *)dev_nodes;
  ctx->dev_nodes_bytes = (size_t)ctx->bvh_info.node_bytes;
  auto ptr = [&](const char *name) -> uint64_t {
    const HostArray *h = find_global(ctx, name);
    return h ? h->dptr : 0;
  };
  ds.prim_tri_verts = (const float4 *)ptr("__prim_tri_verts");
  ds.prim_tri_index = (const uint32_t *)ptr("__prim_tri_index");
  ds.prim_index = (const uint32_t *)ptr("__prim_index");
  ds.prim_object = (const uint32_t *)ptr("__prim_object");
  ds.objects = (const uint8_t *)ptr("__objects");
  ds.object_flag = (const uint32_t *)ptr("__object_flag");
  ds.tri_shader = (const uint32_t *)ptr("__tri_shader");
  ds.tri_vnormal = (const float4 *)ptr("__tri_vnormal");
  ds.tri_vindex = (const uint4 *)ptr("__tri_vindex");
  ds.lights = (const uint8_t *)ptr("__lights");
  ds.light_distribution = (const uint8_t *)ptr("__light_distribution");
  ds.shaders = (const uint8_t *)ptr("__shaders");
  ds.svm_nodes = (const uint4 *)ptr("__svm_nodes");
  ds.lookup_table = (const float *)ptr("__lookup_table");
  ds.sample_pattern_lut = (const uint32_t *)ptr("__sample_pattern_lut");
  ds.attributes_map = (const uint4 *)ptr("__attributes_map");
  ds.attributes_float = (const float *)ptr("__attributes_float");
  ds.attributes_float2 = (const float2 *)ptr("__attributes_float2");
  ds.attributes_float3 = (const float4 *)ptr("__attributes_float3");
  ds.attributes_uchar4 = (const uchar4 *)ptr("__attributes_uchar4");
  ds.light_background_marginal_cdf = (const float2 *)ptr("__light_background_marginal_cdf");
  ds.light_background_conditional_cdf = (const float2 *)ptr("__light_background_conditional_cdf");
  if (ctx->d_texture_info) {
    cudaFree(ctx->d_texture_info);
    ctx->d_texture_info = nullptr;
  }
  if (!ctx->texture_info.empty()) {
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_texture_info, ctx->texture_info.size()));
    CUDA_TRY(ctx, cudaMemcpy(ctx->d_texture_info, ctx->texture_info.data(),
                             ctx->texture_info.size(), cudaMemcpyHostToDevice));
    ds.texture_info = (const uint8_t *)ctx->d_texture_info;
    ds.num_textures = (uint32_t)(ctx->texture_info.size() / SIZEOF_TEXTURE_INFO);
  }
  memcpy(ds.kdata, ctx->kernel_data.data(), SIZEOF_KERNEL_DATA);
  ctx->constant_block.assign((const uint8_t *)&ds, (const uint8_t *)&ds + sizeof(ds));
  {
    std::lock_guard<std::mutex> lock(g_owner_mutex);
    CUDA_TRY(ctx, cudaMemcpyToSymbol(g_scene, &ds, sizeof(ds)));
    g_constant_owner[ctx->ordinal & (MAX_GPUS - 1)] = ctx;
  }
  ctx->scene_dirty = false;
  return B200_OK;
}

/* Held for the length of a call that launches kernels reading g_scene: serialises the
 * contexts of ONE GPU (they share its constant block and the scope-miss flag) and makes
 * sure the block is this context's.  Contexts of different GPUs never meet here. */
struct DeviceUse {
  std::unique_lock<std::mutex> lock;
  explicit DeviceUse(b200_ctx *ctx) : lock(g_device_mutex[ctx->ordinal & (MAX_GPUS - 1)])
  {
    std::lock_guard<std::mutex> owner_lock(g_owner_mutex);
    b200_ctx *&owner = g_constant_owner[ctx->ordinal & (MAX_GPUS - 1)];
    if (owner != ctx && !ctx->constant_block.empty()) {
      DeviceGuard guard(ctx->ordinal);
      cudaDeviceSynchronize();
      cudaMemcpyToSymbol(g_scene, ctx->constant_block.data(), ctx->constant_block.size());
      owner = ctx;
    }
  }
};

static int launch_grid(const b200_ctx *ctx, int blocks_per_sm)
{
  return ctx->num_sms * blocks_per_sm;
}

/* lanes still busy below which a warp goes back to its queue for more rays */
static int refill_threshold(const b200_ctx *ctx)
{
  return ctx->opt_refill_threshold > 0 ? (int)ctx->opt_refill_threshold : 24;
}

/* traverse.cuh: g_trace_overflow.  Checked once per wavefront batch / trace call. */
static int check_trace_overflow(b200_ctx *ctx)
{
  unsigned int flag = 0;
  CUDA_TRY(ctx, cudaMemcpyFromSymbol(&flag, g_trace_overflow, sizeof(flag)));
  if (!flag)
    return B200_OK;
  const unsigned int zero = 0;
  CUDA_TRY(ctx, cudaMemcpyToSymbol(g_trace_overflow, &zero, sizeof(zero)));
  return fail(ctx, B200_ERR_UNSUPPORTED,
              "BVH8 traversal stack overflow: hits of this call are not reliable");
}

/* A/B (b200_set_option("l2_persist_nodes", 1)): an L2 access-policy window over the BVH8
 * node array on the context's stream - hits stay, misses stream.  Measured on B200: see
 * profiles/r02m_l2_window_ab.txt. */
static int apply_l2_window(b200_ctx *ctx)
{
  cudaStreamAttrValue attr;
  memset(&attr, 0, sizeof(attr));
  if (ctx->opt_l2_persist_nodes && ctx->dev_nodes && ctx->dev_nodes_bytes) {
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->ordinal);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, ctx->ordinal);
    const size_t bytes = std::min<size_t>(ctx->dev_nodes_bytes,
                                          std::min<size_t>((size_t)max_persist, (size_t)max_window));
    CUDA_TRY(ctx, cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, bytes));
    attr.accessPolicyWindow.base_ptr = const_cast<void *>(ctx->dev_nodes);
    attr.accessPolicyWindow.num_bytes = bytes;
    attr.accessPolicyWindow.hitRatio = 1.0f;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  }
  else if (!ctx->l2_window_set) {
    return B200_OK;
  }
  CUDA_TRY(ctx, cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
  ctx->l2_window_set = ctx->opt_l2_persist_nodes != 0;
  return B200_OK;
}

#include "wavefront.cuh"

extern "C" {

int b200_build_bvh(b200_ctx *ctx, b200_bvh_info *info)
{
  if (!ctx)
    return B200_ERR_INVALID;
  DeviceUse device_use(ctx);
  ctx->scene_dirty = true;
  int rc = prepare_scene(ctx);
  if (rc)
    return rc;
  if (info)
    *info = ctx->bvh_info;
  return B200_OK;
}

int b200_trace_batch(b200_ctx *ctx, uint64_t rays, uint64_t hits, uint64_t n, int any_hit)
{
  if (!ctx || !rays || !hits)
    return B200_ERR_INVALID;
  if (n >= 0xffffffe0ull)
    return fail(ctx, B200_ERR_INVALID, "batch too large (max 2^32-32 rays)");
  DeviceUse device_use(ctx);
  int rc = prepare_scene(ctx);
  if (rc)
    return rc;
  DeviceGuard guard(ctx->ordinal);
  CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_counters, 0, CNT_NUM * sizeof(unsigned int), ctx->stream));
  const int grid = launch_grid(ctx, ctx->opt_trace_blocks_per_sm > 0 ? (int)ctx->opt_trace_blocks_per_sm : 8);
  const bool count = ctx->opt_count_traversal != 0;
  const int refill = refill_threshold(ctx);
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  const b200_ray *r = (const b200_ray *)rays;
  b200_hit *h = (b200_hit *)hits;
  if (any_hit) {
    if (count)
      k_trace_batch<true, true><<<grid, TRACE_BLOCK, 0, ctx->stream>>>(r, h, n, ctx->d_counters, refill);
    else
      k_trace_batch<true, false><<<grid, TRACE_BLOCK, 0, ctx->stream>>>(r, h, n, ctx->d_counters, refill);
  }
  else {
    if (count)
      k_trace_batch<false, true><<<grid, TRACE_BLOCK, 0, ctx->stream>>>(r, h, n, ctx->d_counters, refill);
    else
      k_trace_batch<false, false><<<grid, TRACE_BLOCK, 0, ctx->stream>>>(r, h, n, ctx->d_counters, refill);
  }
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, CNT_NUM * sizeof(unsigned int),
                                cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  rc = check_trace_overflow(ctx);
  if (rc)
    return rc;
  float ms = 0.0f;
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  memset(&ctx->stats, 0, sizeof(ctx->stats));
  uint64_t cn = 0, ct = 0, ci = 0;
  memcpy(&cn, &ctx->h_counters[16], 8);
  memcpy(&ct, &ctx->h_counters[18], 8);
  memcpy(&ci, &ctx->h_counters[20], 8);
  if (any_hit) {
    ctx->stats.shadow_rays = n;
    ctx->stats.shadow_nodes = cn;
    ctx->stats.shadow_tris = ct;
    ctx->stats.shadow_instances = ci;
    ctx->stats.shadow_launches = 1;
    ctx->stats.shadow_ms = ms;
  }
  else {
    ctx->stats.primary_rays = n;
    ctx->stats.closest_nodes = cn;
    ctx->stats.closest_tris = ct;
    ctx->stats.closest_instances = ci;
    ctx->stats.closest_launches = 1;
    ctx->stats.closest_ms = ms;
  }
  ctx->stats.kernel_launches = 1;
  ctx->stats.device_ms = ms;
  return B200_OK;
}

int b200_get_stats(b200_ctx *ctx, b200_stats *out)
{
  if (!ctx || !out)
    return B200_ERR_INVALID;
  *out = ctx->stats;
  return B200_OK;
}

int b200_debug_read(b200_ctx *ctx, float *out, size_t n_floats)
{
  if (!ctx || !out || !ctx->d_debug || n_floats > 16 * 32)
    return B200_ERR_INVALID;
  DeviceGuard guard(ctx->ordinal);
  CUDA_TRY(ctx, cudaMemcpy(out, ctx->d_debug, n_floats * sizeof(float), cudaMemcpyDeviceToHost));
  return B200_OK;
}

int b200_set_cancel_callback(b200_ctx *ctx, b200_cancel_fn fn, void *user)
{
  if (!ctx)
    return B200_ERR_INVALID;
  ctx->cancel_fn = fn;
  ctx->cancel_user = user;
  return B200_OK;
}

int b200_set_stream(b200_ctx *ctx, uint64_t cuda_stream)
{
  if (!ctx)
    return B200_ERR_INVALID;
  DeviceGuard guard(ctx->ordinal);
  cudaStreamSynchronize(ctx->stream);
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return B200_OK;
}

int b200_set_option(b200_ctx *ctx, const char *name, int64_t value)
{
  if (!ctx || !name)
    return B200_ERR_INVALID;
  if (strcmp(name, "batch_paths") == 0) {
    ctx->opt_batch_paths = value;
    free_pool(ctx);
  }
  else if (strcmp(name, "count_traversal") == 0)
    ctx->opt_count_traversal = value;
  else if (strcmp(name, "debug_slot") == 0) {
    /* debugging aid: record the shading of one path slot, read back with b200_d2h from
     * the pointer returned through option "debug_ptr" semantics (see tools/) */
    DeviceGuard guard(ctx->ordinal);
    ctx->opt_debug_slot = value;
    if (value >= 0 && !ctx->d_debug) {
      if (cudaMalloc(&ctx->d_debug, 16 * 32 * sizeof(float)) != cudaSuccess)
        return fail(ctx, B200_ERR_CUDA, "debug buffer allocation failed");
    }
    if (ctx->d_debug)
      cudaMemset(ctx->d_debug, 0, 16 * 32 * sizeof(float));
  }
  else if (strcmp(name, "refill_threshold") == 0)
    ctx->opt_refill_threshold = value;
  else if (strcmp(name, "trace_blocks_per_sm") == 0)
    ctx->opt_trace_blocks_per_sm = value;
  else if (strcmp(name, "sync_iterations") == 0)
    ctx->opt_sync_iterations = value;
  else if (strcmp(name, "sort_tiles") == 0)
    ctx->opt_sort_tiles = value;
  else if (strcmp(name, "l2_persist_nodes") == 0)
    ctx->opt_l2_persist_nodes = value;
  else if (strcmp(name, "shade_wide") == 0) {
    ctx->opt_shade_wide = value;
    ctx->shade_probe[0] = ctx->shade_probe[1] = b200_ctx::ShadeProbe();
  }
  else if (strcmp(name, "shade_carveout") == 0) {
    ctx->opt_shade_carveout = value;
    ctx->shade_blocks_per_sm[0] = 0; /* set the kernels up again */
  }
  else if (strcmp(name, "instance_detail_boxes") == 0) {
    ctx->opt_instance_detail_boxes = value;
    ctx->bvh_dirty = true;
    ctx->scene_dirty = true;
  }
  else if (strcmp(name, "loose_instances") == 0) {
    /* A/B: keep the host's instance bounds in the TLAS (BVH2 hosts only) */
    ctx->opt_loose_instances = value;
    ctx->bvh_dirty = true;
    ctx->scene_dirty = true;
  }
  else
    return fail(ctx, B200_ERR_INVALID, std::string("unknown option ") + name);
  return B200_OK;
}

} /* extern "C" */

/* ------------------------------------------------ NCCL film all-reduce */

/* NCCL is bound at first use instead of at load time: a host that never renders on
 * several GPUs in one process (one GPU, or one process per GPU with torch.distributed)
 * does not need the library, and a process that already carries an NCCL (PyTorch's) gets
 * that one. */
namespace {
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
  bool ok = false;
};

struct NcclClique {
  std::vector<int> ordinals;
  std::vector<ncclComm_t> comms;
};

std::mutex g_nccl_mutex;
NcclApi g_nccl;
std::vector<NcclClique> g_nccl_cliques; /* one communicator set per distinct GPU list */

bool nccl_api()
{
  if (g_nccl.ok || !g_nccl.error.empty())
    return g_nccl.ok;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle)
      break;
  }
  if (!g_nccl.handle) {
    g_nccl.error = std::string("libnccl.so.2 not found: ") + dlerror();
    return false;
  }
  bool all = true;
  auto sym = [&](const char *name) {
    void *p = dlsym(g_nccl.handle, name);
    if (!p)
      all = false;
    return p;
  };
  g_nccl.CommInitAll = (decltype(g_nccl.CommInitAll))sym("ncclCommInitAll");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))sym("ncclAllReduce");
  g_nccl.GroupStart = (decltype(g_nccl.GroupStart))sym("ncclGroupStart");
  g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))sym("ncclGroupEnd");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
  if (!all) {
    g_nccl.error = "libnccl.so.2 lacks an expected entry point";
    return false;
  }
  g_nccl.ok = true;
  return true;
}
} /* namespace */

extern "C" int b200_film_allreduce(b200_ctx **ctxs, int n, const uint64_t *films,
                                   size_t n_floats)
{
  if (!ctxs || n <= 0 || !films)
    return B200_ERR_INVALID;
  if (n == 1)
    return B200_OK;
  std::vector<int> ordinals(n);
  for (int i = 0; i < n; i++) {
    if (!ctxs[i] || !films[i])
      return B200_ERR_INVALID;
    ordinals[i] = ctxs[i]->ordinal;
    for (int j = 0; j < i; j++)
      if (ordinals[j] == ordinals[i])
        return fail(ctxs[0], B200_ERR_UNSUPPORTED,
                    "b200_film_allreduce needs one context per GPU (use b200_film_reduce for "
                    "several contexts of one GPU)");
  }
  std::lock_guard<std::mutex> lock(g_nccl_mutex);
  if (!nccl_api())
    return fail(ctxs[0], B200_ERR_UNSUPPORTED, "NCCL: " + g_nccl.error);
  NcclClique *clique = nullptr;
  for (NcclClique &c : g_nccl_cliques)
    if (c.ordinals == ordinals)
      clique = &c;
  if (!clique) {
    NcclClique c;
    c.ordinals = ordinals;
    c.comms.resize(n);
    const ncclResult_t r = g_nccl.CommInitAll(c.comms.data(), n, ordinals.data());
    if (r != ncclSuccess)
      return fail(ctxs[0], B200_ERR_CUDA,
                  std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(r));
    g_nccl_cliques.push_back(c);
    clique = &g_nccl_cliques.back();
  }
  /* one in-place all-reduce per GPU, each on its context's stream (ordered after that
   * GPU's render), submitted as one group */
  ncclResult_t r = g_nccl.GroupStart();
  for (int i = 0; i < n && r == ncclSuccess; i++)
    r = g_nccl.AllReduce((const void *)films[i], (void *)films[i], n_floats, ncclFloat, ncclSum,
                         clique->comms[i], ctxs[i]->stream);
  const ncclResult_t e = g_nccl.GroupEnd();
  if (r == ncclSuccess)
    r = e;
  if (r != ncclSuccess)
    return fail(ctxs[0], B200_ERR_CUDA, std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r));
  for (int i = 0; i < n; i++) {
    DeviceGuard guard(ctxs[i]->ordinal);
    CUDA_TRY(ctxs[i], cudaStreamSynchronize(ctxs[i]->stream));
  }
  return B200_OK;
}

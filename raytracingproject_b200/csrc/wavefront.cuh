/* wavefront.cuh - the path-tracing wavefront: SoA path pool, stage kernels and
 * the host launch loop behind b200_render.  Included by b200_cycles.cu.
 *
 * Stage contract (names of the north star; reference semantics in brackets):
 *   init_from_camera   [split/kernel_path_init.h:24, kernel_path.h:643-680]
 *   intersect_closest  [split/kernel_scene_intersect.h:26, kernel_path.h:57-84]
 *   sort / compaction  [split/kernel_queue_enqueue.h:38, kernel_shader_sort.h:19-95]
 *   shade_background   [split/kernel_indirect_background.h:19, kernel_path.h:115-144]
 *   shade_surface      [split/kernel_shader_setup.h .. kernel_next_iteration_setup.h,
 *                       kernel_path.h:254-321,540-640, kernel_path_surface.h:22-358]
 *   intersect_shadow + shade_shadow [split/kernel_shadow_blocked_dl.h:20-94]
 *   film write         [split/kernel_buffer_update.h:41-171, kernel_passes.h:338-350]
 *
 * Differences from the reference's split kernel, by design: paths live in SoA
 * arrays (one 128-bit record per field group) instead of AoS structs; queues
 * are dense index arrays appended with one warp-aggregated atomic per warp;
 * every stage runs over its queue only (not over the whole pool); hits are
 * really sorted by shader before shading; the film is written once per batch by
 * a float4 read-modify-write with one owner thread per pixel (no atomics); the
 * host reads back one counter per bounce instead of the whole ray_state array.
 */
#ifndef B200_WAVEFRONT_CUH
#define B200_WAVEFRONT_CUH

#include <cuda_fp16.h>

#include "shade.cuh"
#include "passes.cuh"

/* bit 30 of a shadow-queue entry's transparent-bounce word: the path was past its first
 * bounce when it sampled the light (selects the indirect clamp in shadow_light_arrives) */
#define SH_INDIRECT_FLAG 0x40000000
/* float4 records per shadow-queue entry in PathSoA::sh_pass (render passes only) */
#define SH_PASS_QUADS 3
/* transparent shadows: up to this many stepping rounds are queued per iteration without
 * reading the queue length back (wavefront loop) */
#define WF_TS_BLIND_ROUNDS 16

#define WF_MAX_KEYS 4096
#ifndef WF_BLOCK
#  define WF_BLOCK 256
#endif

#include "adaptive.cuh"
/* sort keys below this are histogrammed / ranked in shared memory, the (rare) rest by
 * global atomics */
#define WF_SMALL_KEYS 64

#define WF_RING 4
#define WF_RING_BYTES 64
#define WF_RING_EVENTS 5 /* start, after closest, after shading, after shadow, counters home */
#define WF_BATCH_SLOTS 1024
#define WF_BATCH_STAT_BYTES 128
#define WF_FLAG_SCOPE_MISS 1u
#define WF_FLAG_TRACE_OVERFLOW 2u

struct WFCounters {
  unsigned int n_active; /* paths in q_active (input of intersect_closest) */
  unsigned int n_next;   /* paths appended to q_next by shade_surface */
  unsigned int n_shadow; /* shadow rays appended by shade_surface */
  unsigned int work_closest, work_shadow;
  /* transparent shadows: sizes of the two stepping queues, their traversal cursor */
  unsigned int n_ts[2], work_ts;
  /* device flags folded in by k_iteration_end so that they travel with the 64-byte
   * counter read-back: bit 0 = g_svm_scope_miss, bit 1 = g_trace_overflow */
  unsigned int flags, pad_flags;
  unsigned long long primary_rays, bounce_rays, shadow_rays;
  unsigned long long nodes, tris, instances;          /* intersect_closest */
  unsigned long long sh_nodes, sh_tris, sh_instances; /* intersect_shadow */
  unsigned int hist[WF_MAX_KEYS + 1];
  unsigned int offsets[WF_MAX_KEYS + 2];
  unsigned int cursor[WF_MAX_KEYS + 1];
  unsigned int adaptive_any; /* adaptive filter: some pixel of the tile still samples */
};

struct PathSoA {
  /* ray queue, in QUEUE order (entry qi belongs to path q_active[qi]): the 32-byte
   * "ray in" record the traversal kernel streams through shared memory */
  float4 *ray_P_t;    /* Ray::P, Ray::t */
  float4 *ray_D;      /* Ray::D, visibility mask of this segment */
  float4 *nray_P_t;   /* next bounce's queue, written by shade_surface */
  float4 *nray_D;
  float4 *hit;        /* [qi] Intersection t,u,v, prim */
  int *hit_object;    /* [qi] Intersection::object */
  float *ray_pdf;     /* [path] PathState::ray_pdf */
  float4 *throughput; /* throughput, PathState::ray_t */
  float4 *L;          /* PathRadiance::emission, transparent */
  uint4 *stateA;      /* flag, rng_hash, rng_offset, sample */
  uint4 *stateB;      /* bounce|diffuse<<16, glossy|transmission<<16, transparent, min_ray_pdf */
  float4 *sh_P_t;     /* [shadow queue] shadow ray P, t */
  float4 *sh_D;       /* [shadow queue] D, PATH_RAY_SHADOW_OPAQUE */
  float4 *sh_contrib; /* [shadow queue] throughput * L_light (already clamped) */
  unsigned int *key;  /* [qi] sort key: 0 miss, 1 + shader */
  int *q_active; /* [qi] -> path */
  int *q_next;   /* next bounce's q_active */
  int2 *q_sorted; /* (queue position qi, path index) ordered by key */
  int *q_shadow; /* [shadow queue] -> path */
  /* Render passes beyond the combined one (passes.cuh), null otherwise: the per-path
   * accumulator block, and two more colours per shadow-queue entry - a light sample on a
   * directly visible surface arrives split per BSDF class (sh_contrib = diffuse,
   * sh_pass[3 qi] = glossy | shadow-pass weight, sh_pass[3 qi + 1] = transmission | kind),
   * sh_pass[3 qi + 2] = the unweighted light of the denoiser's shadowing feature */
  float *pass;
  float4 *sh_pass;
  /* 1: q_sorted is ordered by key inside every tile of SORT_TILE queue entries
   * (k_sort_tiles), misses flagged in the sign bit of the queue position; 0: one global
   * counting sort, segment bounds in the counters */
  int sort_tiles;
  /* Transparent shadows (only allocated when integrator.transparent_shadows): shadow rays
   * whose first hit is a transparent surface step from surface to surface
   * (kernel_shadow.h:300-352).  Two ping-pong queues of (ray, index into the shadow
   * queue); the running attenuation lives with the shadow-queue entry. */
  float4 *ts_P_t[2];
  float4 *ts_D[2];
  int *ts_idx[2];
  float4 *ts_thr; /* [shadow queue] attenuation so far, .w = transparent bounce count */
  WFCounters *counters;
  /* debugging aid: when debug != NULL the path in slot debug_slot records 32 floats
   * per bounce (ray, hit, shading point, closures, sampled direction) */
  float *debug;
  int debug_slot;
};

struct PathPool {
  size_t capacity = 0;
  PathSoA soa;
  void *block = nullptr;
  size_t bytes = 0;
  bool has_ts = false; /* transparent-shadow arrays carved */
  bool has_ao = false; /* shadow queue sized for two entries per path (light + AO ray) */
  bool has_passes = false; /* per-path pass accumulators + split shadow contributions */
  WFCounters *h_counters = nullptr; /* pinned */
  uint32_t *sobol_tab = nullptr;    /* SOBOL_TABLE_MAX entries */
  /* The bounce loop runs ahead of the host: iteration `it` copies the head of the
   * counters into ring slot it % WF_RING and records the slot's events; the host reads
   * slot it - 1 while iteration `it` is already queued, so the GPU never waits for it. */
  unsigned char *h_ring = nullptr;  /* pinned, WF_RING * WF_RING_BYTES */
  cudaEvent_t ring_ev[WF_RING][WF_RING_EVENTS] = {};
  /* ray / traversal statistics of each batch (the first WF_BATCH_STAT_BYTES of its
   * counters), summed on the host when the call ends */
  unsigned char *h_batch = nullptr; /* pinned, WF_BATCH_SLOTS * WF_BATCH_STAT_BYTES */
};

struct BatchParams {
  int x, y, w, h;      /* pixel rectangle of this batch */
  int sample0, nsamples;
  int offset, stride;  /* film addressing (buffers.cpp:50-54) */
  int num_keys;
  int count_stats;
  /* adaptive sampling (adaptive.cuh): the film, to leave converged pixels out; null = off */
  const float *adaptive_film;
  int pass_stride, adaptive_aux;
};

/* ------------------------------------------------------------ helpers */

/* one atomic per warp: lanes with `pred` get consecutive slots */
CY_DEV unsigned int warp_append(unsigned int *counter, bool pred)
{
  const unsigned int mask = __ballot_sync(__activemask(), pred);
  if (!pred)
    return 0;
  const unsigned int lane = threadIdx.x & 31u;
  const unsigned int leader = __ffs(mask) - 1u;
  unsigned int base = 0;
  if (lane == leader)
    base = atomicAdd(counter, __popc(mask));
  base = __shfl_sync(mask, base, leader);
  return base + __popc(mask & ((1u << lane) - 1u));
}

/* One atomic per BLOCK for up to two queues at once: the queue counters are single
 * addresses, and same-address atomics with a return value serialise in L2 at about
 * 2 ns each - per warp that was the whole cost of init_from_camera and a third of
 * shade_surface.  Must be reached by every thread of the block (32 NW threads).
 * Lanes get consecutive slots in thread order, so what a block appends stays in the
 * (sorted, spatially coherent) order it was read in. */
template<int NW = WF_BLOCK / 32>
CY_DEV void block_append2(unsigned int *counter_a, bool pred_a, unsigned int *counter_b,
                          bool pred_b, unsigned int *slot_a, unsigned int *slot_b)
{
  static_assert(NW <= 32, "one warp scans the per-warp counts of a queue");
  __shared__ unsigned int s_tab[2][NW];
  const unsigned int lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  const unsigned int lt_mask = (1u << lane) - 1u;
  const unsigned int ma = __ballot_sync(0xffffffffu, pred_a);
  const unsigned int mb = __ballot_sync(0xffffffffu, pred_b);
  if (lane == 0) {
    s_tab[0][w] = __popc(ma);
    s_tab[1][w] = __popc(mb);
  }
  __syncthreads();
  if (NW <= 16) {
    if (w == 0) {
      /* lanes 0..NW-1 scan queue a, lanes 16..16+NW-1 queue b */
      const unsigned int q = lane >> 4, j = lane & 15u;
      const unsigned int v = (j < NW) ? s_tab[q][j] : 0u;
      unsigned int inc = v;
#pragma unroll
      for (int o = 1; o < NW; o <<= 1) {
        const unsigned int u = __shfl_up_sync(0xffffffffu, inc, o, 16);
        if (j >= (unsigned)o)
          inc += u;
      }
      unsigned int base = 0;
      if (j == NW - 1 && inc != 0u)
        base = atomicAdd(q ? counter_b : counter_a, inc);
      base = __shfl_sync(0xffffffffu, base, NW - 1, 16);
      if (j < NW)
        s_tab[q][j] = base + inc - v;
    }
  }
  else if (w < 2) {
    /* more than 16 warps: warp 0 scans queue a, warp 1 queue b */
    const unsigned int q = w, j = lane;
    const unsigned int v = (j < NW) ? s_tab[q][j] : 0u;
    unsigned int inc = v;
#pragma unroll
    for (int o = 1; o < NW; o <<= 1) {
      const unsigned int u = __shfl_up_sync(0xffffffffu, inc, o);
      if (j >= (unsigned)o)
        inc += u;
    }
    unsigned int base = 0;
    if (j == NW - 1 && inc != 0u)
      base = atomicAdd(q ? counter_b : counter_a, inc);
    base = __shfl_sync(0xffffffffu, base, NW - 1);
    if (j < NW)
      s_tab[q][j] = base + inc - v;
  }
  __syncthreads();
  *slot_a = s_tab[0][w] + __popc(ma & lt_mask);
  *slot_b = s_tab[1][w] + __popc(mb & lt_mask);
  __syncthreads(); /* the table is reused by the next call */
}

CY_DEV void state_load(const PathSoA &p, int i, PathStateG &s)
{
  const uint4 a = p.stateA[i];
  const uint4 b = p.stateB[i];
  s.flag = a.x;
  s.rng_hash = a.y;
  s.rng_offset = (int)a.z;
  s.sample = (int)a.w;
  s.bounce = (int)(b.x & 0xffffu);
  s.diffuse_bounce = (int)(b.x >> 16);
  s.glossy_bounce = (int)(b.y & 0xffffu);
  s.transmission_bounce = (int)(b.y >> 16);
  s.transparent_bounce = (int)b.z;
  s.min_ray_pdf = __uint_as_float(b.w);
}
CY_DEV void state_store(const PathSoA &p, int i, const PathStateG &s)
{
  p.stateA[i] = make_uint4(s.flag, s.rng_hash, (unsigned)s.rng_offset, (unsigned)s.sample);
  p.stateB[i] = make_uint4((unsigned)s.bounce | ((unsigned)s.diffuse_bounce << 16),
                           (unsigned)s.glossy_bounce | ((unsigned)s.transmission_bounce << 16),
                           (unsigned)s.transparent_bounce, __float_as_uint(s.min_ray_pdf));
}

/* kernel_path_state.h:190-203 */
CY_DEV uint32_t path_state_ray_visibility(uint32_t state_flag)
{
  uint32_t flag = state_flag & CY_PATH_RAY_ALL_VISIBILITY;
  if (flag & CY_PATH_RAY_TRANSMIT)
    flag &= ~(CY_PATH_RAY_DIFFUSE | CY_PATH_RAY_GLOSSY);
  if (state_flag & CY_PATH_RAY_VOLUME_SCATTER)
    flag |= CY_PATH_RAY_DIFFUSE;
  return flag;
}

/* kernel_path_state.h:251-260 */
CY_DEV bool path_state_ao_bounce(const PathStateG &s)
{
  const int ao_bounces = kd_int(KD_INT_AO_BOUNCES);
  if (s.bounce <= ao_bounces)
    return false;
  int bounce = s.bounce - s.transmission_bounce - (s.glossy_bounce > 0);
  return (bounce > ao_bounces);
}

/* kernel_path_state.h:205-241 (no shadow catcher) */
CY_DEV float path_state_continuation_probability(const PathStateG &s, f3 throughput)
{
  if (s.flag & CY_PATH_RAY_TERMINATE_IMMEDIATE) {
    return 0.0f;
  }
  else if (s.flag & CY_PATH_RAY_TRANSPARENT) {
    if (s.transparent_bounce <= kd_int(KD_INT_TRANSPARENT_MIN_BOUNCE))
      return 1.0f;
  }
  else {
    if (s.bounce <= kd_int(KD_INT_MIN_BOUNCE))
      return 1.0f;
  }
  /* branch_factor is 1 for the non-branched integrator */
  return fminf(sqrtf(max3(fabs3(throughput)) * 1.0f), 1.0f);
}

/* kernel_path_state.h:72-169 (surface labels) */
CY_DEV void path_state_next(PathStateG &s, int label)
{
  if (label & CY_LABEL_TRANSPARENT) {
    s.flag |= CY_PATH_RAY_TRANSPARENT;
    s.transparent_bounce++;
    if (s.transparent_bounce >= kd_int(KD_INT_TRANSPARENT_MAX_BOUNCE))
      s.flag |= CY_PATH_RAY_TERMINATE_IMMEDIATE;
    if (!kd_int(KD_INT_TRANSPARENT_SHADOWS))
      s.flag |= CY_PATH_RAY_MIS_SKIP;
    s.rng_offset += CY_PRNG_BOUNCE_NUM;
    return;
  }
  s.bounce++;
  if (s.bounce >= kd_int(KD_INT_MAX_BOUNCE))
    s.flag |= CY_PATH_RAY_TERMINATE_AFTER_TRANSPARENT;
  s.flag &= ~(CY_PATH_RAY_ALL_VISIBILITY | CY_PATH_RAY_MIS_SKIP);

  if (label & CY_LABEL_REFLECT) {
    s.flag |= CY_PATH_RAY_REFLECT;
    s.flag &= ~CY_PATH_RAY_TRANSPARENT_BACKGROUND;
    if (label & CY_LABEL_DIFFUSE) {
      s.diffuse_bounce++;
      if (s.diffuse_bounce >= kd_int(KD_INT_MAX_DIFFUSE_BOUNCE))
        s.flag |= CY_PATH_RAY_TERMINATE_AFTER_TRANSPARENT;
    }
    else {
      s.glossy_bounce++;
      if (s.glossy_bounce >= kd_int(KD_INT_MAX_GLOSSY_BOUNCE))
        s.flag |= CY_PATH_RAY_TERMINATE_AFTER_TRANSPARENT;
    }
  }
  else {
    s.flag |= CY_PATH_RAY_TRANSMIT;
    if (!(label & CY_LABEL_TRANSMIT_TRANSPARENT))
      s.flag &= ~CY_PATH_RAY_TRANSPARENT_BACKGROUND;
    s.transmission_bounce++;
    if (s.transmission_bounce >= kd_int(KD_INT_MAX_TRANSMISSION_BOUNCE))
      s.flag |= CY_PATH_RAY_TERMINATE_AFTER_TRANSPARENT;
  }
  if (label & CY_LABEL_DIFFUSE) {
    s.flag |= CY_PATH_RAY_DIFFUSE | CY_PATH_RAY_DIFFUSE_ANCESTOR;
  }
  else if (label & CY_LABEL_GLOSSY) {
    s.flag |= CY_PATH_RAY_GLOSSY;
  }
  else {
    s.flag |= CY_PATH_RAY_GLOSSY | CY_PATH_RAY_SINGULAR | CY_PATH_RAY_MIS_SKIP;
  }
  s.rng_offset += CY_PRNG_BOUNCE_NUM;
}

/* Pixel order inside a batch: 8x4 tiles, so the 32 lanes of a warp start with a
 * compact bundle of camera rays (coherent traversal, 128-byte film rows); plain
 * rows when the rectangle is not a multiple of the tile. */
CY_DEV void batch_pixel(const BatchParams &bp, unsigned int pix, int *x, int *y)
{
  if (((bp.w & 7) | (bp.h & 3)) == 0) {
    const unsigned int t = pix >> 5, l = pix & 31u;
    const unsigned int tiles_x = (unsigned)bp.w >> 3;
    *x = bp.x + (int)((t % tiles_x) * 8u + (l & 7u));
    *y = bp.y + (int)((t / tiles_x) * 4u + (l >> 3));
  }
  else {
    *x = bp.x + (int)(pix % (unsigned)bp.w);
    *y = bp.y + (int)(pix / (unsigned)bp.w);
  }
}

/* ----------------------------------------------------------- Sobol table */

#define SOBOL_TABLE_MAX (1u << 16) /* entries: 256 KB */

__global__ void k_sobol_table(uint32_t *tab, int s0, unsigned int ns, unsigned int nd)
{
  const unsigned int n = ns * nd;
  for (unsigned int e = blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += gridDim.x * blockDim.x)
    tab[e] = sobol_dimension(s0 + (int)(e / nd), (int)(e % nd));
}

/* ------------------------------------------------------ init_from_camera */

__global__ void __launch_bounds__(WF_BLOCK)
    k_init_from_camera(PathSoA p, BatchParams bp)
{
  const unsigned int npix = (unsigned)bp.w * (unsigned)bp.h;
  const unsigned int n = npix * (unsigned)bp.nsamples;
  /* block-uniform loop: block_append2 has barriers */
  for (unsigned int i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x) {
    const unsigned int i = i0 + threadIdx.x;
    float t = 0.0f;
    f3 P = zero3(), D = zero3();
    uint32_t vis = 0;
    if (i < n) {
      const unsigned int pix = i % npix;
      const int s = (int)(i / npix);
      int x, y;
      batch_pixel(bp, pix, &x, &y);
      const int sample = bp.sample0 + s;

      uint32_t rng_hash;
      t = camera_ray(x, y, sample, &rng_hash, &P, &D);
      /* a pixel the adaptive sampler has stopped traces nothing (kernel_path.h:660-666) */
      bool stopped = false;
      if (bp.adaptive_film) {
        const long long index = (long long)bp.offset + x + (long long)y * bp.stride;
        stopped = bp.adaptive_film[index * bp.pass_stride + bp.adaptive_aux + 3] > 0.0f;
        if (stopped)
          t = 0.0f;
      }

      /* path_state_init - kernel_path_state.h:19-70 */
      PathStateG st;
      st.flag = CY_PATH_RAY_CAMERA | CY_PATH_RAY_MIS_SKIP | CY_PATH_RAY_TRANSPARENT_BACKGROUND;
      if (p.pass && film_has_denoising())
        st.flag |= CY_PATH_RAY_STORE_SHADOW_INFO;
      st.rng_hash = rng_hash;
      st.rng_offset = CY_PRNG_BASE_NUM;
      st.sample = sample;
      st.bounce = st.diffuse_bounce = st.glossy_bounce = st.transmission_bounce = 0;
      st.transparent_bounce = 0;
      st.min_ray_pdf = FLT_MAX;
      state_store(p, i, st);
      vis = path_state_ray_visibility(st.flag);

      /* ray_pdf < 0 marks a pixel sample that writes nothing to the adaptive film */
      p.ray_pdf[i] = (bp.adaptive_film && t == 0.0f) ? -1.0f : 0.0f;
      p.throughput[i] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);    /* ray_t = 0 */
      /* kernel_path_trace returns before kernel_write_result when ray.t == 0
       * (kernel_path.h:668-670): transparent = 1 makes the film add (0,0,0,0). */
      p.L[i] = make_float4(0.0f, 0.0f, 0.0f, (t == 0.0f) ? 1.0f : 0.0f);
      if (p.pass && t == 0.0f) /* the block was cleared for the batch */
        p.pass[(size_t)i * PASS_WORDS + PB_UNTRACED] = 1.0f;
      if (p.pass && film_has_denoising()) {
        float *pb = p.pass + (size_t)i * PASS_WORDS;
        pb[PB_DN_WEIGHT] = 1.0f;
        pb_set3(pb, PB_DN_THROUGHPUT, mk3(1.0f, 1.0f, 1.0f));
      }
    }
    unsigned int slot, unused;
    block_append2(&p.counters->n_active, t != 0.0f, &p.counters->n_active, false, &slot, &unused);
    if (t != 0.0f) {
      p.q_active[slot] = (int)i;
      p.ray_P_t[slot] = make_float4(P.x, P.y, P.z, t);
      p.ray_D[slot] = make_float4(D.x, D.y, D.z, __uint_as_float(vis));
    }
  }
}

/* ----------------------------------------------------- intersect_closest */

struct ClosestJob {
  static constexpr bool QUEUE_RAYS = true; /* two arrays of 16-byte records in queue order */
  PathSoA p;
  __device__ __forceinline__ const float4 *ray_P(unsigned int qi) const
  {
    return p.ray_P_t + qi;
  }
  __device__ __forceinline__ const float4 *ray_D(unsigned int qi) const
  {
    return p.ray_D + qi;
  }
  __device__ __forceinline__ void store(unsigned int qi, const TraceHit &h, bool found)
  {
    p.hit[qi] = make_float4(h.t, h.u, h.v, __int_as_float(h.prim));
    p.hit_object[qi] = h.object;
    unsigned int key = 0;
    if (found) {
      const unsigned int tri = __ldg(&g_scene.prim_index[h.prim]);
      key = 1u + (__ldg(&g_scene.tri_shader[tri]) & CY_SHADER_MASK);
      if (key > WF_MAX_KEYS)
        key = WF_MAX_KEYS;
    }
    p.key[qi] = key;
  }
};

template<bool COUNT>
__global__ void __launch_bounds__(TRACE_BLOCK, TRACE_MIN_BLOCKS) k_intersect_closest(PathSoA p, int refill_threshold)
{
  const unsigned lane = threadIdx.x & 31u;
  TraceCounters cnt;
  cnt.nodes = cnt.tris = cnt.instances = 0;
  ClosestJob job;
  job.p = p;
  trace_persistent<false, COUNT>(job, p.counters->n_active, &p.counters->work_closest,
                                 refill_threshold, cnt);
  if (COUNT) {
    for (int o = 16; o > 0; o >>= 1) {
      cnt.nodes += __shfl_xor_sync(0xffffffffu, cnt.nodes, o);
      cnt.tris += __shfl_xor_sync(0xffffffffu, cnt.tris, o);
      cnt.instances += __shfl_xor_sync(0xffffffffu, cnt.instances, o);
    }
    if (lane == 0) {
      atomicAdd(&p.counters->nodes, (unsigned long long)cnt.nodes);
      atomicAdd(&p.counters->tris, (unsigned long long)cnt.tris);
      atomicAdd(&p.counters->instances, (unsigned long long)cnt.instances);
    }
  }
}

/* ----------------------------------------------- sort by shader (counting) */

/* histogram of the sort keys: per-block shared histogram (one shared atomic per
 * distinct key per warp), one global atomic per key and block */
__global__ void __launch_bounds__(WF_BLOCK) k_sort_count(PathSoA p)
{
  __shared__ unsigned int s_cnt[WF_SMALL_KEYS];
  WFCounters *c = p.counters;
  const unsigned int n = c->n_active;
  const unsigned int lane = threadIdx.x & 31u;
  if (threadIdx.x < WF_SMALL_KEYS)
    s_cnt[threadIdx.x] = 0u;
  __syncthreads();
  const unsigned int n_round = (n + 31u) & ~31u; /* whole warps stay converged */
  for (unsigned int qi = blockIdx.x * blockDim.x + threadIdx.x; qi < n_round;
       qi += gridDim.x * blockDim.x) {
    const unsigned int key = (qi < n) ? p.key[qi] : 0xffffffffu;
    const unsigned int peers = __match_any_sync(0xffffffffu, key);
    if (lane == (unsigned)(__ffs(peers) - 1) && qi < n) {
      if (key < WF_SMALL_KEYS)
        atomicAdd(&s_cnt[key], __popc(peers));
      else
        atomicAdd(&c->hist[key], __popc(peers));
    }
  }
  __syncthreads();
  if (threadIdx.x < WF_SMALL_KEYS && s_cnt[threadIdx.x] != 0u)
    atomicAdd(&c->hist[threadIdx.x], s_cnt[threadIdx.x]);
}

__global__ void k_sort_scan(PathSoA p, int num_keys)
{
  /* single block: exclusive scan of the key histogram */
  __shared__ unsigned int sh[WF_MAX_KEYS + 1];
  WFCounters *c = p.counters;
  for (int k = threadIdx.x; k <= num_keys; k += blockDim.x)
    sh[k] = c->hist[k];
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int run = 0;
    for (int k = 0; k <= num_keys; k++) {
      const unsigned int v = sh[k];
      c->offsets[k] = run;
      c->cursor[k] = 0;
      run += v;
    }
    c->offsets[num_keys + 1] = run;
  }
}

/* Counting-sort scatter.  A block ranks a tile of SORT_ITEMS * WF_BLOCK queue entries
 * in shared memory (one shared atomic per distinct key per warp, __match_any_sync) and
 * reserves its share of every key's output range with ONE global atomic per key and
 * tile, so the few hot counters (one shader, "miss") are not hammered by every warp.
 * The order inside a key is arbitrary - nothing downstream depends on it (the film is
 * summed per pixel in sample order). */
#define SORT_ITEMS 8
__global__ void __launch_bounds__(WF_BLOCK) k_sort_scatter(PathSoA p)
{
  __shared__ unsigned int s_cnt[WF_SMALL_KEYS], s_base[WF_SMALL_KEYS];
  WFCounters *c = p.counters;
  const unsigned int n = c->n_active;
  const unsigned int lane = threadIdx.x & 31u;
  const unsigned int lt_mask = (1u << lane) - 1u;
  const unsigned int tile_size = SORT_ITEMS * WF_BLOCK;
  for (unsigned int tile = blockIdx.x * tile_size; tile < n; tile += gridDim.x * tile_size) {
    if (threadIdx.x < WF_SMALL_KEYS)
      s_cnt[threadIdx.x] = 0u;
    __syncthreads();
    unsigned int key[SORT_ITEMS], rank[SORT_ITEMS];
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++) {
      const unsigned int qi = tile + k * WF_BLOCK + threadIdx.x;
      key[k] = (qi < n) ? p.key[qi] : 0xffffffffu;
      const unsigned int peers = __match_any_sync(0xffffffffu, key[k]);
      const unsigned int leader = __ffs(peers) - 1u;
      unsigned int base = 0;
      if (lane == leader && qi < n) {
        if (key[k] < WF_SMALL_KEYS)
          base = atomicAdd(&s_cnt[key[k]], __popc(peers));
        else /* many-shader scenes: straight to the global cursor, final position */
          base = atomicAdd(&c->cursor[key[k]], __popc(peers));
      }
      base = __shfl_sync(0xffffffffu, base, leader);
      rank[k] = base + __popc(peers & lt_mask);
    }
    __syncthreads();
    if (threadIdx.x < WF_SMALL_KEYS && s_cnt[threadIdx.x] != 0u)
      s_base[threadIdx.x] = atomicAdd(&c->cursor[threadIdx.x], s_cnt[threadIdx.x]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++) {
      const unsigned int qi = tile + k * WF_BLOCK + threadIdx.x;
      if (qi < n) {
        const unsigned int kk = key[k];
        const unsigned int pos = c->offsets[kk] + rank[k] + (kk < WF_SMALL_KEYS ? s_base[kk] : 0u);
        p.q_sorted[pos] = make_int2((int)qi, p.q_active[qi]);
      }
    }
    __syncthreads();
  }
}


/* Sort by shader INSIDE tiles of SORT_TILE consecutive queue entries - an experiment kept
 * behind b200_set_option("sort_tiles", 1), NOT the default.
 *
 * Idea: a global sort by shader makes warps run one shader program, but consecutive
 * threads then hold paths from anywhere in the queue; sorting only within a 2048-entry
 * window would keep a block's gathers inside one window (every sector used in full while
 * it is in L1) at the price of a few mixed warps per tile.  Measured on B200 (tools/
 * r02_run12.sh, profiles/r02i_sort_tiles_ab.txt): Cornell 128 spp 372 -> 567 ms, startup
 * scene 51 -> 65 ms, terrain 77 -> 88 ms, instanced 349 -> 362 ms - clearly slower.  With
 * the global sort the whole device runs ONE shader program at a time (the shade kernel is
 * ~100 KB of SASS; per-tile order has every SM cycling through all programs at once and
 * the instruction-cache stalls, already 3.7 per issue, take over), and the locality the
 * tiles were meant to add is largely there already: k_sort_scatter reserves a key's
 * output range per tile, so a warp's 32 entries of the global order come from one or two
 * 2048-entry windows anyway.  Entry = (queue position | miss flag in the sign bit, path
 * index); key 0 (miss) comes first in a tile. */
#define SORT_TILE (SORT_ITEMS * WF_BLOCK)
#define SORT_TILE_MAX_KEYS 16
__global__ void __launch_bounds__(WF_BLOCK) k_sort_tiles(PathSoA p)
{
  __shared__ unsigned int s_cnt[WF_SMALL_KEYS], s_off[WF_SMALL_KEYS];
  WFCounters *c = p.counters;
  const unsigned int n = c->n_active;
  const unsigned int lane = threadIdx.x & 31u;
  const unsigned int lt_mask = (1u << lane) - 1u;
  for (unsigned int tile = blockIdx.x * SORT_TILE; tile < n; tile += gridDim.x * SORT_TILE) {
    if (threadIdx.x < WF_SMALL_KEYS)
      s_cnt[threadIdx.x] = 0u;
    __syncthreads();
    unsigned int key[SORT_ITEMS], rank[SORT_ITEMS];
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++) {
      const unsigned int qi = tile + k * WF_BLOCK + threadIdx.x;
      key[k] = (qi < n) ? min(p.key[qi], (unsigned int)(WF_SMALL_KEYS - 1)) : 0xffffffffu;
      const unsigned int peers = __match_any_sync(0xffffffffu, key[k]);
      const unsigned int leader = __ffs(peers) - 1u;
      unsigned int base = 0;
      if (lane == leader && qi < n)
        base = atomicAdd(&s_cnt[key[k]], __popc(peers));
      base = __shfl_sync(0xffffffffu, base, leader);
      rank[k] = base + __popc(peers & lt_mask);
    }
    __syncthreads();
    if (threadIdx.x < 32) { /* exclusive scan of the 64 counters, two per lane */
      const unsigned int a = s_cnt[2 * lane], b = s_cnt[2 * lane + 1];
      unsigned int incl = a + b;
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o)
          incl += v;
      }
      s_off[2 * lane] = incl - (a + b);
      s_off[2 * lane + 1] = incl - b;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++) {
      const unsigned int qi = tile + k * WF_BLOCK + threadIdx.x;
      if (qi < n) {
        const unsigned int pos = tile + s_off[key[k]] + rank[k];
        p.q_sorted[pos] = make_int2((int)(qi | (key[k] == 0u ? 0x80000000u : 0u)),
                                    p.q_active[qi]);
      }
    }
    __syncthreads();
  }
}

/* --------------------------------------------------- lamp emission (MIS) */

/* kernel_path.h:86-113 + kernel_emission.h:235-286: emission of lamps hit by the
 * ray segment, weighted against light sampling.  Returns the updated ray_t. */
template<bool EXT>
CY_DEV void path_lamp_emission(PathStateG &st, f3 rayP, f3 rayD, float isect_t, f3 throughput,
                               ShaderDataG &emission_sd, f3 &L_emission)
{
  if (kd_int(KD_INT_USE_LAMP_MIS) && !(st.flag & CY_PATH_RAY_CAMERA)) {
    const f3 lP = rayP - st.ray_t * rayD;
    st.ray_t += isect_t;
    const float lt = st.ray_t;
    const int num_lights = kd_int(KD_INT_NUM_ALL_LIGHTS);
    for (int lamp = 0; lamp < num_lights; lamp++) {
      LightSampleG ls;
      if (!lamp_light_eval(lamp, lP, rayD, lt, &ls))
        continue;
      if (ls.shader & CY_SHADER_EXCLUDE_ANY) {
        if (((ls.shader & CY_SHADER_EXCLUDE_DIFFUSE) && (st.flag & CY_PATH_RAY_DIFFUSE)) ||
            ((ls.shader & CY_SHADER_EXCLUDE_GLOSSY) &&
             ((st.flag & (CY_PATH_RAY_GLOSSY | CY_PATH_RAY_REFLECT)) ==
              (CY_PATH_RAY_GLOSSY | CY_PATH_RAY_REFLECT))) ||
            ((ls.shader & CY_SHADER_EXCLUDE_TRANSMIT) && (st.flag & CY_PATH_RAY_TRANSMIT)) ||
            ((ls.shader & CY_SHADER_EXCLUDE_SCATTER) && (st.flag & CY_PATH_RAY_VOLUME_SCATTER)))
          continue;
      }
      f3 lamp_L = direct_emissive_eval<EXT>(emission_sd, path_depths(st), &ls, -rayD, ls.t);
      if (!(st.flag & CY_PATH_RAY_MIS_SKIP)) {
        float mis_weight = power_heuristic(st.ray_pdf, ls.pdf);
        lamp_L *= mis_weight;
      }
      /* path_radiance_accum_emission - kernel_accumulate.h:310-340 */
      f3 contribution = throughput * lamp_L;
      path_radiance_clamp(&contribution, st.bounce - 1);
      L_emission += contribution;
    }
  }
}

/* ----------------------------------------------------- shade_background */

#ifndef BG_MIN_BLOCKS
#  define BG_MIN_BLOCKS 1
#endif
template<bool EXT, bool PASSES = false>
__global__ void __launch_bounds__(WF_BLOCK, BG_MIN_BLOCKS) k_shade_background(PathSoA p)
{
  WFCounters *c = p.counters;
  /* the key 0 segment of the global sort, or the flagged entries of every tile */
  const unsigned int n = p.sort_tiles ? c->n_active : c->offsets[1];
  for (unsigned int qi = blockIdx.x * blockDim.x + threadIdx.x; qi < n;
       qi += gridDim.x * blockDim.x) {
    const int2 qs = p.q_sorted[qi];
    if (p.sort_tiles && qs.x >= 0)
      continue;
    const int qpos = qs.x & 0x7fffffff, i = qs.y;
    const float4 r0 = p.ray_P_t[qpos];
    const float4 r1 = p.ray_D[qpos];
    const float4 tp = p.throughput[i];
    float4 Lr = p.L[i];
    PathStateG st;
    state_load(p, i, st);
    st.ray_pdf = p.ray_pdf[i];
    st.ray_t = tp.w;
    f3 throughput = mk3(tp);
    f3 L = mk3(Lr);
    const f3 rayP = mk3(r0), rayD = mk3(r1);
    const float isect_t = p.hit[qpos].x; /* = ray t on a miss (bvh_traversal.h:62) */
    /* with light passes a contribution goes to the emission pass (L), to "seen after one
     * bounce" or to "later", by the path's bounce (kernel_accumulate.h:302-340) */
    float *pb = PASSES ? p.pass + (size_t)i * PASS_WORDS : nullptr;
    const bool use_light_pass = PASSES && kd_int(KD_FILM_USE_LIGHT_PASS);
    const int bucket = use_light_pass ? emission_bucket(st.bounce) : -1;

    ShaderDataG esd;
    if (bucket < 0) {
      path_lamp_emission<EXT>(st, rayP, rayD, isect_t, throughput, esd, L);
    }
    else {
      f3 lamp = zero3();
      path_lamp_emission<EXT>(st, rayP, rayD, isect_t, throughput, esd, lamp);
      pb_add3(pb, bucket, lamp);
    }

    /* kernel_path_background - kernel_path.h:115-144 */
    bool eval_bg = true;
    if (kd_int(KD_BG_TRANSPARENT) && (st.flag & CY_PATH_RAY_TRANSPARENT_BACKGROUND)) {
      Lr.w += average(throughput);
      /* the background pass still wants the colour behind a transparent film */
      eval_bg = use_light_pass && light_pass_on(CY_PASS_BACKGROUND);
    }
    if (eval_bg) {
      if (path_state_ao_bounce(st))
        throughput *= kd_float(KD_BG_AO_BOUNCES_FACTOR);
      f3 L_background = indirect_background<EXT>(esd, st, rayD);
      /* path_radiance_accum_background - kernel_accumulate.h:478-515 */
      f3 contribution = throughput * L_background;
      if (PASSES && film_has_denoising()) {
        if (st.flag & CY_PATH_RAY_STORE_SHADOW_INFO) {
          pb_add3(pb, PB_PATH_TOTAL, contribution);
          pb_add3(pb, PB_PATH_TOTAL_SHADED, contribution * 1.0f); /* shadow_transparency */
        }
        pb_add3(pb, PB_DN_ALBEDO,
                pb_get3(pb, PB_DN_THROUGHPUT) * pb[PB_DN_WEIGHT] * L_background);
      }
      path_radiance_clamp(&contribution, st.bounce - 1);
      if (!use_light_pass)
        L += contribution;
      else if (st.flag & CY_PATH_RAY_TRANSPARENT_BACKGROUND)
        pb_add3(pb, PB_BACKGROUND, contribution);
      else
        pb_add3(pb, (st.bounce == 1) ? PB_DIRECT_EMISSION : PB_INDIRECT, contribution);
    }
    p.L[i] = make_float4(L.x, L.y, L.z, Lr.w);
  }
}

/* -------------------------------------------------------- shade_surface */

#ifndef SHADE_MIN_BLOCKS
#  define SHADE_MIN_BLOCKS 2
#endif
#ifndef SHADE_MIN_BLOCKS_EXT
#  define SHADE_MIN_BLOCKS_EXT 2
#endif
/* Dynamic shared memory of the surface-shading kernel, per block:
 *   [SHADE_STAGE_WORDS][WF_BLOCK] floats   the staged shadow-ray record of each thread,
 * a column per thread: a warp touches one 128-byte row per word, no conflicts. */
#define SHADE_STAGE_WORDS 12
#define SHADE_SMEM_BYTES_OF(block) (SHADE_STAGE_WORDS * (block) * sizeof(float))
#define SHADE_SMEM_BYTES SHADE_SMEM_BYTES_OF(WF_BLOCK)
/* with render passes the record carries three more colours and two flags */
#define SHADE_STAGE_WORDS_PASSES 24
#define SHADE_SMEM_BYTES_PASSES (SHADE_STAGE_WORDS_PASSES * WF_BLOCK * sizeof(float))

/* kernel_path_shader_apply .. kernel_path_surface_bounce (kernel_path.h:254-321, 540-640;
 * kernel_path_surface.h:22-125, 270-358) for the hits of one bounce, in shader order.
 *
 * What a thread keeps where: the lobes of its shading point in its arena (lobes.cuh);
 * the shadow ray it wants traced (origin, direction, contribution - 12 words) in
 * its staging column from the moment the light connection is done, so the BSDF sampling
 * that follows does not carry them in registers; slots in the next-bounce and shadow
 * queues come from one block-wide reservation at the end of the round, after which the
 * staged record is copied out to its slot. */
/* WIDE = 1 / 2: the same code as ONE block of 512 / 1024 threads per SM instead of two of
 * 256.  Measured on B200 (profiles/r02r_shade_budget_ab.txt, r02s_block_size_ab.txt):
 * where every hit runs the multiscatter random walk the kernel is bound by instruction
 * fetch (684 KB - 1.1 MB of SASS against a 32 KB instruction cache per SM; no_instruction
 * stalls 10 per issue) - the block-wide barrier of the queue append keeps the warps of one
 * block in the same stretch of code, two blocks run out of phase and fetch twice.  512
 * threads keep the 128 registers: +28 % (lean) / +32 % (full) on the startup scene, 0-3 %
 * lost where most hits are plain diffuse.  1024 threads leave 64 registers: 32 warps share
 * every fetch, another +10 % on the startup scene in spite of the spills, -15 % on the
 * Cornell box.  The host times the shapes on the first batches of a scene and keeps the
 * fastest (b200_render, "shade_wide").  A tighter register budget at 256 threads instead
 * (3 blocks at 80 registers) gains as much on the startup scene on a good run, but with a
 * large run-to-run spread, and loses 14-25 % on the Cornell box. */
#define SHADE_BLOCK_OF(wide) ((wide) == 0 ? WF_BLOCK : ((wide) == 1 ? 512 : 1024))
/* the lean GGX kernel (368 KB) keeps two blocks of 256: a wide block lost 1-3 % on every
 * scene measured (profiles/r02s_block_size_ab.txt); -DGGX_LEAN_SHAPE=1 / 2 for the A/B */
#ifndef GGX_LEAN_SHAPE
#  define GGX_LEAN_SHAPE 0
#endif
template<bool EXT, bool MS = EXT, bool PASSES = false, int WIDE = 0>
__global__ void __launch_bounds__(SHADE_BLOCK_OF(WIDE),
                                  WIDE ? 1 : (EXT ? SHADE_MIN_BLOCKS_EXT : SHADE_MIN_BLOCKS))
    k_shade_surface(PathSoA p, int num_keys)
{
  constexpr int BLOCK = SHADE_BLOCK_OF(WIDE);
  extern __shared__ float s_shade[];
  float *const stage = s_shade + threadIdx.x;
  float4 arena_words[ARENA_QUADS];
  LobeArena arena;
  arena.q = arena_words;

  WFCounters *c = p.counters;
  const unsigned int begin = p.sort_tiles ? 0u : c->offsets[1];
  const unsigned int end = p.sort_tiles ? c->n_active : c->offsets[num_keys + 1];
  const unsigned int total = end - begin;
  /* the grid-stride loop is uniform per block: block_append2 has barriers */
  const unsigned int rounds = (total + gridDim.x * blockDim.x - 1) / (gridDim.x * blockDim.x);
  for (unsigned int r = 0; r < rounds; r++) {
    const unsigned int k = r * gridDim.x * blockDim.x + blockIdx.x * blockDim.x + threadIdx.x;
    bool want_next = false, want_shadow = false;
    int i = -1;
    float4 out_ray_P = make_float4(0.0f, 0.0f, 0.0f, 0.0f), out_ray_D = out_ray_P;
    int2 qs = make_int2(-1, -1);
    if (k < total)
      qs = p.q_sorted[begin + k]; /* tile order: a miss carries the sign bit */
    if (qs.x >= 0) {
      const int qpos = qs.x;
      i = qs.y;
      const float4 r0 = p.ray_P_t[qpos];
      const float4 r1 = p.ray_D[qpos];
      const float4 tp = p.throughput[i];
      const float4 hit = p.hit[qpos];
      const int hit_object = p.hit_object[qpos];
      const float4 Lr = p.L[i];
      PathStateG st;
      state_load(p, i, st);
      st.ray_pdf = p.ray_pdf[i];
      st.ray_t = tp.w;
      f3 throughput = mk3(tp);
      f3 L = mk3(Lr);
      float ray_t = r0.w;

      ShaderDataG sd;
      /* light passes (passes.cuh): the path's accumulator block; where emission seen by
       * this segment goes (the emission pass = L, "after one bounce", or "later") */
      float *pb = PASSES ? p.pass + (size_t)i * PASS_WORDS : nullptr;
      const bool use_light_pass = PASSES && kd_int(KD_FILM_USE_LIGHT_PASS);
      const int bucket = use_light_pass ? emission_bucket(st.bounce) : -1;
      /* lamps crossed before the hit (kernel_path.h:537) - uses `sd` as scratch */
      if (bucket < 0) {
        path_lamp_emission<EXT>(st, mk3(r0), mk3(r1), hit.x, throughput, sd, L);
      }
      else {
        f3 lamp = zero3();
        path_lamp_emission<EXT>(st, mk3(r0), mk3(r1), hit.x, throughput, sd, lamp);
        pb_add3(pb, bucket, lamp);
      }

      bool alive = !path_state_ao_bounce(st); /* kernel_path.h:560-562 */
      if (alive) {
        shader_setup_from_ray(sd, __float_as_int(hit.w), hit_object, hit.x, hit.y, hit.z, mk3(r0),
                              mk3(r1));
        shader_eval_surface<EXT, MS>(sd, arena, path_depths(st), st.flag,
                                 st.rng_hash + (uint32_t)st.rng_offset +
                                     (uint32_t)st.sample * 0xb4bc3953u);
        shader_prepare_lobes<EXT>(sd, arena, st.bounce + st.transparent_bounce == 0);

        if (PASSES) /* kernel_write_data_passes - kernel_path.h:569, kernel_passes.h:174-282 */
          pass_write_data(sd, arena, st, throughput, pb);

        /* kernel_path_shader_apply - kernel_path.h:254-321 (no holdout / shadow catcher:
         * refused at bind time) */
        if (kd_float(KD_INT_FILTER_GLOSSY) != FLT_MAX) {
          const float blur_pdf = kd_float(KD_INT_FILTER_GLOSSY) * st.min_ray_pdf;
          if (blur_pdf < 1.0f)
            shader_blur_lobes(arena, sqrtf(1.0f - blur_pdf) * 0.5f);
        }
        if (sd.flag & CY_SD_EMISSION) {
          /* indirect_primitive_emission - kernel_emission.h:214-233 */
          const float cosNO = fabsf(dot(sd.Ng, sd.I));
          const float res = (cosNO > 0.0f) ? 1.0f : 0.0f;
          f3 emission = mk3(res, res, res) * sd.closure_emission_background;
          if (!(st.flag & CY_PATH_RAY_MIS_SKIP) && (sd.flag & CY_SD_USE_MIS)) {
            /* this triangle is also in the light distribution: weight the BSDF-sampled
             * hit against the pdf light sampling would have had for it */
            const float pdf = triangle_light_pdf(sd.object, sd.prim, sd.P, sd.Ng, sd.I,
                                                 sd.ray_length);
            emission *= power_heuristic(st.ray_pdf, pdf);
          }
          f3 contribution = throughput * emission;
          path_radiance_clamp(&contribution, st.bounce - 1);
          if (bucket < 0)
            L += contribution;
          else
            pb_add3(pb, bucket, contribution);
        }

        /* russian roulette - kernel_path.h:578-589 */
        const float probability = path_state_continuation_probability(st, throughput);
        if (probability == 0.0f) {
          alive = false;
        }
        else if (probability != 1.0f) {
          const float terminate = path_state_rng_1D<EXT>(st, CY_PRNG_TERMINATE);
          if (terminate >= probability)
            alive = false;
          else
            throughput /= probability;
        }
      }

      const bool denoising = PASSES && film_has_denoising();
      if (denoising && alive) /* kernel_path.h:592-595 */
        denoising_update_features(sd, arena, pb);

      if (EXT && alive && kd_int(KD_INT_USE_AMBIENT_OCCLUSION)) {
        /* kernel_path_ao (kernel_path.h:328-372): one cosine-weighted ray of length
         * ao_distance around the averaged diffuse normal; it joins the shadow queue as a
         * second entry of this path, its contribution is throughput * ao_bsdf */
        float bsdf_u, bsdf_v;
        path_state_rng_2D<EXT>(st, CY_PRNG_BSDF_U, &bsdf_u, &bsdf_v);
        const float ao_factor = kd_float(KD_BG_AO_FACTOR);
        f3 ao_bsdf = zero3(), ao_N = zero3();
        int at = 0;
        for (int n = 0; n < arena.n; n++) { /* shader_bsdf_ao, kernel_shader.h */
          const uint32_t kind = lobe_kind_at(arena, at);
          if (lobe_is_diffuse(kind)) {
            const f3 w = lobe_weight_at(arena, at);
            ao_bsdf += w * ao_factor;
            ao_N += lobe_normal_at(arena, at) * fabsf(average(w));
          }
          at += lobe_words(kind);
        }
        ao_N = is_zero(ao_N) ? sd.N : normalize(ao_N);
        f3 ao_D;
        float ao_pdf;
        sample_cos_hemisphere(ao_N, bsdf_u, bsdf_v, &ao_D, &ao_pdf);
        if (dot(sd.Ng, ao_D) > 0.0f && ao_pdf != 0.0f) {
          const f3 aP = ray_offset(sd.P, sd.Ng);
          const f3 contribution = throughput * ao_bsdf;
          const bool store_shadow = denoising && (st.flag & CY_PATH_RAY_STORE_SHADOW_INFO);
          if (store_shadow)
            pb_add3(pb, PB_PATH_TOTAL, contribution);
          /* one returning atomic per warp */
          const unsigned int m = __activemask();
          const unsigned int lane = threadIdx.x & 31u;
          unsigned int base = 0;
          if (lane == (unsigned int)(__ffs(m) - 1))
            base = atomicAdd(&c->n_shadow, (unsigned int)__popc(m));
          base = __shfl_sync(m, base, __ffs(m) - 1);
          const unsigned int slot = base + __popc(m & ((1u << lane) - 1u));
          p.q_shadow[slot] = i;
          p.sh_P_t[slot] = make_float4(aP.x, aP.y, aP.z, kd_float(KD_BG_AO_DISTANCE));
          p.sh_D[slot] = make_float4(ao_D.x, ao_D.y, ao_D.z,
                                     __uint_as_float(CY_PATH_RAY_SHADOW_OPAQUE));
          p.sh_contrib[slot] = make_float4(contribution.x, contribution.y, contribution.z, 0.0f);
          if (PASSES) {
            /* path_radiance_accum_ao (kernel_accumulate.h:342-394): the AO pass gets
             * alpha * throughput on a directly visible surface; the contribution goes to
             * direct diffuse there, to the indirect light later.  .w of the second record:
             * 2 / 3 = AO ray at / past the first surface */
            const f3 transparency = (sd.flag & CY_SD_TRANSPARENT) ?
                                        sd.closure_transparent_extinction :
                                        zero3();
            f3 alpha = mk3(1.0f, 1.0f, 1.0f) - transparency;
            alpha = mk3(fminf(fmaxf(alpha.x, 0.0f), 1.0f), fminf(fmaxf(alpha.y, 0.0f), 1.0f),
                        fminf(fmaxf(alpha.z, 0.0f), 1.0f));
            const f3 ao_pass = alpha * throughput;
            p.sh_pass[SH_PASS_QUADS * (size_t)slot] = make_float4(ao_pass.x, ao_pass.y, ao_pass.z,
                                                                  0.0f);
            p.sh_pass[SH_PASS_QUADS * (size_t)slot + 1] = make_float4(
                0.0f, 0.0f, 0.0f, st.bounce == 0 ? 2.0f : 3.0f);
            p.sh_pass[SH_PASS_QUADS * (size_t)slot + 2] =
                store_shadow ? make_float4(contribution.x, contribution.y, contribution.z, 0.0f) :
                               make_float4(0.0f, 0.0f, 0.0f, 0.0f);
          }
        }
      }

      if (alive) {
        /* direct light - kernel_path_surface.h:22-125 with one light sample */
        if (kd_int(KD_INT_USE_DIRECT_LIGHT) && (sd.flag & CY_SD_BSDF_HAS_EVAL)) {
          float light_u, light_v;
          path_state_rng_2D<EXT>(st, CY_PRNG_LIGHT_U, &light_u, &light_v);
          float terminate = 0.0f;
          if (kd_float(KD_INT_LIGHT_INV_RR_THRESHOLD) > 0.0f)
            terminate = path_state_rng_1D<EXT>(st, CY_PRNG_LIGHT_TERMINATE);
          LightSampleG ls;
          if (light_sample(light_u, light_v, sd.P, st.bounce, &ls) && ls.pdf != 0.0f) {
            /* direct_emission - kernel_emission.h:101-212 */
            f3 light_eval;
            {
              /* constant emission needs no shader run; anything else evaluates the
               * light's shader on a scratch shading point */
              f3 ce;
              if (shader_constant_emission_eval(ls.shader, &ce)) {
                if ((ls.prim != CY_PRIM_NONE) && dot(ls.Ng, -ls.D) < 0.0f)
                  ls.Ng = -ls.Ng;
                light_eval = ce * ls.eval_fac;
                if (ls.lamp != CY_LAMP_NONE)
                  light_eval *= kl_float3(light_ptr(ls.lamp), KL_STRENGTH);
              }
              else {
                ShaderDataG scratch;
                light_eval = direct_emissive_eval<EXT>(scratch, path_depths(st), &ls, -ls.D, ls.t);
              }
            }
            if (!is_zero(light_eval)) {
              /* BSDF towards the light: one colour, or - with light passes - one per BSDF
               * class (diffuse / glossy / transmission), which a lamp's ray-visibility
               * flags may zero (kernel_emission.h:146-157) */
              f3 eval = zero3();
              EvalSplit ev;
              bool ok;
              /* light that could arrive, before MIS: the denoiser's shadowing feature
               * (BsdfEval::sum_no_mis, PATH_RAY_STORE_SHADOW_INFO) */
              const bool store_shadow = denoising && (st.flag & CY_PATH_RAY_STORE_SHADOW_INFO);
              f3 no_mis = zero3();
              if (use_light_pass) {
                shader_bsdf_eval_split<EXT, MS>(sd, arena, ls.D, ls.pdf,
                                                (ls.shader & CY_SHADER_USE_MIS) != 0, ev,
                                                store_shadow ? &no_mis : nullptr);
                no_mis *= light_eval / ls.pdf;
                eval_split_mul3(ev, light_eval / ls.pdf);
                if (ls.shader & CY_SHADER_EXCLUDE_ANY) {
                  if (ls.shader & CY_SHADER_EXCLUDE_DIFFUSE)
                    ev.diffuse = zero3();
                  if (ls.shader & CY_SHADER_EXCLUDE_GLOSSY)
                    ev.glossy = zero3();
                  if (ls.shader & CY_SHADER_EXCLUDE_TRANSMIT)
                    ev.transmission = zero3();
                }
                ok = !eval_split_is_zero(ev);
                eval = eval_split_sum(ev);
              }
              else {
                eval = shader_bsdf_eval<EXT, MS>(sd, arena, ls.D, ls.pdf,
                                                 (ls.shader & CY_SHADER_USE_MIS) != 0,
                                                 store_shadow ? &no_mis : nullptr);
                no_mis *= light_eval / ls.pdf;
                eval *= light_eval / ls.pdf;
                ok = !is_zero(eval);
              }
              if (ok && kd_float(KD_INT_LIGHT_INV_RR_THRESHOLD) > 0.0f) {
                const float probability = max3(fabs3(eval)) *
                                          kd_float(KD_INT_LIGHT_INV_RR_THRESHOLD);
                if (probability < 1.0f) {
                  if (terminate >= probability)
                    ok = false;
                  else {
                    eval *= 1.0f / probability;
                    no_mis *= 1.0f / probability;
                    if (use_light_pass)
                      eval_split_mul(ev, 1.0f / probability);
                  }
                }
              }
              if (ok) {
                /* path_radiance_accum_light (kernel_accumulate.h:402-459), shadow = 1:
                 * without light passes one clamped colour; with them the clamp factor of
                 * the total scales every class, a directly visible surface keeps the
                 * classes apart (A, B, C = diffuse, glossy, transmission), a later one
                 * adds the total to the indirect light (A) */
                f3 contribution, part_b = zero3(), part_c = zero3();
                float shadow_add = 0.0f;
                const f3 light_total = throughput * no_mis; /* zero unless store_shadow */
                if (store_shadow)
                  pb_add3(pb, PB_PATH_TOTAL, light_total);
                /* with transparent shadows the clamp waits for the attenuation
                 * (shadow_light_arrives) */
                const bool defer_clamp = kd_int(KD_INT_TRANSPARENT_SHADOWS) != 0 &&
                                         (ls.shader & CY_SHADER_CAST_SHADOW) != 0;
                if (use_light_pass) {
                  f3 shaded_throughput = throughput;
                  f3 full = shaded_throughput * eval_split_sum(ev);
                  if (!defer_clamp) {
                    const float limit = (st.bounce > 0) ? kd_float(KD_INT_SAMPLE_CLAMP_INDIRECT) :
                                                          kd_float(KD_INT_SAMPLE_CLAMP_DIRECT);
                    const float sum = reduce_add(fabs3(full));
                    if (sum > limit) {
                      const float clamp_factor = limit / sum;
                      full *= clamp_factor;
                      shaded_throughput *= clamp_factor;
                    }
                  }
                  if (st.bounce == 0) {
                    contribution = shaded_throughput * ev.diffuse;
                    part_b = shaded_throughput * ev.glossy;
                    part_c = shaded_throughput * ev.transmission;
                    /* the shadow pass counts lamps only (kernel_emission.h:208) */
                    shadow_add = (ls.prim == CY_PRIM_NONE && ls.type != CY_LIGHT_BACKGROUND) ?
                                     1.0f :
                                     0.0f;
                  }
                  else {
                    contribution = full;
                  }
                }
                else {
                  contribution = throughput * eval;
                  if (!defer_clamp)
                    path_radiance_clamp(&contribution, st.bounce);
                }
                if (ls.shader & CY_SHADER_CAST_SHADOW) {
                  const bool transmit = (dot(sd.Ng, ls.D) < 0.0f);
                  const f3 sP = ray_offset(sd.P, transmit ? -sd.Ng : sd.Ng);
                  f3 sD;
                  float st_t;
                  if (ls.t == FLT_MAX) {
                    sD = ls.D;
                    st_t = ls.t;
                  }
                  else {
                    sD = ray_offset(ls.P, ls.Ng) - sP;
                    sD = normalize_len(sD, &st_t);
                  }
                  stage[0 * BLOCK] = sP.x;
                  stage[1 * BLOCK] = sP.y;
                  stage[2 * BLOCK] = sP.z;
                  stage[3 * BLOCK] = st_t;
                  stage[4 * BLOCK] = sD.x;
                  stage[5 * BLOCK] = sD.y;
                  stage[6 * BLOCK] = sD.z;
                  stage[7 * BLOCK] = contribution.x;
                  stage[8 * BLOCK] = contribution.y;
                  stage[9 * BLOCK] = contribution.z;
                  stage[10 * BLOCK] = __int_as_float(
                      st.transparent_bounce | (st.bounce > 0 ? SH_INDIRECT_FLAG : 0));
                  if (PASSES) {
                    stage[12 * BLOCK] = part_b.x;
                    stage[13 * BLOCK] = part_b.y;
                    stage[14 * BLOCK] = part_b.z;
                    stage[15 * BLOCK] = shadow_add;
                    stage[16 * BLOCK] = part_c.x;
                    stage[17 * BLOCK] = part_c.y;
                    stage[18 * BLOCK] = part_c.z;
                    stage[19 * BLOCK] = (use_light_pass && st.bounce == 0) ? 1.0f : 0.0f;
                    stage[20 * BLOCK] = light_total.x;
                    stage[21 * BLOCK] = light_total.y;
                    stage[22 * BLOCK] = light_total.z;
                  }
                  want_shadow = true;
                }
                else if (!use_light_pass) {
                  /* ray.t = 0: shadow_blocked returns false immediately */
                  L += contribution;
                  if (store_shadow)
                    pb_add3(pb, PB_PATH_TOTAL_SHADED, light_total);
                }
                else if (st.bounce == 0) {
                  pb_add3(pb, PB_DIRECT_DIFFUSE, contribution);
                  pb_add3(pb, PB_DIRECT_GLOSSY, part_b);
                  pb_add3(pb, PB_DIRECT_TRANSMISSION, part_c);
                  pb_add3(pb, PB_SHADOW, mk3(shadow_add, shadow_add, shadow_add));
                  if (store_shadow)
                    pb_add3(pb, PB_PATH_TOTAL_SHADED, light_total);
                }
                else {
                  pb_add3(pb, PB_INDIRECT, contribution);
                  if (store_shadow)
                    pb_add3(pb, PB_PATH_TOTAL_SHADED, light_total);
                }
              }
            }
          }
        }

        /* kernel_path_surface_bounce - kernel_path_surface.h:270-358 */
        if (sd.flag & CY_SD_BSDF) {
          float bsdf_u, bsdf_v;
          path_state_rng_2D<EXT>(st, CY_PRNG_BSDF_U, &bsdf_u, &bsdf_v);
          f3 bsdf_eval = zero3(), omega_in = zero3();
          float bsdf_pdf;
          EvalSplit bs;
          int label;
          bool scattered;
          if (use_light_pass) {
            label = shader_bsdf_sample_split<EXT, MS>(sd, arena, bsdf_u, bsdf_v, bs, &omega_in,
                                                      &bsdf_pdf);
            scattered = !(bsdf_pdf == 0.0f || eval_split_is_zero(bs));
          }
          else {
            label = shader_bsdf_sample<EXT, MS>(sd, arena, bsdf_u, bsdf_v, &bsdf_eval, &omega_in,
                                                &bsdf_pdf);
            scattered = !(bsdf_pdf == 0.0f || is_zero(bsdf_eval));
          }
          if (scattered) {
            /* LABEL_TRANSMIT_TRANSPARENT (closure/bsdf.h:466-475) needs transparent glass,
             * which check_scope refuses (threshold < 0 here) */
            /* path_radiance_bsdf_bounce (kernel_accumulate.h:235-268) */
            if (!use_light_pass) {
              throughput *= bsdf_eval * (1.0f / bsdf_pdf);
            }
            else if (st.bounce == 0 && !(label & CY_LABEL_TRANSPARENT)) {
              /* first bounce off a directly visible surface: the throughput per BSDF class
               * is remembered, the path goes on with their sum */
              const f3 value = throughput * (1.0f / bsdf_pdf);
              const f3 sd_ = bs.diffuse * value, sg_ = bs.glossy * value,
                       st_ = bs.transmission * value;
              throughput = sd_ + sg_ + st_ + zero3();
              pb_set3(pb, PB_STATE_DIFFUSE, sd_);
              pb_set3(pb, PB_STATE_GLOSSY, sg_);
              pb_set3(pb, PB_STATE_TRANSMISSION, st_);
              pb_set3(pb, PB_STATE_DIRECT, throughput);
            }
            else {
              throughput *= (eval_split_sum(bs) + bs.transparent) * (1.0f / bsdf_pdf);
            }
            if (!(label & CY_LABEL_TRANSPARENT)) {
              st.ray_pdf = bsdf_pdf;
              st.ray_t = 0.0f;
              st.min_ray_pdf = fminf(bsdf_pdf, st.min_ray_pdf);
            }
            path_state_next(st, label);
            /* kernel_path_state.h:164-168: once the feature is closed nothing more is
             * counted for the shadowing ratio - decided at the end of path_state_next,
             * which a transparent bounce leaves early (:84-99) */
            if (denoising && !(label & CY_LABEL_TRANSPARENT) && pb[PB_DN_WEIGHT] == 0.0f)
              st.flag &= ~CY_PATH_RAY_STORE_SHADOW_INFO;
            const f3 nP = ray_offset(sd.P, (label & CY_LABEL_TRANSMIT) ? -sd.Ng : sd.Ng);
            const f3 nD = normalize(omega_in);
            if (st.bounce == 0)
              ray_t -= sd.ray_length;
            else
              ray_t = FLT_MAX;
            want_next = true;
            if (p.debug && i == p.debug_slot && st.bounce + st.transparent_bounce >= 1 &&
                st.bounce + st.transparent_bounce <= 16) {
              float *dbg = p.debug + 32 * (st.bounce + st.transparent_bounce - 1);
              dbg[0] = r0.x, dbg[1] = r0.y, dbg[2] = r0.z, dbg[3] = r0.w;
              dbg[4] = r1.x, dbg[5] = r1.y, dbg[6] = r1.z, dbg[7] = hit.x;
              dbg[8] = (float)__float_as_int(hit.w), dbg[9] = (float)hit_object, dbg[10] = sd.P.x;
              dbg[11] = sd.P.y, dbg[12] = sd.P.z, dbg[13] = sd.N.x, dbg[14] = sd.N.y;
              dbg[15] = sd.N.z, dbg[16] = (float)(sd.flag & 0xffff), dbg[17] = (float)arena.n;
              dbg[18] = (float)lobe_id(lobe_kind_at(arena, 0)),
              dbg[19] = lobe_sample_weight_at(arena, 0);
              dbg[20] = 0.0f, dbg[21] = 0.0f;
              dbg[22] = (float)label, dbg[23] = bsdf_pdf, dbg[24] = omega_in.x;
              dbg[25] = omega_in.y, dbg[26] = omega_in.z, dbg[27] = throughput.x;
              dbg[28] = throughput.y, dbg[29] = throughput.z, dbg[30] = bsdf_u, dbg[31] = bsdf_v;
            }
            /* visibility and AO-bounce clipping of the NEXT segment, decided here so the
             * traversal kernel needs nothing but the 32-byte ray (kernel_path.h:66-71) */
            uint32_t vis = path_state_ray_visibility(st.flag);
            if (path_state_ao_bounce(st)) {
              vis = CY_PATH_RAY_SHADOW;
              ray_t = kd_float(KD_BG_AO_DISTANCE);
            }
            out_ray_P = make_float4(nP.x, nP.y, nP.z, ray_t);
            out_ray_D = make_float4(nD.x, nD.y, nD.z, __uint_as_float(vis));
            p.ray_pdf[i] = st.ray_pdf;
            p.throughput[i] = make_float4(throughput.x, throughput.y, throughput.z, st.ray_t);
            state_store(p, i, st);
          }
        }
      }
      p.L[i] = make_float4(L.x, L.y, L.z, Lr.w);
    }
    unsigned int s_next, s_sh;
    block_append2<BLOCK / 32>(&c->n_next, want_next, &c->n_shadow, want_shadow, &s_next, &s_sh);
    if (want_next) {
      p.q_next[s_next] = i;
      p.nray_P_t[s_next] = out_ray_P;
      p.nray_D[s_next] = out_ray_D;
    }
    if (want_shadow) {
      p.q_shadow[s_sh] = i;
      p.sh_P_t[s_sh] = make_float4(stage[0 * BLOCK], stage[1 * BLOCK], stage[2 * BLOCK],
                                   stage[3 * BLOCK]);
      p.sh_D[s_sh] = make_float4(stage[4 * BLOCK], stage[5 * BLOCK], stage[6 * BLOCK],
                                 __uint_as_float(CY_PATH_RAY_SHADOW_OPAQUE));
      p.sh_contrib[s_sh] = make_float4(stage[7 * BLOCK], stage[8 * BLOCK],
                                       stage[9 * BLOCK], stage[10 * BLOCK]);
      if (PASSES) {
        float4 *sp = p.sh_pass + SH_PASS_QUADS * (size_t)s_sh;
        sp[0] = make_float4(stage[12 * BLOCK], stage[13 * BLOCK], stage[14 * BLOCK],
                            stage[15 * BLOCK]);
        sp[1] = make_float4(stage[16 * BLOCK], stage[17 * BLOCK], stage[18 * BLOCK],
                            stage[19 * BLOCK]);
        sp[2] = make_float4(stage[20 * BLOCK], stage[21 * BLOCK], stage[22 * BLOCK],
                            0.0f);
      }
    }
  }
}

/* -------------------------------------- intersect_shadow + shade_shadow */

/* AO: the light ray and the ambient-occlusion ray of one path are in the same launch,
 * so unoccluded contributions are added atomically (a compile-time variant: the test
 * at run time cost the plain kernel 7 %) */
/* Scenes with transparent shadows: a light contribution is clamped when it ARRIVES, after
 * the attenuation of the surfaces it crossed (path_radiance_accum_light clamps
 * throughput * shadow * eval, kernel_accumulate.h:402-459), so the shading kernel hands it
 * over unclamped and marks which limit applies. */
template<bool PASSES>
CY_DEV void shadow_light_arrives(const PathSoA &p, unsigned int sh, f3 shadow)
{
  const int i = p.q_shadow[sh];
  const float4 cn = p.sh_contrib[sh];
  const float limit = (__float_as_int(cn.w) & SH_INDIRECT_FLAG) ?
                          kd_float(KD_INT_SAMPLE_CLAMP_INDIRECT) :
                          kd_float(KD_INT_SAMPLE_CLAMP_DIRECT);
  if (PASSES && film_has_denoising()) {
    /* path_total_shaded += shadow * light (kernel_accumulate.h:412-416) */
    const f3 light = mk3(p.sh_pass[SH_PASS_QUADS * (size_t)sh + 2]);
    pb_add3(p.pass + (size_t)i * PASS_WORDS, PB_PATH_TOTAL_SHADED, shadow * light);
  }
  if (PASSES && kd_int(KD_FILM_USE_LIGHT_PASS)) {
    float *pb = p.pass + (size_t)i * PASS_WORDS;
    const float4 b = p.sh_pass[SH_PASS_QUADS * (size_t)sh], cc = p.sh_pass[SH_PASS_QUADS * (size_t)sh + 1];
    if (cc.w != 0.0f) {
      f3 A = mk3(cn) * shadow, B = mk3(b) * shadow, C = mk3(cc) * shadow;
      const float sum = reduce_add(fabs3(A + B + C));
      if (sum > limit) {
        const float f = limit / sum;
        A *= f;
        B *= f;
        C *= f;
      }
      pb_add3(pb, PB_DIRECT_DIFFUSE, A);
      pb_add3(pb, PB_DIRECT_GLOSSY, B);
      pb_add3(pb, PB_DIRECT_TRANSMISSION, C);
      pb_add3(pb, PB_SHADOW, shadow * b.w);
    }
    else {
      f3 full = mk3(cn) * shadow;
      const float sum = reduce_add(fabs3(full));
      if (sum > limit)
        full *= limit / sum;
      pb_add3(pb, PB_INDIRECT, full);
    }
  }
  else {
    f3 c = mk3(cn) * shadow;
    const float sum = reduce_add(fabs3(c));
    if (sum > limit)
      c *= limit / sum;
    float4 L = p.L[i];
    L.x += c.x;
    L.y += c.y;
    L.z += c.z;
    p.L[i] = L;
  }
}

template<bool TRANSPARENT, bool AO = false, bool PASSES = false> struct ShadowJob {
  static constexpr bool QUEUE_RAYS = true;
  PathSoA p;
  __device__ __forceinline__ const float4 *ray_P(unsigned int qi) const
  {
    return p.sh_P_t + qi;
  }
  __device__ __forceinline__ const float4 *ray_D(unsigned int qi) const
  {
    return p.sh_D + qi; /* .w = PATH_RAY_SHADOW_OPAQUE: shadow_blocked_opaque, kernel_shadow.h:90 */
  }
  __device__ __forceinline__ void store(unsigned int qi, const TraceHit &h, bool blocked)
  {
    if (!blocked && TRANSPARENT) {
      shadow_light_arrives<PASSES>(p, qi, mk3(1.0f, 1.0f, 1.0f));
    }
    else if (!blocked) {
      /* shade_shadow - path_radiance_accum_light, kernel_accumulate.h:402-459 */
      const int i = p.q_shadow[qi];
      const float4 cn = p.sh_contrib[qi];
      if (PASSES && film_has_denoising()) {
        /* the denoiser's shadowing feature: the light arrived (shadow = 1, or ao = 1) */
        float *pb = p.pass + (size_t)i * PASS_WORDS;
        const f3 light = mk3(p.sh_pass[SH_PASS_QUADS * (size_t)qi + 2]);
        if (AO)
          pb_atomic_add3(pb, PB_PATH_TOTAL_SHADED, light);
        else
          pb_add3(pb, PB_PATH_TOTAL_SHADED, light);
      }
      if (PASSES && AO) {
        /* light and AO rays of a path share the launch: atomic adds.  Without light
         * passes everything is one colour in L. */
        float *pb = p.pass + (size_t)i * PASS_WORDS;
        const float4 b = p.sh_pass[SH_PASS_QUADS * (size_t)qi], cc = p.sh_pass[SH_PASS_QUADS * (size_t)qi + 1];
        if (!kd_int(KD_FILM_USE_LIGHT_PASS)) {
          float *L = (float *)&p.L[i];
          atomicAdd(L + 0, cn.x);
          atomicAdd(L + 1, cn.y);
          atomicAdd(L + 2, cn.z);
        }
        else if (cc.w == 1.0f) {
          pb_atomic_add3(pb, PB_DIRECT_DIFFUSE, mk3(cn));
          pb_atomic_add3(pb, PB_DIRECT_GLOSSY, mk3(b));
          pb_atomic_add3(pb, PB_DIRECT_TRANSMISSION, mk3(cc));
          pb_atomic_add3(pb, PB_SHADOW, mk3(b.w, b.w, b.w));
        }
        else if (cc.w == 2.0f) {
          pb_atomic_add3(pb, PB_AO, mk3(b));
          pb_atomic_add3(pb, PB_DIRECT_DIFFUSE, mk3(cn));
        }
        else {
          pb_atomic_add3(pb, PB_INDIRECT, mk3(cn)); /* indirect light, or AO past bounce 0 */
        }
      }
      else if (PASSES && kd_int(KD_FILM_USE_LIGHT_PASS)) {
        /* light passes: a directly visible surface adds per BSDF class (and counts the
         * lamp in the shadow pass), a later one adds to the indirect light */
        float *pb = p.pass + (size_t)i * PASS_WORDS;
        const float4 b = p.sh_pass[SH_PASS_QUADS * (size_t)qi], cc = p.sh_pass[SH_PASS_QUADS * (size_t)qi + 1];
        if (cc.w != 0.0f) {
          pb_add3(pb, PB_DIRECT_DIFFUSE, mk3(cn));
          pb_add3(pb, PB_DIRECT_GLOSSY, mk3(b));
          pb_add3(pb, PB_DIRECT_TRANSMISSION, mk3(cc));
          pb_add3(pb, PB_SHADOW, mk3(b.w, b.w, b.w));
        }
        else {
          pb_add3(pb, PB_INDIRECT, mk3(cn));
        }
      }
      else if (AO) {
        float *L = (float *)&p.L[i];
        atomicAdd(L + 0, cn.x);
        atomicAdd(L + 1, cn.y);
        atomicAdd(L + 2, cn.z);
      }
      else {
        float4 L = p.L[i];
        L.x += cn.x;
        L.y += cn.y;
        L.z += cn.z;
        p.L[i] = L;
      }
    }
    else if (TRANSPARENT) {
      /* shadow_blocked_transparent_stepped (kernel_shadow.h:354-368): the first hit found
       * is transparent -> walk the ray surface by surface; opaque -> blocked */
      const unsigned int tri = __ldg(&g_scene.prim_index[h.prim]);
      const uint32_t flags = shader_flags((int)__ldg(&g_scene.tri_shader[tri]));
      const float4 cn = p.sh_contrib[qi];
      const int bounce = __float_as_int(cn.w) & ~SH_INDIRECT_FLAG;
      if ((flags & CY_SD_HAS_TRANSPARENT_SHADOW) &&
          bounce < kd_int(KD_INT_TRANSPARENT_MAX_BOUNCE)) {
        const unsigned int slot = atomicAdd(&p.counters->n_ts[0], 1u);
        const float4 d = p.sh_D[qi];
        p.ts_P_t[0][slot] = p.sh_P_t[qi];
        p.ts_D[0][slot] = make_float4(d.x, d.y, d.z,
                                      __uint_as_float(CY_PATH_RAY_SHADOW_TRANSPARENT));
        p.ts_idx[0][slot] = (int)qi;
        p.ts_thr[qi] = make_float4(1.0f, 1.0f, 1.0f, __int_as_float(bounce));
      }
    }
  }
};

/* closest-hit traversal of one stepping queue; hits land in the (idle) hit arrays */
struct TransparentShadowJob {
  static constexpr bool QUEUE_RAYS = true;
  PathSoA p;
  int cur;
  __device__ __forceinline__ const float4 *ray_P(unsigned int qi) const
  {
    return p.ts_P_t[cur] + qi;
  }
  __device__ __forceinline__ const float4 *ray_D(unsigned int qi) const
  {
    return p.ts_D[cur] + qi;
  }
  __device__ __forceinline__ void store(unsigned int qi, const TraceHit &h, bool)
  {
    p.hit[qi] = make_float4(h.t, h.u, h.v, __int_as_float(h.prim));
    p.hit_object[qi] = h.object;
  }
};

__global__ void __launch_bounds__(TRACE_BLOCK, TRACE_MIN_BLOCKS)
    k_intersect_shadow_step(PathSoA p, int cur, int refill_threshold)
{
  TraceCounters cnt;
  cnt.nodes = cnt.tris = cnt.instances = 0;
  TransparentShadowJob job;
  job.p = p;
  job.cur = cur;
  trace_persistent<false, false>(job, p.counters->n_ts[cur], &p.counters->work_ts,
                                 refill_threshold, cnt);
}

/* One step of the transparent-shadow walk for every ray of queue `cur`
 * (shadow_blocked_transparent_stepped_loop + shadow_handle_transparent_isect,
 * kernel_shadow.h:46-88, 300-352): nothing hit -> the light arrives, attenuated; an
 * opaque surface -> blocked; a transparent one -> evaluate its shader as a shadow ray,
 * multiply the attenuation by its transparency and continue behind it. */
template<bool EXT, bool PASSES = false>
__global__ void __launch_bounds__(WF_BLOCK) k_shade_shadow_step(PathSoA p, int cur)
{
  WFCounters *c = p.counters;
  const unsigned int n = c->n_ts[cur];
  const int max_bounce = kd_int(KD_INT_TRANSPARENT_MAX_BOUNCE);
  const unsigned int rounds = (n + gridDim.x * blockDim.x - 1) / (gridDim.x * blockDim.x);
  for (unsigned int r = 0; r < rounds; r++) {
    const unsigned int qi = r * gridDim.x * blockDim.x + blockIdx.x * blockDim.x + threadIdx.x;
    bool go_on = false;
    float4 out_P = make_float4(0.0f, 0.0f, 0.0f, 0.0f), out_D = out_P;
    int sh = 0;
    if (qi < n) {
      sh = p.ts_idx[cur][qi];
      const float4 hit = p.hit[qi];
      const int prim = __float_as_int(hit.w);
      float4 thr = p.ts_thr[sh];
      if (prim < 0) {
        /* reached the light */
        shadow_light_arrives<PASSES>(p, (unsigned int)sh, mk3(thr.x, thr.y, thr.z));
      }
      else {
        const unsigned int tri = __ldg(&g_scene.prim_index[prim]);
        const uint32_t flags = shader_flags((int)__ldg(&g_scene.tri_shader[tri]));
        if (flags & CY_SD_HAS_TRANSPARENT_SHADOW) {
          const float4 r0 = p.ts_P_t[cur][qi];
          const float4 r1 = p.ts_D[cur][qi];
          const f3 rayP = mk3(r0), rayD = mk3(r1);
          int bounce = __float_as_int(thr.w);
          ShaderDataG sd;
          shader_setup_from_ray(sd, prim, p.hit_object[qi], hit.x, hit.y, hit.z, rayP, rayD);
          /* path_state_modify_bounce(state, true) around the evaluation; only the
           * transparent counter is known here, which is what a shadow shader can vary by */
          PathDepths depths;
          depths.bounce = 1;
          depths.diffuse = depths.glossy = depths.transmission = 0;
          depths.transparent = (short)bounce;
          shader_eval_emission<EXT>(sd, depths, CY_PATH_RAY_SHADOW);
          f3 t = (sd.flag & CY_SD_TRANSPARENT) ? sd.closure_transparent_extinction : zero3();
          f3 nthr = mk3(thr.x, thr.y, thr.z) * t;
          if (!is_zero(nthr)) {
            bounce++;
            if (bounce < max_bounce) {
              /* move the ray behind the surface, keep aiming at the same end point */
              const float4 s0 = p.sh_P_t[sh];
              const float4 s1 = p.sh_D[sh];
              f3 nP = ray_offset(sd.P, -sd.Ng);
              f3 nD = rayD;
              float nt = r0.w;
              if (nt != FLT_MAX) {
                const f3 Pend = mk3(s0) + mk3(s1) * s0.w;
                nD = normalize_len(Pend - nP, &nt);
              }
              out_P = make_float4(nP.x, nP.y, nP.z, nt);
              out_D = make_float4(nD.x, nD.y, nD.z,
                                  __uint_as_float(CY_PATH_RAY_SHADOW_TRANSPARENT));
              p.ts_thr[sh] = make_float4(nthr.x, nthr.y, nthr.z, __int_as_float(bounce));
              go_on = true;
            }
          }
        }
      }
    }
    unsigned int slot, unused;
    block_append2(&c->n_ts[cur ^ 1], go_on, &c->n_ts[cur ^ 1], false, &slot, &unused);
    if (go_on) {
      p.ts_P_t[cur ^ 1][slot] = out_P;
      p.ts_D[cur ^ 1][slot] = out_D;
      p.ts_idx[cur ^ 1][slot] = sh;
    }
  }
}

/* between two steps: the consumed queue becomes the empty target of the next step */
__global__ void k_shadow_step_end(PathSoA p, int cur)
{
  WFCounters *c = p.counters;
  c->shadow_rays += c->n_ts[cur];
  c->n_ts[cur] = 0;
  c->work_ts = 0;
}

template<bool COUNT, bool TRANSPARENT, bool AO = false, bool PASSES = false>
__global__ void __launch_bounds__(TRACE_BLOCK, TRACE_MIN_BLOCKS)
    k_intersect_shadow(PathSoA p, int refill_threshold)
{
  const unsigned lane = threadIdx.x & 31u;
  TraceCounters cnt;
  cnt.nodes = cnt.tris = cnt.instances = 0;
  ShadowJob<TRANSPARENT, AO, PASSES> job;
  job.p = p;
  trace_persistent<true, COUNT>(job, p.counters->n_shadow, &p.counters->work_shadow,
                                refill_threshold, cnt);
  if (COUNT) {
    for (int o = 16; o > 0; o >>= 1) {
      cnt.nodes += __shfl_xor_sync(0xffffffffu, cnt.nodes, o);
      cnt.tris += __shfl_xor_sync(0xffffffffu, cnt.tris, o);
      cnt.instances += __shfl_xor_sync(0xffffffffu, cnt.instances, o);
    }
    if (lane == 0) {
      atomicAdd(&p.counters->sh_nodes, (unsigned long long)cnt.nodes);
      atomicAdd(&p.counters->sh_tris, (unsigned long long)cnt.tris);
      atomicAdd(&p.counters->sh_instances, (unsigned long long)cnt.instances);
    }
  }
}

/* per-iteration bookkeeping: q_next becomes q_active (pointers are swapped on
 * the host), counters roll over */
__global__ void k_iteration_end(PathSoA p, int num_keys, int iteration)
{
  WFCounters *c = p.counters;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    /* the first segment of every path is the camera ray (PATH_RAY_CAMERA) */
    if (iteration == 0)
      c->primary_rays += c->n_active;
    else
      c->bounce_rays += c->n_active;
    c->shadow_rays += c->n_shadow;
    c->n_active = c->n_next;
    c->n_next = 0;
    c->n_shadow = 0;
    c->work_closest = 0;
    c->work_shadow = 0;
    c->n_ts[0] = c->n_ts[1] = 0;
    c->work_ts = 0;
    c->flags = (g_svm_scope_miss ? WF_FLAG_SCOPE_MISS : 0u) |
               (g_trace_overflow ? WF_FLAG_TRACE_OVERFLOW : 0u);
  }
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k <= num_keys; k += gridDim.x * blockDim.x)
    c->hist[k] = 0;
}

/* ---------------------------------------------------------- film write */

/* kernel_write_result / kernel_write_pass_float4 (kernel_passes.h:338-350,
 * kernel_write_passes.h:49-65) for the whole batch: one thread owns a pixel,
 * sums its samples in sample order (the order the CPU adds them) and does one
 * float4 read-modify-write - no atomics. */
__global__ void __launch_bounds__(WF_BLOCK)
    k_film_accumulate(PathSoA p, BatchParams bp, float *film, int pass_stride, int pass_combined)
{
  const unsigned int npix = (unsigned)bp.w * (unsigned)bp.h;
  for (unsigned int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < npix;
       pix += gridDim.x * blockDim.x) {
    int x, y;
    batch_pixel(bp, pix, &x, &y);
    const long long index = (long long)bp.offset + x + (long long)y * bp.stride;
    float4 *dst = (float4 *)(film + index * pass_stride + pass_combined);
    float4 acc = *dst;
    for (int s = 0; s < bp.nsamples; s++) {
      const float4 L = p.L[(size_t)s * npix + pix];
      /* path_radiance_clamp_and_sum - kernel_accumulate.h:622-688 */
      float3 Ls = make_float3(L.x, L.y, L.z);
      const float sum = fabsf(Ls.x) + fabsf(Ls.y) + fabsf(Ls.z);
      if (!isfinite_safe(sum))
        Ls = make_float3(0.0f, 0.0f, 0.0f);
      const float alpha = 1.0f - L.w;
      acc.x += Ls.x;
      acc.y += Ls.y;
      acc.z += Ls.z;
      acc.w += alpha;
    }
    *dst = acc;
  }
}

/* kernel_write_result with adaptive sampling (kernel_passes.h:338-350, 392-425): the
 * combined pass, twice the even samples into the aux buffer, and the sample count kept
 * negative while the tile is in progress. */
__global__ void __launch_bounds__(WF_BLOCK)
    k_film_accumulate_adaptive(PathSoA p, BatchParams bp, float *film, int pass_stride, int aux,
                               int sample_count)
{
  const unsigned int npix = (unsigned)bp.w * (unsigned)bp.h;
  const int pattern = kd_int(KD_INT_SAMPLING_PATTERN);
  const bool write_aux = kd_float(KD_INT_ADAPTIVE_THRESHOLD) > 0.0f;
  for (unsigned int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < npix;
       pix += gridDim.x * blockDim.x) {
    int x, y;
    batch_pixel(bp, pix, &x, &y);
    const long long index = (long long)bp.offset + x + (long long)y * bp.stride;
    float *buf = film + index * pass_stride;
    float4 acc = *(float4 *)buf;
    float4 a = *(float4 *)(buf + aux);
    float count = sample_count ? buf[sample_count] : 0.0f;
    for (int s = 0; s < bp.nsamples; s++) {
      const size_t path = (size_t)s * npix + pix;
      if (p.ray_pdf[path] < 0.0f)
        continue; /* stopped pixel, or no camera ray: nothing is written */
      const float4 L = p.L[path];
      float3 Ls = make_float3(L.x, L.y, L.z);
      const float sum = fabsf(Ls.x) + fabsf(Ls.y) + fabsf(Ls.z);
      if (!isfinite_safe(sum))
        Ls = make_float3(0.0f, 0.0f, 0.0f);
      acc.x += Ls.x;
      acc.y += Ls.y;
      acc.z += Ls.z;
      acc.w += 1.0f - L.w;
      if (write_aux && sample_is_even(pattern, bp.sample0 + s)) {
        a.x += Ls.x * 2.0f;
        a.y += Ls.y * 2.0f;
        a.z += Ls.z * 2.0f;
      }
      count = -fabsf(count) - 1.0f;
    }
    *(float4 *)buf = acc;
    *(float4 *)(buf + aux) = a;
    if (sample_count)
      buf[sample_count] = count;
  }
}

/* kernel_write_result + kernel_write_light_passes + the data passes
 * (kernel_passes.h:285-350, 174-224; path_radiance_clamp_and_sum,
 * kernel_accumulate.h:537-560, 640-700) for the whole batch when the film holds more than
 * the combined pass: one thread owns a pixel and folds its samples into every pass in
 * sample order. */
__global__ void __launch_bounds__(WF_BLOCK)
    k_film_accumulate_passes(PathSoA p, BatchParams bp, float *film, int pass_stride)
{
  const unsigned int npix = (unsigned)bp.w * (unsigned)bp.h;
  const int flag = kd_int(KD_FILM_PASS_FLAG), light_flag = kd_int(KD_FILM_LIGHT_PASS_FLAG);
  const bool use_light_pass = kd_int(KD_FILM_USE_LIGHT_PASS) != 0;
  auto add3 = [](float *dst, f3 v) {
    dst[0] += v.x;
    dst[1] += v.y;
    dst[2] += v.z;
  };
  auto lp = [&](int pass_type) { return (light_flag & (1 << (pass_type % 32))) != 0; };
  for (unsigned int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < npix;
       pix += gridDim.x * blockDim.x) {
    int x, y;
    batch_pixel(bp, pix, &x, &y);
    const long long index = (long long)bp.offset + x + (long long)y * bp.stride;
    float *buf = film + index * pass_stride;
    for (int s = 0; s < bp.nsamples; s++) {
      /* a pixel sample whose camera ray was never traced (t = 0, outside the lens of a
       * panoramic camera) writes nothing at all: kernel_path_trace returns early */
      const size_t path = (size_t)s * npix + pix;
      const float4 Lr = p.L[path];
      const float *pb = p.pass + path * PASS_WORDS;
      if (pb[PB_UNTRACED] != 0.0f)
        continue;
      const f3 emission = mk3(Lr);
      const float transparent = Lr.w;
      f3 L_sum;
      f3 dd = zero3(), dg = zero3(), dt = zero3(), id = zero3(), ig = zero3(), it = zero3();
      if (use_light_pass) {
        const f3 s_d = pb_get3(pb, PB_STATE_DIFFUSE), s_g = pb_get3(pb, PB_STATE_GLOSSY),
                 s_t = pb_get3(pb, PB_STATE_TRANSMISSION), s_all = pb_get3(pb, PB_STATE_DIRECT);
        /* path_radiance_sum_indirect: light that arrived after the first bounce is divided
         * by that bounce's total throughput and handed to the classes in proportion */
        const f3 de = safe_divide_color(pb_get3(pb, PB_DIRECT_EMISSION), s_all);
        dd = pb_get3(pb, PB_DIRECT_DIFFUSE) + s_d * de;
        dg = pb_get3(pb, PB_DIRECT_GLOSSY) + s_g * de;
        dt = pb_get3(pb, PB_DIRECT_TRANSMISSION) + s_t * de;
        const f3 ind = safe_divide_color(pb_get3(pb, PB_INDIRECT), s_all);
        id = s_d * ind;
        ig = s_g * ind;
        it = s_t * ind;
        f3 L_direct = dd + dg + dt + zero3() + emission;
        const f3 L_indirect = id + ig + it + zero3();
        if (!kd_int(KD_BG_TRANSPARENT))
          L_direct += pb_get3(pb, PB_BACKGROUND);
        L_sum = L_direct + L_indirect;
      }
      else {
        L_sum = emission;
      }
      const float sum = fabsf(L_sum.x) + fabsf(L_sum.y) + fabsf(L_sum.z);
      bool finite = isfinite_safe(sum);
      if (!finite) {
        L_sum = zero3();
        dd = dg = dt = id = ig = it = zero3();
      }
      const float alpha = 1.0f - transparent;
      if (flag & (1 << CY_PASS_COMBINED)) {
        float *dst = buf + kd_int(KD_FILM_PASS_COMBINED);
        dst[0] += L_sum.x;
        dst[1] += L_sum.y;
        dst[2] += L_sum.z;
        dst[3] += alpha;
      }
      if (use_light_pass) {
        if (lp(CY_PASS_DIFFUSE_INDIRECT))
          add3(buf + kd_int(KD_FILM_PASS_DIFFUSE_INDIRECT), id);
        if (lp(CY_PASS_GLOSSY_INDIRECT))
          add3(buf + kd_int(KD_FILM_PASS_GLOSSY_INDIRECT), ig);
        if (lp(CY_PASS_TRANSMISSION_INDIRECT))
          add3(buf + kd_int(KD_FILM_PASS_TRANSMISSION_INDIRECT), it);
        if (lp(CY_PASS_DIFFUSE_DIRECT))
          add3(buf + kd_int(KD_FILM_PASS_DIFFUSE_DIRECT), dd);
        if (lp(CY_PASS_GLOSSY_DIRECT))
          add3(buf + kd_int(KD_FILM_PASS_GLOSSY_DIRECT), dg);
        if (lp(CY_PASS_TRANSMISSION_DIRECT))
          add3(buf + kd_int(KD_FILM_PASS_TRANSMISSION_DIRECT), dt);
        if (lp(CY_PASS_EMISSION))
          add3(buf + kd_int(KD_FILM_PASS_EMISSION), finite ? emission : zero3());
        if (lp(CY_PASS_BACKGROUND))
          add3(buf + kd_int(KD_FILM_PASS_BACKGROUND), pb_get3(pb, PB_BACKGROUND));
        if (lp(CY_PASS_AO))
          add3(buf + kd_int(KD_FILM_PASS_AO), pb_get3(pb, PB_AO));
        if (lp(CY_PASS_DIFFUSE_COLOR))
          add3(buf + kd_int(KD_FILM_PASS_DIFFUSE_COLOR), pb_get3(pb, PB_COLOR_DIFFUSE));
        if (lp(CY_PASS_GLOSSY_COLOR))
          add3(buf + kd_int(KD_FILM_PASS_GLOSSY_COLOR), pb_get3(pb, PB_COLOR_GLOSSY));
        if (lp(CY_PASS_TRANSMISSION_COLOR))
          add3(buf + kd_int(KD_FILM_PASS_TRANSMISSION_COLOR), pb_get3(pb, PB_COLOR_TRANSMISSION));
        if (lp(CY_PASS_SHADOW)) {
          float *dst = buf + kd_int(KD_FILM_PASS_SHADOW);
          add3(dst, pb_get3(pb, PB_SHADOW));
          dst[3] += kd_float(KD_FILM_PASS_SHADOW_SCALE);
        }
        if (lp(CY_PASS_MIST))
          buf[kd_int(KD_FILM_PASS_MIST)] += 1.0f - pb[PB_MIST];
      }
      if (film_has_denoising()) {
        /* kernel_write_result, kernel_passes.h:355-389 */
        float *dn = buf + kd_int(KD_FILM_PASS_DENOISING_DATA);
        {
          /* kernel_write_denoising_shadow: even and odd samples in two buffers */
          float *sh = dn + (sample_is_even(kd_int(KD_INT_SAMPLING_PATTERN), bp.sample0 + s) ?
                                CY_DENOISING_PASS_SHADOW_B :
                                CY_DENOISING_PASS_SHADOW_A);
          const float path_total = ensure_finite(average(pb_get3(pb, PB_PATH_TOTAL)));
          const float path_total_shaded = ensure_finite(
              average(pb_get3(pb, PB_PATH_TOTAL_SHADED)));
          sh[0] += path_total;
          sh[1] += path_total_shaded;
          const float value = path_total_shaded / fmaxf(path_total, 1e-7f);
          sh[2] += value * value;
        }
        f3 noisy;
        if (kd_int(KD_FILM_PASS_DENOISING_CLEAN)) {
          /* path_radiance_split_denoising, kernel_accumulate.h:690-727 */
          const int cf = kd_int(KD_FILM_DENOISING_FLAGS);
          f3 clean = (finite ? emission : zero3()) + pb_get3(pb, PB_BACKGROUND);
          noisy = zero3() + zero3();
          if (cf & CY_DENOISING_CLEAN_DIFFUSE_DIR) clean += dd; else noisy += dd;
          if (cf & CY_DENOISING_CLEAN_DIFFUSE_IND) clean += id; else noisy += id;
          if (cf & CY_DENOISING_CLEAN_GLOSSY_DIR) clean += dg; else noisy += dg;
          if (cf & CY_DENOISING_CLEAN_GLOSSY_IND) clean += ig; else noisy += ig;
          if (cf & CY_DENOISING_CLEAN_TRANSMISSION_DIR) clean += dt; else noisy += dt;
          if (cf & CY_DENOISING_CLEAN_TRANSMISSION_IND) clean += it; else noisy += it;
          noisy = ensure_finite3(noisy);
          add3(buf + kd_int(KD_FILM_PASS_DENOISING_CLEAN), ensure_finite3(clean));
        }
        else {
          noisy = ensure_finite3(L_sum);
        }
        add3(dn + CY_DENOISING_PASS_COLOR, noisy);
        add3(dn + CY_DENOISING_PASS_COLOR_VAR, noisy * noisy);
        const f3 dnn = pb_get3(pb, PB_DN_NORMAL), dna = pb_get3(pb, PB_DN_ALBEDO);
        add3(dn + CY_DENOISING_PASS_NORMAL, dnn);
        add3(dn + CY_DENOISING_PASS_NORMAL_VAR, dnn * dnn);
        add3(dn + CY_DENOISING_PASS_ALBEDO, dna);
        add3(dn + CY_DENOISING_PASS_ALBEDO_VAR, dna * dna);
        dn[CY_DENOISING_PASS_DEPTH] += pb[PB_DN_DEPTH];
        dn[CY_DENOISING_PASS_DEPTH_VAR] += pb[PB_DN_DEPTH] * pb[PB_DN_DEPTH];
      }
      if (pb[PB_HAS_DATA] != 0.0f) {
        if (bp.sample0 + s == 0) {
          if (flag & (1 << CY_PASS_DEPTH))
            buf[kd_int(KD_FILM_PASS_DEPTH)] += pb[PB_DEPTH];
          if (flag & (1 << CY_PASS_OBJECT_ID))
            buf[kd_int(KD_FILM_PASS_OBJECT_ID)] += pb[PB_OBJECT_ID];
          if (flag & (1 << CY_PASS_MATERIAL_ID))
            buf[kd_int(KD_FILM_PASS_MATERIAL_ID)] += pb[PB_MATERIAL_ID];
        }
        if (flag & (1 << CY_PASS_NORMAL))
          add3(buf + kd_int(KD_FILM_PASS_NORMAL), pb_get3(pb, PB_NORMAL));
        if (flag & (1 << CY_PASS_UV))
          add3(buf + kd_int(KD_FILM_PASS_UV), pb_get3(pb, PB_UV));
      }
    }
  }
}

/* kernel_film.h:90-130 - float film -> display bytes / halfs */
CY_DEV float color_linear_to_srgb(float c)
{
  if (c < 0.0031308f)
    return (c < 0.0f) ? 0.0f : c * 12.92f;
  else
    return 1.055f * powf(c, 1.0f / 2.4f) - 0.055f;
}
struct FilmDisplay {
  int pass_stride;
  int display_pass_stride;   /* float offset of the displayed pass inside a pixel */
  int use_pass_alpha;        /* KernelFilm::use_display_pass_alpha */
  int use_exposure;          /* KernelFilm::use_display_exposure */
  float exposure;
};

/* float -> half the way the reference CPU kernels store display pixels
 * (util/util_half.h:80-100, scalar branch): clamp to [0, 65504], flush what would be a
 * half denormal to 0, TRUNCATE the mantissa - bit-exact with the oracle, unlike
 * __float2half which rounds to nearest. */
CY_DEV unsigned short float_to_half_display(float f)
{
  const float c = (f > 0.0f) ? ((f < 65504.0f) ? f : 65504.0f) : 0.0f;
  const int absolute = __float_as_int(c) & 0x7FFFFFFF;
  const int Z = absolute + (int)0xC8000000u;
  const int result = (absolute < 0x38800000) ? 0 : Z;
  return (unsigned short)((result >> 13) & 0x7FFF);
}

/* kernel_film_convert_to_byte / _to_half_float (kernel_film.h:19-130) for the combined
 * display pass (display_pass_components == 4, no divide pass) */
__global__ void k_film_convert(const float *film, void *rgba, int half_float, float sample_scale,
                               int x0, int y0, int w, int h, int offset, int stride,
                               FilmDisplay fd)
{
  const int x = x0 + blockIdx.x * blockDim.x + threadIdx.x;
  const int y = y0 + blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= x0 + w || y >= y0 + h)
    return;
  const long long index = (long long)offset + x + (long long)y * stride;
  const float4 in = *(const float4 *)(film + fd.display_pass_stride + index * fd.pass_stride);
  /* film_get_pass_result */
  float4 r = make_float4(in.x, in.y, in.z, fd.use_pass_alpha ? in.w : 1.0f / sample_scale);
  if (fd.use_exposure) {
    r.x *= fd.exposure;
    r.y *= fd.exposure;
    r.z *= fd.exposure;
  }
  if (half_float) {
    ushort4 v;
    v.x = float_to_half_display(r.x * sample_scale);
    v.y = float_to_half_display(r.y * sample_scale);
    v.z = float_to_half_display(r.z * sample_scale);
    v.w = float_to_half_display(r.w * sample_scale);
    *((ushort4 *)rgba + index) = v;
  }
  else {
    /* film_map + film_float_to_byte */
    uchar4 v;
    v.x = (unsigned char)(saturate(color_linear_to_srgb(r.x * sample_scale)) * 255.0f);
    v.y = (unsigned char)(saturate(color_linear_to_srgb(r.y * sample_scale)) * 255.0f);
    v.z = (unsigned char)(saturate(color_linear_to_srgb(r.z * sample_scale)) * 255.0f);
    v.w = (unsigned char)(saturate(r.w * sample_scale) * 255.0f);
    *((uchar4 *)rgba + index) = v;
  }
}

__global__ void k_film_add(float4 *dst, const float4 *src, size_t n4)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4;
       i += (size_t)gridDim.x * blockDim.x) {
    float4 a = dst[i];
    const float4 b = src[i];
    a.x += b.x;
    a.y += b.y;
    a.z += b.z;
    a.w += b.w;
    dst[i] = a;
  }
}


/* ------------------------------------------------------- DeviceTask::SHADER */

/* kernel_background_evaluate (kernel_bake.h:474-510) - what the host runs once per scene
 * to build the world importance map (shade_background_pixels, render/light.cpp:38-102):
 * input[i] = (u, v) of an equirectangular pixel as float bits, output[i] += the world
 * shader's colour in that direction. */
template<bool EXT>
__global__ void __launch_bounds__(WF_BLOCK) k_background_evaluate(const uint4 *input,
                                                                  float4 *output, int first,
                                                                  int count)
{
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
    const int i = first + k;
    const uint4 in = input[i];
    const f3 D = equirectangular_to_direction(__uint_as_float(in.x), __uint_as_float(in.y));
    const int shader = kd_int(KD_BG_SURFACE_SHADER);
    f3 color = zero3();
    if (!shader_constant_emission_eval(shader, &color)) {
      ShaderDataG sd;
      sd.P = D;
      sd.N = -D;
      sd.Ng = -D;
      sd.I = -D;
      sd.shader = shader;
      sd.flag = shader_flags(shader);
      sd.object_flag = 0;
      sd.ray_length = 0.0f;
      sd.object = -1;
      sd.prim = CY_PRIM_NONE;
      sd.lamp = -1;
      sd.type = 0;
      sd.u = sd.v = 0.0f;
      sd.dPdu = zero3();
      PathDepths depths;
      depths.bounce = depths.diffuse = depths.glossy = depths.transmission = 0;
      depths.transparent = 0;
      shader_eval_emission<EXT>(sd, depths, CY_PATH_RAY_EMISSION);
      if (sd.flag & CY_SD_EMISSION)
        color = sd.closure_emission_background;
    }
    float4 o = output[i];
    o.x += color.x;
    o.y += color.y;
    o.z += color.z;
    output[i] = o;
  }
}

/* ============================================================ host side */

static void free_pool(b200_ctx *ctx)
{
  if (!ctx->pool)
    return;
  DeviceGuard guard(ctx->ordinal);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->pool->block)
    cudaFree(ctx->pool->block);
  if (ctx->pool->h_counters)
    cudaFreeHost(ctx->pool->h_counters);
  if (ctx->pool->h_ring)
    cudaFreeHost(ctx->pool->h_ring);
  if (ctx->pool->h_batch)
    cudaFreeHost(ctx->pool->h_batch);
  for (int r = 0; r < WF_RING; r++)
    for (int e = 0; e < WF_RING_EVENTS; e++)
      if (ctx->pool->ring_ev[r][e])
        cudaEventDestroy(ctx->pool->ring_ev[r][e]);
  if (ctx->pool->sobol_tab)
    cudaFree(ctx->pool->sobol_tab);
  delete ctx->pool;
  ctx->pool = nullptr;
  ctx->pool_bytes = 0;
}

#define PATH_POOL_BYTES_PER_PATH 228 /* 12 float4 + 8 words per path, see the carve list */

static int ensure_pool(b200_ctx *ctx, size_t capacity, bool transparent_shadows, bool ao,
                       bool passes)
{
  if (ctx->pool && ctx->pool->capacity >= capacity &&
      (ctx->pool->has_ts || !transparent_shadows) && (ctx->pool->has_ao || !ao) &&
      (ctx->pool->has_passes || !passes))
    return B200_OK;
  free_pool(ctx);
  PathPool *pool = new PathPool();
  pool->capacity = capacity;
  pool->has_ts = transparent_shadows;
  pool->has_ao = ao;
  pool->has_passes = passes;
  /* carve one allocation; every array 256-byte aligned */
  size_t off = 0;
  auto carve = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    return o;
  };
  const size_t n = capacity;
  const size_t nsh = ao ? 2 * n : n; /* ambient occlusion: a second shadow ray per path */
  size_t o_nrayP = carve(n * 16), o_nrayD = carve(n * 16), o_rpdf = carve(n * 4);
  size_t o_rayP = carve(n * 16), o_rayD = carve(n * 16), o_hit = carve(n * 16),
         o_hobj = carve(n * 4), o_thr = carve(n * 16), o_L = carve(n * 16), o_sA = carve(n * 16),
         o_sB = carve(n * 16), o_shP = carve(nsh * 16), o_shD = carve(nsh * 16), o_shC = carve(nsh * 16),
         o_key = carve(n * 4), o_qa = carve(n * 4), o_qn = carve(n * 4), o_qs = carve(n * 8),
         o_qsh = carve(nsh * 4), o_cnt = carve(sizeof(WFCounters));
  /* transparent-shadow stepping queues: 88 bytes per path more, only when needed */
  const size_t nts = transparent_shadows ? n : 0;
  size_t o_tsP0 = carve(nts * 16), o_tsP1 = carve(nts * 16), o_tsD0 = carve(nts * 16),
         o_tsD1 = carve(nts * 16), o_tsI0 = carve(nts * 4), o_tsI1 = carve(nts * 4),
         o_tsT = carve(nts * 16);
  /* render passes: PASS_WORDS floats per path, two more colours per shadow entry */
  const size_t npass = passes ? n : 0;
  size_t o_pass = carve(npass * PASS_WORDS * sizeof(float)), o_shpass = carve((passes ? nsh : 0) * SH_PASS_QUADS * 16);
  cudaError_t e = cudaMalloc(&pool->block, off);
  if (e != cudaSuccess) {
    cudaGetLastError();
    delete pool;
    return fail(ctx, e == cudaErrorMemoryAllocation ? B200_ERR_OOM : B200_ERR_CUDA,
                std::string("path pool allocation: ") + cudaGetErrorString(e));
  }
  pool->bytes = off;
  ctx->pool_bytes = off;
  char *b = (char *)pool->block;
  PathSoA &s = pool->soa;
  s.ray_P_t = (float4 *)(b + o_rayP);
  s.ray_D = (float4 *)(b + o_rayD);
  s.nray_P_t = (float4 *)(b + o_nrayP);
  s.nray_D = (float4 *)(b + o_nrayD);
  s.ray_pdf = (float *)(b + o_rpdf);
  s.hit = (float4 *)(b + o_hit);
  s.hit_object = (int *)(b + o_hobj);
  s.throughput = (float4 *)(b + o_thr);
  s.L = (float4 *)(b + o_L);
  s.stateA = (uint4 *)(b + o_sA);
  s.stateB = (uint4 *)(b + o_sB);
  s.sh_P_t = (float4 *)(b + o_shP);
  s.sh_D = (float4 *)(b + o_shD);
  s.sh_contrib = (float4 *)(b + o_shC);
  s.key = (unsigned int *)(b + o_key);
  s.q_active = (int *)(b + o_qa);
  s.q_next = (int *)(b + o_qn);
  s.q_sorted = (int2 *)(b + o_qs);
  s.q_shadow = (int *)(b + o_qsh);
  s.ts_P_t[0] = transparent_shadows ? (float4 *)(b + o_tsP0) : nullptr;
  s.ts_P_t[1] = transparent_shadows ? (float4 *)(b + o_tsP1) : nullptr;
  s.ts_D[0] = transparent_shadows ? (float4 *)(b + o_tsD0) : nullptr;
  s.ts_D[1] = transparent_shadows ? (float4 *)(b + o_tsD1) : nullptr;
  s.ts_idx[0] = transparent_shadows ? (int *)(b + o_tsI0) : nullptr;
  s.ts_idx[1] = transparent_shadows ? (int *)(b + o_tsI1) : nullptr;
  s.ts_thr = transparent_shadows ? (float4 *)(b + o_tsT) : nullptr;
  s.counters = (WFCounters *)(b + o_cnt);
  s.pass = passes ? (float *)(b + o_pass) : nullptr;
  s.sh_pass = passes ? (float4 *)(b + o_shpass) : nullptr;
  bool ok = cudaMallocHost(&pool->h_counters, sizeof(WFCounters)) == cudaSuccess &&
            cudaMallocHost(&pool->h_ring, WF_RING * WF_RING_BYTES) == cudaSuccess &&
            cudaMallocHost(&pool->h_batch, WF_BATCH_SLOTS * WF_BATCH_STAT_BYTES) == cudaSuccess &&
            cudaMalloc(&pool->sobol_tab, SOBOL_TABLE_MAX * sizeof(uint32_t)) == cudaSuccess;
  for (int r = 0; ok && r < WF_RING; r++)
    for (int e = 0; ok && e < WF_RING_EVENTS; e++)
      ok = cudaEventCreate(&pool->ring_ev[r][e]) == cudaSuccess;
  ctx->pool = pool;
  if (!ok) {
    free_pool(ctx);
    return fail(ctx, B200_ERR_CUDA, "pinned counter / Sobol table allocation failed");
  }
  return B200_OK;
}

static bool svm_mix_supported_host(uint32_t type)
{
  switch (type) {
    case CY_NODE_MIX_BLEND:
    case CY_NODE_MIX_ADD:
    case CY_NODE_MIX_MUL:
    case CY_NODE_MIX_SCREEN:
    case CY_NODE_MIX_OVERLAY:
    case CY_NODE_MIX_SUB:
    case CY_NODE_MIX_DIV:
    case CY_NODE_MIX_DIFF:
    case CY_NODE_MIX_DARK:
    case CY_NODE_MIX_LIGHT:
    case CY_NODE_MIX_DODGE:
    case CY_NODE_MIX_BURN:
    case CY_NODE_MIX_HUE:
    case CY_NODE_MIX_SAT:
    case CY_NODE_MIX_VAL:
    case CY_NODE_MIX_COLOR:
    case CY_NODE_MIX_SOFT:
    case CY_NODE_MIX_LINEAR:
    case CY_NODE_MIX_CLAMP:
      return true;
    default:
      return false;
  }
}

/* Opcodes and closure ids the kernels implement (svm.h switch subset). */
static bool svm_validate(const uint32_t *nodes, size_t n_nodes, std::string &why,
                         uint32_t *features, int *max_image_slot)
{
  *features = 0;
  /* Constants the program wrote with NODE_VALUE_F since the last node of any other kind:
   * the SVM compiler emits a closure node's unlinked inputs as such a run right before
   * it (SVMCompiler::stack_assign from ShaderNode::compile), which is enough to tell a
   * Principled BSDF whose sheen is a constant zero from one that may have sheen. */
  bool const_known[256] = {};
  float const_value[256] = {};
  size_t const_run_end = 0; /* node index the table is valid for */
  /* Walk the stream linearly; NODE_CLOSURE_BSDF and NODE_VALUE_V carry data nodes
   * that must be skipped exactly as the interpreter does. */
  size_t i = 0;
  while (i < n_nodes) {
    const uint32_t op = nodes[4 * i];
    if (op == CY_NODE_VALUE_F) {
      if (const_run_end != i)
        memset(const_known, 0, sizeof(const_known));
      const uint32_t slot = nodes[4 * i + 2] & 0xff;
      const_known[slot] = true;
      memcpy(&const_value[slot], &nodes[4 * i + 1], 4);
      const_run_end = i + 1;
    }
    else if (op == CY_NODE_VALUE_V && const_run_end == i) {
      const uint32_t slot = nodes[4 * i + 1] & 0xff; /* overwrites three slots */
      for (uint32_t k = slot; k < slot + 3 && k < 256; k++)
        const_known[k] = false;
      const_run_end = i + 2;
    }
    switch (op) {
      case CY_NODE_END:
      case CY_NODE_SHADER_JUMP:
      case CY_NODE_CLOSURE_EMISSION:
      case CY_NODE_CLOSURE_BACKGROUND:
      case CY_NODE_CLOSURE_SET_WEIGHT:
      case CY_NODE_CLOSURE_WEIGHT:
      case CY_NODE_EMISSION_WEIGHT:
      case CY_NODE_MIX_CLOSURE:
      case CY_NODE_JUMP_IF_ZERO:
      case CY_NODE_JUMP_IF_ONE:
      case CY_NODE_VALUE_F:
      case CY_NODE_CONVERT:
      case CY_NODE_FRESNEL:
      case CY_NODE_LAYER_WEIGHT:
      case CY_NODE_MATH:
      case CY_NODE_INVERT:
      case CY_NODE_GAMMA:
      case CY_NODE_BRIGHTCONTRAST:
      case CY_NODE_SEPARATE_VECTOR:
      case CY_NODE_COMBINE_VECTOR:
      case CY_NODE_LIGHT_PATH:    /* both run in the lean interpreter and read */
      case CY_NODE_LIGHT_FALLOFF: /* nothing but the path state / the light sample */
        i += 1;
        break;
      /* from here on: the nodes of svm_eval_extended_node (shade.cuh) */
      case CY_NODE_TANGENT:    /* the only nodes of this group that read mesh */
      case CY_NODE_NORMAL_MAP: /* attributes (tangent / generated coordinates) */
        *features |= SVM_USES_ATTRIBUTES;
        /* fall through */
      case CY_NODE_MAPPING:
      case CY_NODE_TEX_CHECKER:
      case CY_NODE_TEX_GRADIENT:
      case CY_NODE_HSV:
      case CY_NODE_VECTOR_ROTATE:
      case CY_NODE_VECTOR_TRANSFORM:
      case CY_NODE_OBJECT_INFO:
      case CY_NODE_CAMERA:
      case CY_NODE_TEX_WHITE_NOISE:
      case CY_NODE_BLACKBODY:
      case CY_NODE_WAVELENGTH:
        *features |= SVM_USES_EXTENDED_NODES;
        i += 1;
        break;
      case CY_NODE_SEPARATE_HSV:
      case CY_NODE_COMBINE_HSV:
      case CY_NODE_NORMAL:
        *features |= SVM_USES_EXTENDED_NODES;
        i += 2;
        break;
      case CY_NODE_MAP_RANGE:
        *features |= SVM_USES_EXTENDED_NODES;
        i += 3;
        break;
      case CY_NODE_GEOMETRY: /* the tangent reads the generated-coordinates attribute */
        if (nodes[4 * i + 1] == 2 /* NODE_GEOM_T */)
          *features |= SVM_USES_TANGENT;
        i += 1;
        break;
      case CY_NODE_ATTR:
        *features |= SVM_USES_ATTRIBUTES | SVM_USES_EXTENDED_NODES;
        i += 1;
        break;
      case CY_NODE_TEX_COORD: {
        const uint32_t type = nodes[4 * i + 1];
        *features |= SVM_USES_EXTENDED_NODES;
        if (type == CY_NODE_TEXCO_WINDOW)
          *features |= SVM_USES_WINDOW_COORDINATES;
        else if (type != CY_NODE_TEXCO_NORMAL && type != CY_NODE_TEXCO_OBJECT &&
                 type != CY_NODE_TEXCO_CAMERA && type != CY_NODE_TEXCO_REFLECTION) {
          why = "texture coordinate type " + std::to_string(type) +
                " (instancer / volume coordinates) is outside the hot-path scope";
          return false;
        }
        /* object coordinates of another object carry its transform in three nodes */
        i += (type == CY_NODE_TEXCO_OBJECT && nodes[4 * i + 3] != 0) ? 4 : 1;
        break;
      }
      case CY_NODE_TEX_MAGIC:
        *features |= SVM_USES_EXTENDED_NODES;
        i += 2;
        break;
      case CY_NODE_TEX_IMAGE: {
        /* followed by its UDIM tile nodes (two (tile, slot) pairs each) when it has any;
         * otherwise the slot is the negated count */
        const int tiles = (int)nodes[4 * i + 1];
        *features |= SVM_USES_EXTENDED_NODES | SVM_USES_IMAGES;
        if (max_image_slot) {
          if (tiles <= 0)
            *max_image_slot = std::max(*max_image_slot, -tiles);
          for (int t = 0; t < tiles && i + 1 + (size_t)t < n_nodes; t++) {
            const uint32_t *tn = nodes + 4 * (i + 1 + (size_t)t);
            *max_image_slot = std::max(*max_image_slot, std::max((int)tn[1], (int)tn[3]));
          }
        }
        i += 1 + (size_t)(tiles > 0 ? tiles : 0);
        break;
      }
      case CY_NODE_TEX_IMAGE_BOX:
      case CY_NODE_TEX_ENVIRONMENT:
        *features |= SVM_USES_EXTENDED_NODES | SVM_USES_IMAGES;
        if (max_image_slot)
          *max_image_slot = std::max(*max_image_slot, (int)nodes[4 * i + 1]);
        i += 1;
        break;
      case CY_NODE_MIN_MAX:
      case CY_NODE_TEX_NOISE:
      case CY_NODE_TEX_WAVE:
      case CY_NODE_TEX_MUSGRAVE:
      case CY_NODE_TEX_VORONOI:
        *features |= SVM_USES_EXTENDED_NODES;
        i += 3;
        break;
      case CY_NODE_TEXTURE_MAPPING:
      case CY_NODE_TEX_BRICK:
        *features |= SVM_USES_EXTENDED_NODES;
        i += 4;
        break;
      case CY_NODE_VALUE_V:
      case CY_NODE_CLAMP: /* + one node of default values */
        i += 2;
        break;
      case CY_NODE_RGB_RAMP:
      case CY_NODE_RGB_CURVES:
      case CY_NODE_VECTOR_CURVES: {
        /* the instruction, one node holding the table size, then the table itself */
        if (i + 1 >= n_nodes) {
          why = "truncated ramp node";
          return false;
        }
        const size_t table_size = nodes[4 * (i + 1)];
        if (table_size < 2 || i + 2 + table_size > n_nodes) {
          why = "ramp table runs past the end of the SVM program";
          return false;
        }
        i += 2 + table_size;
        break;
      }
      case CY_NODE_VECTOR_MATH: /* the three-input operator carries an extra node */
        i += (nodes[4 * i + 1] == CY_NODE_VECTOR_MATH_WRAP) ? 2 : 1;
        break;
      case CY_NODE_MIX: {
        if (i + 1 >= n_nodes) {
          why = "truncated Mix node";
          return false;
        }
        const uint32_t blend = nodes[4 * (i + 1) + 1];
        if (!svm_mix_supported_host(blend)) {
          why = "unknown MixRGB blend mode " + std::to_string(blend);
          return false;
        }
        i += 2;
        break;
      }
      case CY_NODE_CLOSURE_BSDF: {
        const uint32_t type = nodes[4 * i + 1] & 0xff;
        if (type == CY_CLOSURE_BSDF_PRINCIPLED_ID) {
          if (i + 2 >= n_nodes) {
            why = "truncated Principled BSDF node";
            return false;
          }
          const uint32_t distribution = nodes[4 * (i + 2) + 1];
          const uint32_t subsurface_method = nodes[4 * (i + 2) + 2];
          (void)subsurface_method;
          /* subsurface > 0 makes the reference emit a BSSRDF closure (random-walk / Burley
           * scattering, outside the scope): the Principled node is accepted only when its
           * subsurface input is provably zero - a node constant, or a NODE_VALUE_F of the
           * run the SVM compiler emits right before the closure node */
          {
            const uint32_t ss_slot = (nodes[4 * i + 1] >> 16) & 0xff;
            float ss = 1.0f;
            if (ss_slot == (uint32_t)CY_SVM_STACK_INVALID)
              memcpy(&ss, &nodes[4 * i + 3], 4);
            else if (const_run_end == i && const_known[ss_slot])
              ss = const_value[ss_slot];
            if (!(ss <= 1e-5f)) {
              why = "Principled BSDF with subsurface scattering (a subsurface input that is "
                    "not a constant zero) is outside the hot-path scope";
              return false;
            }
          }
          /* sheen: handled by the full interpreter only (shade.cuh svm_eval_nodes) */
          const uint32_t sheen_slot = nodes[4 * (i + 1) + 3] & 0xff;
          if (!(const_run_end == i && const_known[sheen_slot] &&
                const_value[sheen_slot] <= 1e-5f /* CLOSURE_WEIGHT_CUTOFF */))
            *features |= SVM_USES_EXTENDED_NODES;
          /* Multiscatter GGX (the node's default): the random-walk lobes live in the full
           * kernels */
          if (distribution != CY_CLOSURE_BSDF_MICROFACET_GGX_GLASS_ID)
            *features |= SVM_USES_MULTISCATTER;
          i += 6;
        }
        else if (type == CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_ID ||
                 type == CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_GLASS_ID) {
          /* Glossy / Anisotropic / Glass BSDF nodes with Multiscatter GGX; a tangent
           * input (the Anisotropic node) runs in the full interpreter */
          *features |= SVM_USES_MULTISCATTER;
          if (type == CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_ID && i + 1 < n_nodes &&
              nodes[4 * (i + 1) + 1] != (uint32_t)CY_SVM_STACK_INVALID)
            *features |= SVM_USES_EXTENDED_NODES;
          i += 2;
        }
        else if (type == CY_CLOSURE_BSDF_REFLECTION_ID ||
                 type == CY_CLOSURE_BSDF_MICROFACET_GGX_ID) {
          /* the Anisotropic BSDF node (a tangent input) runs in the full interpreter */
          if (i + 1 >= n_nodes) {
            why = "truncated glossy BSDF node";
            return false;
          }
          if (nodes[4 * (i + 1) + 1] != (uint32_t)CY_SVM_STACK_INVALID)
            *features |= SVM_USES_EXTENDED_NODES;
          i += 2;
        }
        else if (type == CY_CLOSURE_BSDF_DIFFUSE_ID || type == CY_CLOSURE_BSDF_TRANSLUCENT_ID ||
                 type == CY_CLOSURE_BSDF_TRANSPARENT_ID ||
                 type == CY_CLOSURE_BSDF_REFRACTION_ID ||
                 type == CY_CLOSURE_BSDF_MICROFACET_GGX_REFRACTION_ID ||
                 type == CY_CLOSURE_BSDF_SHARP_GLASS_ID ||
                 type == CY_CLOSURE_BSDF_MICROFACET_GGX_GLASS_ID) {
          i += 2;
        }
        else {
          why = "SVM closure type " + std::to_string(type) + " is outside the hot-path scope";
          return false;
        }
        break;
      }
      default:
        why = "SVM node opcode " + std::to_string(op) + " at " + std::to_string(i) +
              " is outside the hot-path scope (supported: closures diffuse/principled-GGX/"
              "glossy-GGX/emission/background, mix closure, value, geometry, convert, fresnel, "
              "layer weight, math, vector math, mix, invert, gamma, bright/contrast, "
              "separate/combine, clamp, light path, light falloff, RGB ramp, curves, attribute, "
              "texture coordinate, mapping, noise/wave/magic/checker/brick/gradient/white noise/musgrave/voronoi "
              "textures, HSV, map range, normal, vector rotate/transform, object info, camera)";
        return false;
    }
  }
  return true;
}

/* Features of KernelData the kernels do not implement are refused, never
 * silently approximated. */
/* one past the highest image slot that has pixels */
static int bound_image_slots(const b200_ctx *ctx)
{
  int n = 0;
  for (size_t t = 0; t + SIZEOF_TEXTURE_INFO <= ctx->texture_info.size(); t += SIZEOF_TEXTURE_INFO) {
    uint64_t data = 0;
    memcpy(&data, ctx->texture_info.data() + t + TI_DATA, 8);
    if (data)
      n = (int)(t / SIZEOF_TEXTURE_INFO) + 1;
  }
  return n;
}

/* the film holds more than the combined pass, or a lamp's ray-visibility flags need the
 * light split per BSDF class: the kernels with the pass accumulators run (passes.cuh) */
static bool film_wants_passes(const b200_ctx *ctx)
{
  return kd_host<int>(ctx, KD_FILM_USE_LIGHT_PASS) != 0 ||
         kd_host<int>(ctx, KD_FILM_PASS_DENOISING_DATA) != 0 ||
         (kd_host<int>(ctx, KD_FILM_PASS_FLAG) &
          ~((1 << CY_PASS_COMBINED) | (1 << CY_PASS_ADAPTIVE_AUX_BUFFER) |
            (1 << CY_PASS_SAMPLE_COUNT))) != 0;
}

static int check_scope(b200_ctx *ctx)
{
  auto I = [&](int off) { return kd_host<int>(ctx, off); };
  auto F = [&](int off) { return kd_host<float>(ctx, off); };
  std::string why;
  if (I(KD_CAM_TYPE) != CY_CAMERA_PERSPECTIVE && I(KD_CAM_TYPE) != CY_CAMERA_ORTHOGRAPHIC &&
      I(KD_CAM_TYPE) != CY_CAMERA_PANORAMA)
    why = "unknown camera type";
  else if (F(KD_CAM_INTEROCULAR_OFFSET) != 0.0f)
    why = "stereo cameras are outside the hot-path scope";
  else if (F(KD_CAM_SHUTTERTIME) != -1.0f || I(KD_CAM_NUM_MOTION_STEPS) != 0 ||
           I(KD_BVH_HAVE_MOTION))
    why = "motion blur is outside the hot-path scope";
  else if (I(KD_BVH_HAVE_CURVES))
    why = "hair curves are outside the hot-path scope";
  else if (I(KD_INT_SAMPLING_PATTERN) != CY_SAMPLING_PATTERN_SOBOL &&
           I(KD_INT_SAMPLING_PATTERN) != CY_SAMPLING_PATTERN_CMJ &&
           I(KD_INT_SAMPLING_PATTERN) != CY_SAMPLING_PATTERN_PMJ)
    why = "unknown sampling pattern";
  else if (I(KD_INT_SAMPLING_PATTERN) == CY_SAMPLING_PATTERN_PMJ &&
           (!find_global(ctx, "__sample_pattern_lut") ||
            find_global(ctx, "__sample_pattern_lut")->bytes <
                (size_t)CY_NUM_PMJ_PATTERNS * CY_NUM_PMJ_SAMPLES * 2 * 4))
    why = "the PMJ sampling pattern needs the __sample_pattern_lut table";
  else if (I(KD_INT_BRANCHED))
    why = "branched path tracing is outside the hot-path scope";
  else if (I(KD_INT_USE_VOLUMES))
    why = "volumes are outside the hot-path scope";
  else if (I(KD_INT_USE_AMBIENT_OCCLUSION) && I(KD_INT_TRANSPARENT_SHADOWS))
    why = "ambient occlusion together with transparent shadows is outside the hot-path scope";
  else if ((ctx->svm_features & SVM_USES_IMAGES) && ctx->svm_max_image_slot >= bound_image_slots(ctx))
    why = "the shader program samples image slot " + std::to_string(ctx->svm_max_image_slot) +
          " but only " + std::to_string(bound_image_slots(ctx)) +
          " image slots are bound (tex_alloc / b200_texture_set)";
  else if (F(KD_BG_PORTAL_WEIGHT) > 0.0f || I(KD_BG_NUM_PORTALS) > 0)
    why = "light portals are outside the hot-path scope";
  else if (I(KD_BG_USE_MIS) && F(KD_BG_SUN_WEIGHT) > 0.0f)
    why = "the sky texture's sun disc (background sun sampling) is outside the hot-path scope";
  else if (I(KD_BG_USE_MIS) && F(KD_BG_MAP_WEIGHT) > 0.0f &&
           (!find_global(ctx, "__light_background_marginal_cdf") ||
            !find_global(ctx, "__light_background_conditional_cdf") ||
            find_global(ctx, "__light_background_marginal_cdf")->bytes <
                (size_t)(I(KD_BG_MAP_RES_Y) + 1) * 8 ||
            find_global(ctx, "__light_background_conditional_cdf")->bytes <
                (size_t)(I(KD_BG_MAP_RES_X) + 1) * I(KD_BG_MAP_RES_Y) * 8))
    why = "background importance sampling needs the world map CDF arrays";
  else if (I(KD_FILM_CRYPTOMATTE_PASSES))
    why = "cryptomatte passes are outside the hot-path scope";
  else if (I(KD_FILM_PASS_DENOISING_DATA) &&
           (I(KD_FILM_PASS_DENOISING_DATA) + CY_DENOISING_PASS_SIZE_BASE > I(KD_FILM_PASS_STRIDE) ||
            (I(KD_FILM_PASS_DENOISING_CLEAN) &&
             I(KD_FILM_PASS_DENOISING_CLEAN) + CY_DENOISING_PASS_SIZE_CLEAN >
                 I(KD_FILM_PASS_STRIDE))))
    why = "the denoising data passes do not fit the film's pass stride";
  else if (I(KD_FILM_PASS_DENOISING_CLEAN) &&
           (!I(KD_FILM_PASS_DENOISING_DATA) || !I(KD_FILM_USE_LIGHT_PASS)))
    why = "the denoising clean pass needs the denoising data passes and light passes";
  else if (I(KD_FILM_PASS_ADAPTIVE_AUX_BUFFER) &&
           (I(KD_INT_SAMPLING_PATTERN) != CY_SAMPLING_PATTERN_PMJ || film_wants_passes(ctx) ||
            I(KD_FILM_PASS_ADAPTIVE_AUX_BUFFER) % 4 != 0))
    why = "adaptive sampling is in scope with the PMJ pattern and the combined pass only";
  else if (I(KD_FILM_PASS_SAMPLE_COUNT) && !I(KD_FILM_PASS_ADAPTIVE_AUX_BUFFER))
    why = "the sample-count pass is in scope together with adaptive sampling only";
  else if (I(KD_FILM_PASS_FLAG) & ~((1 << CY_PASS_COMBINED) | (1 << CY_PASS_DEPTH) |
                                    (1 << CY_PASS_NORMAL) | (1 << CY_PASS_UV) |
                                    (1 << CY_PASS_OBJECT_ID) | (1 << CY_PASS_MATERIAL_ID) |
                                    (1 << CY_PASS_ADAPTIVE_AUX_BUFFER) |
                                    (1 << CY_PASS_SAMPLE_COUNT)))
    why = "of the data passes depth, normal, UV, object id and material id are in scope "
          "(motion, AOV and render-time passes are not)";
  else if (I(KD_FILM_LIGHT_PASS_FLAG) &
           ((1 << (CY_PASS_VOLUME_DIRECT % 32)) | (1 << (CY_PASS_VOLUME_INDIRECT % 32))))
    why = "the volume light passes are outside the hot-path scope";
  else if (!(I(KD_FILM_PASS_FLAG) & (1u << CY_PASS_COMBINED)))
    why = "the combined pass must be enabled";
  else if (I(KD_FILM_PASS_STRIDE) % 4 != 0)
    why = "pass_stride must be a multiple of 4";
  else if (F(KD_BG_TRANSPARENT_ROUGHNESS_SQ_THRESHOLD) >= 0.0f)
    why = "transparent glass is outside the hot-path scope";
  else if (I(KD_INT_MAX_CLOSURES) > MAX_CLOSURES_GPU)
    why = "more than 32 closures per shader";
  else if ((ctx->svm_features & SVM_USES_WINDOW_COORDINATES) &&
           I(KD_CAM_TYPE) != CY_CAMERA_PERSPECTIVE)
    why = "window texture coordinates with a non-perspective camera are outside the hot-path "
          "scope";
  else if (ctx->has_subd_patches &&
           ((ctx->svm_features & SVM_USES_ATTRIBUTES) ||
            ((ctx->svm_features & SVM_USES_TANGENT) && ctx->has_generated_attr)))
    why = "attributes on subdivision patches are outside the hot-path scope";
  const HostArray *sh = find_global(ctx, "__shaders");
  if (why.empty() && sh && sh->bytes / SIZEOF_KERNEL_SHADER > WF_MAX_KEYS - 1)
    why = "more than 4095 shaders";
  if (!why.empty())
    return fail(ctx, B200_ERR_UNSUPPORTED, why);
  return B200_OK;
}

/* The surface-shading kernels use more dynamic shared memory than the default limit;
 * their grids are sized to what is resident (blocks per SM x SMs), each block loops. */
static int shade_kernel_setup(b200_ctx *ctx)
{
  if (ctx->shade_blocks_per_sm[0] > 0)
    return B200_OK;
  DeviceGuard guard(ctx->ordinal);
  const void *kernels[8] = {(const void *)k_shade_surface<false, false, false, GGX_LEAN_SHAPE>,
                            (const void *)k_shade_surface<false, true>,
                            (const void *)k_shade_surface<true, true>,
                            (const void *)k_shade_surface<true, true, true>,
                            (const void *)k_shade_surface<false, true, false, 1>,
                            (const void *)k_shade_surface<true, true, false, 1>,
                            (const void *)k_shade_surface<false, true, false, 2>,
                            (const void *)k_shade_surface<true, true, false, 2>};
  for (int k = 0; k < 8; k++) {
    const int block = (k == 0) ? SHADE_BLOCK_OF(GGX_LEAN_SHAPE) :
                      (k < 4)  ? WF_BLOCK :
                                 SHADE_BLOCK_OF(k < 6 ? 1 : 2);
    const size_t smem = (k == 3) ? SHADE_SMEM_BYTES_PASSES : SHADE_SMEM_BYTES_OF(block);
    CUDA_TRY(ctx, cudaFuncSetAttribute(kernels[k], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    int blocks = 0;
    CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kernels[k], block,
                                                                smem));
    if (blocks < 1)
      return fail(ctx, B200_ERR_CUDA, "surface-shading kernel does not fit on an SM");
    /* A/B: how much of the L1 / shared array is asked for as shared memory (percent) */
    if (ctx->opt_shade_carveout > 0)
      CUDA_TRY(ctx, cudaFuncSetAttribute(kernels[k], cudaFuncAttributePreferredSharedMemoryCarveout,
                                         (int)ctx->opt_shade_carveout));
    /* a few waves of blocks per SM even the tail out */
    ctx->shade_blocks_per_sm[k] = blocks * 2;
  }
  return B200_OK;
}

extern "C" {

int b200_render(b200_ctx *ctx, const b200_work_tile *tile, volatile const int *cancel)
{
  if (!ctx || !tile || !tile->buffer)
    return B200_ERR_INVALID;
  if (tile->w <= 0 || tile->h <= 0 || tile->num_samples <= 0)
    return B200_OK;
  DeviceUse device_use(ctx);
  int rc = prepare_scene(ctx);
  if (rc)
    return rc;
  DeviceGuard guard(ctx->ordinal);
  rc = apply_l2_window(ctx);
  if (rc)
    return rc;

  /* Paths per wavefront batch.  Every kernel of a bounce is one launch over the whole
   * batch and the persistent traversal kernels pay a drain phase per launch (warps
   * running out of rays), so the batch is made as large as the work and the memory
   * allow: up to 32 Mi paths (7 GB of the 180 GB), never more than a quarter of what
   * is free, never more than the task has. */
  size_t capacity = (size_t)ctx->opt_batch_paths;
  if (capacity == 0) {
    const size_t task_paths = (size_t)tile->w * tile->h * (size_t)tile->num_samples;
    capacity = std::min<size_t>(std::max<size_t>(task_paths, (size_t)1 << 20), (size_t)1 << 25);
    size_t free_b = 0, total_b = 0;
    if (!(ctx->pool && ctx->pool->capacity >= capacity) &&
        cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
      const size_t held = ctx->pool ? ctx->pool->capacity * PATH_POOL_BYTES_PER_PATH : 0;
      const size_t per_path = PATH_POOL_BYTES_PER_PATH +
                              (kd_host<int>(ctx, KD_INT_TRANSPARENT_SHADOWS) ? 88 : 0) +
                              (kd_host<int>(ctx, KD_INT_USE_AMBIENT_OCCLUSION) ? 52 : 0) +
                              (film_wants_passes(ctx) ? PASS_WORDS * 4 + SH_PASS_QUADS * 16 : 0);
      const size_t fit = (free_b + held) / 4 / per_path;
      capacity = std::max<size_t>(std::min(capacity, fit), (size_t)1 << 16);
    }
  }
  const bool transparent_shadows = kd_host<int>(ctx, KD_INT_TRANSPARENT_SHADOWS) != 0;
  /* the full shading kernels: extended SVM nodes or sheen (svm_validate), an object with a
   * shadow terminator offset (terminator terms of bsdf_eval / bsdf_sample), the table
   * sampling pattern, ambient occlusion - or a lean batch that met a shader it could not run (a closure whose
   * linked normal differs from the shading normal), see the retry below */
  bool svm_ext = ctx->force_svm_ext || (ctx->svm_features & SVM_USES_EXTENDED_NODES) != 0 ||
                       ctx->has_terminator_offset ||
                       kd_host<int>(ctx, KD_INT_SAMPLING_PATTERN) == CY_SAMPLING_PATTERN_PMJ ||
                       kd_host<int>(ctx, KD_INT_USE_AMBIENT_OCCLUSION) != 0;
  const bool use_ao = kd_host<int>(ctx, KD_INT_USE_AMBIENT_OCCLUSION) != 0;
  /* render passes beyond the combined one: the full kernels with the pass accumulators */
  const bool passes = film_wants_passes(ctx);
  if (passes)
    svm_ext = true;
  /* lean kernels with the multi-scatter lobes: the Principled default distribution */
  const bool multiscatter = (ctx->svm_features & SVM_USES_MULTISCATTER) != 0;
  rc = ensure_pool(ctx, std::max<size_t>(capacity, (size_t)tile->w), transparent_shadows, use_ao,
                   passes);
  if (rc)
    return rc;
  PathPool *pool = ctx->pool;
  const HostArray *sh = find_global(ctx, "__shaders");
  const int num_keys = sh ? (int)(sh->bytes / SIZEOF_KERNEL_SHADER) : 1;
  const int pass_stride = kd_host<int>(ctx, KD_FILM_PASS_STRIDE);
  const int pass_combined = kd_host<int>(ctx, KD_FILM_PASS_COMBINED);
  const bool count = ctx->opt_count_traversal != 0;
  const int grid_wide = launch_grid(ctx, 8);
  const int grid_trace = launch_grid(
      ctx, ctx->opt_trace_blocks_per_sm > 0 ? (int)ctx->opt_trace_blocks_per_sm : 8);
  const int refill = refill_threshold(ctx);
  cudaStream_t st = ctx->stream;
  rc = shade_kernel_setup(ctx);
  if (rc)
    return rc;

  b200_stats stats;
  memset(&stats, 0, sizeof(stats));
  float closest_ms = 0.0f, shade_ms = 0.0f, shadow_ms = 0.0f;
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev5, st));

  /* How far the host reads behind the device.  The bounce loop ends when a counter says
   * the queue is empty; instead of stopping the stream for that read every iteration, the
   * host queues iteration `it`, and only then looks at the counters iteration it - 1
   * copied into its ring slot.  The price is one iteration over empty queues per batch
   * (8 launches that find nothing to do); the device never idles.  Transparent shadows
   * queue a fixed number of stepping rounds (see ts_rounds) and are pipelined the same
   * way. */
  /* transparent shadows: rounds of the stepping loop queued per iteration without a
   * counter read (0 = the host asks after every step) */
  const int ts_max_bounce = kd_host<int>(ctx, KD_INT_TRANSPARENT_MAX_BOUNCE);
  const int ts_rounds = (transparent_shadows && !ctx->opt_sync_iterations &&
                         ts_max_bounce <= WF_TS_BLIND_ROUNDS) ? std::max(ts_max_bounce, 1) : 0;
  const int lag = ((transparent_shadows && ts_rounds == 0) || ctx->opt_sync_iterations) ? 0 : 1;
  int batch_slot = 0;
  auto sum_batch_stats = [&]() -> int {
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    stats.host_syncs += 1;
    for (int k = 0; k < batch_slot; k++) {
      const WFCounters *h = (const WFCounters *)(pool->h_batch + (size_t)k * WF_BATCH_STAT_BYTES);
      stats.primary_rays += h->primary_rays;
      stats.bounce_rays += h->bounce_rays;
      stats.shadow_rays += h->shadow_rays;
      stats.closest_nodes += h->nodes;
      stats.closest_tris += h->tris;
      stats.closest_instances += h->instances;
      stats.shadow_nodes += h->sh_nodes;
      stats.shadow_tris += h->sh_tris;
      stats.shadow_instances += h->sh_instances;
    }
    batch_slot = 0;
    return B200_OK;
  };

  /* one bounce of the whole batch: 8 launches, the head of the counters into the
   * iteration's ring slot, events around the three phases */
  /* which block shape of the lean multiscatter kernel and of the full kernel this batch
   * runs (see k_shade_surface, WIDE): forced, decided, or - while probing - alternating by
   * batch */
  int shade_wide = 0;
  auto enqueue_iteration = [&](const PathSoA &soa, int it) -> int {
    cudaEvent_t *ev = pool->ring_ev[it % WF_RING];
    const int grid_shade = ctx->num_sms *
                           ctx->shade_blocks_per_sm[svm_ext ? (shade_wide ? 3 + 2 * shade_wide : 2) :
                                                    (multiscatter ?
                                                         (shade_wide ? 2 + 2 * shade_wide : 1) :
                                                         0)];
    CUDA_TRY(ctx, cudaEventRecord(ev[0], st));
    if (count)
      k_intersect_closest<true><<<grid_trace, TRACE_BLOCK, 0, st>>>(soa, refill);
    else
      k_intersect_closest<false><<<grid_trace, TRACE_BLOCK, 0, st>>>(soa, refill);
    CUDA_TRY(ctx, cudaEventRecord(ev[1], st));
    if (soa.sort_tiles) {
      k_sort_tiles<<<grid_wide, WF_BLOCK, 0, st>>>(soa);
      stats.kernel_launches += 6;
    }
    else {
      k_sort_count<<<grid_wide, WF_BLOCK, 0, st>>>(soa);
      k_sort_scan<<<1, 256, 0, st>>>(soa, num_keys);
      k_sort_scatter<<<grid_wide, WF_BLOCK, 0, st>>>(soa);
      stats.kernel_launches += 8;
    }
    if (passes) {
      k_shade_background<true, true><<<grid_wide, WF_BLOCK, 0, st>>>(soa);
      k_shade_surface<true, true, true>
          <<<ctx->num_sms * ctx->shade_blocks_per_sm[3], WF_BLOCK, SHADE_SMEM_BYTES_PASSES, st>>>(
              soa, num_keys);
    }
    else if (svm_ext) {
      k_shade_background<true><<<grid_wide, WF_BLOCK, 0, st>>>(soa);
      if (shade_wide == 2)
        k_shade_surface<true, true, false, 2>
            <<<grid_shade, SHADE_BLOCK_OF(2), SHADE_SMEM_BYTES_OF(SHADE_BLOCK_OF(2)), st>>>(
                soa, num_keys);
      else if (shade_wide == 1)
        k_shade_surface<true, true, false, 1>
            <<<grid_shade, SHADE_BLOCK_OF(1), SHADE_SMEM_BYTES_OF(SHADE_BLOCK_OF(1)), st>>>(
                soa, num_keys);
      else
        k_shade_surface<true, true><<<grid_shade, WF_BLOCK, SHADE_SMEM_BYTES, st>>>(soa,
                                                                                    num_keys);
    }
    else {
      k_shade_background<false><<<grid_wide, WF_BLOCK, 0, st>>>(soa);
      if (multiscatter && shade_wide == 2)
        k_shade_surface<false, true, false, 2>
            <<<grid_shade, SHADE_BLOCK_OF(2), SHADE_SMEM_BYTES_OF(SHADE_BLOCK_OF(2)), st>>>(
                soa, num_keys);
      else if (multiscatter && shade_wide == 1)
        k_shade_surface<false, true, false, 1>
            <<<grid_shade, SHADE_BLOCK_OF(1), SHADE_SMEM_BYTES_OF(SHADE_BLOCK_OF(1)), st>>>(
                soa, num_keys);
      else if (multiscatter)
        k_shade_surface<false, true><<<grid_shade, WF_BLOCK, SHADE_SMEM_BYTES, st>>>(soa,
                                                                                     num_keys);
      else
        k_shade_surface<false, false, false, GGX_LEAN_SHAPE>
            <<<grid_shade, SHADE_BLOCK_OF(GGX_LEAN_SHAPE),
               SHADE_SMEM_BYTES_OF(SHADE_BLOCK_OF(GGX_LEAN_SHAPE)), st>>>(soa, num_keys);
    }
    CUDA_TRY(ctx, cudaEventRecord(ev[2], st));
    if (passes && use_ao) /* AO never with transparent shadows (check_scope) */
      k_intersect_shadow<false, false, true, true><<<grid_trace, TRACE_BLOCK, 0, st>>>(soa,
                                                                                       refill);
    else if (passes && transparent_shadows)
      k_intersect_shadow<false, true, false, true><<<grid_trace, TRACE_BLOCK, 0, st>>>(soa,
                                                                                       refill);
    else if (passes)
      k_intersect_shadow<false, false, false, true><<<grid_trace, TRACE_BLOCK, 0, st>>>(soa,
                                                                                        refill);
    else if (use_ao) /* never with transparent shadows (check_scope) */
      k_intersect_shadow<false, false, true><<<grid_trace, TRACE_BLOCK, 0, st>>>(soa, refill);
    else if (transparent_shadows)
      k_intersect_shadow<false, true><<<grid_trace, TRACE_BLOCK, 0, st>>>(soa, refill);
    else if (count)
      k_intersect_shadow<true, false><<<grid_trace, TRACE_BLOCK, 0, st>>>(soa, refill);
    else
      k_intersect_shadow<false, false><<<grid_trace, TRACE_BLOCK, 0, st>>>(soa, refill);
    CUDA_TRY(ctx, cudaEventRecord(ev[3], st));
    if (transparent_shadows && ts_rounds > 0) {
      /* shadow rays stopped by a transparent surface walk on, one surface per step.  A ray
       * crosses at most transparent_max_bounce surfaces, so that many rounds are queued
       * without asking how many rays are left: a round over an empty queue is three launches
       * that return at once, and the host stays out of the loop. */
      int cur = 0;
      for (int round = 0; round < ts_rounds; round++) {
        k_intersect_shadow_step<<<grid_trace, TRACE_BLOCK, 0, st>>>(soa, cur, refill);
        if (passes)
          k_shade_shadow_step<true, true><<<grid_wide, WF_BLOCK, 0, st>>>(soa, cur);
        else if (svm_ext)
          k_shade_shadow_step<true><<<grid_wide, WF_BLOCK, 0, st>>>(soa, cur);
        else
          k_shade_shadow_step<false><<<grid_wide, WF_BLOCK, 0, st>>>(soa, cur);
        k_shadow_step_end<<<1, 1, 0, st>>>(soa, cur);
        stats.kernel_launches += 3;
        cur ^= 1;
      }
      CUDA_TRY(ctx, cudaGetLastError());
    }
    else if (transparent_shadows) {
      /* more surfaces allowed than is worth queueing blind: ask after every step */
      CUDA_TRY(ctx, cudaMemcpyAsync(pool->h_counters, soa.counters, 64, cudaMemcpyDeviceToHost,
                                    st));
      CUDA_TRY(ctx, cudaStreamSynchronize(st));
      stats.host_syncs += 1;
      int cur = 0;
      while (pool->h_counters->n_ts[cur] != 0) {
        k_intersect_shadow_step<<<grid_trace, TRACE_BLOCK, 0, st>>>(soa, cur, refill);
        if (passes)
          k_shade_shadow_step<true, true><<<grid_wide, WF_BLOCK, 0, st>>>(soa, cur);
        else if (svm_ext)
          k_shade_shadow_step<true><<<grid_wide, WF_BLOCK, 0, st>>>(soa, cur);
        else
          k_shade_shadow_step<false><<<grid_wide, WF_BLOCK, 0, st>>>(soa, cur);
        k_shadow_step_end<<<1, 1, 0, st>>>(soa, cur);
        stats.kernel_launches += 3;
        CUDA_TRY(ctx, cudaMemcpyAsync(pool->h_counters, soa.counters, 64,
                                      cudaMemcpyDeviceToHost, st));
        CUDA_TRY(ctx, cudaStreamSynchronize(st));
        stats.host_syncs += 1;
        CUDA_TRY(ctx, cudaGetLastError());
        cur ^= 1;
      }
    }
    k_iteration_end<<<8, 256, 0, st>>>(soa, num_keys, it);
    /* the reference copies the whole ray_state array every 16 iterations
     * (device_split_kernel.cpp:302-318); here 64 bytes per bounce */
    CUDA_TRY(ctx, cudaMemcpyAsync(pool->h_ring + (size_t)(it % WF_RING) * WF_RING_BYTES,
                                  soa.counters, WF_RING_BYTES, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaEventRecord(ev[4], st));
    CUDA_TRY(ctx, cudaGetLastError());
    return B200_OK;
  };
  /* wait for iteration `it` to have finished (later ones may be queued behind it) and
   * read what it left: queue length, flags, phase times */
  auto harvest_iteration = [&](int it, bool had_work, unsigned int *n_active,
                               unsigned int *flags) -> int {
    cudaEvent_t *ev = pool->ring_ev[it % WF_RING];
    CUDA_TRY(ctx, cudaEventSynchronize(ev[4]));
    stats.host_waits += 1;
    const WFCounters *h = (const WFCounters *)(pool->h_ring +
                                               (size_t)(it % WF_RING) * WF_RING_BYTES);
    *n_active = h->n_active;
    *flags = h->flags;
    if (had_work) {
      float a = 0.0f, b = 0.0f, d = 0.0f;
      cudaEventElapsedTime(&a, ev[0], ev[1]);
      cudaEventElapsedTime(&b, ev[1], ev[2]);
      cudaEventElapsedTime(&d, ev[2], ev[3]);
      closest_ms += a;
      shade_ms += b;
      shadow_ms += d;
      stats.closest_launches += 1;
      stats.shadow_launches += 1;
      stats.iterations += 1;
    }
    return B200_OK;
  };

  /* adaptive sampling (adaptive.cuh): Session's AdaptiveSampling (session.cpp:1077-1080)
   * read from KernelData; a batch never crosses a filter point */
  const int adaptive_aux = kd_host<int>(ctx, KD_FILM_PASS_ADAPTIVE_AUX_BUFFER);
  const bool adaptive = adaptive_aux != 0 &&
                        kd_host<int>(ctx, KD_INT_SAMPLING_PATTERN) == CY_SAMPLING_PATTERN_PMJ;
  const int adaptive_step = std::max(1, kd_host<int>(ctx, KD_INT_ADAPTIVE_STEP));
  const int adaptive_min = kd_host<int>(ctx, KD_INT_ADAPTIVE_MIN_SAMPLES);
  AdaptiveTile atile;
  atile.film = (float *)tile->buffer;
  atile.x = tile->x, atile.y = tile->y, atile.w = tile->w, atile.h = tile->h;
  atile.offset = tile->offset, atile.stride = tile->stride, atile.pass_stride = pass_stride;
  atile.aux = adaptive_aux;
  atile.sample_count = kd_host<int>(ctx, KD_FILM_PASS_SAMPLE_COUNT);
  bool adaptive_done = false;

  /* bands of rows so that one sample of a band fits the pool */
  const int band_h = (int)std::max<size_t>(1, std::min<size_t>((size_t)tile->h,
                                                               pool->capacity / (size_t)tile->w));
  if (adaptive && band_h < tile->h)
    return fail(ctx, B200_ERR_UNSUPPORTED,
                "adaptive sampling needs one sample of the whole tile to fit the path pool");
  for (int by = 0; by < tile->h; by += band_h) {
    const int bh = std::min(band_h, tile->h - by);
    const size_t npix = (size_t)tile->w * bh;
    int spb = (int)std::max<size_t>(1, pool->capacity / npix);
    if (adaptive) { /* a divisor of the step, so that batches end where the filter runs */
      spb = std::min(spb, adaptive_step);
      while (adaptive_step % spb != 0)
        spb--;
    }
    for (int s0 = 0; s0 < tile->num_samples && !adaptive_done; s0 += spb) {
      if ((cancel && *cancel) || (ctx->cancel_fn && ctx->cancel_fn(ctx->cancel_user)))
        return fail(ctx, B200_ERR_CANCELLED, "cancelled");
      BatchParams bp;
      bp.x = tile->x;
      bp.y = tile->y + by;
      bp.w = tile->w;
      bp.h = bh;
      bp.sample0 = tile->start_sample + s0;
      bp.nsamples = std::min(spb, tile->num_samples - s0);
      bp.offset = tile->offset;
      bp.stride = tile->stride;
      bp.num_keys = num_keys;
      bp.count_stats = count;
      bp.adaptive_film = adaptive ? (const float *)tile->buffer : nullptr;
      bp.pass_stride = pass_stride;
      bp.adaptive_aux = adaptive_aux;

      /* kernels with a wide variant: the lean multiscatter one and the full one (not the
       * lean GGX kernel - the cap lost on every scene measured - nor the passes kernel) */
      auto probe_kind = [&]() { return passes ? -1 : (svm_ext ? 1 : (multiscatter ? 0 : -1)); };
      auto pick_shape = [&]() {
        const int kind = probe_kind();
        if (kind < 0)
          shade_wide = 0;
        else if (ctx->opt_shade_wide >= 0)
          shade_wide = (int)std::min<int64_t>(ctx->opt_shade_wide, 2);
        else if (ctx->shade_probe[kind].choice >= 0)
          shade_wide = ctx->shade_probe[kind].choice;
        else
          shade_wide = (int)(ctx->shade_probe[kind].batches++ % 3);
      };
      pick_shape();
      const float shade_ms_before = shade_ms;
      for (int attempt = 0;; attempt++) {
        PathSoA soa = pool->soa;
        soa.sort_tiles = (num_keys <= SORT_TILE_MAX_KEYS && ctx->opt_sort_tiles) ? 1 : 0;
        soa.debug = ctx->d_debug;
        soa.debug_slot = (int)ctx->opt_debug_slot;
        CUDA_TRY(ctx, cudaMemsetAsync(soa.counters, 0, sizeof(WFCounters), st));
        if (passes)
          CUDA_TRY(ctx, cudaMemsetAsync(soa.pass, 0,
                                        (size_t)bp.w * bp.h * bp.nsamples * PASS_WORDS *
                                            sizeof(float),
                                        st));
        if (attempt == 0) {
          /* the Sobol points of this batch's samples, for every dimension a path can reach */
          SobolTable sobol = {nullptr, bp.sample0, 0u, 0u};
          const HostArray *lut = find_global(ctx, "__sample_pattern_lut");
          if (kd_host<int>(ctx, KD_INT_SAMPLING_PATTERN) == CY_SAMPLING_PATTERN_SOBOL && lut) {
            const unsigned int reach =
                CY_PRNG_BASE_NUM + (unsigned int)(kd_host<int>(ctx, KD_INT_MAX_BOUNCE) +
                                                  kd_host<int>(ctx, KD_INT_TRANSPARENT_MAX_BOUNCE) +
                                                  3) * CY_PRNG_BOUNCE_NUM;
            const unsigned int nd = std::min<unsigned int>(reach, (unsigned int)(lut->bytes / 128));
            if (nd > 0 && (size_t)nd * bp.nsamples <= SOBOL_TABLE_MAX) {
              sobol.tab = pool->sobol_tab;
              sobol.ns = (unsigned int)bp.nsamples;
              sobol.nd = nd;
              k_sobol_table<<<64, 256, 0, st>>>(pool->sobol_tab, bp.sample0, sobol.ns, nd);
              stats.kernel_launches += 1;
            }
          }
          CUDA_TRY(ctx, cudaMemcpyToSymbolAsync(g_sobol, &sobol, sizeof(sobol), 0,
                                                cudaMemcpyHostToDevice, st));
        }
        k_init_from_camera<<<grid_wide, WF_BLOCK, 0, st>>>(soa, bp);
        stats.kernel_launches += 1;

        const int max_iterations = 4096;
        unsigned int n_active = 1u, prev_n_active = 1u, flags = 0u;
        int harvested = 0;
        for (int it = 0; it < max_iterations && n_active != 0u; it++) {
          rc = enqueue_iteration(soa, it);
          if (rc)
            return rc;
          std::swap(soa.q_active, soa.q_next);
          std::swap(soa.ray_P_t, soa.nray_P_t);
          std::swap(soa.ray_D, soa.nray_D);
          while (harvested <= it - lag && n_active != 0u) {
            prev_n_active = n_active;
            rc = harvest_iteration(harvested, prev_n_active != 0u, &n_active, &flags);
            if (rc)
              return rc;
            harvested++;
          }
        }
        /* iterations queued behind the one that emptied the queue find nothing to do */
        if (flags & WF_FLAG_TRACE_OVERFLOW) {
          const unsigned int zero = 0;
          CUDA_TRY(ctx, cudaMemcpyToSymbol(g_trace_overflow, &zero, sizeof(zero)));
          return fail(ctx, B200_ERR_UNSUPPORTED,
                      "BVH8 traversal stack overflow: hits of this batch are not reliable");
        }
        if (!svm_ext && (flags & WF_FLAG_SCOPE_MISS)) {
          /* Nothing of this batch has reached the film yet.  A lean kernel met a shader
           * only the full interpreter runs: the batch is thrown away and traced again with
           * the full kernels, and the context stays on them until the program changes. */
          const unsigned int zero = 0;
          CUDA_TRY(ctx, cudaMemcpyToSymbol(g_svm_scope_miss, &zero, sizeof(zero)));
          if (attempt > 0)
            return fail(ctx, B200_ERR_UNSUPPORTED, "SVM scope miss in the full kernels");
          svm_ext = true;
          ctx->force_svm_ext = true;
          pick_shape();
          continue;
        }
        if (adaptive)
          k_film_accumulate_adaptive<<<grid_wide, WF_BLOCK, 0, st>>>(
              soa, bp, (float *)tile->buffer, pass_stride, adaptive_aux, atile.sample_count);
        else if (passes)
          k_film_accumulate_passes<<<grid_wide, WF_BLOCK, 0, st>>>(soa, bp,
                                                                   (float *)tile->buffer,
                                                                   pass_stride);
        else
          k_film_accumulate<<<grid_wide, WF_BLOCK, 0, st>>>(soa, bp, (float *)tile->buffer,
                                                            pass_stride, pass_combined);
        stats.kernel_launches += 1;
        CUDA_TRY(ctx, cudaMemcpyAsync(pool->h_batch + (size_t)batch_slot * WF_BATCH_STAT_BYTES,
                                      soa.counters, WF_BATCH_STAT_BYTES, cudaMemcpyDeviceToHost,
                                      st));
        CUDA_TRY(ctx, cudaGetLastError());
        stats.batches += 1;
        if (probe_kind() >= 0 && ctx->opt_shade_wide < 0 &&
            ctx->shade_probe[probe_kind()].choice < 0 && attempt == 0) {
          /* every iteration with work has been harvested when the bounce loop ends: the
           * batch's shading time is complete.  Decide once both shapes have shaded 4 Mi
           * paths; until then (small tiles) keep alternating. */
          b200_ctx::ShadeProbe &pr = ctx->shade_probe[probe_kind()];
          const int v = shade_wide;
          /* the first batch of each shape pays for loading its kernel: not counted */
          if (pr.batches > 3) {
            pr.ms[v] += (double)(shade_ms - shade_ms_before);
            pr.paths[v] += (double)npix * bp.nsamples;
          }
          const double enough = (double)(1 << 22);
          if (pr.paths[0] >= enough && pr.paths[1] >= enough && pr.paths[2] >= enough) {
            int best = 0;
            for (int k = 1; k < 3; k++)
              if (pr.ms[k] / pr.paths[k] < pr.ms[best] / pr.paths[best])
                best = k;
            pr.choice = best;
          }
        }
        if (++batch_slot == WF_BATCH_SLOTS) {
          rc = sum_batch_stats();
          if (rc)
            return rc;
        }
        if (adaptive) {
          /* AdaptiveSampling::need_filter (device_task.cpp:184-192) for the last sample of
           * the batch: convergence test of every pixel, then the two dilation sweeps; the
           * tile is finished when no pixel is left sampling */
          const int last = bp.sample0 + bp.nsamples - 1;
          if (last > adaptive_min && (last & (adaptive_step - 1)) == (adaptive_step - 1)) {
            unsigned int *any = &soa.counters->adaptive_any;
            CUDA_TRY(ctx, cudaMemsetAsync(any, 0, sizeof(unsigned int), st));
            k_adaptive_stopping<<<grid_wide, WF_BLOCK, 0, st>>>(atile, last);
            k_adaptive_filter_x<<<(tile->h + 63) / 64, 64, 0, st>>>(atile, any);
            k_adaptive_filter_y<<<(tile->w + 63) / 64, 64, 0, st>>>(atile, any);
            stats.kernel_launches += 3;
            unsigned int h_any = 1;
            CUDA_TRY(ctx, cudaMemcpyAsync(&pool->h_counters->adaptive_any, any, sizeof(unsigned int),
                                          cudaMemcpyDeviceToHost, st));
            CUDA_TRY(ctx, cudaStreamSynchronize(st));
            stats.host_syncs += 1;
            h_any = pool->h_counters->adaptive_any;
            if (!h_any)
              adaptive_done = true;
          }
        }
        break;
      } /* attempt */
    }
  }
  if (adaptive) {
    /* adaptive_sampling_post (device_cpu.cpp:864-886): tile.sample is the end of the task
     * whether the tile stopped early or not */
    k_adaptive_scale_samples<<<grid_wide, WF_BLOCK, 0, st>>>(atile, tile->start_sample,
                                                             tile->start_sample + tile->num_samples);
    stats.kernel_launches += 1;
  }
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev6, st));
  rc = sum_batch_stats();
  if (rc)
    return rc;
  {
    float total = 0.0f;
    cudaEventElapsedTime(&total, ctx->ev5, ctx->ev6);
    stats.device_ms = total;
    stats.closest_ms = closest_ms;
    stats.shade_ms = shade_ms;
    stats.shadow_ms = shadow_ms;
  }
  stats.svm_extended = svm_ext ? 1 : 0;
  {
    const int kind = passes ? -1 : (svm_ext ? 1 : (multiscatter ? 0 : -1));
    stats.shade_wide = kind < 0 ? 0 :
                        (ctx->opt_shade_wide >= 0 ? std::min<int64_t>(ctx->opt_shade_wide, 2) :
                                                     (int64_t)ctx->shade_probe[kind].choice);
  }
  ctx->stats = stats;
  return B200_OK;
}

/* DeviceTask::SHADER with SHADER_EVAL_BACKGROUND (CUDADevice::shader,
 * device_cuda_impl.cpp:2019-2093): the world shader evaluated for `shader_w` directions
 * starting at `shader_x`.  Runs before the scene's BVH exists (the light manager updates
 * ahead of the geometry), so only the constant block's shader arrays are needed. */
int b200_shader_eval_background(b200_ctx *ctx, uint64_t input, uint64_t output, int shader_x,
                                int shader_w)
{
  if (!ctx || !input || !output || shader_w < 0)
    return B200_ERR_INVALID;
  if (!ctx->have_data)
    return fail(ctx, B200_ERR_NOT_READY, "KernelData not uploaded");
  DeviceUse device_use(ctx);
  int rc = prepare_scene(ctx, true);
  if (rc)
    return rc;
  ctx->scene_dirty = true; /* the scope check is still due at the first render */
  DeviceGuard guard(ctx->ordinal);
  const int grid = std::max(1, std::min(launch_grid(ctx, 8), (shader_w + WF_BLOCK - 1) / WF_BLOCK));
  k_background_evaluate<true><<<grid, WF_BLOCK, 0, ctx->stream>>>(
      (const uint4 *)input, (float4 *)output, shader_x, shader_w);
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}

int b200_film_convert(b200_ctx *ctx, uint64_t film, uint64_t rgba, int half_float,
                      float sample_scale, int x, int y, int w, int h, int offset, int stride)
{
  if (!ctx || !film || !rgba)
    return B200_ERR_INVALID;
  if (!ctx->have_data)
    return fail(ctx, B200_ERR_NOT_READY, "KernelData not uploaded");
  DeviceGuard guard(ctx->ordinal);
  if (w <= 0 || h <= 0)
    return B200_OK;
  if (kd_host<int>(ctx, KD_FILM_DISPLAY_PASS_COMPONENTS) != 4 ||
      kd_host<int>(ctx, KD_FILM_DISPLAY_DIVIDE_PASS_STRIDE) != -1)
    return fail(ctx, B200_ERR_UNSUPPORTED,
                "film_convert: only the 4-component combined display pass is implemented");
  FilmDisplay fd;
  fd.pass_stride = kd_host<int>(ctx, KD_FILM_PASS_STRIDE);
  fd.display_pass_stride = kd_host<int>(ctx, KD_FILM_DISPLAY_PASS_STRIDE);
  fd.use_pass_alpha = kd_host<int>(ctx, KD_FILM_USE_DISPLAY_PASS_ALPHA);
  fd.use_exposure = kd_host<int>(ctx, KD_FILM_USE_DISPLAY_EXPOSURE);
  fd.exposure = kd_host<float>(ctx, KD_FILM_EXPOSURE);
  dim3 block(16, 16), grid((w + 15) / 16, (h + 15) / 16);
  k_film_convert<<<grid, block, 0, ctx->stream>>>((const float *)film, (void *)rgba, half_float,
                                                  sample_scale, x, y, w, h, offset, stride, fd);
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}

int b200_film_reduce(b200_ctx **ctxs, int n, const uint64_t *films, size_t n_floats)
{
  if (!ctxs || n <= 0 || !films)
    return B200_ERR_INVALID;
  if (n == 1)
    return B200_OK;
  if (n_floats % 4)
    return fail(ctxs[0], B200_ERR_INVALID, "film size must be a multiple of 4 floats");
  b200_ctx *root = ctxs[0];
  DeviceGuard guard(root->ordinal);
  /* the staging film lives with the root context and is reused by every call */
  if (root->reduce_tmp_bytes < n_floats * sizeof(float)) {
    CUDA_TRY(root, cudaStreamSynchronize(root->stream));
    if (root->reduce_tmp)
      cudaFree(root->reduce_tmp);
    root->reduce_tmp = nullptr;
    root->reduce_tmp_bytes = 0;
    CUDA_TRY(root, cudaMalloc(&root->reduce_tmp, n_floats * sizeof(float)));
    root->reduce_tmp_bytes = n_floats * sizeof(float);
  }
  void *tmp = root->reduce_tmp;
  for (int i = 1; i < n; i++) {
    /* the peer's film must be complete before it is copied */
    {
      DeviceGuard peer(ctxs[i]->ordinal);
      CUDA_TRY(ctxs[i], cudaStreamSynchronize(ctxs[i]->stream));
    }
    /* peer copy over NVLink, then one vectorised add on the root */
    cudaError_t e = cudaMemcpyPeerAsync(tmp, root->ordinal, (const void *)films[i],
                                        ctxs[i]->ordinal, n_floats * sizeof(float), root->stream);
    if (e != cudaSuccess)
      return fail(root, B200_ERR_CUDA, std::string("peer copy: ") + cudaGetErrorString(e));
    k_film_add<<<launch_grid(root, 4), 256, 0, root->stream>>>((float4 *)films[0],
                                                               (const float4 *)tmp, n_floats / 4);
  }
  cudaError_t e = cudaStreamSynchronize(root->stream);
  if (e != cudaSuccess)
    return fail(root, B200_ERR_CUDA, std::string("film reduce: ") + cudaGetErrorString(e));
  return B200_OK;
}

} /* extern "C" */

#endif /* B200_WAVEFRONT_CUH */

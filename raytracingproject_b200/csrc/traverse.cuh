/* traverse.cuh - stack traversal of the compressed BVH8 (bvh8.h) for one ray.
 *
 * Semantics follow the reference's BVH2 walker, so that hit records mean the
 * same thing:
 *   - scene_intersect / bvh_intersect       kernel/bvh/bvh.h:154-237,
 *                                            kernel/bvh/bvh_traversal.h:34-227
 *   - triangle test (generic scalar branch)  util/util_math_intersect.h:88-195,
 *                                            geom/geom_triangle_intersect.h:25-72
 *   - instance push / pop                    geom/geom_object.h:412-460
 *   - shadow early-out (PATH_RAY_SHADOW_OPAQUE) bvh_traversal.h:144-147
 * What is new is the structure walked: an 8-wide quantised node fetched with
 * five 128-bit loads, octant-ordered child visiting with a hit bit-mask, and
 * (node group, triangle group) stack entries after Ylitie et al. 2017.
 *
 * The triangle test is written without FMA contraction (the file is compiled
 * with -fmad=false) in the reference's operation order: u, v, t and the
 * accept/reject decision are bit-identical to the oracle's for the same ray in
 * the same space.  The box test uses explicit fmaf and is only conservative.
 */
#ifndef B200_TRAVERSE_CUH
#define B200_TRAVERSE_CUH

#include <cuda_fp16.h>

#include "device_scene.cuh"

#define BVH8_STACK_SIZE 64
#define BVH8_SENTINEL 0xffffffffu
#ifndef PREFETCH
#  define PREFETCH 0
#endif
#ifndef CHUNK_DIV
#  define CHUNK_DIV 4u
#endif
#ifndef CHUNK_MAX
#  define CHUNK_MAX 512u
#endif

struct TraceHit {
  float t, u, v;
  int prim;   /* -1 = miss */
  int object; /* -1 = OBJECT_NONE */
};

/* Set when a traversal had to drop a stack entry (the stack is sized for the depth the
 * BVH8 builder accepts, so this cannot happen for a tree it produced; if it ever does the
 * host refuses the result instead of returning hits that may be wrong). */
__device__ unsigned int g_trace_overflow;

struct TraceCounters {
  uint32_t nodes, tris, instances;
};

/* geom_object.h:414-420 */
CY_DEV f3 bvh_clamp_direction(f3 dir)
{
  const float ooeps = 8.271806E-25f;
  return mk3((fabsf(dir.x) > ooeps) ? dir.x : copysignf(ooeps, dir.x),
             (fabsf(dir.y) > ooeps) ? dir.y : copysignf(ooeps, dir.y),
             (fabsf(dir.z) > ooeps) ? dir.z : copysignf(ooeps, dir.z));
}

CY_DEV uint32_t sign_extend_s8x4(uint32_t x)
{
  uint32_t r;
  asm("prmt.b32 %0, %1, 0x0, 0x0000BA98;" : "=r"(r) : "r"(x));
  return r;
}

/* bytes 2*jj and 2*jj+1 of x as the floats (1024 + b0, 1024 + b1) */
CY_DEV float2 bytes_to_float2(uint32_t x, int jj)
{
  const uint32_t pair = __byte_perm(x, 0x64646464u, jj ? 0x4342u : 0x4140u);
  return __half22float2(*reinterpret_cast<const __half2 *>(&pair));
}

/* util_math_intersect.h:88-195, scalar branch.  Returns true and u,v,t on a hit
 * closer than ray_t.  Exact arithmetic (cymath.cuh, x-functions): the signs of U, V, W,
 * the depth test and u, v, t are the reference's bits, whatever the build flags. */
CY_DEV bool ray_triangle_intersect(
    f3 P, f3 dir, float ray_t, f3 tri_a, f3 tri_b, f3 tri_c, float *isect_u, float *isect_v,
    float *isect_t)
{
  const f3 v0 = xsub3(tri_c, P);
  const f3 v1 = xsub3(tri_a, P);
  const f3 v2 = xsub3(tri_b, P);

  const f3 e0 = xsub3(v2, v0);
  const f3 e1 = xsub3(v0, v1);
  const f3 e2 = xsub3(v1, v2);

  const float U = xdot(xcross(xadd3(v2, v0), e0), dir);
  const float V = xdot(xcross(xadd3(v0, v1), e1), dir);
  const float W = xdot(xcross(xadd3(v1, v2), e2), dir);

  const float minUVW = fminf(U, fminf(V, W));
  const float maxUVW = fmaxf(U, fmaxf(V, W));
  if (minUVW < 0.0f && maxUVW > 0.0f)
    return false;

  const f3 Ng1 = xcross(e1, e0);
  const f3 Ng = xadd3(Ng1, Ng1);
  const float den = xdot(Ng, dir);
  if (den == 0.0f)
    return false;

  const float T = xdot(v0, Ng);
  const int sign_den = (__float_as_int(den) & 0x80000000);
  const float sign_T = xor_signmask(T, sign_den);
  if ((sign_T < 0.0f) || (sign_T > xmul(ray_t, xor_signmask(den, sign_den))))
    return false;

  const float inv_den = xdiv(1.0f, den);
  *isect_u = xmul(U, inv_den);
  *isect_v = xmul(V, inv_den);
  *isect_t = xmul(T, inv_den);
  return true;
}

CY_DEV float rcp_approx(float x)
{
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

struct RaySpace {
  f3 P, dir, idir;
  uint32_t oct_inv4;
};

CY_DEV void ray_space_setup(RaySpace &rs, f3 P, f3 D)
{
  rs.P = P;
  rs.dir = bvh_clamp_direction(D);
  /* 1/dir only feeds the (conservative, padded) box test, never the triangle test: the
   * one-instruction approximate reciprocal (about 1 ulp) instead of three IEEE divisions */
  rs.idir = mk3(rcp_approx(rs.dir.x), rcp_approx(rs.dir.y), rcp_approx(rs.dir.z));
  rs.oct_inv4 = ((rs.dir.x < 0.0f) ? 0u : 0x04040404u) | ((rs.dir.y < 0.0f) ? 0u : 0x02020202u) |
                ((rs.dir.z < 0.0f) ? 0u : 0x01010101u);
}

/* Intersects the 8 quantised child boxes of node `node_index`; returns the hit
 * mask: bits 24..31 inner children (ordered front to back for this octant),
 * bits 0..23 leaf records relative to prim_base. */
CY_DEV uint32_t bvh8_node_intersect(const RaySpace &rs,
                                    float tmax,
                                    uint32_t node_index,
                                    uint32_t &child_base,
                                    uint32_t &prim_base,
                                    uint32_t &imask)
{
  const uint4 *np = g_scene.nodes + (size_t)node_index * 5;
  const uint4 n0 = __ldg(np + 0);
  const uint4 n1 = __ldg(np + 1);
  const uint4 n2 = __ldg(np + 2);
  const uint4 n3 = __ldg(np + 3);
  const uint4 n4 = __ldg(np + 4);

  const uint32_t e_imask = n0.w;
  imask = e_imask >> 24;
  child_base = n1.x;
  prim_base = n1.y;

  /* Plane distances  t = q * adj + org  for the quantised byte q, without integer ->
   * float conversions (I2F runs on the XU pipe at a quarter of the FP32 rate and was
   * the busiest pipe of this kernel): two bytes at a time are dropped under the
   * exponent byte 0x64 by one byte-permute, which makes the fp16 pair (1024 + q0,
   * 1024 + q1) exactly; fp16 -> fp32 is an FMA-pipe op, and the 1024 is folded into
   * the constant term.  The test is kept conservative by an ABSOLUTE pad of ~8 ulp of
   * the largest intermediate (covers the rounding of org, adj, the approximate 1/dir and
   * the fma). */
  const float adjx = __uint_as_float((e_imask & 0xffu) << 23) * rs.idir.x;
  const float adjy = __uint_as_float(((e_imask >> 8) & 0xffu) << 23) * rs.idir.y;
  const float adjz = __uint_as_float(((e_imask >> 16) & 0xffu) << 23) * rs.idir.z;
  const float orgx = fmaf(-1024.0f, adjx, (__uint_as_float(n0.x) - rs.P.x) * rs.idir.x);
  const float orgy = fmaf(-1024.0f, adjy, (__uint_as_float(n0.y) - rs.P.y) * rs.idir.y);
  const float orgz = fmaf(-1024.0f, adjz, (__uint_as_float(n0.z) - rs.P.z) * rs.idir.z);
  const float padx = fmaf(1280.0f, fabsf(adjx), fabsf(orgx)) * 4.8e-7f;
  const float pady = fmaf(1280.0f, fabsf(adjy), fabsf(orgy)) * 4.8e-7f;
  const float padz = fmaf(1280.0f, fabsf(adjz), fabsf(orgz)) * 4.8e-7f;
  const float onx = orgx - padx, ony = orgy - pady, onz = orgz - padz; /* near: earlier */
  const float ofx = orgx + padx, ofy = orgy + pady, ofz = orgz + padz; /* far: later */

  const bool negx = rs.dir.x < 0.0f, negy = rs.dir.y < 0.0f, negz = rs.dir.z < 0.0f;
  uint32_t hitmask = 0;

#pragma unroll
  for (int h = 0; h < 2; h++) {
    const uint32_t meta4 = h ? n1.w : n1.z;
    const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
    const uint32_t inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
    const uint32_t bit_index4 = (meta4 ^ (rs.oct_inv4 & inner_mask4)) & 0x1f1f1f1fu;
    const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;

    const uint32_t qlox = h ? n2.y : n2.x, qloy = h ? n2.w : n2.z;
    const uint32_t qloz = h ? n3.y : n3.x, qhix = h ? n3.w : n3.z;
    const uint32_t qhiy = h ? n4.y : n4.x, qhiz = h ? n4.w : n4.z;

    const uint32_t xn = negx ? qhix : qlox, xf = negx ? qlox : qhix;
    const uint32_t yn = negy ? qhiy : qloy, yf = negy ? qloy : qhiy;
    const uint32_t zn = negz ? qhiz : qloz, zf = negz ? qloz : qhiz;

#pragma unroll
    for (int jj = 0; jj < 2; jj++) {
      const float2 fxn = bytes_to_float2(xn, jj), fyn = bytes_to_float2(yn, jj);
      const float2 fzn = bytes_to_float2(zn, jj), fxf = bytes_to_float2(xf, jj);
      const float2 fyf = bytes_to_float2(yf, jj), fzf = bytes_to_float2(zf, jj);
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const int j = 2 * jj + k;
        const float tnx = fmaf(k ? fxn.y : fxn.x, adjx, onx);
        const float tny = fmaf(k ? fyn.y : fyn.x, adjy, ony);
        const float tnz = fmaf(k ? fzn.y : fzn.x, adjz, onz);
        const float tfx = fmaf(k ? fxf.y : fxf.x, adjx, ofx);
        const float tfy = fmaf(k ? fyf.y : fyf.x, adjy, ofy);
        const float tfz = fmaf(k ? fzf.y : fzf.x, adjz, ofz);
        const float cmin = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f));
        const float cmax = fminf(fminf(tfx, tfy), fminf(tfz, tmax));
        if (cmin <= cmax) {
          const uint32_t bits = (child_bits4 >> (8 * j)) & 0xffu;
          const uint32_t index = (bit_index4 >> (8 * j)) & 0xffu;
          hitmask |= bits << index;
        }
      }
    }
  }
  return hitmask;
}

/* The (node group, leaf group) stack of one lane.  SMEM_STACK > 0 keeps the lowest
 * SMEM_STACK entries in shared memory (a column per thread, entry k of thread t at
 * s[k * TRACE_BLOCK + t]: conflict-free 64-bit accesses), deeper ones in local memory.
 * A/B on B200: profiles/r02o_smem_stack_ab.txt. */
#ifndef SMEM_STACK
#  define SMEM_STACK 0
#endif
struct TraceStack {
  uint2 *local;
#if SMEM_STACK > 0
  uint2 *shared; /* this thread's column */
#endif
  __device__ __forceinline__ uint2 get(int i) const
  {
#if SMEM_STACK > 0
    if (i < SMEM_STACK)
      return shared[i * 128];
    return local[i - SMEM_STACK];
#else
    return local[i];
#endif
  }
  __device__ __forceinline__ void set(int i, uint2 e) const
  {
#if SMEM_STACK > 0
    if (i < SMEM_STACK)
      shared[i * 128] = e;
    else
      local[i - SMEM_STACK] = e;
#else
    local[i] = e;
#endif
  }
};

/* Resumable traversal of one ray, cut into the three phases the persistent warp
 * driver below interleaves:
 *   node_phase()   one BVH8 node (or a leaf group popped earlier) -> pending leaf
 *                  records in Gt
 *   leaf phase     WARP-COOPERATIVE (trace_persistent): the pending records of all
 *                  lanes are pooled and spread over the 32 lanes, so the triangle
 *                  test runs once per step with many lanes on instead of once per
 *                  record with the two or three lanes that happen to own one
 *   finish_step()  instance push for records the leaf phase flagged, stack pop
 * Kernels keep one Traversal per lane and refill finished lanes from their queue
 * (persistent threads with dynamic fetch).
 *
 * ANY_HIT = false: closest hit, `hit` is the Intersection.  ANY_HIT = true:
 * occlusion (shadow early-out), only hit.prim >= 0 is meaningful. */
template<bool ANY_HIT, bool COUNT> struct Traversal {
  /* the (node group, leaf group) stack is a separate local array handed to the
   * phases: keeping it out of the struct lets every scalar member live in a register */
  int sp;
  RaySpace rs;
  float tmax; /* current limit = distance of the closest hit so far (hit.t is filled at the end) */
  uint32_t visibility;
  TraceHit hit;
  int cur_object; /* OBJECT_NONE (-1) while in world space */
  bool inst_hit;
  /* The world-space ray is NOT kept in registers: the instance push / pop re-read it
   * from the ray queue; the world-space limit and the direction scale saved at an
   * instance push live in the stack entry under the sentinel. */
  uint2 G;  /* node group: (child base, hits << 24 | imask) */
  uint2 Gt; /* pending leaf group: (record base, record hit bits) */

  __device__ __forceinline__ void start(f3 P, f3 D, float tmax_, uint32_t visibility_)
  {
    sp = 0;
    tmax = tmax_;
    visibility = visibility_;
    ray_space_setup(rs, P, D);
    hit.u = 0.0f;
    hit.v = 0.0f;
    hit.prim = -1;
    hit.object = -1;
    cur_object = -1;
    inst_hit = false;
    G = make_uint2(g_scene.bvh_root, 0x80000000u);
    Gt = make_uint2(0u, 0u);
  }

  __device__ __forceinline__ void push(const TraceStack &stack, uint2 e)
  {
    if (sp < BVH8_STACK_SIZE)
      stack.set(sp++, e);
    else
      g_trace_overflow = 1u;
  }

  __device__ __forceinline__ void node_phase(const TraceStack &stack, TraceCounters &cnt)
  {
    if (G.y & 0xff000000u) {
      const uint32_t hits_imask = G.y;
      const uint32_t child_bit_index = 31u - (uint32_t)__clz((int)hits_imask);
      G.y &= ~(1u << child_bit_index);
      if (G.y & 0xff000000u)
        push(stack, G);
      const uint32_t slot_index = (child_bit_index - 24u) ^ (rs.oct_inv4 & 0xffu);
      const uint32_t relative_index = __popc(hits_imask & ~(0xffffffffu << slot_index) & 0xffu);
      const uint32_t node_index = G.x + relative_index;

      uint32_t child_base, prim_base, imask;
      const uint32_t hitmask = bvh8_node_intersect(rs, tmax, node_index, child_base, prim_base,
                                                   imask);
      if (COUNT)
        cnt.nodes++;
      G = make_uint2(child_base, (hitmask & 0xff000000u) | imask);
      Gt = make_uint2(prim_base, hitmask & 0x00ffffffu);
      /* the child this ray visits next is known now: pull its 80 bytes towards L1
       * while the leaf phase runs */
      if (PREFETCH && (G.y & 0xff000000u)) {
        const uint32_t cbi = 31u - (uint32_t)__clz((int)G.y);
        const uint32_t slot = (cbi - 24u) ^ (rs.oct_inv4 & 0xffu);
        const uint32_t rel = __popc(G.y & ~(0xffffffffu << slot) & 0xffu);
        const char *np = (const char *)(g_scene.nodes + (size_t)(G.x + rel) * 5);
        asm volatile("prefetch.global.L1 [%0];" ::"l"(np));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(np + 64));
      }
    }
    else {
      Gt = G;
      G = make_uint2(0u, 0u);
    }
  }

  /* a triangle of this ray's pending group was hit (found by whichever lane ran the test) */
  __device__ __forceinline__ void accept(float t, float u, float v, int prim)
  {
    hit.prim = prim;
    hit.object = cur_object;
    hit.u = u;
    hit.v = v;
    tmax = t;
    if (cur_object >= 0)
      inst_hit = true;
  }

  /* `inst_bits`: records of the group at Gt.x the leaf phase found to be visible
   * instances.  Returns true when the traversal is complete. */
  template<class Job>
  __device__ __forceinline__ bool finish_step(
      const TraceStack &stack, uint32_t inst_bits, TraceCounters &cnt, const Job &job,
      unsigned int qi)
  {
    if (inst_bits != 0u) {
      /* instance push - geom_object.h:427-443 */
      const uint32_t bit = (uint32_t)__ffs((int)inst_bits) - 1u;
      inst_bits &= inst_bits - 1u;
      const float4 ra = __ldg(g_scene.records + (size_t)(Gt.x + bit) * 3);
      const int object = ~__float_as_int(ra.w);
      if (COUNT)
        cnt.instances++;
      if (G.y & 0xff000000u)
        push(stack, G);
      if (inst_bits != 0u)
        push(stack, make_uint2(Gt.x, inst_bits));

      /* exact arithmetic: the object-space ray and the rescaled limit decide hits */
      const tfm34 itfm = object_itfm(object);
      float len;
      const f3 oP = xtransform_point(itfm, mk3(__ldg(job.ray_P(qi))));
      const f3 oD = xnormalize_len(xtransform_direction(itfm, mk3(__ldg(job.ray_D(qi)))), &len);
      ray_space_setup(rs, oP, oD);
      push(stack, make_uint2(__float_as_uint(tmax), __float_as_uint(len)));
      push(stack, make_uint2(BVH8_SENTINEL, 0u));
      inst_hit = false;
      if (tmax != FLT_MAX)
        tmax = xmul(tmax, len);
      cur_object = object;

      G = make_uint2(__float_as_uint(ra.x), 0x80000000u);
      return false;
    }

    /* pop */
    if ((G.y & 0xff000000u) == 0u) {
      while (true) {
        if (sp == 0)
          return true;
        G = stack.get(--sp);
        if (G.x == BVH8_SENTINEL) {
          /* instance pop - geom_object.h:447-460 */
          const uint2 saved = stack.get(--sp); /* (world-space limit, direction scale) */
          if (inst_hit)
            tmax = xdiv(tmax, __uint_as_float(saved.y));
          else
            tmax = __uint_as_float(saved.x);
          ray_space_setup(rs, mk3(__ldg(job.ray_P(qi))), mk3(__ldg(job.ray_D(qi))));
          cur_object = -1;
          inst_hit = false;
          continue;
        }
        break;
      }
    }
    return false;
  }
};

#define TRACE_BLOCK 128
#define TRACE_WARPS (TRACE_BLOCK / 32)
#ifndef TRACE_MIN_BLOCKS
#  define TRACE_MIN_BLOCKS 8 /* <= 64 registers: 32 resident warps per SM */
#endif

CY_DEV void cp_async16(void *smem_dst, const void *gmem_src)
{
  const unsigned int dst = (unsigned int)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem_src) : "memory");
}
CY_DEV void cp_async_commit()
{
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template<int N> CY_DEV void cp_async_wait()
{
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

/* 1-D bulk asynchronous copy (the TMA unit, cp.async.bulk -> UBLKCP) with an mbarrier that
 * counts the bytes: one elected lane moves a whole 512-byte tile of ray records per
 * instruction where the LDGSTS path above spends one instruction per lane and record.
 * Built, correct (hit-id tests green) and measured on B200 against the LDGSTS path
 * (tools/r02_run19.sh, profiles/r02n_bulk_staging_ab.txt): closest-hit 4.92 vs 5.31 Grays/s
 * on the terrain, 1.49 vs 1.58 on the instanced scene, 8.3 vs 9.1 on the Cornell box - 5-11 %
 * SLOWER.  A warp's tile is 1 KB: too small for the bulk engine's fixed latency to pay, and
 * the 32 lanes spin on the mbarrier where cp.async.wait_group parks the warp on a
 * scoreboard.  Kept behind -DRAY_STAGE_BULK=1, not the default. */
#ifndef RAY_STAGE_BULK
#  define RAY_STAGE_BULK 0
#endif
CY_DEV void mbar_init(unsigned long long *bar, unsigned int count)
{
  const unsigned int a = (unsigned int)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
CY_DEV void mbar_expect_tx(unsigned long long *bar, unsigned int bytes)
{
  const unsigned int a = (unsigned int)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes)
               : "memory");
}
CY_DEV void mbar_arrive(unsigned long long *bar)
{
  const unsigned int a = (unsigned int)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
CY_DEV void mbar_wait(unsigned long long *bar, unsigned int parity)
{
  const unsigned int a = (unsigned int)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(a),
      "r"(parity)
      : "memory");
}
CY_DEV void bulk_copy_g2s(void *smem_dst, const void *gmem_src, unsigned int bytes,
                          unsigned long long *bar)
{
  const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
  const unsigned int b = (unsigned int)__cvta_generic_to_shared(bar);
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(d),
      "l"(gmem_src), "r"(bytes), "r"(b)
      : "memory");
}

/* Persistent-thread driver shared by every traversal kernel.
 *
 * Rays arrive in QUEUE ORDER as two 16-byte records (P.xyz,t | D.xyz,visibility) -
 * the 32 B "ray in" of the roofline model.  Each warp stages them through shared
 * memory: it claims 32 queue entries with one atomic, pulls them in with cp.async
 * (LDGSTS, fully coalesced) into one half of a double buffer while the lanes
 * traverse, and finished lanes refill from the staged rays at shared-memory latency
 * instead of stalling the warp on a dependent global load.  Each lane owns a
 * Traversal; a refill round is taken when fewer than `refill_threshold` lanes are
 * still busy.  `Job` supplies
 *   const float4 *ray_P(unsigned qi), *ray_D(unsigned qi)   addresses of the records
 *   void store(unsigned qi, const TraceHit &hit, bool found)
 */
template<bool ANY_HIT, bool COUNT, class Job>
__device__ __forceinline__ void trace_persistent(Job &job, unsigned int n, unsigned int *cursor,
                                                 int refill_threshold, TraceCounters &cnt)
{
  __shared__ __align__(128) float4 s_ray[TRACE_WARPS][2][2][32]; /* [warp][buffer][P|D][entry] */
  /* bulk staging: one mbarrier per warp and buffer, its phase parity kept per warp */
  __shared__ __align__(8) unsigned long long s_bar[TRACE_WARPS][2];
  constexpr bool BULK = RAY_STAGE_BULK && Job::QUEUE_RAYS;
  unsigned int parity0 = 0u, parity1 = 0u;
  /* cooperative leaf phase: pooled (record, owner lane | bit << 8) work list */
  __shared__ uint2 s_list[TRACE_WARPS][224];
  const unsigned lane = threadIdx.x & 31u;
  const unsigned warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;

  uint2 stack_local[BVH8_STACK_SIZE];
  TraceStack stack;
  stack.local = stack_local;
#if SMEM_STACK > 0
  __shared__ uint2 s_stack[SMEM_STACK * TRACE_BLOCK];
  stack.shared = s_stack + threadIdx.x;
#endif
  Traversal<ANY_HIT, COUNT> tr;
  bool active = false;
  unsigned int my_qi = 0;

  /* warp-uniform staging state */
  unsigned int base0 = 0, base1 = 0, avail0 = 0, avail1 = 0;
  unsigned int cur = 0, consumed = 0;
  bool drained = false;

  /* Each warp claims a CHUNK of consecutive queue entries (neighbouring pixel tiles /
   * neighbouring paths) and takes its 32-ray refills from it, so the rays a warp holds
   * stay close together in the scene and its lanes keep fetching the same BVH nodes.
   * The chunk shrinks with the work that is left (guided self-scheduling) so that the
   * queue still drains evenly over the warps. */
  const unsigned int n_warps = gridDim.x * TRACE_WARPS;
  unsigned int chunk_next = 0, chunk_end = 0;
  auto fill = [&](unsigned int b) {
    if (chunk_next >= chunk_end) {
      unsigned int bs0 = 0, want = 0;
      if (lane == 0) {
        const unsigned int left = (chunk_end < n) ? n - chunk_end : 0u;
        want = min(max((left / (CHUNK_DIV * n_warps)) & ~31u, 32u), CHUNK_MAX);
        bs0 = atomicAdd(cursor, want);
      }
      chunk_next = __shfl_sync(0xffffffffu, bs0, 0);
      chunk_end = chunk_next + __shfl_sync(0xffffffffu, want, 0);
    }
    const unsigned int bs = chunk_next;
    chunk_next += 32u;
    const unsigned int av = (bs < n) ? min(32u, n - bs) : 0u;
    if (BULK) {
      /* the queue is two arrays of 16-byte records: a tile is two contiguous runs */
      if (lane == 0) {
        if (av > 0) {
          mbar_expect_tx(&s_bar[warp][b], av * 32u);
          bulk_copy_g2s(&s_ray[warp][b][0][0], job.ray_P(bs), av * 16u, &s_bar[warp][b]);
          bulk_copy_g2s(&s_ray[warp][b][1][0], job.ray_D(bs), av * 16u, &s_bar[warp][b]);
        }
        else {
          mbar_arrive(&s_bar[warp][b]); /* nothing to move: the phase completes at once */
        }
      }
    }
    else {
      if (lane < av) {
        cp_async16(&s_ray[warp][b][0][lane], job.ray_P(bs + lane));
        cp_async16(&s_ray[warp][b][1][lane], job.ray_D(bs + lane));
      }
      cp_async_commit();
    }
    if (b == 0) {
      base0 = bs;
      avail0 = av;
    }
    else {
      base1 = bs;
      avail1 = av;
    }
  };

  /* wait until the fill of buffer `b` queued last has landed */
  auto landed = [&](unsigned int b) {
    if (BULK) {
      mbar_wait(&s_bar[warp][b], b ? parity1 : parity0);
      if (b)
        parity1 ^= 1u;
      else
        parity0 ^= 1u;
    }
    else {
      cp_async_wait<1>(); /* all but the newest group: the older buffer */
    }
  };
  if (BULK) {
    if (lane == 0) {
      mbar_init(&s_bar[warp][0], 1u);
      mbar_init(&s_bar[warp][1], 1u);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
  }

  fill(0);
  fill(1);
  landed(0);
  __syncwarp();
  if (avail0 == 0)
    drained = true;

  while (true) {
    /* ---- refill idle lanes from the staged rays ---- */
    while (!drained) {
      const unsigned int need = __ballot_sync(0xffffffffu, !active);
      if (need == 0u)
        break;
      const unsigned int cur_base = cur ? base1 : base0;
      const unsigned int cur_avail = cur ? avail1 : avail0;
      const unsigned int left = cur_avail - consumed;
      const unsigned int take = min((unsigned int)__popc(need), left);
      const unsigned int rank = __popc(need & lt_mask);
      if (!active && rank < take) {
        const unsigned int e = consumed + rank;
        const float4 r0 = s_ray[warp][cur][0][e];
        const float4 r1 = s_ray[warp][cur][1][e];
        const unsigned int qi = cur_base + e;
        const f3 D = mk3(r1);
        /* inactive (t = 0) or invalid rays (scene_intersect_valid, bvh/bvh.h:146-152) */
        if (r0.w != 0.0f && isfinite_safe(r0.x) && isfinite_safe(r1.x) &&
            len_squared(D) != 0.0f) {
          tr.start(mk3(r0), D, r0.w, __float_as_uint(r1.w));
          my_qi = qi;
          active = true;
        }
        else {
          TraceHit miss;
          miss.t = r0.w;
          miss.u = miss.v = 0.0f;
          miss.prim = -1;
          miss.object = -1;
          job.store(qi, miss, false);
        }
      }
      consumed += take;
      if (consumed == cur_avail) {
        /* current half is used up: restage it, switch to the other half */
        __syncwarp();
        fill(cur);
        landed(cur ^ 1u);
        __syncwarp();
        cur ^= 1u;
        consumed = 0;
        if ((cur ? avail1 : avail0) == 0u)
          drained = true;
      }
    }

    if (!__any_sync(0xffffffffu, active))
      break;

    /* ---- traverse until too few lanes are busy ---- */
    while (true) {
      if (active)
        tr.node_phase(stack, cnt);

      /* ---- cooperative leaf phase (warp-convergent) ---- */
      uint32_t inst_bits = 0u;
      bool finished = false;
      while (true) {
        const uint32_t pend = active ? tr.Gt.y : 0u;
        /* up to seven records per lane and pass (a leaf holds at most three, a node can
         * expose several leaves), pooled in lane order: offsets from three ballots, no
         * scan, no atomics */
        const unsigned int npend = min((unsigned int)__popc(pend), 7u);
        const unsigned int b0 = __ballot_sync(0xffffffffu, (npend & 1u) != 0u);
        const unsigned int b1 = __ballot_sync(0xffffffffu, (npend & 2u) != 0u);
        const unsigned int b2 = __ballot_sync(0xffffffffu, (npend & 4u) != 0u);
        if ((b0 | b1 | b2) == 0u)
          break;
        const unsigned int total = __popc(b0) + 2u * __popc(b1) + 4u * __popc(b2);
        const unsigned int off = __popc(b0 & lt_mask) + 2u * __popc(b1 & lt_mask) +
                                 4u * __popc(b2 & lt_mask);
        if (npend != 0u) {
          uint2 *dst = s_list[warp] + off;
          uint32_t bits = pend;
          for (unsigned int k = 0; k < npend; k++) {
            const uint32_t bit = (uint32_t)__ffs((int)bits) - 1u;
            dst[k] = make_uint2(tr.Gt.x + bit, lane | (bit << 8));
            bits &= bits - 1u;
          }
          tr.Gt.y = bits;
        }
        __syncwarp();

        for (unsigned int base = 0; base < total; base += 32u) {
          const bool has = base + lane < total;
          const uint2 ent = has ? s_list[warp][base + lane] : make_uint2(0u, lane);
          const unsigned int owner = ent.y & 31u;
          /* the owner's ray, in the space it is traversing in */
          const float opx = __shfl_sync(0xffffffffu, tr.rs.P.x, owner);
          const float opy = __shfl_sync(0xffffffffu, tr.rs.P.y, owner);
          const float opz = __shfl_sync(0xffffffffu, tr.rs.P.z, owner);
          const float odx = __shfl_sync(0xffffffffu, tr.rs.dir.x, owner);
          const float ody = __shfl_sync(0xffffffffu, tr.rs.dir.y, owner);
          const float odz = __shfl_sync(0xffffffffu, tr.rs.dir.z, owner);
          const float otmax = __shfl_sync(0xffffffffu, tr.tmax, owner);
          const uint32_t ovis = __shfl_sync(0xffffffffu, tr.visibility, owner);
          bool cand = false, inst = false;
          float t = 0.0f, u = 0.0f, v = 0.0f;
          int tag = 0;
          if (has) {
            /* all three quarters of the record at once: triangles dominate, an
             * instance record just ignores the last two */
            const float4 *rp = g_scene.records + (size_t)ent.x * 3;
            const float4 ra = __ldg(rp + 0);
            const float4 rb = __ldg(rp + 1);
            const float4 rc = __ldg(rp + 2);
            tag = __float_as_int(ra.w);
            if (tag >= 0) {
              if (COUNT)
                cnt.tris++;
              cand = ray_triangle_intersect(mk3(opx, opy, opz), mk3(odx, ody, odz), otmax,
                                            mk3(ra), mk3(rb), mk3(rc), &u, &v, &t) &&
                     (__float_as_uint(rb.w) & ovis) != 0u;
            }
            else {
              inst = (__float_as_uint(ra.y) & ovis) != 0u;
            }
          }
          /* results go back to the owners by shuffle, candidates in list order - the
           * order a single lane would have tested them in (a later record with the
           * same t replaces an earlier one, like the serial loop) */
          unsigned int hm = __ballot_sync(0xffffffffu, cand);
          if (hm != 0u) {
            /* first only (owner, t) of every candidate travels; an owner remembers the
             * lane of the hit it keeps and fetches u, v, prim from it once */
            int win = -1;
            do {
              const int src = __ffs((int)hm) - 1;
              hm &= hm - 1u;
              const unsigned int o = __shfl_sync(0xffffffffu, owner, src);
              const float ht = __shfl_sync(0xffffffffu, t, src);
              if (lane == o && ht <= tr.tmax) {
                tr.tmax = ht;
                win = src;
              }
            } while (hm != 0u);
            const int from = (win >= 0) ? win : (int)lane;
            const float hu = __shfl_sync(0xffffffffu, u, from);
            const float hv = __shfl_sync(0xffffffffu, v, from);
            const int hp = __shfl_sync(0xffffffffu, tag, from);
            if (win >= 0) {
              tr.accept(tr.tmax, hu, hv, hp);
              if (ANY_HIT)
                finished = true;
            }
          }
          /* visible instance records: every owner picks the flags of its own list
           * entries out of one ballot (entry off + k was tested by lane off + k - base) */
          const unsigned int im = __ballot_sync(0xffffffffu, inst);
          if (im != 0u && npend != 0u) {
            const int rel = (int)off - (int)base;
            uint32_t mine = (rel >= 0) ? (rel < 32 ? im >> rel : 0u) : (rel > -32 ? im << (-rel) : 0u);
            mine &= (1u << npend) - 1u;
            uint32_t pb = pend;
            while (mine != 0u) { /* k-th flag <-> k-th lowest set bit of pend */
              if (mine & 1u)
                inst_bits |= pb & (0u - pb);
              pb &= pb - 1u;
              mine >>= 1;
            }
          }
        }
        if (ANY_HIT && finished)
          tr.Gt.y = 0u;
        /* usually everything fitted into this pass */
        if (!__any_sync(0xffffffffu, active && tr.Gt.y != 0u))
          break;
        __syncwarp();
      }

      if (active) {
        if (finished || tr.finish_step(stack, inst_bits, cnt, job, my_qi)) {
          tr.hit.t = tr.tmax;
          job.store(my_qi, tr.hit, tr.hit.prim >= 0);
          active = false;
        }
      }
      const unsigned int busy = __ballot_sync(0xffffffffu, active);
      if (busy == 0u)
        break;
      if (!drained && __popc(busy) < refill_threshold)
        break;
    }
  }
  if (BULK) {
    /* the fill queued last (never consumed) must have landed before the block may go */
    landed(cur ^ 1u);
  }
  else {
    cp_async_wait<0>();
  }
}

#endif /* B200_TRAVERSE_CUH */

/* lobes.cuh - where the BSDF lobes of a shading point live on this device.
 *
 * The reference keeps a shading point's closures as an array of fixed 80-byte structs
 * inside ShaderData, up to 64 of them (kernel_types.h:879-1021, closure/alloc.h).  Here a
 * shading point owns a WORD ARENA: lobes are variable-length records packed back to back
 * - an 8-word header plus what the lobe kind needs (a Lambert lobe is 8 words, a GGX lobe
 * 12, a Fresnel-tinted GGX lobe 16) - in units of four words, so every access is a
 * 128-bit load or store and a typical shading point (one to three lobes) touches a
 * hundred bytes instead of walking 80-byte slots.  The closure limit of the scene
 * (KernelIntegrator::max_closures) is honoured exactly like the reference does.
 *
 * Where the arena lives is the kernel's choice (LobeArena::q points at it): thread-local
 * memory in the shipped kernels.  A shared-memory arena (column per thread, word w of
 * thread t at arena[w][t]) was built and measured on B200: it lost 7-17 % on the Cornell
 * workload - the 53 KB of shared memory per block halve what is left of L1 for the SVM
 * stack and the scene gathers, and the shared arena is read word by word where the local
 * one is read 128 bits at a time ("measured and rejected").
 *
 * Lobe record (words):
 *   0..2 weight   3 sample_weight   4..6 N   7 kind = closure id | LOBE_HAS_TANGENT
 *   then by kind, padded to a multiple of four, see lobe_tail_words():
 *     Oren-Nayar            a, b
 *     Principled diffuse    roughness
 *     sharp refraction      ior
 *     GGX / GGX refraction  alpha_x, alpha_y, ior
 *     GGX Fresnel           alpha_x, alpha_y, ior, -, cspec0.xyz
 *     GGX clearcoat         alpha, strength           (ior 1.5, cspec0 0.04 are constants)
 *     multi-scatter GGX     alpha_x, alpha_y, ior, -, color.xyz, - [, cspec0.xyz if Fresnel]
 *     multi-scatter glass   alpha, ior, -, -, color.xyz, - [, cspec0.xyz if Fresnel]
 *   then T.xyz when LOBE_HAS_TANGENT (anisotropic lobes only).
 *
 * Closure ids are the reference's ClosureType values (generated into cycles_abi.h): the
 * path-state logic tests id ranges exactly like kernel_types.h's CLOSURE_IS_* macros.
 * Free of warp intrinsics: compiled for the host by tests/host_check. */
#ifndef B200_LOBES_CUH
#define B200_LOBES_CUH

#include "cymath.cuh"

/* room for the closure limit this device accepts: MAX_CLOSURES_GPU lobes of <= 24 words */
#define ARENA_QUADS (32 * 6)

#define LOBE_HEADER_WORDS 8
#define LOBE_HAS_TANGENT 0x100u
#define LOBE_ID_MASK 0xffu

struct Lobe {
  f3 weight;
  float sample_weight;
  f3 N;
  uint32_t kind;
  float ax, ay, ior; /* roughness pair and index of refraction (kind dependent) */
  float aux;         /* Oren-Nayar b | Principled-diffuse roughness | clearcoat strength */
  f3 cspec0;         /* Fresnel lobes: reflectance at normal incidence */
  f3 color;          /* multi-scatter lobes: per-bounce albedo */
  f3 T;              /* anisotropic lobes: tangent */
};

struct LobeArena {
  float4 *q; /* the arena, in quads of four words (null when left == 0 for good) */
  int n;     /* lobes stored */
  int used;  /* words used (a multiple of four) */
  int left;  /* closure budget left - ShaderData::num_closure_left (closure/alloc.h:19-68) */
};

CY_DEV int lobe_id(uint32_t kind)
{
  return (int)(kind & LOBE_ID_MASK);
}

/* words after the header for a lobe of this kind (a multiple of four) */
CY_DEV int lobe_tail_words(uint32_t kind)
{
  int w;
  switch (lobe_id(kind)) {
    case CY_CLOSURE_BSDF_OREN_NAYAR_ID:
    case CY_CLOSURE_BSDF_PRINCIPLED_DIFFUSE_ID:
    case CY_CLOSURE_BSDF_REFRACTION_ID:
    case CY_CLOSURE_BSDF_MICROFACET_GGX_ID:
    case CY_CLOSURE_BSDF_MICROFACET_GGX_REFRACTION_ID:
    case CY_CLOSURE_BSDF_MICROFACET_GGX_CLEARCOAT_ID:
      w = 4;
      break;
    case CY_CLOSURE_BSDF_MICROFACET_GGX_FRESNEL_ID:
    case CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_ID:
    case CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_GLASS_ID:
      w = 8;
      break;
    case CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_FRESNEL_ID:
    case CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_GLASS_FRESNEL_ID:
      w = 12;
      break;
    default:
      w = 0;
  }
  return w + ((kind & LOBE_HAS_TANGENT) ? 4 : 0);
}
CY_DEV int lobe_words(uint32_t kind)
{
  return LOBE_HEADER_WORDS + lobe_tail_words(kind);
}

CY_DEV void arena_reset(LobeArena &a, int budget)
{
  a.n = 0;
  a.used = 0;
  a.left = budget;
}

/* ------------------------------------------------------------ header access */

/* header fields of the lobe at word offset `at` (a multiple of four) */
CY_DEV uint32_t lobe_kind_at(const LobeArena &a, int at)
{
  return __float_as_uint(a.q[(at >> 2) + 1].w);
}
CY_DEV float lobe_sample_weight_at(const LobeArena &a, int at)
{
  return a.q[at >> 2].w;
}
CY_DEV void lobe_set_sample_weight_at(LobeArena &a, int at, float v)
{
  a.q[at >> 2].w = v;
}
CY_DEV f3 lobe_weight_at(const LobeArena &a, int at)
{
  return mk3(a.q[at >> 2]);
}
CY_DEV f3 lobe_normal_at(const LobeArena &a, int at)
{
  return mk3(a.q[(at >> 2) + 1]);
}
CY_DEV void lobe_add_weight_at(LobeArena &a, int at, f3 w, float sample_weight)
{
  float4 h = a.q[at >> 2];
  h.x += w.x;
  h.y += w.y;
  h.z += w.z;
  h.w += sample_weight;
  a.q[at >> 2] = h;
}

CY_DEV bool lobe_kind_is_multi_glass(uint32_t kind)
{
  return lobe_id(kind) == CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_GLASS_ID ||
         lobe_id(kind) == CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_GLASS_FRESNEL_ID;
}
CY_DEV bool lobe_kind_is_multi(uint32_t kind)
{
  return lobe_id(kind) == CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_ID ||
         lobe_id(kind) == CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_FRESNEL_ID ||
         lobe_kind_is_multi_glass(kind);
}
CY_DEV bool lobe_kind_has_cspec0(uint32_t kind)
{
  return lobe_id(kind) == CY_CLOSURE_BSDF_MICROFACET_GGX_FRESNEL_ID ||
         lobe_id(kind) == CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_FRESNEL_ID ||
         lobe_id(kind) == CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_GLASS_FRESNEL_ID;
}

/* widen the roughness of a GGX-family lobe in place (bsdf_blur, closure/bsdf.h:706-737) */
CY_DEV void lobe_blur_at(LobeArena &a, int at, float roughness)
{
  const uint32_t kind = lobe_kind_at(a, at);
  const int id = lobe_id(kind);
  float4 &p = a.q[(at >> 2) + 2];
  if (id == CY_CLOSURE_BSDF_MICROFACET_GGX_ID || id == CY_CLOSURE_BSDF_MICROFACET_GGX_FRESNEL_ID ||
      id == CY_CLOSURE_BSDF_MICROFACET_GGX_REFRACTION_ID ||
      id == CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_ID ||
      id == CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_FRESNEL_ID) {
    p.x = fmaxf(roughness, p.x);
    p.y = fmaxf(roughness, p.y);
  }
  else if (id == CY_CLOSURE_BSDF_MICROFACET_GGX_CLEARCOAT_ID || lobe_kind_is_multi_glass(kind)) {
    p.x = fmaxf(roughness, p.x);
  }
}

/* --------------------------------------------------------- whole records */

/* Appends the lobe; returns its word offset, or -1 when the arena is full (cannot happen
 * within the closure limit check_scope accepts). */
CY_DEV int lobe_store(LobeArena &a, const Lobe &l)
{
  const int at = a.used;
  const int words = lobe_words(l.kind);
  if (at + words > ARENA_QUADS * 4)
    return -1;
  float4 *q = a.q + (at >> 2);
  q[0] = make_float4(l.weight.x, l.weight.y, l.weight.z, l.sample_weight);
  q[1] = make_float4(l.N.x, l.N.y, l.N.z, __uint_as_float(l.kind));
  q += 2;
  const int id = lobe_id(l.kind);
  if (lobe_tail_words(l.kind & LOBE_ID_MASK) != 0) {
    /* first parameter quad: the scalars of the kind */
    if (lobe_kind_is_multi_glass(l.kind))
      *q++ = make_float4(l.ax, l.ior, 0.0f, 0.0f);
    else if (id == CY_CLOSURE_BSDF_MICROFACET_GGX_CLEARCOAT_ID ||
             id == CY_CLOSURE_BSDF_OREN_NAYAR_ID)
      *q++ = make_float4(l.ax, l.aux, 0.0f, 0.0f);
    else if (id == CY_CLOSURE_BSDF_PRINCIPLED_DIFFUSE_ID)
      *q++ = make_float4(l.aux, 0.0f, 0.0f, 0.0f);
    else if (id == CY_CLOSURE_BSDF_REFRACTION_ID)
      *q++ = make_float4(l.ior, 0.0f, 0.0f, 0.0f);
    else
      *q++ = make_float4(l.ax, l.ay, l.ior, 0.0f);
    if (lobe_kind_is_multi(l.kind))
      *q++ = make_float4(l.color.x, l.color.y, l.color.z, 0.0f);
    if (lobe_kind_has_cspec0(l.kind))
      *q++ = make_float4(l.cspec0.x, l.cspec0.y, l.cspec0.z, 0.0f);
  }
  if (l.kind & LOBE_HAS_TANGENT)
    *q = make_float4(l.T.x, l.T.y, l.T.z, 0.0f);
  a.used = at + words;
  a.n++;
  return at;
}

/* Reads the lobe at word offset `at`; fields its kind does not carry get their neutral
 * values (isotropic alpha pair, zero tangent ...). */
CY_DEV Lobe lobe_fetch(const LobeArena &a, int at)
{
  Lobe l;
  const float4 *q = a.q + (at >> 2);
  const float4 h0 = q[0], h1 = q[1];
  l.weight = mk3(h0);
  l.sample_weight = h0.w;
  l.N = mk3(h1);
  l.kind = __float_as_uint(h1.w);
  l.ax = l.ay = 0.0f;
  l.ior = 0.0f;
  l.aux = 0.0f;
  l.cspec0 = zero3();
  l.color = zero3();
  l.T = zero3();
  q += 2;
  const int id = lobe_id(l.kind);
  if (lobe_tail_words(l.kind & LOBE_ID_MASK) != 0) {
    const float4 p = *q++;
    if (lobe_kind_is_multi_glass(l.kind)) {
      l.ax = l.ay = p.x;
      l.ior = p.y;
    }
    else if (id == CY_CLOSURE_BSDF_MICROFACET_GGX_CLEARCOAT_ID) {
      l.ax = l.ay = p.x;
      l.aux = p.y;
      l.ior = 1.5f;
      l.cspec0 = mk3(0.04f, 0.04f, 0.04f);
    }
    else if (id == CY_CLOSURE_BSDF_OREN_NAYAR_ID) {
      l.ax = p.x;
      l.aux = p.y;
    }
    else if (id == CY_CLOSURE_BSDF_PRINCIPLED_DIFFUSE_ID) {
      l.aux = p.x;
    }
    else if (id == CY_CLOSURE_BSDF_REFRACTION_ID) {
      l.ior = p.x;
    }
    else {
      l.ax = p.x;
      l.ay = p.y;
      l.ior = p.z;
    }
    if (lobe_kind_is_multi(l.kind))
      l.color = mk3(*q++);
    if (lobe_kind_has_cspec0(l.kind))
      l.cspec0 = mk3(*q++);
  }
  if (l.kind & LOBE_HAS_TANGENT)
    l.T = mk3(*q);
  return l;
}

/* CLOSURE_IS_BSDF_OR_BSSRDF / CLOSURE_IS_BSDF / CLOSURE_IS_BSDF_DIFFUSE / _MICROFACET of
 * kernel/svm/svm_types.h:573-612, on ids */
CY_DEV bool lobe_is_sampled(uint32_t kind)
{
  return lobe_id(kind) <= CY_CLOSURE_BSSRDF_PRINCIPLED_RANDOM_WALK_ID;
}
CY_DEV bool lobe_is_bsdf(uint32_t kind)
{
  return lobe_id(kind) <= CY_CLOSURE_BSDF_TRANSPARENT_ID;
}
CY_DEV bool lobe_is_diffuse(uint32_t kind)
{
  return lobe_id(kind) >= CY_CLOSURE_BSDF_DIFFUSE_ID &&
         lobe_id(kind) <= CY_CLOSURE_BSDF_TRANSLUCENT_ID;
}
#endif /* B200_LOBES_CUH */

/* bsdf.cuh - evaluation and sampling of one lobe (lobes.cuh), dispatching on its kind.
 *
 * Reference semantics (blender/intern/cycles/kernel/closure): bsdf.h bsdf_eval /
 * bsdf_sample (which side of the geometric normal is evaluated, the softened-terminator
 * factors for bent normals and the per-object terminator offset), bsdf_diffuse.h,
 * bsdf_oren_nayar.h, bsdf_principled_diffuse.h, bsdf_principled_sheen.h,
 * bsdf_reflection.h, bsdf_refraction.h, bsdf_transparent.h; kernel_montecarlo.h for the
 * hemisphere mappings.  All BSDF values include the cosine of the incoming direction.
 * Included by shade.cuh; host-compilable. */
#ifndef B200_BSDF_CUH
#define B200_BSDF_CUH

#include "lobes.cuh"

/* ------------------------------------------------ hemisphere mappings */

/* polar mapping of the unit square onto the unit disk */
CY_DEV float2 square_to_disk_polar(float u, float v)
{
  const float phi = CY_2PI_F * u;
  const float r = sqrtf(v);
  return make_float2(r * cosf(phi), r * sinf(phi));
}

/* cosine-weighted direction about N; pdf = cos / pi */
CY_DEV void sample_cos_hemisphere(f3 N, float randu, float randv, f3 *omega_in, float *pdf)
{
  const float2 d = square_to_disk_polar(randu, randv);
  const float costheta = sqrtf(fmaxf(1.0f - d.x * d.x - d.y * d.y, 0.0f));
  f3 T, B;
  make_orthonormals(N, &T, &B);
  *omega_in = d.x * T + d.y * B + costheta * N;
  *pdf = costheta * CY_1_PI_F;
}

/* uniform direction about N; pdf = 1 / 2pi */
CY_DEV void sample_uniform_hemisphere(f3 N, float randu, float randv, f3 *omega_in, float *pdf)
{
  const float z = randu;
  const float r = sqrtf(fmaxf(0.0f, 1.0f - z * z));
  const float phi = CY_2PI_F * randv;
  f3 T, B;
  make_orthonormals(N, &T, &B);
  *omega_in = (r * cosf(phi)) * T + (r * sinf(phi)) * B + z * N;
  *pdf = 0.5f * CY_1_PI_F;
}

#include "microfacet.cuh"
#include "microfacet_multi.cuh"

/* ------------------------------------------------------ diffuse family */

/* Oren-Nayar in Cycles' two-coefficient form: value = n.l * (a + b * t) */
CY_DEV void oren_nayar_coefficients(float roughness, float *a, float *b)
{
  const float sigma = saturate(roughness);
  const float div = 1.0f / (CY_M_PI_F + ((3.0f * CY_M_PI_F - 4.0f) / 6.0f) * sigma);
  *a = 1.0f * div;
  *b = sigma * div;
}
CY_DEV float oren_nayar_value(const Lobe &l, f3 v, f3 wi)
{
  const float nl = fmaxf(dot(l.N, wi), 0.0f);
  const float nv = fmaxf(dot(l.N, v), 0.0f);
  float t = dot(wi, v) - nl * nv;
  if (t > 0.0f)
    t /= fmaxf(nl, nv) + FLT_MIN;
  return nl * (l.ax + l.aux * t);
}

/* Burley's diffuse with the retro-reflective rim (Principled BSDF) */
CY_DEV float disney_diffuse_value(const Lobe &l, f3 V, f3 L)
{
  const float NdotL = fmaxf(dot(l.N, L), 0.0f);
  const float NdotV = fmaxf(dot(l.N, V), 0.0f);
  const f3 H = normalize(L + V);
  const float LdotH = dot(L, H);
  const float FL = schlick_weight(NdotL), FV = schlick_weight(NdotV);
  const float Fd90 = 0.5f + 2.0f * LdotH * LdotH * l.aux;
  const float Fd = (1.0f * (1.0f - FL) + Fd90 * FL) * (1.0f * (1.0f - FV) + Fd90 * FV);
  return CY_1_PI_F * NdotL * Fd;
}

/* Principled sheen: Schlick weight of the half-vector angle; zero (and pdf zero) below
 * the horizon of either direction */
CY_DEV bool sheen_value(const Lobe &l, f3 V, f3 L, float *value)
{
  const float NdotL = dot(l.N, L), NdotV = dot(l.N, V);
  if (NdotL < 0.0f || NdotV < 0.0f)
    return false;
  const f3 H = normalize(L + V);
  *value = schlick_weight(dot(L, H)) * NdotL;
  return true;
}

/* -------------------------------------------- softened shadow terminator */

CY_DEV float bump_shadowing_term(f3 Ng, f3 N, f3 I)
{
  const float g = safe_divide(dot(Ng, I), dot(N, I) * dot(Ng, N));
  if (g >= 1.0f)
    return 1.0f;
  if (g < 0.0f)
    return 0.0f;
  const float g2 = sqr(g);
  return -g2 * g + g2 + g;
}
CY_DEV float shift_cos_in(float cos_in, float frequency_multiplier)
{
  cos_in = fminf(cos_in, 1.0f);
  const float angle = fast_acosf(cos_in);
  return fmaxf(cosf(angle * frequency_multiplier), 0.0f) / cos_in;
}

/* Decided once per shading point, after the shader ran: does any lobe need the
 * terminator factors at all?  Almost never - and testing it per lobe inside the
 * eval / sample loops cost the Cornell workload 11 %. */
CY_DEV void bsdf_terminator_terms_setup(ShaderDataG &sd, const LobeArena &arena)
{
  int need = (sd.terminator_freq > 1.0f) ? 1 : 0;
  int at = 0;
  for (int i = 0; i < arena.n; i++) {
    const uint32_t kind = lobe_kind_at(arena, at);
    if (lobe_is_diffuse(kind) && !isequal3(lobe_normal_at(arena, at), sd.N))
      need = 1;
    at += lobe_words(kind);
  }
  sd.terminator_terms = need;
}

/* ----------------------------------------------------------- evaluation */

/* EXT = false: the lobes the lean interpreter can create (it hands shaders with sheen or
 * bent normals to the full one, shade.cuh svm_eval_nodes).  MS: the kernel carries the
 * multi-scatter lobes (two out-of-line calls); its own switch, because even unreachable
 * code in the lean kernels costs the common shaders time (measured: +11 % on the GGX
 * Cornell box, +4 % on the terrain when every lean kernel carried the calls). */
template<bool EXT, bool MS = EXT>
CY_DEV f3 bsdf_eval(ShaderDataG &sd, const Lobe &l, f3 omega_in, float *pdf)
{
  const bool same_side = dot(sd.Ng, omega_in) >= 0.0f;
  const int id = lobe_id(l.kind);
  float v = 0.0f;
  f3 value = zero3();
  switch (id) {
    case CY_CLOSURE_BSDF_DIFFUSE_ID:
      if (same_side) {
        v = fmaxf(dot(l.N, omega_in), 0.0f) * CY_1_PI_F;
        *pdf = v;
        value = mk3(v, v, v);
      }
      break;
    case CY_CLOSURE_BSDF_TRANSLUCENT_ID:
      if (!same_side) {
        v = fmaxf(-dot(l.N, omega_in), 0.0f) * CY_1_PI_F;
        *pdf = v;
        value = mk3(v, v, v);
      }
      break;
    case CY_CLOSURE_BSDF_OREN_NAYAR_ID:
      if (same_side) {
        if (dot(l.N, omega_in) > 0.0f) {
          *pdf = 0.5f * CY_1_PI_F;
          v = oren_nayar_value(l, sd.I, omega_in);
          value = mk3(v, v, v);
        }
        else {
          *pdf = 0.0f;
        }
      }
      break;
    case CY_CLOSURE_BSDF_PRINCIPLED_DIFFUSE_ID:
      if (same_side) {
        if (dot(l.N, omega_in) > 0.0f) {
          *pdf = fmaxf(dot(l.N, omega_in), 0.0f) * CY_1_PI_F;
          v = disney_diffuse_value(l, sd.I, omega_in);
          value = mk3(v, v, v);
        }
        else {
          *pdf = 0.0f;
        }
      }
      break;
    case CY_CLOSURE_BSDF_PRINCIPLED_SHEEN_ID:
      if (EXT && same_side) {
        *pdf = 0.0f;
        if (dot(l.N, omega_in) > 0.0f && sheen_value(l, sd.I, omega_in, &v)) {
          *pdf = fmaxf(dot(l.N, omega_in), 0.0f) * CY_M_1_PI_F;
          value = mk3(v, v, v);
        }
      }
      break;
    case CY_CLOSURE_BSDF_MICROFACET_GGX_ID:
    case CY_CLOSURE_BSDF_MICROFACET_GGX_FRESNEL_ID:
    case CY_CLOSURE_BSDF_MICROFACET_GGX_CLEARCOAT_ID:
    case CY_CLOSURE_BSDF_MICROFACET_GGX_REFRACTION_ID:
      value = ggx_eval(l, sd.I, omega_in, same_side, pdf);
      break;
    case CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_ID:
    case CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_FRESNEL_ID:
    case CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_GLASS_ID:
    case CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_GLASS_FRESNEL_ID:
      if (MS)
        value = multi_ggx_eval(l, sd.I, omega_in, same_side, pdf, &sd.lcg_state);
      break;
    default: /* sharp lobes, transparent, placeholders: nothing to evaluate */
      break;
  }
  if (EXT && sd.terminator_terms) {
    if (lobe_is_diffuse(l.kind) && !isequal3(l.N, sd.N))
      value *= bump_shadowing_term(same_side ? sd.N : -sd.N, l.N, omega_in);
    if (same_side && sd.terminator_freq > 1.0f)
      value *= shift_cos_in(dot(omega_in, l.N), sd.terminator_freq);
  }
  return value;
}

/* ------------------------------------------------------------ sampling */

/* a cosine- or uniformly-sampled diffuse direction counts only on the side the lobe
 * scatters to */
CY_DEV bool keep_side(f3 Ng, f3 omega_in, bool reflect)
{
  const float c = dot(Ng, omega_in);
  return reflect ? (c > 0.0f) : (c < 0.0f);
}

template<bool EXT, bool MS = EXT>
CY_DEV int bsdf_sample_lobe(ShaderDataG &sd, const Lobe &l, float randu, float randv, f3 *value,
                            f3 *omega_in, float *pdf)
{
  float v = 0.0f;
  switch (lobe_id(l.kind)) {
    case CY_CLOSURE_BSDF_DIFFUSE_ID:
      sample_cos_hemisphere(l.N, randu, randv, omega_in, pdf);
      if (keep_side(sd.Ng, *omega_in, true))
        *value = mk3(*pdf, *pdf, *pdf);
      else
        *pdf = 0.0f;
      return CY_LABEL_REFLECT | CY_LABEL_DIFFUSE;
    case CY_CLOSURE_BSDF_TRANSLUCENT_ID:
      sample_cos_hemisphere(-l.N, randu, randv, omega_in, pdf);
      if (keep_side(sd.Ng, *omega_in, false))
        *value = mk3(*pdf, *pdf, *pdf);
      else
        *pdf = 0.0f;
      return CY_LABEL_TRANSMIT | CY_LABEL_DIFFUSE;
    case CY_CLOSURE_BSDF_OREN_NAYAR_ID:
      sample_uniform_hemisphere(l.N, randu, randv, omega_in, pdf);
      if (keep_side(sd.Ng, *omega_in, true)) {
        v = oren_nayar_value(l, sd.I, *omega_in);
        *value = mk3(v, v, v);
      }
      else {
        *pdf = 0.0f;
        *value = zero3();
      }
      return CY_LABEL_REFLECT | CY_LABEL_DIFFUSE;
    case CY_CLOSURE_BSDF_PRINCIPLED_DIFFUSE_ID:
      sample_cos_hemisphere(l.N, randu, randv, omega_in, pdf);
      if (keep_side(sd.Ng, *omega_in, true)) {
        v = disney_diffuse_value(l, sd.I, *omega_in);
        *value = mk3(v, v, v);
      }
      else {
        *pdf = 0.0f;
      }
      return CY_LABEL_REFLECT | CY_LABEL_DIFFUSE;
    case CY_CLOSURE_BSDF_PRINCIPLED_SHEEN_ID:
      if (!EXT) {
        *pdf = 0.0f;
        return CY_LABEL_NONE;
      }
      sample_cos_hemisphere(l.N, randu, randv, omega_in, pdf);
      if (keep_side(sd.Ng, *omega_in, true)) {
        if (sheen_value(l, sd.I, *omega_in, &v))
          *value = mk3(v, v, v);
        else
          *pdf = 0.0f;
      }
      else {
        *pdf = 0.0f;
      }
      return CY_LABEL_REFLECT | CY_LABEL_DIFFUSE;
    case CY_CLOSURE_BSDF_REFLECTION_ID: {
      /* perfect mirror: one direction, "some high number" as density */
      const float cosNO = dot(l.N, sd.I);
      if (cosNO > 0.0f) {
        *omega_in = (2.0f * cosNO) * l.N - sd.I;
        if (dot(sd.Ng, *omega_in) > 0.0f) {
          *pdf = 1e6f;
          *value = mk3(1e6f, 1e6f, 1e6f);
        }
      }
      return CY_LABEL_REFLECT | CY_LABEL_SINGULAR;
    }
    case CY_CLOSURE_BSDF_REFRACTION_ID: {
      const DielectricSplit split = dielectric_split(l.ior, l.N, sd.I);
      if (!split.inside && split.reflectance != 1.0f) {
        *pdf = 1e6f;
        *value = mk3(1e6f, 1e6f, 1e6f);
        *omega_in = split.refracted;
      }
      return CY_LABEL_TRANSMIT | CY_LABEL_SINGULAR;
    }
    case CY_CLOSURE_BSDF_TRANSPARENT_ID:
      /* straight through */
      *omega_in = -sd.I;
      *pdf = 1.0f;
      *value = one3();
      return CY_LABEL_TRANSMIT | CY_LABEL_TRANSPARENT;
    case CY_CLOSURE_BSDF_MICROFACET_GGX_ID:
    case CY_CLOSURE_BSDF_MICROFACET_GGX_FRESNEL_ID:
    case CY_CLOSURE_BSDF_MICROFACET_GGX_CLEARCOAT_ID:
    case CY_CLOSURE_BSDF_MICROFACET_GGX_REFRACTION_ID:
      return ggx_sample(l, sd.Ng, sd.I, randu, randv, value, omega_in, pdf);
    case CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_ID:
    case CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_FRESNEL_ID:
    case CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_GLASS_ID:
    case CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_GLASS_FRESNEL_ID:
      if (MS)
        return multi_ggx_sample(l, sd.I, randu, randv, value, omega_in, pdf, &sd.lcg_state);
      *pdf = 0.0f;
      return CY_LABEL_NONE;
    default:
      *pdf = 0.0f;
      return CY_LABEL_NONE;
  }
}

/* the lobe's own sampler, then the terminator factors on the reflection side */
template<bool EXT, bool MS = EXT>
CY_DEV int bsdf_sample(ShaderDataG &sd, const Lobe &l, float randu, float randv, f3 *value,
                       f3 *omega_in, float *pdf)
{
  const int label = bsdf_sample_lobe<EXT, MS>(sd, l, randu, randv, value, omega_in, pdf);
  if (EXT && sd.terminator_terms && !(label & CY_LABEL_TRANSMIT)) {
    if (sd.terminator_freq > 1.0f)
      *value *= shift_cos_in(dot(*omega_in, l.N), sd.terminator_freq);
    if ((label & CY_LABEL_DIFFUSE) && !isequal3(l.N, sd.N))
      *value *= bump_shadowing_term(sd.N, l.N, *omega_in);
  }
  return label;
}

#endif

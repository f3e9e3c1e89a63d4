/* bsdf.cuh - closure evaluation and sampling (kernel/closure/bsdf.h dispatch).
 * Included by shade.cuh. */
#ifndef B200_BSDF_CUH
#define B200_BSDF_CUH

/* kernel_montecarlo.h:57-66 */
CY_DEV void sample_cos_hemisphere(f3 N, float randu, float randv, f3 *omega_in, float *pdf)
{
  /* to_unit_disk - kernel_montecarlo.h:39-46 */
  float phi = CY_2PI_F * randu;
  float r = sqrtf(randv);
  randu = r * cosf(phi);
  randv = r * sinf(phi);
  float costheta = sqrtf(fmaxf(1.0f - randu * randu - randv * randv, 0.0f));
  f3 T, B;
  make_orthonormals(N, &T, &B);
  *omega_in = randu * T + randv * B + costheta * N;
  *pdf = costheta * CY_1_PI_F;
}

/* closure/bsdf_diffuse.h:55-110 */
CY_DEV f3 bsdf_diffuse_eval_reflect(const Closure &sc, f3 omega_in, float *pdf)
{
  float cos_pi = fmaxf(dot(sc.N, omega_in), 0.0f) * CY_1_PI_F;
  *pdf = cos_pi;
  return mk3(cos_pi, cos_pi, cos_pi);
}
CY_DEV int bsdf_diffuse_sample(const Closure &sc, f3 Ng, float randu, float randv, f3 *eval,
                               f3 *omega_in, float *pdf)
{
  sample_cos_hemisphere(sc.N, randu, randv, omega_in, pdf);
  if (dot(Ng, *omega_in) > 0.0f)
    *eval = mk3(*pdf, *pdf, *pdf);
  else
    *pdf = 0.0f;
  return CY_LABEL_REFLECT | CY_LABEL_DIFFUSE;
}

/* kernel_montecarlo.h:69-82 */
CY_DEV void sample_uniform_hemisphere(f3 N, float randu, float randv, f3 *omega_in, float *pdf)
{
  float z = randu;
  float r = sqrtf(fmaxf(0.0f, 1.0f - z * z));
  float phi = CY_2PI_F * randv;
  float x = r * cosf(phi);
  float y = r * sinf(phi);
  f3 T, B;
  make_orthonormals(N, &T, &B);
  *omega_in = x * T + y * B + z * N;
  *pdf = 0.5f * CY_1_PI_F;
}

/* closure/bsdf_oren_nayar.h: Diffuse BSDF with roughness > 0.  The two precomputed
 * coefficients a, b live in alpha_x, alpha_y of the closure record. */
CY_DEV uint32_t bsdf_oren_nayar_setup(Closure *bsdf, float roughness)
{
  bsdf->type = CY_CLOSURE_BSDF_OREN_NAYAR_ID;
  const float sigma = saturate(roughness);
  const float div = 1.0f / (CY_M_PI_F + ((3.0f * CY_M_PI_F - 4.0f) / 6.0f) * sigma);
  bsdf->roughness = roughness;
  bsdf->alpha_x = 1.0f * div;
  bsdf->alpha_y = sigma * div;
  return CY_SD_BSDF | CY_SD_BSDF_HAS_EVAL;
}
CY_DEV f3 bsdf_oren_nayar_get_intensity(const Closure &sc, f3 n, f3 v, f3 l)
{
  float nl = fmaxf(dot(n, l), 0.0f);
  float nv = fmaxf(dot(n, v), 0.0f);
  float t = dot(l, v) - nl * nv;
  if (t > 0.0f)
    t /= fmaxf(nl, nv) + FLT_MIN;
  float is = nl * (sc.alpha_x + sc.alpha_y * t);
  return mk3(is, is, is);
}
CY_DEV f3 bsdf_oren_nayar_eval_reflect(const Closure &sc, f3 I, f3 omega_in, float *pdf)
{
  if (dot(sc.N, omega_in) > 0.0f) {
    *pdf = 0.5f * CY_1_PI_F;
    return bsdf_oren_nayar_get_intensity(sc, sc.N, I, omega_in);
  }
  *pdf = 0.0f;
  return zero3();
}
CY_DEV int bsdf_oren_nayar_sample(const Closure &sc, f3 Ng, f3 I, float randu, float randv,
                                  f3 *eval, f3 *omega_in, float *pdf)
{
  sample_uniform_hemisphere(sc.N, randu, randv, omega_in, pdf);
  if (dot(Ng, *omega_in) > 0.0f) {
    *eval = bsdf_oren_nayar_get_intensity(sc, sc.N, I, *omega_in);
  }
  else {
    *pdf = 0.0f;
    *eval = zero3();
  }
  return CY_LABEL_REFLECT | CY_LABEL_DIFFUSE;
}

/* closure/bsdf_diffuse.h:112-170: Translucent BSDF (Lambert on the far side) */
CY_DEV f3 bsdf_translucent_eval_transmit(const Closure &sc, f3 omega_in, float *pdf)
{
  float cos_pi = fmaxf(-dot(sc.N, omega_in), 0.0f) * CY_1_PI_F;
  *pdf = cos_pi;
  return mk3(cos_pi, cos_pi, cos_pi);
}
CY_DEV int bsdf_translucent_sample(const Closure &sc, f3 Ng, float randu, float randv, f3 *eval,
                                   f3 *omega_in, float *pdf)
{
  sample_cos_hemisphere(-sc.N, randu, randv, omega_in, pdf);
  if (dot(Ng, *omega_in) < 0)
    *eval = mk3(*pdf, *pdf, *pdf);
  else
    *pdf = 0;
  return CY_LABEL_TRANSMIT | CY_LABEL_DIFFUSE;
}

#include "bsdf_principled.cuh"

/* closure/bsdf_reflection.h, bsdf_refraction.h: the singular (sharp) closures - one
 * possible direction, evaluation is zero, the sample carries "some high number" */
CY_DEV int bsdf_reflection_sample(const Closure &sc, f3 Ng, f3 I, f3 *eval, f3 *omega_in,
                                  float *pdf)
{
  const f3 N = sc.N;
  float cosNO = dot(N, I);
  if (cosNO > 0) {
    *omega_in = (2 * cosNO) * N - I;
    if (dot(Ng, *omega_in) > 0) {
      *pdf = 1e6f;
      *eval = mk3(1e6f, 1e6f, 1e6f);
    }
  }
  return CY_LABEL_REFLECT | CY_LABEL_SINGULAR;
}
CY_DEV int bsdf_refraction_sample(const Closure &sc, f3 I, f3 *eval, f3 *omega_in, float *pdf)
{
  f3 R, T;
  bool inside;
  float fresnel = fresnel_dielectric(sc.ior, sc.N, I, &R, &T, &inside);
  if (!inside && fresnel != 1.0f) {
    *pdf = 1e6f;
    *eval = mk3(1e6f, 1e6f, 1e6f);
    *omega_in = T;
  }
  return CY_LABEL_TRANSMIT | CY_LABEL_SINGULAR;
}

/* closure/bsdf.h bsdf_eval: reflect side when dot(Ng, omega_in) >= 0 */
/* closure/bsdf.h:82-111: softened terminator for closures whose normal is not the
 * shading normal (a linked Normal input), and the per-object shadow terminator offset */
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
CY_DEV bool closure_is_bsdf_diffuse(int type)
{
  return type >= CY_CLOSURE_BSDF_DIFFUSE_ID && type <= CY_CLOSURE_BSDF_TRANSLUCENT_ID;
}

/* Decided once per shading point, after the shader ran: does any closure need the
 * terminator terms at all?  Almost never - and testing it per closure inside the
 * eval / sample loops cost the Cornell workload 11 %. */
CY_DEV void bsdf_terminator_terms_setup(ShaderDataG &sd)
{
  int need = (sd.terminator_freq > 1.0f) ? 1 : 0;
  for (int i = 0; i < sd.num_closure; i++)
    if (closure_is_bsdf_diffuse(sd.closure[i].type) && !isequal3(sd.closure[i].N, sd.N))
      need = 1;
  sd.terminator_terms = need;
}

/* EXT = false: the closures the lean interpreter can create (it hands shaders with sheen
 * to the full one, shade.cuh svm_eval_nodes) */
template<bool EXT>
CY_DEV f3 bsdf_eval(const ShaderDataG &sd, const Closure &sc, f3 omega_in, float *pdf)
{
  f3 eval = zero3();
  if (dot(sd.Ng, omega_in) >= 0.0f) {
    switch (sc.type) {
      case CY_CLOSURE_BSDF_DIFFUSE_ID:
        eval = bsdf_diffuse_eval_reflect(sc, omega_in, pdf);
        break;
      case CY_CLOSURE_BSDF_PRINCIPLED_DIFFUSE_ID:
        eval = bsdf_principled_diffuse_eval_reflect(sc, sd.I, omega_in, pdf);
        break;
      case CY_CLOSURE_BSDF_PRINCIPLED_SHEEN_ID:
        if (EXT)
          eval = bsdf_principled_sheen_eval_reflect(sc, sd.I, omega_in, pdf);
        break;
      case CY_CLOSURE_BSDF_OREN_NAYAR_ID:
        eval = bsdf_oren_nayar_eval_reflect(sc, sd.I, omega_in, pdf);
        break;
      case CY_CLOSURE_BSDF_MICROFACET_GGX_ID:
      case CY_CLOSURE_BSDF_MICROFACET_GGX_FRESNEL_ID:
      case CY_CLOSURE_BSDF_MICROFACET_GGX_CLEARCOAT_ID:
      case CY_CLOSURE_BSDF_MICROFACET_GGX_REFRACTION_ID:
        eval = bsdf_microfacet_ggx_eval_reflect(sc, sd.I, omega_in, pdf);
        break;
      default:
        break;
    }
    if (EXT && sd.terminator_terms) {
      if (closure_is_bsdf_diffuse(sc.type) && !isequal3(sc.N, sd.N))
        eval *= bump_shadowing_term(sd.N, sc.N, omega_in);
      if (sd.terminator_freq > 1.0f)
        eval *= shift_cos_in(dot(omega_in, sc.N), sd.terminator_freq);
    }
  }
  else {
    switch (sc.type) {
      case CY_CLOSURE_BSDF_TRANSLUCENT_ID:
        eval = bsdf_translucent_eval_transmit(sc, omega_in, pdf);
        break;
      case CY_CLOSURE_BSDF_MICROFACET_GGX_ID:
      case CY_CLOSURE_BSDF_MICROFACET_GGX_FRESNEL_ID:
      case CY_CLOSURE_BSDF_MICROFACET_GGX_CLEARCOAT_ID:
      case CY_CLOSURE_BSDF_MICROFACET_GGX_REFRACTION_ID:
        eval = bsdf_microfacet_ggx_eval_transmit(sc, sd.I, omega_in, pdf);
        break;
      default:
        break;
    }
    if (EXT && sd.terminator_terms && closure_is_bsdf_diffuse(sc.type) &&
        !isequal3(sc.N, sd.N))
      eval *= bump_shadowing_term(-sd.N, sc.N, omega_in);
  }
  return eval;
}

/* closure/bsdf.h bsdf_sample */
template<bool EXT>
CY_DEV int bsdf_sample_closure(const ShaderDataG &sd, const Closure &sc, float randu,
                               float randv, f3 *eval, f3 *omega_in, float *pdf)
{
  switch (sc.type) {
    case CY_CLOSURE_BSDF_DIFFUSE_ID:
      return bsdf_diffuse_sample(sc, sd.Ng, randu, randv, eval, omega_in, pdf);
    case CY_CLOSURE_BSDF_PRINCIPLED_DIFFUSE_ID:
      return bsdf_principled_diffuse_sample(sc, sd.Ng, sd.I, randu, randv, eval, omega_in, pdf);
    case CY_CLOSURE_BSDF_PRINCIPLED_SHEEN_ID:
      if (EXT)
        return bsdf_principled_sheen_sample(sc, sd.Ng, sd.I, randu, randv, eval, omega_in, pdf);
      *pdf = 0.0f;
      return CY_LABEL_NONE;
    case CY_CLOSURE_BSDF_OREN_NAYAR_ID:
      return bsdf_oren_nayar_sample(sc, sd.Ng, sd.I, randu, randv, eval, omega_in, pdf);
    case CY_CLOSURE_BSDF_TRANSLUCENT_ID:
      return bsdf_translucent_sample(sc, sd.Ng, randu, randv, eval, omega_in, pdf);
    case CY_CLOSURE_BSDF_REFLECTION_ID:
      return bsdf_reflection_sample(sc, sd.Ng, sd.I, eval, omega_in, pdf);
    case CY_CLOSURE_BSDF_REFRACTION_ID:
      return bsdf_refraction_sample(sc, sd.I, eval, omega_in, pdf);
    case CY_CLOSURE_BSDF_TRANSPARENT_ID:
      /* closure/bsdf_transparent.h:103-125: straight through */
      *omega_in = -sd.I;
      *pdf = 1;
      *eval = one3();
      return CY_LABEL_TRANSMIT | CY_LABEL_TRANSPARENT;
    case CY_CLOSURE_BSDF_MICROFACET_GGX_ID:
    case CY_CLOSURE_BSDF_MICROFACET_GGX_FRESNEL_ID:
    case CY_CLOSURE_BSDF_MICROFACET_GGX_CLEARCOAT_ID:
    case CY_CLOSURE_BSDF_MICROFACET_GGX_REFRACTION_ID:
      return bsdf_microfacet_ggx_sample(sc, sd.Ng, sd.I, randu, randv, eval, omega_in, pdf);
    default:
      *pdf = 0.0f;
      return CY_LABEL_NONE;
  }
}

/* closure/bsdf.h bsdf_sample: the closure's own sampler, then the terminator terms on
 * the reflection side (closure/bsdf.h:466-489) */
template<bool EXT>
CY_DEV int bsdf_sample(const ShaderDataG &sd, const Closure &sc, float randu, float randv,
                       f3 *eval, f3 *omega_in, float *pdf)
{
  const int label = bsdf_sample_closure<EXT>(sd, sc, randu, randv, eval, omega_in, pdf);
  if (EXT && sd.terminator_terms && !(label & CY_LABEL_TRANSMIT)) {
    if (sd.terminator_freq > 1.0f)
      *eval *= shift_cos_in(dot(*omega_in, sc.N), sd.terminator_freq);
    if ((label & CY_LABEL_DIFFUSE) && !isequal3(sc.N, sd.N))
      *eval *= bump_shadowing_term(sd.N, sc.N, *omega_in);
  }
  return label;
}

#endif

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

#include "bsdf_principled.cuh"

/* closure/bsdf.h bsdf_eval: reflect side when dot(Ng, omega_in) >= 0 */
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
      case CY_CLOSURE_BSDF_MICROFACET_GGX_ID:
      case CY_CLOSURE_BSDF_MICROFACET_GGX_FRESNEL_ID:
      case CY_CLOSURE_BSDF_MICROFACET_GGX_CLEARCOAT_ID:
      case CY_CLOSURE_BSDF_MICROFACET_GGX_REFRACTION_ID:
        eval = bsdf_microfacet_ggx_eval_reflect(sc, sd.I, omega_in, pdf);
        break;
      default:
        break;
    }
  }
  else {
    switch (sc.type) {
      case CY_CLOSURE_BSDF_MICROFACET_GGX_ID:
      case CY_CLOSURE_BSDF_MICROFACET_GGX_FRESNEL_ID:
      case CY_CLOSURE_BSDF_MICROFACET_GGX_CLEARCOAT_ID:
      case CY_CLOSURE_BSDF_MICROFACET_GGX_REFRACTION_ID:
        eval = bsdf_microfacet_ggx_eval_transmit(sc, sd.I, omega_in, pdf);
        break;
      default:
        break;
    }
  }
  return eval;
}

/* closure/bsdf.h bsdf_sample */
CY_DEV int bsdf_sample(const ShaderDataG &sd, const Closure &sc, float randu, float randv,
                       f3 *eval, f3 *omega_in, float *pdf)
{
  switch (sc.type) {
    case CY_CLOSURE_BSDF_DIFFUSE_ID:
      return bsdf_diffuse_sample(sc, sd.Ng, randu, randv, eval, omega_in, pdf);
    case CY_CLOSURE_BSDF_PRINCIPLED_DIFFUSE_ID:
      return bsdf_principled_diffuse_sample(sc, sd.Ng, sd.I, randu, randv, eval, omega_in, pdf);
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

#endif

/* bsdf_principled.cuh - Principled-diffuse and GGX microfacet closures
 * (kernel/closure/bsdf_principled_diffuse.h:34-135, bsdf_microfacet.h:150-790,
 * bsdf_util.h:38-149).  Included by bsdf.cuh. */
#ifndef B200_BSDF_PRINCIPLED_CUH
#define B200_BSDF_PRINCIPLED_CUH

/* bsdf_util.h:38-102, without differentials */
CY_DEV float fresnel_dielectric(float eta, f3 N, f3 I, f3 *R, f3 *T, bool *is_inside)
{
  float cos = dot(N, I), neta;
  f3 Nn;
  if (cos > 0) {
    neta = 1 / eta;
    Nn = N;
    *is_inside = false;
  }
  else {
    cos = -cos;
    neta = eta;
    Nn = -N;
    *is_inside = true;
  }
  *R = (2 * cos) * Nn - I;
  float arg = 1 - (neta * neta * (1 - (cos * cos)));
  if (arg < 0) {
    *T = zero3();
    return 1;
  }
  else {
    float dnp = fmaxf(sqrtf(arg), 1e-7f);
    float nK = (neta * cos) - dnp;
    *T = -(neta * I) + (nK * Nn);
    float cosTheta1 = cos;
    float cosTheta2 = -dot(Nn, *T);
    float pPara = (cosTheta1 - eta * cosTheta2) / (cosTheta1 + eta * cosTheta2);
    float pPerp = (eta * cosTheta1 - cosTheta2) / (eta * cosTheta1 + cosTheta2);
    return 0.5f * (pPara * pPara + pPerp * pPerp);
  }
}

/* bsdf_util.h:130-149 */
CY_DEV float schlick_fresnel(float u)
{
  float m = clampf(1.0f - u, 0.0f, 1.0f);
  float m2 = m * m;
  return m2 * m2 * m;
}
CY_DEV f3 interpolate_fresnel_color(f3 L, f3 H, float ior, float F0, f3 cspec0)
{
  float F0_norm = 1.0f / (1.0f - F0);
  float FH = (fresnel_dielectric_cos(dot(L, H), ior) - F0) * F0_norm;
  return cspec0 * (1.0f - FH) + one3() * FH;
}

/* ---- principled diffuse: bsdf_principled_diffuse.h:36-135 ---- */
CY_DEV f3 calculate_principled_diffuse_brdf(const Closure &bsdf, f3 N, f3 V, f3 L, f3 H,
                                            float *pdf)
{
  float NdotL = fmaxf(dot(N, L), 0.0f);
  float NdotV = fmaxf(dot(N, V), 0.0f);
  if (NdotL < 0 || NdotV < 0) {
    *pdf = 0.0f;
    return zero3();
  }
  float LdotH = dot(L, H);
  float FL = schlick_fresnel(NdotL), FV = schlick_fresnel(NdotV);
  const float Fd90 = 0.5f + 2.0f * LdotH * LdotH * bsdf.roughness;
  float Fd = (1.0f * (1.0f - FL) + Fd90 * FL) * (1.0f * (1.0f - FV) + Fd90 * FV);
  float value = CY_1_PI_F * NdotL * Fd;
  return mk3(value, value, value);
}
CY_DEV f3 bsdf_principled_diffuse_eval_reflect(const Closure &bsdf, f3 I, f3 omega_in, float *pdf)
{
  f3 N = bsdf.N;
  f3 V = I;
  f3 L = omega_in;
  f3 H = normalize(L + V);
  if (dot(N, omega_in) > 0.0f) {
    *pdf = fmaxf(dot(N, omega_in), 0.0f) * CY_1_PI_F;
    return calculate_principled_diffuse_brdf(bsdf, N, V, L, H, pdf);
  }
  *pdf = 0.0f;
  return zero3();
}
CY_DEV int bsdf_principled_diffuse_sample(const Closure &bsdf, f3 Ng, f3 I, float randu,
                                          float randv, f3 *eval, f3 *omega_in, float *pdf)
{
  f3 N = bsdf.N;
  sample_cos_hemisphere(N, randu, randv, omega_in, pdf);
  if (dot(Ng, *omega_in) > 0) {
    f3 H = normalize(I + *omega_in);
    *eval = calculate_principled_diffuse_brdf(bsdf, N, I, *omega_in, H, pdf);
  }
  else {
    *pdf = 0.0f;
  }
  return CY_LABEL_REFLECT | CY_LABEL_DIFFUSE;
}

/* ---- Principled sheen (closure/bsdf_principled_sheen.h:47-134) ---- */
#ifndef SHEEN_INLINE
#  define SHEEN_INLINE 0
#endif
#if SHEEN_INLINE
#  define SHEEN_FN CY_DEV
#else
#  define SHEEN_FN __device__ __noinline__
#endif


CY_DEV f3 principled_sheen_brdf(f3 N, f3 V, f3 L, f3 H, float *pdf)
{
  const float NdotL = dot(N, L), NdotV = dot(N, V);
  if (NdotL < 0 || NdotV < 0) {
    *pdf = 0.0f;
    return zero3();
  }
  const float value = schlick_fresnel(dot(L, H)) * NdotL;
  return mk3(value, value, value);
}
SHEEN_FN f3 bsdf_principled_sheen_eval_reflect(const Closure &bsdf, f3 I, f3 omega_in, float *pdf)
{
  const f3 N = bsdf.N;
  const f3 H = normalize(omega_in + I);
  if (dot(N, omega_in) > 0.0f) {
    *pdf = fmaxf(dot(N, omega_in), 0.0f) * CY_M_1_PI_F;
    return principled_sheen_brdf(N, I, omega_in, H, pdf);
  }
  *pdf = 0.0f;
  return zero3();
}
SHEEN_FN int bsdf_principled_sheen_sample(const Closure &bsdf, f3 Ng, f3 I, float randu,
                                        float randv, f3 *eval, f3 *omega_in, float *pdf)
{
  const f3 N = bsdf.N;
  sample_cos_hemisphere(N, randu, randv, omega_in, pdf);
  if (dot(Ng, *omega_in) > 0) {
    const f3 H = normalize(I + *omega_in);
    *eval = principled_sheen_brdf(N, I, *omega_in, H, pdf);
  }
  else {
    *pdf = 0.0f;
  }
  return CY_LABEL_REFLECT | CY_LABEL_DIFFUSE;
}

/* ---- GGX microfacet ---- */

/* kernel_montecarlo.h:50-54 */
CY_DEV void make_orthonormals_tangent(f3 N, f3 T, f3 *a, f3 *b)
{
  *b = normalize(cross(N, T));
  *a = cross(*b, N);
}

/* bsdf_microfacet.h:143-193 */
CY_DEV void microfacet_ggx_sample_slopes(const float cos_theta_i, const float sin_theta_i,
                                         float randu, float randv, float *slope_x,
                                         float *slope_y, float *G1i)
{
  if (cos_theta_i >= 0.99999f) {
    const float r = sqrtf(randu / (1.0f - randu));
    const float phi = CY_2PI_F * randv;
    *slope_x = r * cosf(phi);
    *slope_y = r * sinf(phi);
    *G1i = 1.0f;
    return;
  }
  const float tan_theta_i = sin_theta_i / cos_theta_i;
  const float G1_inv = 0.5f * (1.0f + safe_sqrtf(1.0f + tan_theta_i * tan_theta_i));
  *G1i = 1.0f / G1_inv;
  const float A = 2.0f * randu * G1_inv - 1.0f;
  const float AA = A * A;
  const float tmp = 1.0f / (AA - 1.0f);
  const float B = tan_theta_i;
  const float BB = B * B;
  const float D = safe_sqrtf(BB * (tmp * tmp) - (AA - BB) * tmp);
  const float slope_x_1 = B * tmp - D;
  const float slope_x_2 = B * tmp + D;
  *slope_x = (A < 0.0f || slope_x_2 * tan_theta_i > 1.0f) ? slope_x_1 : slope_x_2;
  float S;
  if (randv > 0.5f) {
    S = 1.0f;
    randv = 2.0f * (randv - 0.5f);
  }
  else {
    S = -1.0f;
    randv = 2.0f * (0.5f - randv);
  }
  const float z = (randv * (randv * (randv * 0.27385f - 0.73369f) + 0.46341f)) /
                  (randv * (randv * (randv * 0.093073f + 0.309420f) - 1.000000f) + 0.597999f);
  *slope_y = S * z * safe_sqrtf(1.0f + (*slope_x) * (*slope_x));
}

/* bsdf_microfacet.h:195-245 (GGX branch) */
CY_DEV f3 microfacet_sample_stretched(f3 omega_i, float alpha_x, float alpha_y, float randu,
                                      float randv, float *G1i)
{
  f3 omega_i_ = mk3(alpha_x * omega_i.x, alpha_y * omega_i.y, omega_i.z);
  omega_i_ = normalize(omega_i_);
  float costheta_ = 1.0f;
  float sintheta_ = 0.0f;
  float cosphi_ = 1.0f;
  float sinphi_ = 0.0f;
  if (omega_i_.z < 0.99999f) {
    costheta_ = omega_i_.z;
    sintheta_ = safe_sqrtf(1.0f - costheta_ * costheta_);
    float invlen = 1.0f / sintheta_;
    cosphi_ = omega_i_.x * invlen;
    sinphi_ = omega_i_.y * invlen;
  }
  float slope_x, slope_y;
  microfacet_ggx_sample_slopes(costheta_, sintheta_, randu, randv, &slope_x, &slope_y, G1i);
  float tmp = cosphi_ * slope_x - sinphi_ * slope_y;
  slope_y = sinphi_ * slope_x + cosphi_ * slope_y;
  slope_x = tmp;
  slope_x = alpha_x * slope_x;
  slope_y = alpha_y * slope_y;
  return normalize(mk3(-slope_x, -slope_y, 1.0f));
}

/* bsdf_microfacet.h:253-266 */
CY_DEV f3 reflection_color(const Closure &bsdf, f3 L, f3 H)
{
  f3 F = one3();
  bool use_fresnel = (bsdf.type == CY_CLOSURE_BSDF_MICROFACET_GGX_FRESNEL_ID ||
                      bsdf.type == CY_CLOSURE_BSDF_MICROFACET_GGX_CLEARCOAT_ID);
  if (use_fresnel) {
    float F0 = fresnel_dielectric_cos(1.0f, bsdf.ior);
    F = interpolate_fresnel_color(L, H, bsdf.ior, F0, bsdf.cspec0);
  }
  return F;
}
CY_DEV float D_GTR1(float NdotH, float alpha)
{
  if (alpha >= 1.0f)
    return CY_1_PI_F;
  float alpha2 = alpha * alpha;
  float t = 1.0f + (alpha2 - 1.0f) * NdotH * NdotH;
  return (alpha2 - 1.0f) / (CY_PI_F * logf(alpha2) * t);
}

/* bsdf_microfacet.h:397-497 */
CY_DEV f3 bsdf_microfacet_ggx_eval_reflect(const Closure &bsdf, f3 I, f3 omega_in, float *pdf)
{
  float alpha_x = bsdf.alpha_x;
  float alpha_y = bsdf.alpha_y;
  bool m_refractive = bsdf.type == CY_CLOSURE_BSDF_MICROFACET_GGX_REFRACTION_ID;
  f3 N = bsdf.N;
  if (m_refractive || alpha_x * alpha_y <= 1e-7f)
    return zero3();
  float cosNO = dot(N, I);
  float cosNI = dot(N, omega_in);
  if (cosNI > 0 && cosNO > 0) {
    f3 m = normalize(omega_in + I);
    float alpha2 = alpha_x * alpha_y;
    float D, G1o, G1i;
    if (alpha_x == alpha_y) {
      float cosThetaM = dot(N, m);
      float cosThetaM2 = cosThetaM * cosThetaM;
      float cosThetaM4 = cosThetaM2 * cosThetaM2;
      float tanThetaM2 = (1 - cosThetaM2) / cosThetaM2;
      if (bsdf.type == CY_CLOSURE_BSDF_MICROFACET_GGX_CLEARCOAT_ID) {
        D = D_GTR1(cosThetaM, bsdf.alpha_x);
        alpha2 = 0.0625f;
      }
      else {
        D = alpha2 / (CY_PI_F * cosThetaM4 * (alpha2 + tanThetaM2) * (alpha2 + tanThetaM2));
      }
      G1o = 2 / (1 + safe_sqrtf(1 + alpha2 * (1 - cosNO * cosNO) / (cosNO * cosNO)));
      G1i = 2 / (1 + safe_sqrtf(1 + alpha2 * (1 - cosNI * cosNI) / (cosNI * cosNI)));
    }
    else {
      f3 X, Y, Z = N;
      make_orthonormals_tangent(Z, bsdf.T, &X, &Y);
      f3 local_m = mk3(dot(X, m), dot(Y, m), dot(Z, m));
      float slope_x = -local_m.x / (local_m.z * alpha_x);
      float slope_y = -local_m.y / (local_m.z * alpha_y);
      float slope_len = 1 + slope_x * slope_x + slope_y * slope_y;
      float cosThetaM = local_m.z;
      float cosThetaM2 = cosThetaM * cosThetaM;
      float cosThetaM4 = cosThetaM2 * cosThetaM2;
      D = 1 / ((slope_len * slope_len) * CY_PI_F * alpha2 * cosThetaM4);
      float tanThetaO2 = (1 - cosNO * cosNO) / (cosNO * cosNO);
      float cosPhiO = dot(I, X);
      float sinPhiO = dot(I, Y);
      float alphaO2 = (cosPhiO * cosPhiO) * (alpha_x * alpha_x) +
                      (sinPhiO * sinPhiO) * (alpha_y * alpha_y);
      alphaO2 /= cosPhiO * cosPhiO + sinPhiO * sinPhiO;
      G1o = 2 / (1 + safe_sqrtf(1 + alphaO2 * tanThetaO2));
      float tanThetaI2 = (1 - cosNI * cosNI) / (cosNI * cosNI);
      float cosPhiI = dot(omega_in, X);
      float sinPhiI = dot(omega_in, Y);
      float alphaI2 = (cosPhiI * cosPhiI) * (alpha_x * alpha_x) +
                      (sinPhiI * sinPhiI) * (alpha_y * alpha_y);
      alphaI2 /= cosPhiI * cosPhiI + sinPhiI * sinPhiI;
      G1i = 2 / (1 + safe_sqrtf(1 + alphaI2 * tanThetaI2));
    }
    float G = G1o * G1i;
    float common = D * 0.25f / cosNO;
    f3 F = reflection_color(bsdf, omega_in, m);
    if (bsdf.type == CY_CLOSURE_BSDF_MICROFACET_GGX_CLEARCOAT_ID)
      F *= 0.25f * bsdf.clearcoat;
    f3 out = F * G * common;
    *pdf = G1o * common;
    return out;
  }
  return zero3();
}

/* bsdf_microfacet.h:499-560 */
CY_DEV f3 bsdf_microfacet_ggx_eval_transmit(const Closure &bsdf, f3 I, f3 omega_in, float *pdf)
{
  float alpha_x = bsdf.alpha_x;
  float alpha_y = bsdf.alpha_y;
  float m_eta = bsdf.ior;
  bool m_refractive = bsdf.type == CY_CLOSURE_BSDF_MICROFACET_GGX_REFRACTION_ID;
  f3 N = bsdf.N;
  if (!m_refractive || alpha_x * alpha_y <= 1e-7f)
    return zero3();
  float cosNO = dot(N, I);
  float cosNI = dot(N, omega_in);
  if (cosNO <= 0 || cosNI >= 0)
    return zero3();
  f3 ht = -(m_eta * omega_in + I);
  f3 Ht = normalize(ht);
  float cosHO = dot(Ht, I);
  float cosHI = dot(Ht, omega_in);
  float D, G1o, G1i;
  float alpha2 = alpha_x * alpha_y;
  float cosThetaM = dot(N, Ht);
  float cosThetaM2 = cosThetaM * cosThetaM;
  float tanThetaM2 = (1 - cosThetaM2) / cosThetaM2;
  float cosThetaM4 = cosThetaM2 * cosThetaM2;
  D = alpha2 / (CY_PI_F * cosThetaM4 * (alpha2 + tanThetaM2) * (alpha2 + tanThetaM2));
  G1o = 2 / (1 + safe_sqrtf(1 + alpha2 * (1 - cosNO * cosNO) / (cosNO * cosNO)));
  G1i = 2 / (1 + safe_sqrtf(1 + alpha2 * (1 - cosNI * cosNI) / (cosNI * cosNI)));
  float G = G1o * G1i;
  float Ht2 = dot(ht, ht);
  float common = D * (m_eta * m_eta) / (cosNO * Ht2);
  float out = G * fabsf(cosHI * cosHO) * common;
  *pdf = G1o * fabsf(cosHO * cosHI) * common;
  return mk3(out, out, out);
}

/* bsdf_microfacet.h:562-790 */
CY_DEV int bsdf_microfacet_ggx_sample(const Closure &bsdf, f3 Ng, f3 I, float randu, float randv,
                                      f3 *eval, f3 *omega_in, float *pdf)
{
  float alpha_x = bsdf.alpha_x;
  float alpha_y = bsdf.alpha_y;
  bool m_refractive = bsdf.type == CY_CLOSURE_BSDF_MICROFACET_GGX_REFRACTION_ID;
  f3 N = bsdf.N;
  int label;
  float cosNO = dot(N, I);
  if (cosNO > 0) {
    f3 X, Y, Z = N;
    if (alpha_x == alpha_y)
      make_orthonormals(Z, &X, &Y);
    else
      make_orthonormals_tangent(Z, bsdf.T, &X, &Y);
    f3 local_I = mk3(dot(X, I), dot(Y, I), cosNO);
    f3 local_m;
    float G1o;
    local_m = microfacet_sample_stretched(local_I, alpha_x, alpha_y, randu, randv, &G1o);
    f3 m = X * local_m.x + Y * local_m.y + Z * local_m.z;
    float cosThetaM = local_m.z;
    if (!m_refractive) {
      float cosMO = dot(m, I);
      label = CY_LABEL_REFLECT | CY_LABEL_GLOSSY;
      if (cosMO > 0) {
        *omega_in = 2 * cosMO * m - I;
        if (dot(Ng, *omega_in) > 0) {
          if (alpha_x * alpha_y <= 1e-7f) {
            *pdf = 1e6f;
            *eval = mk3(1e6f, 1e6f, 1e6f);
            bool use_fresnel = (bsdf.type == CY_CLOSURE_BSDF_MICROFACET_GGX_FRESNEL_ID ||
                                bsdf.type == CY_CLOSURE_BSDF_MICROFACET_GGX_CLEARCOAT_ID);
            if (use_fresnel)
              *eval *= reflection_color(bsdf, *omega_in, m);
            label = CY_LABEL_REFLECT | CY_LABEL_SINGULAR;
          }
          else {
            float alpha2 = alpha_x * alpha_y;
            float D, G1i;
            if (alpha_x == alpha_y) {
              float cosThetaM2 = cosThetaM * cosThetaM;
              float cosThetaM4 = cosThetaM2 * cosThetaM2;
              float tanThetaM2 = 1 / (cosThetaM2)-1;
              float cosNI = dot(N, *omega_in);
              if (bsdf.type == CY_CLOSURE_BSDF_MICROFACET_GGX_CLEARCOAT_ID) {
                D = D_GTR1(cosThetaM, bsdf.alpha_x);
                alpha2 = 0.0625f;
                G1o = 2 / (1 + safe_sqrtf(1 + alpha2 * (1 - cosNO * cosNO) / (cosNO * cosNO)));
              }
              else {
                D = alpha2 / (CY_PI_F * cosThetaM4 * (alpha2 + tanThetaM2) * (alpha2 + tanThetaM2));
              }
              G1i = 2 / (1 + safe_sqrtf(1 + alpha2 * (1 - cosNI * cosNI) / (cosNI * cosNI)));
            }
            else {
              f3 lm = mk3(dot(X, m), dot(Y, m), dot(Z, m));
              float slope_x = -lm.x / (lm.z * alpha_x);
              float slope_y = -lm.y / (lm.z * alpha_y);
              float slope_len = 1 + slope_x * slope_x + slope_y * slope_y;
              float cosThetaMa = lm.z;
              float cosThetaM2 = cosThetaMa * cosThetaMa;
              float cosThetaM4 = cosThetaM2 * cosThetaM2;
              D = 1 / ((slope_len * slope_len) * CY_PI_F * alpha2 * cosThetaM4);
              float cosNI = dot(N, *omega_in);
              float tanThetaI2 = (1 - cosNI * cosNI) / (cosNI * cosNI);
              float cosPhiI = dot(*omega_in, X);
              float sinPhiI = dot(*omega_in, Y);
              float alphaI2 = (cosPhiI * cosPhiI) * (alpha_x * alpha_x) +
                              (sinPhiI * sinPhiI) * (alpha_y * alpha_y);
              alphaI2 /= cosPhiI * cosPhiI + sinPhiI * sinPhiI;
              G1i = 2 / (1 + safe_sqrtf(1 + alphaI2 * tanThetaI2));
            }
            float common = (G1o * D) * 0.25f / cosNO;
            *pdf = common;
            f3 F = reflection_color(bsdf, *omega_in, m);
            *eval = G1i * common * F;
          }
          if (bsdf.type == CY_CLOSURE_BSDF_MICROFACET_GGX_CLEARCOAT_ID)
            *eval *= 0.25f * bsdf.clearcoat;
        }
      }
    }
    else {
      label = CY_LABEL_TRANSMIT | CY_LABEL_GLOSSY;
      f3 R, T;
      float m_eta = bsdf.ior, fresnel;
      bool inside;
      fresnel = fresnel_dielectric(m_eta, m, I, &R, &T, &inside);
      if (!inside && fresnel != 1.0f) {
        *omega_in = T;
        if (alpha_x * alpha_y <= 1e-7f || fabsf(m_eta - 1.0f) < 1e-4f) {
          *pdf = 1e6f;
          *eval = mk3(1e6f, 1e6f, 1e6f);
          label = CY_LABEL_TRANSMIT | CY_LABEL_SINGULAR;
        }
        else {
          float alpha2 = alpha_x * alpha_y;
          float cosThetaM2 = cosThetaM * cosThetaM;
          float cosThetaM4 = cosThetaM2 * cosThetaM2;
          float tanThetaM2 = 1 / (cosThetaM2)-1;
          float D = alpha2 / (CY_PI_F * cosThetaM4 * (alpha2 + tanThetaM2) * (alpha2 + tanThetaM2));
          float cosNI = dot(N, *omega_in);
          float G1i = 2 / (1 + safe_sqrtf(1 + alpha2 * (1 - cosNI * cosNI) / (cosNI * cosNI)));
          float cosHI = dot(m, *omega_in);
          float cosHO = dot(m, I);
          float Ht2 = m_eta * cosHI + cosHO;
          Ht2 *= Ht2;
          float common = (G1o * D) * (m_eta * m_eta) / (cosNO * Ht2);
          float out = G1i * fabsf(cosHI * cosHO) * common;
          *pdf = cosHO * fabsf(cosHI) * common;
          *eval = mk3(out, out, out);
        }
      }
    }
  }
  else {
    label = (m_refractive) ? CY_LABEL_TRANSMIT | CY_LABEL_GLOSSY :
                             CY_LABEL_REFLECT | CY_LABEL_GLOSSY;
  }
  return label;
}

#endif

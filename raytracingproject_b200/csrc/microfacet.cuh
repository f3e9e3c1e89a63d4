/* microfacet.cuh - the GGX lobes (single scattering): Smith-GGX terms as small pure
 * functions, one evaluator and one visible-normal sampler built from them.
 *
 * Semantics to match (reference = blender/intern/cycles/kernel/closure):
 *   bsdf_microfacet.h   MicrofacetBsdf with GGX distribution: reflection, refraction, the
 *                       Fresnel-tinted and the clearcoat (GTR1) variants, isotropic and
 *                       anisotropic, incl. the "singular" limit alpha_x * alpha_y <= 1e-7
 *   bsdf_util.h         dielectric Fresnel with refraction direction, Schlick weight,
 *                       the Principled tint interpolation
 * The random-number -> direction mapping is the published visible-normal sampling of
 * Heitz & d'Eon 2014 (stretch, sample slopes of a unit-roughness surface, rotate,
 * unstretch), the same mapping the reference uses, so a path that draws the same numbers
 * leaves in the same direction.  What is different is the shape: the reference spells the
 * D / G1 terms out five times (two evaluators, three sampler branches); here each term is
 * one function of (alpha^2, cosine) and the lobe variants differ only in which terms they
 * combine.  Included by bsdf.cuh; host-compilable (tests/host_check). */
#ifndef B200_MICROFACET_CUH
#define B200_MICROFACET_CUH

/* ------------------------------------------------------------ Fresnel */

/* (1 - u)^5 clamped */
CY_DEV float schlick_weight(float u)
{
  const float m = clampf(1.0f - u, 0.0f, 1.0f);
  const float m2 = m * m;
  return m2 * m2 * m;
}

/* Principled specular tint: blends cspec0 towards white by how far the dielectric
 * Fresnel term has moved from its normal-incidence value F0 */
CY_DEV f3 fresnel_tint(f3 L, f3 H, float ior, float F0, f3 cspec0)
{
  const float scale = 1.0f / (1.0f - F0);
  const float FH = (fresnel_dielectric_cos(dot(L, H), ior) - F0) * scale;
  return cspec0 * (1.0f - FH) + one3() * FH;
}

/* Both halves of a dielectric interface for the incident direction I about N: mirror
 * direction, refracted direction, reflectance.  `inside` when I arrives from below N;
 * reflectance 1 and a zero refraction direction on total internal reflection. */
struct DielectricSplit {
  f3 reflected, refracted;
  float reflectance;
  bool inside;
};
CY_DEV DielectricSplit dielectric_split(float eta, f3 N, f3 I)
{
  DielectricSplit s;
  float c = dot(N, I);
  float rel; /* relative index seen by I */
  f3 Nf;     /* N on I's side */
  s.inside = !(c > 0.0f);
  if (s.inside) {
    c = -c;
    rel = eta;
    Nf = -N;
  }
  else {
    rel = 1.0f / eta;
    Nf = N;
  }
  s.reflected = (2.0f * c) * Nf - I;
  const float under_root = 1.0f - (rel * rel * (1.0f - (c * c)));
  if (under_root < 0.0f) {
    s.refracted = zero3();
    s.reflectance = 1.0f;
    return s;
  }
  const float ct = fmaxf(sqrtf(under_root), 1e-7f);
  s.refracted = -(rel * I) + ((rel * c) - ct) * Nf;
  const float c2 = -dot(Nf, s.refracted);
  const float r_par = (c - eta * c2) / (c + eta * c2);
  const float r_perp = (eta * c - c2) / (eta * c + c2);
  s.reflectance = 0.5f * (r_par * r_par + r_perp * r_perp);
  return s;
}

/* --------------------------------------------------- Smith-GGX terms */

/* frame whose Y is perpendicular to the tangent T (anisotropic lobes) */
CY_DEV void tangent_frame(f3 N, f3 T, f3 *X, f3 *Y)
{
  *Y = normalize(cross(N, T));
  *X = cross(*Y, N);
}

/* masking of one direction: cosine c to the normal, squared roughness a2 along it */
CY_DEV float ggx_g1(float a2, float c)
{
  return 2.0f / (1.0f + safe_sqrtf(1.0f + a2 * (1.0f - c * c) / (c * c)));
}

/* squared roughness an anisotropic lobe presents along direction w */
CY_DEV float ggx_directional_a2(f3 w, f3 X, f3 Y, float ax, float ay)
{
  const float cx = dot(w, X), cy = dot(w, Y);
  const float a2 = (cx * cx) * (ax * ax) + (cy * cy) * (ay * ay);
  return a2 / (cx * cx + cy * cy);
}

/* normal distribution, isotropic, from the cosine of the microfacet normal */
CY_DEV float ggx_d(float a2, float cm)
{
  const float c2 = cm * cm;
  const float t2 = (1.0f - c2) / c2;
  return a2 / (CY_PI_F * (c2 * c2) * (a2 + t2) * (a2 + t2));
}

/* anisotropic, from the microfacet normal in the lobe's frame */
CY_DEV float ggx_d_aniso(f3 m, float ax, float ay)
{
  const float sx = -m.x / (m.z * ax), sy = -m.y / (m.z * ay);
  const float s = 1.0f + sx * sx + sy * sy;
  const float c2 = m.z * m.z;
  return 1.0f / ((s * s) * CY_PI_F * (ax * ay) * (c2 * c2));
}

/* Berry / GTR1 distribution of the clearcoat lobe */
CY_DEV float gtr1_d(float cm, float alpha)
{
  if (alpha >= 1.0f)
    return CY_1_PI_F;
  const float a2 = alpha * alpha;
  const float t = 1.0f + (a2 - 1.0f) * cm * cm;
  return (a2 - 1.0f) / (CY_PI_F * logf(a2) * t);
}

CY_DEV bool lobe_is_clearcoat(const Lobe &l)
{
  return lobe_id(l.kind) == CY_CLOSURE_BSDF_MICROFACET_GGX_CLEARCOAT_ID;
}
CY_DEV bool lobe_has_fresnel_tint(const Lobe &l)
{
  return lobe_id(l.kind) == CY_CLOSURE_BSDF_MICROFACET_GGX_FRESNEL_ID || lobe_is_clearcoat(l);
}
CY_DEV bool lobe_refracts(const Lobe &l)
{
  return lobe_id(l.kind) == CY_CLOSURE_BSDF_MICROFACET_GGX_REFRACTION_ID;
}

/* colour of the reflection off microfacet m towards L: white, or the Principled tint */
CY_DEV f3 ggx_reflection_tint(const Lobe &l, f3 L, f3 m)
{
  if (!lobe_has_fresnel_tint(l))
    return one3();
  const float F0 = fresnel_dielectric_cos(1.0f, l.ior);
  return fresnel_tint(L, m, l.ior, F0, l.cspec0);
}

/* D and the two masking terms of a reflection between `wo` and `wi` about microfacet m.
 * The clearcoat lobe swaps in GTR1 for D and a fixed 0.25 roughness for the masking. */
CY_DEV void ggx_reflection_terms(const Lobe &l, f3 wo, f3 wi, f3 m, float cos_o, float cos_i,
                                 float *D, float *G1o, float *G1i)
{
  if (l.ax == l.ay) {
    float a2 = l.ax * l.ay;
    const float cm = dot(l.N, m);
    if (lobe_is_clearcoat(l)) {
      *D = gtr1_d(cm, l.ax);
      a2 = 0.0625f;
    }
    else {
      *D = ggx_d(a2, cm);
    }
    *G1o = ggx_g1(a2, cos_o);
    *G1i = ggx_g1(a2, cos_i);
  }
  else {
    f3 X, Y;
    tangent_frame(l.N, l.T, &X, &Y);
    *D = ggx_d_aniso(mk3(dot(X, m), dot(Y, m), dot(l.N, m)), l.ax, l.ay);
    *G1o = ggx_g1(ggx_directional_a2(wo, X, Y, l.ax, l.ay), cos_o);
    *G1i = ggx_g1(ggx_directional_a2(wi, X, Y, l.ax, l.ay), cos_i);
  }
}

/* ---------------------------------------------------------- evaluation */

/* Value (cosine included, as all Cycles BSDFs) and sampling pdf of the lobe for the pair
 * (wo = sd.I, wi).  `same_side` = wi is on the geometric-normal side of wo. */
CY_DEV f3 ggx_eval(const Lobe &l, f3 wo, f3 wi, bool same_side, float *pdf)
{
  if (lobe_refracts(l) == same_side || l.ax * l.ay <= 1e-7f)
    return zero3();
  const float cos_o = dot(l.N, wo), cos_i = dot(l.N, wi);
  if (same_side) {
    if (!(cos_i > 0.0f && cos_o > 0.0f))
      return zero3();
    const f3 m = normalize(wi + wo);
    float D, G1o, G1i;
    ggx_reflection_terms(l, wo, wi, m, cos_o, cos_i, &D, &G1o, &G1i);
    const float common = D * 0.25f / cos_o;
    f3 F = ggx_reflection_tint(l, wi, m);
    if (lobe_is_clearcoat(l))
      F *= 0.25f * l.aux;
    *pdf = G1o * common;
    return F * (G1o * G1i) * common;
  }
  /* refraction through the half vector of the interface */
  if (cos_o <= 0.0f || cos_i >= 0.0f)
    return zero3();
  const float eta = l.ior;
  const f3 h = -(eta * wi + wo);
  const f3 m = normalize(h);
  const float a2 = l.ax * l.ay;
  const float D = ggx_d(a2, dot(l.N, m));
  const float G1o = ggx_g1(a2, cos_o), G1i = ggx_g1(a2, cos_i);
  const float common = D * (eta * eta) / (cos_o * dot(h, h));
  const float cc = fabsf(dot(m, wi) * dot(m, wo));
  const float value = (G1o * G1i) * cc * common;
  *pdf = G1o * cc * common;
  return mk3(value, value, value);
}

/* ------------------------------------------------- visible-normal sampling */

/* slope of a unit-roughness GGX surface visible from the direction (sin_v, 0, cos_v);
 * also returns the masking term of that direction */
CY_DEV float2 ggx_unit_slopes(float cos_v, float sin_v, float u1, float u2, float *G1)
{
  if (cos_v >= 0.99999f) { /* normal incidence: the distribution is radially symmetric */
    const float r = sqrtf(u1 / (1.0f - u1));
    const float phi = CY_2PI_F * u2;
    *G1 = 1.0f;
    return make_float2(r * cosf(phi), r * sinf(phi));
  }
  const float tan_v = sin_v / cos_v;
  const float inv_g1 = 0.5f * (1.0f + safe_sqrtf(1.0f + tan_v * tan_v));
  *G1 = 1.0f / inv_g1;
  /* slope along the view direction: inverse of the marginal cdf */
  const float A = 2.0f * u1 * inv_g1 - 1.0f;
  const float AA = A * A;
  const float k = 1.0f / (AA - 1.0f);
  const float tt = tan_v * tan_v;
  const float disc = safe_sqrtf(tt * (k * k) - (AA - tt) * k);
  const float lo = tan_v * k - disc, hi = tan_v * k + disc;
  const float sx = (A < 0.0f || hi * tan_v > 1.0f) ? lo : hi;
  /* slope across: rational fit of the conditional cdf, symmetric about zero */
  const bool upper = u2 > 0.5f;
  const float t = upper ? 2.0f * (u2 - 0.5f) : 2.0f * (0.5f - u2);
  const float z = (t * (t * (t * 0.27385f - 0.73369f) + 0.46341f)) /
                  (t * (t * (t * 0.093073f + 0.309420f) - 1.000000f) + 0.597999f);
  const float sy = (upper ? 1.0f : -1.0f) * z * safe_sqrtf(1.0f + sx * sx);
  return make_float2(sx, sy);
}

/* microfacet normal visible from `v` (in the lobe's frame) for roughness (ax, ay) */
CY_DEV f3 ggx_visible_normal(f3 v, float ax, float ay, float u1, float u2, float *G1)
{
  const f3 s = normalize(mk3(ax * v.x, ay * v.y, v.z)); /* stretched view direction */
  float cos_v = 1.0f, sin_v = 0.0f, cphi = 1.0f, sphi = 0.0f;
  if (s.z < 0.99999f) {
    cos_v = s.z;
    sin_v = safe_sqrtf(1.0f - cos_v * cos_v);
    const float inv = 1.0f / sin_v;
    cphi = s.x * inv;
    sphi = s.y * inv;
  }
  const float2 unit = ggx_unit_slopes(cos_v, sin_v, u1, u2, G1);
  /* rotate back about the normal, then unstretch */
  const float sx = ax * (cphi * unit.x - sphi * unit.y);
  const float sy = ay * (sphi * unit.x + cphi * unit.y);
  return normalize(mk3(-sx, -sy, 1.0f));
}

/* Samples the lobe: direction, value and pdf; returns the scattering label.  pdf stays
 * zero when no direction is produced. */
CY_DEV int ggx_sample(const Lobe &l, f3 Ng, f3 wo, float u1, float u2, f3 *value, f3 *wi,
                      float *pdf)
{
  const bool refracts = lobe_refracts(l);
  const int glossy = (refracts ? CY_LABEL_TRANSMIT : CY_LABEL_REFLECT) | CY_LABEL_GLOSSY;
  const float cos_o = dot(l.N, wo);
  if (!(cos_o > 0.0f))
    return glossy;
  f3 X, Y;
  if (l.ax == l.ay)
    make_orthonormals(l.N, &X, &Y);
  else
    tangent_frame(l.N, l.T, &X, &Y);
  float G1o;
  const f3 lm = ggx_visible_normal(mk3(dot(X, wo), dot(Y, wo), cos_o), l.ax, l.ay, u1, u2, &G1o);
  const f3 m = X * lm.x + Y * lm.y + l.N * lm.z;
  const bool singular = l.ax * l.ay <= 1e-7f;

  if (!refracts) {
    const float cos_mo = dot(m, wo);
    if (!(cos_mo > 0.0f))
      return glossy;
    *wi = 2.0f * cos_mo * m - wo;
    if (!(dot(Ng, *wi) > 0.0f))
      return glossy;
    int label = glossy;
    if (singular) {
      /* "some high number": a mirror has no finite density */
      *pdf = 1e6f;
      *value = mk3(1e6f, 1e6f, 1e6f);
      if (lobe_has_fresnel_tint(l))
        *value *= ggx_reflection_tint(l, *wi, m);
      label = CY_LABEL_REFLECT | CY_LABEL_SINGULAR;
    }
    else {
      const float cos_i = dot(l.N, *wi);
      float D, G1i;
      if (l.ax == l.ay) {
        float a2 = l.ax * l.ay;
        if (lobe_is_clearcoat(l)) {
          D = gtr1_d(lm.z, l.ax);
          a2 = 0.0625f;
          G1o = ggx_g1(a2, cos_o); /* the sampler's masking term was for alpha, not 0.25 */
        }
        else {
          D = ggx_d(a2, lm.z);
        }
        G1i = ggx_g1(a2, cos_i);
      }
      else {
        D = ggx_d_aniso(mk3(dot(X, m), dot(Y, m), dot(l.N, m)), l.ax, l.ay);
        G1i = ggx_g1(ggx_directional_a2(*wi, X, Y, l.ax, l.ay), cos_i);
      }
      const float common = (G1o * D) * 0.25f / cos_o;
      *pdf = common;
      *value = G1i * common * ggx_reflection_tint(l, *wi, m);
    }
    if (lobe_is_clearcoat(l))
      *value *= 0.25f * l.aux;
    return label;
  }

  const DielectricSplit split = dielectric_split(l.ior, m, wo);
  if (split.inside || split.reflectance == 1.0f)
    return glossy;
  *wi = split.refracted;
  if (singular || fabsf(l.ior - 1.0f) < 1e-4f) {
    *pdf = 1e6f;
    *value = mk3(1e6f, 1e6f, 1e6f);
    return CY_LABEL_TRANSMIT | CY_LABEL_SINGULAR;
  }
  const float a2 = l.ax * l.ay;
  const float D = ggx_d(a2, lm.z);
  const float G1i = ggx_g1(a2, dot(l.N, *wi));
  const float cos_hi = dot(m, *wi), cos_ho = dot(m, wo);
  const float h = l.ior * cos_hi + cos_ho;
  const float common = (G1o * D) * (l.ior * l.ior) / (cos_o * (h * h));
  const float v = G1i * fabsf(cos_hi * cos_ho) * common;
  *pdf = cos_ho * fabsf(cos_hi) * common;
  *value = mk3(v, v, v);
  return glossy;
}

#endif /* B200_MICROFACET_CUH */

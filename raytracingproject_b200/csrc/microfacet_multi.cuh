/* microfacet_multi.cuh - multiple-scattering GGX lobes: the Principled BSDF's default
 * "Multiscatter GGX" distribution and the Glossy / Glass nodes' multiscatter option.
 *
 * The model is Heitz, Hanika, d'Eon, Dachsbacher 2016, "Multiple-Scattering Microfacet
 * BSDFs with the Smith Model": a random walk between the microfacets of a Smith surface -
 * sample the height at which the walker meets the surface next (or let it escape), sample
 * a visible microfacet normal there, scatter off it (mirror reflection, or a dielectric
 * interface that reflects or refracts), repeat.  Sampling follows one walk and returns the
 * direction it leaves in; evaluation follows one walk and, at every vertex, adds what
 * scatters from there into the requested direction ("next event estimation"), so a value
 * is a one-sample estimate whose mean is the BSDF.
 *
 * Reference semantics: kernel/closure/bsdf_microfacet_multi.h (distribution, phase
 * functions, height sampling, the pdf approximations, closure entry points) and
 * bsdf_microfacet_multi_impl.h (the two walks).  The walk draws its numbers from the
 * shading point's LCG (ShaderData::lcg_state, seeded in shader_eval_surface when a lobe
 * sets SD_BSDF_NEEDS_LCG, kernel_shader.h:1109-1111) in the reference's order - height,
 * normal v, normal u, [interface choice] per vertex - so a path that starts from the same
 * seed follows the same walk.  Here the glossy and the glass walk are ONE template over
 * the interface type and the walker is a struct (the reference instantiates a header
 * twice through macros and threads ten loose variables through it).
 * Included by bsdf.cuh; host-compilable. */
#ifndef B200_MICROFACET_MULTI_CUH
#define B200_MICROFACET_MULTI_CUH

#define MF_MAX_ORDER 10 /* scattering events followed before the walk is cut */

CY_DEV float lcg_next_float(uint32_t *state)
{
  *state = 1103515245u * (*state) + 12345u; /* mod 2^32 */
  return (float)(*state) * (1.0f / (float)0xFFFFFFFF);
}
CY_DEV uint32_t lcg_seed(uint32_t seed)
{
  return 1103515245u * seed + 12345u;
}

/* ----------------------------------------- GGX on the local frame (z = normal) */

CY_DEV float mf_d(f3 m, float ax, float ay)
{
  if (ax == ay) {
    const float c2 = m.z * m.z;
    const float a2 = ax * ax;
    const float t = (1.0f - c2) + a2 * c2;
    return a2 / fmaxf(CY_M_PI_F * t * t, 1e-7f);
  }
  const float sx = -m.x / ax, sy = -m.y / ay;
  const float t = m.z * m.z + sx * sx + sy * sy;
  return 1.0f / fmaxf(CY_M_PI_F * t * t * ax * ay, 1e-7f);
}

/* Smith Lambda of direction w */
CY_DEV float mf_lambda(f3 w, float ax, float ay)
{
  if (w.z > 0.9999f)
    return 0.0f;
  if (w.z < -0.9999f)
    return -0.9999f;
  const float inv_z2 = 1.0f / fmaxf(w.z * w.z, 1e-7f);
  const float px = w.x * ax, py = w.y * ay;
  float v = sqrtf(1.0f + (px * px + py * py) * inv_z2);
  if (w.z <= 0.0f)
    v = -v;
  return 0.5f * (v - 1.0f);
}

/* uniform height distribution on [-1, 1]: cdf and inverse */
CY_DEV float mf_height_cdf(float h)
{
  return saturate(0.5f * (h + 1.0f));
}
CY_DEV float mf_height_from_cdf(float c)
{
  return 2.0f * saturate(c) - 1.0f;
}

/* probability that a ray along w from height-cdf C1 escapes the surface */
CY_DEV float mf_escape(f3 w, float C1, float lambda)
{
  if (w.z > 0.9999f)
    return 1.0f;
  if (w.z < 1e-5f)
    return 0.0f;
  return powf(C1, lambda);
}

/* slope of the unit-roughness surface visible under cosine c (the multi-scatter flavour
 * of the visible-normal sampler: guarded against grazing and degenerate inputs) */
CY_DEV float2 mf_unit_slopes(float c, float u1, float u2)
{
  if (c > 0.9999f || fabsf(c) < 1e-6f) {
    const float r = sqrtf(u1 / fmaxf(1.0f - u1, 1e-7f));
    const float phi = CY_M_2PI_F * u2;
    return make_float2(r * cosf(phi), r * sinf(phi));
  }
  const float s = safe_sqrtf(1.0f - c * c);
  const float tan_v = s / c;
  const float proj = 0.5f * (c + 1.0f);
  if (proj < 0.0001f)
    return make_float2(0.0f, 0.0f);
  const float A = 2.0f * u1 * proj / c - 1.0f;
  float k = A * A - 1.0f;
  if (fabsf(k) < 1e-7f)
    return make_float2(0.0f, 0.0f);
  k = 1.0f / k;
  const float disc = safe_sqrtf(tan_v * tan_v * k * k - (A * A - tan_v * tan_v) * k);
  const float hi = tan_v * k + disc;
  const float sx = (A < 0.0f || hi > 1.0f / tan_v) ? (tan_v * k - disc) : hi;
  const bool upper = u2 >= 0.5f;
  const float t = upper ? 2.0f * (u2 - 0.5f) : 2.0f * (0.5f - u2);
  const float z = (t * (t * (t * 0.27385f - 0.73369f) + 0.46341f)) /
                  (t * (t * (t * 0.093073f + 0.309420f) - 1.0f) + 0.597999f);
  const float sy = z * sqrtf(1.0f + sx * sx);
  return make_float2(sx, upper ? sy : -sy);
}

CY_DEV f3 mf_visible_normal(f3 v, float ax, float ay, float u1, float u2)
{
  const f3 st = normalize(mk3(ax * v.x, ay * v.y, v.z));
  const float2 unit = mf_unit_slopes(st.z, u1, u2);
  const f3 az = safe_normalize(mk3(st.x, st.y, 0.0f)); /* (cos phi, sin phi, 0) */
  const float sx = ax * (az.x * unit.x - az.y * unit.y);
  const float sy = ay * (az.y * unit.x + az.x * unit.y);
  return normalize(mk3(-sx, -sy, 1.0f));
}

/* ------------------------------------------------------ phase functions */

/* mirror microfacets: density of scattering from direction w (pointing INTO the surface
 * side the walker moves along) to wo */
CY_DEV float mf_phase_mirror(f3 w, float lambda, f3 wo, float ax, float ay)
{
  if (w.z > 0.9999f)
    return 0.0f;
  const f3 h = normalize(wo - w);
  if (h.z < 0.0f)
    return 0.0f;
  const float area = (w.z < -0.9999f) ? 1.0f : lambda * w.z;
  const float c = dot(-w, h);
  if (c < 0.0f)
    return 0.0f;
  return fmaxf(0.0f, c) * 0.25f / fmaxf(area * c, 1e-7f) * mf_d(h, ax, ay);
}

/* dielectric microfacets; `wo_outside`: wo leaves on the side w came from */
CY_DEV float mf_phase_dielectric(f3 w, float lambda, f3 wo, bool wo_outside, float alpha,
                                 float eta)
{
  if (w.z > 0.9999f)
    return 0.0f;
  const float area = (w.z < -0.9999f) ? 1.0f : lambda * w.z;
  if (wo_outside) {
    const f3 h = normalize(wo - w);
    if (h.z < 0.0f)
      return 0.0f;
    const float c = dot(-w, h);
    return fresnel_dielectric_cos(c, eta) * fmaxf(0.0f, c) * mf_d(h, alpha, alpha) * 0.25f /
           (area * c);
  }
  f3 h = normalize(wo * eta - w);
  if (h.z < 0.0f)
    h = -h;
  const float c = dot(-w, h), co = dot(wo, h);
  if (c < 0.0f)
    return 0.0f;
  const float denom = c + eta * co;
  return (1.0f - fresnel_dielectric_cos(c, eta)) * fmaxf(0.0f, c) * fmaxf(0.0f, -co) *
         mf_d(h, alpha, alpha) / (area * denom * denom);
}

/* scatter off microfacet m coming along wi (pointing away from it): reflect, or - for a
 * dielectric - refract with probability 1 - Fresnel; *same_side tells which */
CY_DEV f3 mf_scatter_dielectric(f3 wi, float eta, f3 m, float u, bool *same_side)
{
  const float c = dot(wi, m);
  if (u < fresnel_dielectric_cos(c, eta)) {
    *same_side = true;
    return -wi + 2.0f * m * c;
  }
  *same_side = false;
  const float inv_eta = 1.0f / eta;
  const float ct = -safe_sqrtf(1.0f - (1.0f - c * c) * inv_eta * inv_eta);
  return normalize(m * (c * inv_eta + ct) - wi * inv_eta);
}

/* -------------------------------------------------------------- the walker */

struct MicroWalker {
  f3 w;         /* travelling direction, local frame */
  float h;      /* height */
  float C1;     /* height cdf */
  float G1;     /* escape probability along w from here */
  float lambda; /* Smith Lambda of w */
  bool outside; /* which side of a dielectric interface the walker is on */
};

/* advance to the next intersection with the surface; false = the walker escaped */
CY_DEV bool mf_walk_to_surface(MicroWalker &k, float u)
{
  if (k.w.z > 0.9999f)
    return false;
  if (k.w.z < -0.9999f) {
    k.C1 *= u;
    k.h = mf_height_from_cdf(k.C1);
    k.G1 = mf_escape(k.w, k.C1, k.lambda);
  }
  else if (fabsf(k.w.z) >= 0.0001f) {
    if (u > 1.0f - k.G1)
      return false;
    if (k.lambda >= 0.0f)
      k.C1 = 1.0f;
    else
      k.C1 *= powf(1.0f - u, -1.0f / k.lambda);
    k.h = mf_height_from_cdf(k.C1);
    k.G1 = mf_escape(k.w, k.C1, k.lambda);
  }
  return true;
}

/* Stochastic value of the lobe for (wi, wo), both in the local frame.  GLASS = dielectric
 * interface (reflection for wo_outside, refraction otherwise), else mirror facets. */
template<bool GLASS>
CY_DEV f3 mf_walk_eval(f3 wi, f3 wo, bool wo_outside, f3 color, float ax, float ay,
                       uint32_t *lcg, float eta, bool tinted, f3 cspec0)
{
  /* start from the shallower direction: less variance, and the BSDF is reciprocal */
  bool swapped = false;
  if (GLASS && wi.z * wo.z < 0.0f) {
    if (-wo.z < wi.z) {
      swapped = true;
      const f3 t = -wo;
      wo = -wi;
      wi = t;
    }
  }
  else if (wo.z < wi.z) {
    swapped = true;
    const f3 t = wo;
    wo = wi;
    wi = t;
  }
  if (wi.z < 1e-5f || (wo.z < 1e-5f && wo_outside) || (wo.z > -1e-5f && !wo_outside))
    return zero3();

  MicroWalker k;
  k.w = -wi;
  k.h = 1.0f;
  k.C1 = 1.0f;
  k.G1 = 0.0f;
  k.outside = true;
  k.lambda = mf_lambda(k.w, ax, ay);
  const f3 wo_up = wo_outside ? wo : -wo;
  const float lambda_o = mf_lambda(wo_up, ax, ay);

  /* single scattering in closed form */
  const f3 half = normalize(wi + wo);
  f3 value;
  if (GLASS) {
    float v = mf_phase_dielectric(k.w, k.lambda, wo, wo_outside, ax, eta);
    if (wo_outside)
      v *= -k.lambda / (lambda_o - k.lambda);
    else
      v *= -k.lambda *
           expf(lgammaf(-k.lambda) + lgammaf(lambda_o + 1.0f) - lgammaf(-k.lambda + lambda_o + 1.0f));
    value = mk3(v, v, v);
  }
  else {
    const float G2 = 1.0f / (1.0f - (k.lambda + 1.0f) + lambda_o);
    const float v = G2 * 0.25f / wi.z * mf_d(half, ax, ay);
    value = mk3(v, v, v);
  }
  f3 throughput = one3();
  const float F0 = fresnel_dielectric_cos(1.0f, eta);
  if (tinted) {
    throughput = fresnel_tint(wi, half, eta, F0, cspec0);
    value *= throughput;
  }

  for (int order = 0; order < MF_MAX_ORDER; order++) {
    if (!mf_walk_to_surface(k, lcg_next_float(lcg)))
      break;
    const float u2 = lcg_next_float(lcg);
    const float u1 = lcg_next_float(lcg);
    const f3 m = mf_visible_normal(-k.w, ax, ay, u1, u2);

    /* what scatters from this vertex towards wo */
    if (order > 0 || (GLASS && tinted)) {
      f3 phase;
      if (GLASS) {
        const float p = k.outside ?
                            mf_phase_dielectric(k.w, k.lambda, wo, wo_outside, ax, eta) :
                            mf_phase_dielectric(k.w, k.lambda, -wo, !wo_outside, ax, 1.0f / eta);
        phase = mk3(p, p, p);
      }
      else {
        const float p = mf_phase_mirror(k.w, k.lambda, wo, ax, ay);
        phase = mk3(p, p, p) * throughput;
      }
      const f3 add = throughput * phase *
                     mf_escape(wo_up, mf_height_cdf((k.outside == wo_outside) ? k.h : -k.h),
                               lambda_o);
      if (order == 0)
        value = add; /* tinted glass: the first vertex replaces the closed form */
      else
        value += add;
    }
    if (order + 1 < MF_MAX_ORDER) {
      if (GLASS) {
        bool same_side;
        const f3 from = -k.w;
        k.w = mf_scatter_dielectric(from, k.outside ? eta : 1.0f / eta, m, lcg_next_float(lcg),
                                    &same_side);
        if (!same_side) {
          k.outside = !k.outside;
          k.w = -k.w;
          k.h = -k.h;
        }
        if (tinted && !same_side)
          throughput *= color;
        else if (tinted && order > 0)
          throughput *= fresnel_tint(from, m, eta, F0, cspec0);
      }
      else {
        if (tinted && order > 0)
          throughput *= fresnel_tint(-k.w, m, eta, F0, cspec0);
        const f3 from = -k.w;
        k.w = -from + 2.0f * m * dot(from, m);
      }
      k.lambda = mf_lambda(k.w, ax, ay);
      if (!tinted)
        throughput *= color;
      k.C1 = mf_height_cdf(k.h);
      k.G1 = mf_escape(k.w, k.C1, k.lambda);
    }
  }
  if (swapped)
    value *= fabsf(wi.z / wo.z);
  return value;
}

/* Follows one walk from wi; returns the throughput and the direction it left in. */
template<bool GLASS>
CY_DEV f3 mf_walk_sample(f3 wi, f3 *wo, f3 color, float ax, float ay, uint32_t *lcg, float eta,
                         bool tinted, f3 cspec0)
{
  MicroWalker k;
  k.w = -wi;
  k.lambda = mf_lambda(k.w, ax, ay);
  k.h = 1.0f;
  k.C1 = 1.0f;
  k.G1 = 0.0f;
  k.outside = true;
  f3 throughput = one3();
  const float F0 = fresnel_dielectric_cos(1.0f, eta);
  if (tinted)
    throughput = fresnel_tint(wi, normalize(wi + k.w), eta, F0, cspec0);

  for (int order = 0; order < MF_MAX_ORDER; order++) {
    if (!mf_walk_to_surface(k, lcg_next_float(lcg))) {
      *wo = k.outside ? k.w : -k.w;
      return throughput;
    }
    const float u2 = lcg_next_float(lcg);
    const float u1 = lcg_next_float(lcg);
    const f3 m = mf_visible_normal(-k.w, ax, ay, u1, u2);
    /* the first bounce's albedo is already in the lobe's weight */
    if (!tinted && order > 0)
      throughput *= color;
    const f3 from = -k.w;
    if (GLASS) {
      bool same_side;
      k.w = mf_scatter_dielectric(from, k.outside ? eta : 1.0f / eta, m, lcg_next_float(lcg),
                                  &same_side);
      if (!same_side) {
        k.h = -k.h;
        k.w = -k.w;
        k.outside = !k.outside;
      }
      if (tinted) {
        if (!same_side) {
          throughput *= color;
        }
        else {
          const f3 t = fresnel_tint(from, m, eta, F0, cspec0);
          throughput = (order == 0) ? t : throughput * t;
        }
      }
    }
    else {
      if (tinted) {
        const f3 t = fresnel_tint(from, m, eta, F0, cspec0);
        throughput = (order == 0) ? t : throughput * t;
      }
      k.w = -from + 2.0f * m * dot(from, m);
    }
    k.lambda = mf_lambda(k.w, ax, ay);
    k.G1 = mf_escape(k.w, k.C1, k.lambda);
  }
  *wo = mk3(0.0f, 0.0f, 1.0f);
  return zero3();
}

/* ------------------------------------------- pdf: single scattering + a diffuse rest */

/* fitted albedo of single-scattering GGX (what the walk's first bounce returns) */
CY_DEV float mf_single_scatter_albedo(float r)
{
  float albedo = 0.806495f * expf(-1.98712f * r * r) + 0.199531f;
  albedo -= ((((((1.76741f * r - 8.43891f) * r + 15.784f) * r - 14.398f) * r + 6.45221f) * r -
              1.19722f) * r + 0.027803f) * r + 0.00568739f;
  return saturate(albedo);
}
CY_DEV float mf_transmission_albedo(float a, float ior)
{
  if (ior < 1.0f)
    ior = 1.0f / ior;
  a = saturate(a);
  ior = clampf(ior, 1.0f, 3.0f);
  const float I_1 = 0.0476898f * expf(-0.978352f * (ior - 0.65657f) * (ior - 0.65657f)) -
                    0.033756f * ior + 0.993261f;
  const float R_1 = (((0.116991f * a - 0.270369f) * a + 0.0501366f) * a - 0.00411511f) * a +
                    1.00008f;
  const float I_2 = (((-2.08704f * ior + 26.3298f) * ior - 127.906f) * ior + 292.958f) * ior -
                    287.946f + 199.803f / (ior * ior) - 101.668f / (ior * ior * ior);
  const float R_2 = ((((5.3725f * a - 24.9307f) * a + 22.7437f) * a - 3.40751f) * a +
                     0.0986325f) * a + 0.00493504f;
  return saturate(1.0f + I_2 * R_2 * 0.0019127f - (1.0f - I_1) * (1.0f - R_1) * 9.3205f);
}

CY_DEV float mf_pdf_glossy(f3 wi, f3 wo, float ax, float ay)
{
  const float D = mf_d(normalize(wi + wo), ax, ay);
  const float lambda = mf_lambda(wi, ax, ay);
  const float single = 0.25f * D / fmaxf((1.0f + lambda) * wi.z, 1e-7f);
  const float rest = wo.z * CY_M_1_PI_F;
  const float albedo = mf_single_scatter_albedo((ax == ay) ? ax : sqrtf(ax * ay));
  return albedo * single + (1.0f - albedo) * rest;
}
CY_DEV float mf_pdf_glass(f3 wi, f3 wo, float alpha, float eta)
{
  const bool reflective = (wi.z * wo.z > 0.0f);
  float hl;
  f3 h = normalize_len(wi + (reflective ? wo : (wo * eta)), &hl);
  if (h.z < 0.0f)
    h = -h;
  const f3 up = (wi.z < 0.0f) ? -wi : wi;
  const float lambda = mf_lambda(up, alpha, alpha);
  const float D = mf_d(h, alpha, alpha);
  const float fresnel = fresnel_dielectric_cos(dot(up, h), eta);
  const float rest = fabsf(wo.z * CY_M_1_PI_F);
  if (reflective) {
    const float single = 0.25f * D / fmaxf((1.0f + lambda) * up.z, 1e-7f);
    const float albedo = mf_single_scatter_albedo(alpha);
    return fresnel * (albedo * single + (1.0f - albedo) * rest);
  }
  const float single = fabsf(dot(up, h) * dot(wo, h) * D * eta * eta /
                             fmaxf((1.0f + lambda) * up.z * hl * hl, 1e-7f));
  const float albedo = mf_transmission_albedo(alpha, eta);
  return (1.0f - fresnel) * (albedo * single + (1.0f - albedo) * rest);
}

/* ------------------------------------------------------- lobe entry points */

CY_DEV bool multi_lobe_is_glass(const Lobe &l)
{
  return lobe_id(l.kind) == CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_GLASS_ID ||
         lobe_id(l.kind) == CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_GLASS_FRESNEL_ID;
}
CY_DEV bool multi_lobe_is_tinted(const Lobe &l)
{
  return lobe_id(l.kind) == CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_FRESNEL_ID ||
         lobe_id(l.kind) == CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_GLASS_FRESNEL_ID;
}
CY_DEV void multi_lobe_frame(const Lobe &l, f3 *X, f3 *Y)
{
  if (l.ax != l.ay && !multi_lobe_is_glass(l)) {
    *Y = normalize(cross(l.N, l.T));
    *X = cross(*Y, l.N);
  }
  else {
    make_orthonormals(l.N, X, Y);
  }
}

/* out of line: reached only by shaders with a multi-scatter lobe; inlined into the
 * eval / sample switches the walks would set the register budget of every shader */
__device__ __noinline__ f3 multi_ggx_eval(const Lobe &l, f3 I, f3 omega_in, bool same_side,
                                          float *pdf, uint32_t *lcg)
{
  const bool glass = multi_lobe_is_glass(l);
  if (!glass && !same_side) {
    *pdf = 0.0f;
    return zero3();
  }
  if (l.ax * l.ay < 1e-7f)
    return zero3();
  f3 X, Y;
  multi_lobe_frame(l, &X, &Y);
  const f3 wi = mk3(dot(I, X), dot(I, Y), dot(I, l.N));
  const f3 wo = mk3(dot(omega_in, X), dot(omega_in, Y), dot(omega_in, l.N));
  if (glass) {
    *pdf = mf_pdf_glass(wi, wo, l.ax, l.ior);
    if (same_side)
      return mf_walk_eval<true>(wi, wo, true, l.color, l.ax, l.ay, lcg, l.ior,
                                multi_lobe_is_tinted(l), l.cspec0);
    return mf_walk_eval<true>(wi, wo, false, l.color, l.ax, l.ay, lcg, l.ior, false, l.color);
  }
  *pdf = mf_pdf_glossy(wi, wo, l.ax, l.ay);
  return mf_walk_eval<false>(wi, wo, true, l.color, l.ax, l.ay, lcg, l.ior,
                             multi_lobe_is_tinted(l), l.cspec0);
}

__device__ __noinline__ int multi_ggx_sample(const Lobe &l, f3 I, float randu, float randv,
                                             f3 *value, f3 *omega_in, float *pdf, uint32_t *lcg)
{
  (void)randv;
  const bool glass = multi_lobe_is_glass(l);
  if (l.ax * l.ay < 1e-7f) {
    /* the smooth limit: a mirror, or a sharp dielectric interface */
    *pdf = 1e6f;
    *value = mk3(1e6f, 1e6f, 1e6f);
    if (!glass) {
      *omega_in = 2.0f * dot(l.N, I) * l.N - I;
      return CY_LABEL_REFLECT | CY_LABEL_SINGULAR;
    }
    const DielectricSplit split = dielectric_split(l.ior, l.N, I);
    if (randu < split.reflectance) {
      *omega_in = split.reflected;
      return CY_LABEL_REFLECT | CY_LABEL_SINGULAR;
    }
    *omega_in = split.refracted;
    return CY_LABEL_TRANSMIT | CY_LABEL_SINGULAR;
  }
  f3 X, Y;
  multi_lobe_frame(l, &X, &Y);
  const f3 wi = mk3(dot(I, X), dot(I, Y), dot(I, l.N));
  f3 wo;
  if (glass) {
    *value = mf_walk_sample<true>(wi, &wo, l.color, l.ax, l.ay, lcg, l.ior,
                                  multi_lobe_is_tinted(l), l.cspec0);
    *pdf = mf_pdf_glass(wi, wo, l.ax, l.ior);
  }
  else {
    *value = mf_walk_sample<false>(wi, &wo, l.color, l.ax, l.ay, lcg, l.ior,
                                   multi_lobe_is_tinted(l), l.cspec0);
    *pdf = mf_pdf_glossy(wi, wo, l.ax, l.ay);
  }
  *value *= *pdf;
  *omega_in = X * wo.x + Y * wo.y + l.N * wo.z;
  if (glass && !(wo.z * wi.z > 0.0f))
    return CY_LABEL_TRANSMIT | CY_LABEL_GLOSSY;
  return CY_LABEL_REFLECT | CY_LABEL_GLOSSY;
}

#endif /* B200_MICROFACET_MULTI_CUH */

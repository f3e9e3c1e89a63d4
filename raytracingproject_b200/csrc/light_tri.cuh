/* light_tri.cuh - emissive triangles as lights ("mesh lights").
 *
 * What it restates (reference = blender/intern/cycles/kernel):
 *   triangle_world_space_vertices   kernel_light.h:302-329 (no motion blur)
 *   triangle_light_pdf_area / _pdf  kernel_light.h:331-412
 *   triangle_light_sample           kernel_light.h:414-581
 *   fast_acosf / fast_sinf / fast_sincosf  util/util_math_fast.h:95-205, 278-293
 *     (SLEEF-style range reduction + polynomials; the reference uses these, not
 *      libm, so both sides evaluate the same float operations and agree bit for bit)
 *
 * Two sampling strategies, chosen by comparing the distance to the triangle's plane
 * with its longest edge: Arvo's stratified sampling of the spherical triangle when the
 * triangle is large as seen from P, uniform area sampling (Heitz's low-distortion
 * square -> triangle map) otherwise.  pdf_triangles (KernelIntegrator) is the
 * probability density per unit emissive area.
 *
 * The functions are __noinline__: they are only reached in scenes that have emissive
 * meshes and must not cost registers in k_shade_surface otherwise.
 */
#ifndef B200_LIGHT_TRI_CUH
#define B200_LIGHT_TRI_CUH

/* round to nearest by adding +-0.5 and truncating (util_math_fast.h:83-93, non-SSE4) */
CY_DEV int fast_rint(float x)
{
  return (int)(x + copysignf(0.5f, x));
}

/* shared argument reduction: x - q*pi in four steps (the constants are pi/4 split into
 * exactly representable pieces, times 4) */
CY_DEV float fast_reduce_pi(float x, int *q_out)
{
  const int q = fast_rint(x * CY_M_1_PI_F);
  const float qf = (float)q;
  x = qf * (-0.78515625f * 4) + x;
  x = qf * (-0.00024187564849853515625f * 4) + x;
  x = qf * (-3.7747668102383613586e-08f * 4) + x;
  x = qf * (-1.2816720341285448015e-12f * 4) + x;
  x = CY_M_PI_2_F - (CY_M_PI_2_F - x); /* crush denormals */
  *q_out = q;
  return x;
}

CY_DEV float fast_sin_poly(float x, float s)
{
  float u = 2.6083159809786593541503e-06f;
  u = u * s + -0.0001981069071916863322258f;
  u = u * s + 0.00833307858556509017944336f;
  u = u * s + -0.166666597127914428710938f;
  u = s * (u * x) + x;
  return u;
}

CY_DEV float fast_cos_poly(float s)
{
  float u = -2.71811842367242206819355e-07f;
  u = u * s + 2.47990446951007470488548e-05f;
  u = u * s + -0.00138888787478208541870117f;
  u = u * s + 0.0416666641831398010253906f;
  u = u * s + -0.5f;
  u = u * s + 1.0f;
  return u;
}

CY_DEV float fast_sinf(float x)
{
  int q;
  x = fast_reduce_pi(x, &q);
  const float s = x * x;
  if (q & 1)
    x = -x;
  float u = fast_sin_poly(x, s);
  if (fabsf(u) > 1.0f)
    u = 0.0f;
  return u;
}

CY_DEV void fast_sincosf(float x, float *sine, float *cosine)
{
  int q;
  x = fast_reduce_pi(x, &q);
  const float s = x * x;
  if (q & 1)
    x = -x;
  float su = fast_sin_poly(x, s);
  float cu = fast_cos_poly(s);
  if (q & 1)
    cu = -cu;
  if (fabsf(su) > 1.0f)
    su = 0.0f;
  if (fabsf(cu) > 1.0f)
    cu = 0.0f;
  *sine = su;
  *cosine = cu;
}

CY_DEV f3 safe_normalize_len(f3 a, float *t)
{
  *t = len(a);
  return (*t != 0.0f) ? a / (*t) : a;
}

CY_DEV float triangle_area(f3 v1, f3 v2, f3 v3)
{
  return len(cross(v3 - v2, v1 - v2)) * 0.5f;
}

/* The three vertices of mesh triangle `prim` in world space.  Returns true when an
 * object transform had to be applied (instanced mesh): the caller then rescales
 * pdf_triangles, which the host computed from these same world-space triangles. */
CY_DEV bool triangle_world_space_vertices(int object, int prim, f3 V[3])
{
  const uint4 tri_vindex = __ldg(&g_scene.tri_vindex[prim]);
  V[0] = mk3(__ldg(&g_scene.prim_tri_verts[tri_vindex.w + 0]));
  V[1] = mk3(__ldg(&g_scene.prim_tri_verts[tri_vindex.w + 1]));
  V[2] = mk3(__ldg(&g_scene.prim_tri_verts[tri_vindex.w + 2]));
  const uint32_t object_flag = __ldg(&g_scene.object_flag[object]);
  if (!(object_flag & CY_SD_OBJECT_TRANSFORM_APPLIED)) {
    const tfm34 tfm = object_tfm(object);
    V[0] = transform_point(tfm, V[0]);
    V[1] = transform_point(tfm, V[1]);
    V[2] = transform_point(tfm, V[2]);
    return true;
  }
  return false;
}

CY_DEV float triangle_light_pdf_area(f3 Ng, f3 I, float t)
{
  const float pdf = kd_float(KD_INT_PDF_TRIANGLES);
  const float cos_pi = fabsf(dot(Ng, I));
  if (cos_pi == 0.0f)
    return 0.0f;
  return t * t * pdf / cos_pi;
}

/* pdf of having sampled the point sd.P on emissive triangle (sd.object, sd.prim) from the
 * point sd.P + sd.I * t - the MIS partner of a BSDF-sampled ray that hit the triangle */
__device__ __noinline__ float triangle_light_pdf(
    int object, int prim, f3 sdP, f3 sdNg, f3 sdI, float t)
{
  f3 V[3];
  const bool has_motion = triangle_world_space_vertices(object, prim, V);

  const f3 e0 = V[1] - V[0];
  const f3 e1 = V[2] - V[0];
  const f3 e2 = V[2] - V[1];
  const float longest_edge_squared = fmaxf(len_squared(e0),
                                           fmaxf(len_squared(e1), len_squared(e2)));
  const f3 N = cross(e0, e1);
  const float distance_to_plane = fabsf(dot(N, sdI * t)) / dot(N, N);

  if (longest_edge_squared > distance_to_plane * distance_to_plane) {
    /* solid angle of the spherical triangle seen from the shading point */
    const f3 Px = sdP + sdI * t;
    const f3 v0_p = V[0] - Px;
    const f3 v1_p = V[1] - Px;
    const f3 v2_p = V[2] - Px;

    const f3 u01 = safe_normalize(cross(v0_p, v1_p));
    const f3 u02 = safe_normalize(cross(v0_p, v2_p));
    const f3 u12 = safe_normalize(cross(v1_p, v2_p));

    const float alpha = fast_acosf(dot(u02, u01));
    const float beta = fast_acosf(-dot(u01, u12));
    const float gamma = fast_acosf(dot(u02, u12));
    const float solid_angle = alpha + beta + gamma - CY_M_PI_F;

    if (solid_angle == 0.0f)
      return 0.0f;
    /* without motion blur the "centre frame" triangle is the same triangle */
    const float area = has_motion ? triangle_area(V[0], V[1], V[2]) : 0.5f * len(N);
    const float pdf = area * kd_float(KD_INT_PDF_TRIANGLES);
    return pdf / solid_angle;
  }
  else {
    float pdf = triangle_light_pdf_area(sdNg, sdI, t);
    if (has_motion) {
      const float area = 0.5f * len(N);
      if (area == 0.0f)
        return 0.0f;
      const float area_pre = triangle_area(V[0], V[1], V[2]);
      pdf = pdf * area_pre / area;
    }
    return pdf;
  }
}

/* Sample a point on emissive triangle (object, prim) as seen from P. */
__device__ __noinline__ void triangle_light_sample(
    int prim, int object, float randu, float randv, LightSampleG *ls, f3 P)
{
  f3 V[3];
  const bool has_motion = triangle_world_space_vertices(object, prim, V);

  const f3 e0 = V[1] - V[0];
  const f3 e1 = V[2] - V[0];
  const f3 e2 = V[2] - V[1];
  const float longest_edge_squared = fmaxf(len_squared(e0),
                                           fmaxf(len_squared(e1), len_squared(e2)));
  const f3 N0 = cross(e0, e1);
  float Nl = 0.0f;
  ls->Ng = safe_normalize_len(N0, &Nl);
  float area = 0.5f * Nl;

  const uint32_t object_flag = __ldg(&g_scene.object_flag[object]);
  if (object_flag & CY_SD_OBJECT_NEGATIVE_SCALE_APPLIED)
    ls->Ng = -ls->Ng;
  ls->eval_fac = 1.0f;
  ls->shader = (int)__ldg(&g_scene.tri_shader[prim]);
  ls->object = object;
  ls->prim = prim;
  ls->lamp = CY_LAMP_NONE;
  ls->shader |= CY_SHADER_USE_MIS;
  ls->type = CY_LIGHT_TRIANGLE;

  const float distance_to_plane = fabsf(dot(N0, V[0] - P) / dot(N0, N0));

  if (longest_edge_squared > distance_to_plane * distance_to_plane) {
    /* Arvo 1995, "Stratified Sampling of Spherical Triangles": project the triangle onto
     * the unit sphere around P, pick the sub-triangle A B C' whose area is randu times the
     * whole, then a point along the arc B C' */
    const f3 v0_p = V[0] - P;
    const f3 v1_p = V[1] - P;
    const f3 v2_p = V[2] - P;

    const f3 u01 = safe_normalize(cross(v0_p, v1_p));
    const f3 u02 = safe_normalize(cross(v0_p, v2_p));
    const f3 u12 = safe_normalize(cross(v1_p, v2_p));

    const f3 A = safe_normalize(v0_p);
    const f3 B = safe_normalize(v1_p);
    const f3 C = safe_normalize(v2_p);

    const float cos_alpha = dot(u02, u01);
    const float cos_beta = -dot(u01, u12);
    const float cos_gamma = dot(u02, u12);

    const float alpha = fast_acosf(cos_alpha);
    const float beta = fast_acosf(cos_beta);
    const float gamma = fast_acosf(cos_gamma);
    const float solid_angle = alpha + beta + gamma - CY_M_PI_F;

    const float cos_c = dot(A, B);
    const float sin_alpha = fast_sinf(alpha);
    const float product = sin_alpha * cos_c;

    const float phi = randu * solid_angle - alpha;
    float s, t;
    fast_sincosf(phi, &s, &t);
    const float u = t - cos_alpha;
    const float v = s + product;

    const f3 U = safe_normalize(C - dot(C, A) * A);

    float q = 1.0f;
    const float det = ((v * s + u * t) * sin_alpha);
    if (det != 0.0f)
      q = ((v * t - u * s) * cos_alpha - v) / det;
    const float temp = fmaxf(1.0f - q * q, 0.0f);

    const f3 C_ = safe_normalize(q * A + sqrtf(temp) * U);

    const float z = 1.0f - randv * (1.0f - dot(C_, B));
    ls->D = z * B + safe_sqrtf(1.0f - z * z) * safe_normalize(C_ - dot(C_, B) * B);

    /* back onto the planar triangle */
    if (!ray_triangle_intersect(P, ls->D, FLT_MAX, V[0], V[1], V[2], &ls->u, &ls->v, &ls->t)) {
      ls->pdf = 0.0f;
      return;
    }
    ls->P = P + ls->D * ls->t;

    if (solid_angle == 0.0f) {
      ls->pdf = 0.0f;
      return;
    }
    if (has_motion)
      area = triangle_area(V[0], V[1], V[2]);
    const float pdf = area * kd_float(KD_INT_PDF_TRIANGLES);
    ls->pdf = pdf / solid_angle;
  }
  else {
    /* Heitz, "A Low-Distortion Map Between Triangle and Square" */
    float u = randu;
    float v = randv;
    if (v > u) {
      u *= 0.5f;
      v -= u;
    }
    else {
      v *= 0.5f;
      u -= v;
    }
    const float t = 1.0f - u - v;
    ls->P = u * V[0] + v * V[1] + t * V[2];
    ls->D = normalize_len(ls->P - P, &ls->t);
    ls->pdf = triangle_light_pdf_area(ls->Ng, -ls->D, ls->t);
    if (has_motion && area != 0.0f) {
      const float area_pre = triangle_area(V[0], V[1], V[2]);
      ls->pdf = ls->pdf * area_pre / area;
    }
    ls->u = u;
    ls->v = v;
  }
}

#endif /* B200_LIGHT_TRI_CUH */

/* light_tri.cuh - emissive triangles as lights ("mesh lights"): sampling a point on one
 * for next-event estimation, and the density of that strategy for a BSDF-sampled ray that
 * happened to hit one (the MIS partner).
 *
 * Semantics to match (reference = blender/intern/cycles/kernel): kernel_light.h:302-581
 * (triangle_world_space_vertices, triangle_light_pdf, triangle_light_sample, no motion
 * blur) - the choice between the two strategies by comparing the distance to the
 * triangle's plane with its longest edge, and pdf_triangles (KernelIntegrator) as the
 * density per unit emissive area.  The strategies themselves are published algorithms:
 * Arvo 1995, "Stratified Sampling of Spherical Triangles" when the triangle is large as
 * seen from the shading point, and a uniform point through Heitz's low-distortion
 * square -> triangle map otherwise.  The trigonometry uses the reference's polynomial
 * sin / cos / acos (util/util_math_fast.h:95-205, 278-293) rather than libm: the solid
 * angle is a small difference of three arccosines, and only the same polynomials keep
 * that difference the same number on both sides.
 *
 * Shape: the triangle is loaded once into an `EmissiveTriangle` (world-space corners,
 * edge-derived quantities); both entry points work from it, and the spherical-triangle
 * geometry they share is one helper.
 *
 * The entry points are __noinline__: they are only reached in scenes that have emissive
 * meshes and must not cost registers in k_shade_surface otherwise. */
#ifndef B200_LIGHT_TRI_CUH
#define B200_LIGHT_TRI_CUH

/* ------------------------------------------- polynomial trigonometry */

/* x - round(x / pi) * pi in four exactly representable steps; *parity = round(x / pi) */
CY_DEV float reduce_to_half_period(float x, int *parity)
{
  const float scaled = x * CY_M_1_PI_F;
  const int q = (int)(scaled + copysignf(0.5f, scaled));
  const float qf = (float)q;
  x = qf * (-0.78515625f * 4) + x;
  x = qf * (-0.00024187564849853515625f * 4) + x;
  x = qf * (-3.7747668102383613586e-08f * 4) + x;
  x = qf * (-1.2816720341285448015e-12f * 4) + x;
  *parity = q;
  return CY_M_PI_2_F - (CY_M_PI_2_F - x); /* crush denormals */
}

/* sine and cosine of x by odd / even polynomials on the reduced argument; results that
 * escape [-1, 1] (huge inputs) are flushed to zero */
CY_DEV void poly_sincos(float x, float *sine, float *cosine)
{
  int q;
  x = reduce_to_half_period(x, &q);
  const float s = x * x;
  if (q & 1)
    x = -x;
  float su = 2.6083159809786593541503e-06f;
  su = su * s + -0.0001981069071916863322258f;
  su = su * s + 0.00833307858556509017944336f;
  su = su * s + -0.166666597127914428710938f;
  su = s * (su * x) + x;
  float cu = -2.71811842367242206819355e-07f;
  cu = cu * s + 2.47990446951007470488548e-05f;
  cu = cu * s + -0.00138888787478208541870117f;
  cu = cu * s + 0.0416666641831398010253906f;
  cu = cu * s + -0.5f;
  cu = cu * s + 1.0f;
  if (q & 1)
    cu = -cu;
  *sine = (fabsf(su) > 1.0f) ? 0.0f : su;
  *cosine = (fabsf(cu) > 1.0f) ? 0.0f : cu;
}
CY_DEV float poly_sin(float x)
{
  float s, c;
  poly_sincos(x, &s, &c);
  return s;
}

/* ------------------------------------------------- the emissive triangle */

struct EmissiveTriangle {
  f3 v0, v1, v2;      /* world space */
  f3 plane_normal;    /* cross(e01, e02), unnormalised */
  float longest_edge2;
  bool transformed;   /* an object transform was applied: pdf_triangles was computed by
                       * the host from these same world-space triangles, see callers */
};

CY_DEV EmissiveTriangle emissive_triangle(int object, int prim)
{
  EmissiveTriangle t;
  const uint4 vi = __ldg(&g_scene.tri_vindex[prim]);
  t.v0 = mk3(__ldg(&g_scene.prim_tri_verts[vi.w + 0]));
  t.v1 = mk3(__ldg(&g_scene.prim_tri_verts[vi.w + 1]));
  t.v2 = mk3(__ldg(&g_scene.prim_tri_verts[vi.w + 2]));
  t.transformed = !(__ldg(&g_scene.object_flag[object]) & CY_SD_OBJECT_TRANSFORM_APPLIED);
  if (t.transformed) {
    const tfm34 tfm = object_tfm(object);
    t.v0 = transform_point(tfm, t.v0);
    t.v1 = transform_point(tfm, t.v1);
    t.v2 = transform_point(tfm, t.v2);
  }
  const f3 e01 = t.v1 - t.v0, e02 = t.v2 - t.v0, e12 = t.v2 - t.v1;
  t.longest_edge2 = fmaxf(len_squared(e01), fmaxf(len_squared(e02), len_squared(e12)));
  t.plane_normal = cross(e01, e02);
  return t;
}

CY_DEV float emissive_triangle_area(const EmissiveTriangle &t)
{
  return len(cross(t.v2 - t.v1, t.v0 - t.v1)) * 0.5f;
}

/* The triangle projected on the unit sphere about `apex`: cosines of its three interior
 * angles (between the great-circle planes through the edges) and its solid angle. */
struct SphericalTriangle {
  f3 A, B, C; /* unit directions to the corners (only sampling needs them) */
  float cos_alpha, cos_beta, cos_gamma;
  float alpha, solid_angle;
};
CY_DEV SphericalTriangle spherical_triangle(const EmissiveTriangle &t, f3 apex)
{
  SphericalTriangle s;
  const f3 a = t.v0 - apex, b = t.v1 - apex, c = t.v2 - apex;
  const f3 n_ab = safe_normalize(cross(a, b));
  const f3 n_ac = safe_normalize(cross(a, c));
  const f3 n_bc = safe_normalize(cross(b, c));
  s.A = safe_normalize(a);
  s.B = safe_normalize(b);
  s.C = safe_normalize(c);
  s.cos_alpha = dot(n_ac, n_ab);
  s.cos_beta = -dot(n_ab, n_bc);
  s.cos_gamma = dot(n_ac, n_bc);
  s.alpha = fast_acosf(s.cos_alpha);
  s.solid_angle = s.alpha + fast_acosf(s.cos_beta) + fast_acosf(s.cos_gamma) - CY_M_PI_F;
  return s;
}

/* density of a uniformly chosen point of the light's area, as seen along I over t */
CY_DEV float emissive_area_pdf(f3 Ng, f3 I, float t)
{
  const float cos_pi = fabsf(dot(Ng, I));
  if (cos_pi == 0.0f)
    return 0.0f;
  return t * t * kd_float(KD_INT_PDF_TRIANGLES) / cos_pi;
}

/* an instanced mesh light: the host's pdf_triangles is per unit of the world-space area,
 * the area density above used the same triangle - rescale by pre / post (both equal
 * without motion blur, kept as the reference does it) */
CY_DEV float emissive_rescale_instanced(const EmissiveTriangle &t, float pdf, float area)
{
  if (!t.transformed)
    return pdf;
  if (area == 0.0f)
    return 0.0f;
  return pdf * emissive_triangle_area(t) / area;
}

/* Density with which light sampling would have produced the point P on emissive triangle
 * (object, prim), seen from P + I * t - the MIS partner of a BSDF-sampled ray. */
__device__ __noinline__ float triangle_light_pdf(int object, int prim, f3 P, f3 Ng, f3 I, float t)
{
  const EmissiveTriangle tri = emissive_triangle(object, prim);
  const f3 N = tri.plane_normal;
  const float plane_distance = fabsf(dot(N, I * t)) / dot(N, N);
  if (tri.longest_edge2 > plane_distance * plane_distance) {
    const SphericalTriangle s = spherical_triangle(tri, P + I * t);
    if (s.solid_angle == 0.0f)
      return 0.0f;
    const float area = tri.transformed ? emissive_triangle_area(tri) : 0.5f * len(N);
    return area * kd_float(KD_INT_PDF_TRIANGLES) / s.solid_angle;
  }
  return emissive_rescale_instanced(tri, emissive_area_pdf(Ng, I, t), 0.5f * len(N));
}

/* Sample a point on emissive triangle (object, prim) as seen from P. */
__device__ __noinline__ void triangle_light_sample(int prim, int object, float randu, float randv,
                                                   LightSampleG *ls, f3 P)
{
  const EmissiveTriangle tri = emissive_triangle(object, prim);
  const f3 N = tri.plane_normal;
  const float Nl = len(N);
  ls->Ng = (Nl != 0.0f) ? N / Nl : N;
  if (__ldg(&g_scene.object_flag[object]) & CY_SD_OBJECT_NEGATIVE_SCALE_APPLIED)
    ls->Ng = -ls->Ng;
  ls->eval_fac = 1.0f;
  ls->shader = (int)__ldg(&g_scene.tri_shader[prim]) | CY_SHADER_USE_MIS;
  ls->object = object;
  ls->prim = prim;
  ls->lamp = CY_LAMP_NONE;
  ls->type = CY_LIGHT_TRIANGLE;

  const float plane_distance = fabsf(dot(N, tri.v0 - P) / dot(N, N));
  if (tri.longest_edge2 > plane_distance * plane_distance) {
    /* Arvo: pick the sub-triangle A B C' holding the fraction randu of the solid angle
     * (C' on the arc A C), then a point on the arc B C' by randv */
    const SphericalTriangle s = spherical_triangle(tri, P);
    const float cos_c = dot(s.A, s.B);
    const float sin_alpha = poly_sin(s.alpha);
    float sin_phi, cos_phi;
    poly_sincos(randu * s.solid_angle - s.alpha, &sin_phi, &cos_phi);
    const float u = cos_phi - s.cos_alpha;
    const float v = sin_phi + sin_alpha * cos_c;
    const f3 U = safe_normalize(s.C - dot(s.C, s.A) * s.A); /* C made orthogonal to A */
    float q = 1.0f;
    const float det = (v * sin_phi + u * cos_phi) * sin_alpha;
    if (det != 0.0f)
      q = ((v * cos_phi - u * sin_phi) * s.cos_alpha - v) / det;
    const f3 Cp = safe_normalize(q * s.A + sqrtf(fmaxf(1.0f - q * q, 0.0f)) * U);
    const float z = 1.0f - randv * (1.0f - dot(Cp, s.B));
    ls->D = z * s.B + safe_sqrtf(1.0f - z * z) * safe_normalize(Cp - dot(Cp, s.B) * s.B);

    /* back onto the planar triangle */
    if (!ray_triangle_intersect(P, ls->D, FLT_MAX, tri.v0, tri.v1, tri.v2, &ls->u, &ls->v,
                                &ls->t) ||
        s.solid_angle == 0.0f) {
      ls->pdf = 0.0f;
      return;
    }
    ls->P = P + ls->D * ls->t;
    const float area = tri.transformed ? emissive_triangle_area(tri) : 0.5f * Nl;
    ls->pdf = area * kd_float(KD_INT_PDF_TRIANGLES) / s.solid_angle;
    return;
  }

  /* far away: uniform over the area.  Heitz's map folds the unit square along its
   * diagonal onto the triangle without the sqrt distortion */
  float u = randu, v = randv;
  if (v > u) {
    u *= 0.5f;
    v -= u;
  }
  else {
    v *= 0.5f;
    u -= v;
  }
  ls->P = u * tri.v0 + v * tri.v1 + (1.0f - u - v) * tri.v2;
  ls->D = normalize_len(ls->P - P, &ls->t);
  ls->pdf = emissive_area_pdf(ls->Ng, -ls->D, ls->t);
  if (tri.transformed && 0.5f * Nl != 0.0f)
    ls->pdf = ls->pdf * emissive_triangle_area(tri) / (0.5f * Nl);
  ls->u = u;
  ls->v = v;
}

#endif /* B200_LIGHT_TRI_CUH */

/* shade.cuh - device-side shading for the wavefront stages: random numbers,
 * camera rays, ShaderData setup, the SVM subset, closures (BSDF eval/sample),
 * light sampling and emission.  Each function restates the reference function
 * cited above it with the same operation order (the file is compiled with
 * -fmad=false), so sample sequences, pdfs and weights agree with the CPU oracle
 * to rounding of the transcendental functions.
 *
 * Scope (b200_cycles.cu:check_scope / svm_validate refuse anything else):
 * perspective, orthographic and panoramic cameras (with depth of field, without motion
 * blur or stereo),
 * static triangles and instances,
 * point / spot / area / distant lamps, emissive triangles (light_tri.cuh), background
 * colour, SVM nodes of
 * svm_nodes.cuh / svm_validate and the Diffuse / Oren-Nayar, Translucent, Principled(GGX),
 * Glossy, Glass and Refraction (GGX or sharp) closures,
 * opaque shadows, combined pass only.
 */
#ifndef B200_SHADE_CUH
#define B200_SHADE_CUH

#include "device_scene.cuh"
#include "traverse.cuh"
#include "shader_data.cuh"

/* ------------------------------------------------------------- path state */

/* inlining of the two largest shared pieces of the shading kernels, overridable for A/B
 * builds (tools/variants.sh): -DMULTI_EVAL_ATTR="__device__ __noinline__" ... */
#ifndef MULTI_EVAL_ATTR
#  define MULTI_EVAL_ATTR CY_DEV
#endif
#ifndef LIGHT_SAMPLE_ATTR
#  define LIGHT_SAMPLE_ATTR CY_DEV
#endif

struct PathStateG {
  uint32_t flag;
  uint32_t rng_hash;
  int rng_offset;
  int sample;
  int bounce, diffuse_bounce, glossy_bounce, transmission_bounce, transparent_bounce;
  float min_ray_pdf, ray_pdf, ray_t;
};

CY_DEV PathDepths path_depths(const PathStateG &s)
{
  PathDepths d;
  d.bounce = (short)s.bounce;
  d.diffuse = (short)s.diffuse_bounce;
  d.glossy = (short)s.glossy_bounce;
  d.transparent = (short)s.transparent_bounce;
  d.transmission = (short)s.transmission_bounce;
  return d;
}

/* ---------------------------------------------------------------- hashing */

/* util/util_hash.h:28-93 (Jenkins lookup3 final) */
CY_DEV uint32_t rot32(uint32_t x, int k)
{
  return (x << k) | (x >> (32 - k));
}
CY_DEV uint32_t hash_uint2(uint32_t kx, uint32_t ky)
{
  uint32_t a, b, c;
  a = b = c = 0xdeadbeefu + (2u << 2) + 13u;
  b += ky;
  a += kx;
  c ^= b;
  c -= rot32(b, 14);
  a ^= c;
  a -= rot32(c, 11);
  b ^= a;
  b -= rot32(a, 25);
  c ^= b;
  c -= rot32(b, 16);
  a ^= c;
  a -= rot32(c, 4);
  b ^= a;
  b -= rot32(a, 14);
  c ^= b;
  c -= rot32(b, 24);
  return c;
}

/* kernel/kernel_jitter.h:122-129 */
CY_DEV uint32_t cmj_hash_simple(uint32_t i, uint32_t p)
{
  i = (i ^ 61u) ^ p;
  i += i << 3;
  i ^= i >> 4;
  i *= 0x27d4eb2du;
  return i;
}

/* Sobol points of the samples a wavefront batch covers, tabulated once per batch by
 * k_sobol_table: every path of a batch asks for the same few (sample, dimension) pairs -
 * a batch holds a handful of sample indices, a bounce touches eight dimensions - and the
 * bit loop over the direction vectors was 9 % of the instructions of the shading kernel.
 * tab == NULL (batch too wide for the table): computed on the fly as before. */
struct SobolTable {
  const uint32_t *tab; /* [sample - s0][dimension] */
  int s0;
  uint32_t ns, nd;
};
__constant__ SobolTable g_sobol;

/* kernel/kernel_random.h:40-50 - Sobol via the uploaded direction vectors */
CY_DEV uint32_t sobol_dimension(int index, int dimension)
{
  uint32_t result = 0;
  uint32_t i = (uint32_t)index + 64u; /* SOBOL_SKIP */
  for (int j = 0, x; (x = __ffs((int)i)); i >>= x) {
    j += x;
    result ^= __ldg(&g_scene.sample_pattern_lut[32 * dimension + j - 1]);
  }
  return result;
}

/* Correlated multi-jittered sampling (Kensler 2013) as kernel/kernel_jitter.h:31-196
 * implements it: a keyed permutation of the sample index plus a hashed jitter. */
CY_DEV uint32_t cmj_permute(uint32_t i, uint32_t l, uint32_t p)
{
  uint32_t w = l - 1;
  const bool pow2 = (l & w) == 0;
  if (!pow2) /* smallest 2^k - 1 covering w */
    w = (1u << (32 - __clz((int)w))) - 1u;
  do {
    i ^= p;
    i *= 0xe170893du;
    i ^= p >> 16;
    i ^= (i & w) >> 4;
    i ^= p >> 8;
    i *= 0x0929eb3fu;
    i ^= p >> 23;
    i ^= (i & w) >> 1;
    i *= 1u | p >> 27;
    i *= 0x6935fa69u;
    i ^= (i & w) >> 11;
    i *= 0x74dcb303u;
    i ^= (i & w) >> 2;
    i *= 0x9e501cc3u;
    i ^= (i & w) >> 2;
    i *= 0xc860a3dfu;
    i &= w;
    i ^= i >> 5;
  } while (!pow2 && i >= l); /* cycle-walk when l is not a power of two */
  return pow2 ? ((i + p) & w) : ((i + p) % l);
}
CY_DEV uint32_t cmj_hash(uint32_t i, uint32_t p)
{
  i ^= p;
  i ^= i >> 17;
  i ^= i >> 10;
  i *= 0xb36534e5u;
  i ^= i >> 12;
  i ^= i >> 21;
  i *= 0x93fc4795u;
  i ^= 0xdf6e307fu;
  i ^= i >> 17;
  i *= 1u | p >> 18;
  return i;
}
CY_DEV float cmj_randfloat(uint32_t i, uint32_t p)
{
  return (float)cmj_hash(i, p) * (1.0f / 4294967808.0f);
}
CY_DEV float cmj_sample_1D(int s, int N, uint32_t p)
{
  const uint32_t x = cmj_permute((uint32_t)s, (uint32_t)N, p * 0x68bc21ebu);
  const float jx = cmj_randfloat((uint32_t)s, p * 0x967a889bu);
  const float invN = 1.0f / N;
  return ((float)x + jx) * invN;
}
CY_DEV void cmj_sample_2D(int s, int N, uint32_t p, float *fx, float *fy)
{
  /* an m x n grid with m ~ sqrt(N); the CPU flavour of cmj_isqrt */
  const int m = (int)(sqrtf((float)N) + 1e-6f);
  const int n = (N - 1) / m + 1;
  const float invN = 1.0f / N;
  const float invm = 1.0f / m;
  const float invn = 1.0f / n;
  s = (int)cmj_permute((uint32_t)s, (uint32_t)N, p * 0x51633e2du);
  const int sdivm = s / m; /* equals the shift / mask the reference uses for 2^k */
  const int smodm = s - sdivm * m;
  const uint32_t sx = cmj_permute((uint32_t)smodm, (uint32_t)m, p * 0x68bc21ebu);
  const uint32_t sy = cmj_permute((uint32_t)sdivm, (uint32_t)n, p * 0x02e5be93u);
  const float jx = cmj_randfloat((uint32_t)s, p * 0x967a889bu);
  const float jy = cmj_randfloat((uint32_t)s, p * 0x368cc8b7u);
  *fx = ((float)sx + ((float)sy + jx) * invn) * invm;
  *fy = ((float)s + jy) * invN;
}

/* kernel_jitter.h:198-232: progressive multi-jitter from the host's table
 * (__sample_pattern_lut: NUM_PMJ_PATTERNS x NUM_PMJ_SAMPLES 2D points as floats in
 * [1, 2)), scrambled per pixel and dimension by xor on the mantissa; beyond the table
 * it falls back to hashed random numbers */
/* out of line: the table patterns are rare next to Sobol, and inlined into the kernels'
 * random-number paths they cost the shading-bound workload 4 % (register allocation) */
__device__ __noinline__ float pmj_sample_1D(int sample, uint32_t rng_hash, int dimension)
{
  if (sample >= CY_NUM_PMJ_SAMPLES)
    return cmj_randfloat((uint32_t)sample, rng_hash + (uint32_t)dimension);
  const uint32_t mask = cmj_hash_simple((uint32_t)dimension, rng_hash) & 0x007fffffu;
  const int index = ((dimension % CY_NUM_PMJ_PATTERNS) * CY_NUM_PMJ_SAMPLES + sample) * 2;
  return __uint_as_float(__ldg(&g_scene.sample_pattern_lut[index]) ^ mask) - 1.0f;
}
__device__ __noinline__ float2 pmj_sample_2D(int sample, uint32_t rng_hash, int dimension)
{
  if (sample >= CY_NUM_PMJ_SAMPLES) {
    const uint32_t p = rng_hash + (uint32_t)dimension;
    return make_float2(cmj_randfloat((uint32_t)sample, p), cmj_randfloat((uint32_t)sample, p + 1));
  }
  const int index = ((dimension % CY_NUM_PMJ_PATTERNS) * CY_NUM_PMJ_SAMPLES + sample) * 2;
  const uint32_t maskx = cmj_hash_simple((uint32_t)dimension, rng_hash) & 0x007fffffu;
  const uint32_t masky = cmj_hash_simple((uint32_t)dimension + 1, rng_hash) & 0x007fffffu;
  return make_float2(
      __uint_as_float(__ldg(&g_scene.sample_pattern_lut[index]) ^ maskx) - 1.0f,
      __uint_as_float(__ldg(&g_scene.sample_pattern_lut[index + 1]) ^ masky) - 1.0f);
}

/* kernel/kernel_random.h:53-127: PMJ, CMJ, or Sobol with a Cranley-Patterson rotation.
 * TABLE = false leaves the table pattern (PMJ) out: the lean shading kernels, which the
 * host never selects for a PMJ scene (b200_render: svm_ext). */
template<bool TABLE = true>
CY_DEV float path_rng_1D(uint32_t rng_hash, int sample, int dimension)
{
  if (TABLE && kd_int(KD_INT_SAMPLING_PATTERN) == CY_SAMPLING_PATTERN_PMJ)
    return pmj_sample_1D(sample, rng_hash, dimension);
  if (kd_int(KD_INT_SAMPLING_PATTERN) == CY_SAMPLING_PATTERN_CMJ)
    return cmj_sample_1D(sample, kd_int(KD_INT_AA_SAMPLES), rng_hash + (uint32_t)dimension);
  uint32_t result;
  const uint32_t si = (uint32_t)(sample - g_sobol.s0);
  if (g_sobol.tab && si < g_sobol.ns && (uint32_t)dimension < g_sobol.nd)
    result = __ldg(&g_sobol.tab[si * g_sobol.nd + (uint32_t)dimension]);
  else
    result = sobol_dimension(sample, dimension);
  float r = (float)result * (1.0f / (float)0xFFFFFFFF);
  uint32_t tmp_rng = cmj_hash_simple((uint32_t)dimension, rng_hash);
  float shift = (float)tmp_rng * (1.0f / (float)0xFFFFFFFF);
  return r + shift - floorf(r + shift);
}
template<bool TABLE = true>
CY_DEV void path_rng_2D(uint32_t rng_hash, int sample, int dimension, float *fx, float *fy)
{
  if (TABLE && kd_int(KD_INT_SAMPLING_PATTERN) == CY_SAMPLING_PATTERN_PMJ) {
    const float2 f = pmj_sample_2D(sample, rng_hash, dimension);
    *fx = f.x;
    *fy = f.y;
    return;
  }
  if (kd_int(KD_INT_SAMPLING_PATTERN) == CY_SAMPLING_PATTERN_CMJ) {
    cmj_sample_2D(sample, kd_int(KD_INT_AA_SAMPLES), rng_hash + (uint32_t)dimension, fx, fy);
    return;
  }
  *fx = path_rng_1D<TABLE>(rng_hash, sample, dimension);
  *fy = path_rng_1D<TABLE>(rng_hash, sample, dimension + 1);
}
template<bool TABLE = true>
CY_DEV float path_state_rng_1D(const PathStateG &s, int dimension)
{
  return path_rng_1D<TABLE>(s.rng_hash, s.sample, s.rng_offset + dimension);
}
template<bool TABLE = true>
CY_DEV void path_state_rng_2D(const PathStateG &s, int dimension, float *fx, float *fy)
{
  path_rng_2D<TABLE>(s.rng_hash, s.sample, s.rng_offset + dimension, fx, fy);
}

/* kernel/kernel_globals.h:213-227 */
CY_DEV float lookup_table_read(float x, int offset, int size)
{
  x = saturate(x) * (size - 1);
  int index = min((int)x, size - 1);
  int nindex = min(index + 1, size - 1);
  float t = x - index;
  float data0 = __ldg(&g_scene.lookup_table[index + offset]);
  if (t == 0.0f)
    return data0;
  float data1 = __ldg(&g_scene.lookup_table[nindex + offset]);
  return (1.0f - t) * data0 + t * data1;
}

/* ------------------------------------------------------------- camera ray */

/* kernel_path_common.h:21-46 + kernel_random.h:129-153 + kernel_camera.h:355-427,
 * 42-170 (perspective, no DOF, no motion).  Returns Ray::t (0 = no ray). */
/* kernel_montecarlo.h:150-194 */
CY_DEV float2 concentric_sample_disk(float u1, float u2)
{
  float phi, r;
  float a = 2.0f * u1 - 1.0f;
  float b = 2.0f * u2 - 1.0f;
  if (a == 0.0f && b == 0.0f) {
    return make_float2(0.0f, 0.0f);
  }
  else if (a * a > b * b) {
    r = a;
    phi = CY_M_PI_4_F * (b / a);
  }
  else {
    r = b;
    phi = CY_M_PI_2_F - CY_M_PI_4_F * (a / b);
  }
  return make_float2(r * cosf(phi), r * sinf(phi));
}
CY_DEV float2 regular_polygon_sample(float corners, float rotation, float u, float v)
{
  /* pick a corner triangle, then a uniform point in it */
  float corner = floorf(u * corners);
  u = u * corners - corner;
  u = sqrtf(u);
  v = v * u;
  u = 1.0f - u;
  float angle = CY_M_PI_F / corners;
  float2 p = make_float2((u + v) * cosf(angle), (u - v) * sinf(angle));
  rotation += corner * 2.0f * angle;
  float cr = cosf(rotation);
  float sr = sinf(rotation);
  return make_float2(cr * p.x - sr * p.y, sr * p.x + cr * p.y);
}
/* kernel_camera.h:21-40; the caller scales by the aperture size */
CY_DEV float2 camera_sample_aperture(float u, float v)
{
  const float blades = kd_float(KD_CAM_BLADES);
  float2 bokeh;
  if (blades == 0.0f)
    bokeh = concentric_sample_disk(u, v);
  else
    bokeh = regular_polygon_sample(blades, kd_float(KD_CAM_BLADESROTATION), u, v);
  bokeh.x *= kd_float(KD_CAM_INV_APERTURE_RATIO); /* anamorphic bokeh */
  return bokeh;
}

/* kernel_projection.h:50-215: raster-plane (u, v) of a panoramic camera -> direction in
 * camera space; a zero vector means "outside the lens" */
CY_DEV f3 panorama_to_direction(float u, float v)
{
  switch (kd_int(KD_CAM_PANORAMA_TYPE)) {
    case CY_PANORAMA_EQUIRECTANGULAR: {
      const float4 range = kd_float4(KD_CAM_EQUIRECTANGULAR_RANGE);
      const float phi = range.x * u + range.y;
      const float theta = range.z * v + range.w;
      const float sin_theta = sinf(theta);
      return mk3(sin_theta * cosf(phi), sin_theta * sinf(phi), cosf(theta));
    }
    case CY_PANORAMA_MIRRORBALL: {
      /* point on the ball, then the reflection of the view axis */
      f3 dir;
      dir.x = 2.0f * u - 1.0f;
      dir.z = 2.0f * v - 1.0f;
      if (dir.x * dir.x + dir.z * dir.z > 1.0f)
        return zero3();
      dir.y = -sqrtf(fmaxf(1.0f - dir.x * dir.x - dir.z * dir.z, 0.0f));
      const f3 I = mk3(0.0f, -1.0f, 0.0f);
      return 2.0f * dot(dir, I) * dir - I;
    }
    case CY_PANORAMA_FISHEYE_EQUIDISTANT: {
      const float fov = kd_float(KD_CAM_FISHEYE_FOV);
      u = (u - 0.5f) * 2.0f;
      v = (v - 0.5f) * 2.0f;
      const float r = sqrtf(u * u + v * v);
      if (r > 1.0f)
        return zero3();
      float phi = acosf(fminf(fmaxf((r != 0.0f) ? u / r : 0.0f, -1.0f), 1.0f));
      const float theta = r * fov * 0.5f;
      if (v < 0.0f)
        phi = -phi;
      return mk3(cosf(theta), -cosf(phi) * sinf(theta), sinf(phi) * sinf(theta));
    }
    default: { /* PANORAMA_FISHEYE_EQUISOLID */
      const float lens = kd_float(KD_CAM_FISHEYE_LENS), fov = kd_float(KD_CAM_FISHEYE_FOV);
      u = (u - 0.5f) * kd_float(KD_CAM_SENSORWIDTH);
      v = (v - 0.5f) * kd_float(KD_CAM_SENSORHEIGHT);
      const float rmax = 2.0f * lens * sinf(fov * 0.25f);
      const float r = sqrtf(u * u + v * v);
      if (r > rmax)
        return zero3();
      float phi = acosf(fminf(fmaxf((r != 0.0f) ? u / r : 0.0f, -1.0f), 1.0f));
      const float theta = 2.0f * asinf(r / (2.0f * lens));
      if (v < 0.0f)
        phi = -phi;
      return mk3(cosf(theta), -cosf(phi) * sinf(theta), sinf(phi) * sinf(theta));
    }
  }
}

CY_DEV float camera_ray(int x, int y, int sample, uint32_t *rng_hash, f3 *P, f3 *D)
{
  *rng_hash = hash_uint2((uint32_t)x, (uint32_t)y);
  *rng_hash ^= (uint32_t)kd_int(KD_INT_SEED);

  float filter_u, filter_v;
  if (sample == 0) {
    filter_u = 0.5f;
    filter_v = 0.5f;
  }
  else {
    path_rng_2D(*rng_hash, sample, CY_PRNG_FILTER_U, &filter_u, &filter_v);
  }

  const int filter_table_offset = kd_int(KD_FILM_FILTER_TABLE_OFFSET);
  float raster_x = x + lookup_table_read(filter_u, filter_table_offset, CY_FILTER_TABLE_SIZE);
  float raster_y = y + lookup_table_read(filter_v, filter_table_offset, CY_FILTER_TABLE_SIZE);

  /* lens sample only when there is an aperture - kernel_path_common.h:34-37 */
  float lens_u = 0.0f, lens_v = 0.0f;
  const float aperturesize = kd_float(KD_CAM_APERTURESIZE);
  if (aperturesize > 0.0f)
    path_rng_2D(*rng_hash, sample, CY_PRNG_LENS_U, &lens_u, &lens_v);

  const int r2c = KD_CAM_RASTERTOCAMERA;
  f3 raster = mk3(raster_x, raster_y, 0.0f);
  f3 Pcamera = transform_perspective(kd_float4(r2c), kd_float4(r2c + 16), kd_float4(r2c + 32),
                                     kd_float4(r2c + 48), raster);

  tfm34 c2w;
  c2w.x = kd_float4(KD_CAM_CAMERATOWORLD);
  c2w.y = kd_float4(KD_CAM_CAMERATOWORLD + 16);
  c2w.z = kd_float4(KD_CAM_CAMERATOWORLD + 32);

  if (kd_int(KD_CAM_TYPE) == CY_CAMERA_PANORAMA) {
    /* camera_sample_panorama - kernel_camera.h:238-351 (no stereo, no motion) */
    f3 Dp = panorama_to_direction(Pcamera.x, Pcamera.y);
    if (is_zero(Dp))
      return 0.0f; /* outside the lens: the path receives no light (ray.t = 0) */
    f3 Pp = zero3();
    if (aperturesize > 0.0f) {
      const float2 lensuv = camera_sample_aperture(lens_u, lens_v);
      const f3 Dfocus = normalize(Dp);
      const f3 Pfocus = Dfocus * kd_float(KD_CAM_FOCALDISTANCE);
      /* orthonormal frame perpendicular to the focus direction */
      const f3 U = normalize(mk3(1.0f, 0.0f, 0.0f) - Dfocus.x * Dfocus);
      const f3 V = normalize(cross(Dfocus, U));
      Pp = U * (lensuv.x * aperturesize) + V * (lensuv.y * aperturesize);
      Dp = normalize(Pfocus - Pp);
    }
    f3 Pw = transform_point(c2w, Pp);
    f3 Dw = normalize(transform_direction(c2w, Dp));
    Pw += kd_float(KD_CAM_NEARCLIP) * Dw;
    *P = Pw;
    *D = Dw;
    return kd_float(KD_CAM_CLIPLENGTH);
  }

  if (kd_int(KD_CAM_TYPE) == CY_CAMERA_ORTHOGRAPHIC) {
    /* camera_sample_orthographic - kernel_camera.h:174-235 */
    f3 Po = Pcamera;
    f3 Do = mk3(0.0f, 0.0f, 1.0f);
    if (aperturesize > 0.0f) {
      const float2 lensuv = camera_sample_aperture(lens_u, lens_v);
      const f3 Pfocus = Do * kd_float(KD_CAM_FOCALDISTANCE);
      const f3 lensuvw = mk3(lensuv.x * aperturesize, lensuv.y * aperturesize, 0.0f);
      Po = Pcamera + lensuvw;
      Do = normalize(Pfocus - lensuvw);
    }
    *P = transform_point(c2w, Po);
    *D = normalize(transform_direction(c2w, Do));
    return kd_float(KD_CAM_CLIPLENGTH);
  }

  /* camera_sample_perspective - kernel_camera.h:42-172 (no motion, no stereo) */
  f3 Pc = zero3();
  f3 Dc = Pcamera;
  if (aperturesize > 0.0f) {
    /* point on the aperture, point on the plane of focus */
    const float2 lensuv = camera_sample_aperture(lens_u, lens_v);
    const float ft = kd_float(KD_CAM_FOCALDISTANCE) / Dc.z;
    const f3 Pfocus = Dc * ft;
    Pc = mk3(lensuv.x * aperturesize, lensuv.y * aperturesize, 0.0f);
    Dc = normalize(Pfocus - Pc);
  }
  f3 Pw = transform_point(c2w, Pc);
  f3 Dw = normalize(transform_direction(c2w, Dc));

  /* clipping - kernel_camera.h:158-167 */
  float z_inv = 1.0f / normalize(Pcamera).z;
  float nearclip = kd_float(KD_CAM_NEARCLIP) * z_inv;
  Pw += nearclip * Dw;
  *P = Pw;
  *D = Dw;
  return kd_float(KD_CAM_CLIPLENGTH) * z_inv;
}

/* bvh/bvh.h:541-587 (__INTERSECTION_REFINE__ branch) */
CY_DEV float ray_offset_1(float p, float ng)
{
  const float epsilon_f = 1e-5f;
  const float epsilon_test = 1.0f;
  const int epsilon_i = 32;
  if (fabsf(p) < epsilon_test) {
    return p + ng * epsilon_f;
  }
  uint32_t ix = __float_as_uint(p);
  ix += ((ix ^ __float_as_uint(ng)) >> 31) ? (uint32_t)(-epsilon_i) : (uint32_t)epsilon_i;
  return __uint_as_float(ix);
}
CY_DEV f3 ray_offset(f3 P, f3 Ng)
{
  return mk3(ray_offset_1(P.x, Ng.x), ray_offset_1(P.y, Ng.y), ray_offset_1(P.z, Ng.z));
}

/* geom/geom_triangle_intersect.h:195-256 */
CY_DEV f3 triangle_refine(int isect_prim, int isect_object, float isect_t, f3 P, f3 D)
{
  float t = isect_t;
  if (isect_object != -1) {
    if (t == 0.0f)
      return P;
    tfm34 itfm = object_itfm(isect_object);
    P = transform_point(itfm, P);
    D = transform_direction(itfm, D * t);
    D = normalize_len(D, &t);
  }
  P = P + D * t;

  const uint32_t tri_vindex = __ldg(&g_scene.prim_tri_index[isect_prim]);
  const float4 tri_a = __ldg(&g_scene.prim_tri_verts[tri_vindex + 0]);
  const float4 tri_b = __ldg(&g_scene.prim_tri_verts[tri_vindex + 1]);
  const float4 tri_c = __ldg(&g_scene.prim_tri_verts[tri_vindex + 2]);
  f3 edge1 = mk3(tri_a.x - tri_c.x, tri_a.y - tri_c.y, tri_a.z - tri_c.z);
  f3 edge2 = mk3(tri_b.x - tri_c.x, tri_b.y - tri_c.y, tri_b.z - tri_c.z);
  f3 tvec = mk3(P.x - tri_c.x, P.y - tri_c.y, P.z - tri_c.z);
  f3 qvec = cross(tvec, edge1);
  f3 pvec = cross(D, edge2);
  float det = dot(edge1, pvec);
  if (det != 0.0f) {
    float rt = dot(edge2, qvec) / det;
    P = P + D * rt;
  }
  if (isect_object != -1) {
    tfm34 tfm = object_tfm(isect_object);
    P = transform_point(tfm, P);
  }
  return P;
}

/* kernel_shader.h:59-153 (static triangles) */
CY_DEV void shader_setup_from_ray(
    ShaderDataG &sd, int isect_prim, int isect_object, float t, float u, float v, f3 rayP, f3 rayD)
{
  sd.object = (isect_object == -1) ? (int)__ldg(&g_scene.prim_object[isect_prim]) : isect_object;
  sd.type = CY_PRIMITIVE_TRIANGLE;
  sd.lamp = -1;
  sd.terminator_freq = object_shadow_terminator_offset(sd.object);
  sd.flag = 0;
  sd.object_flag = __ldg(&g_scene.object_flag[sd.object]);
  sd.prim = (int)__ldg(&g_scene.prim_index[isect_prim]);
  sd.ray_length = t;
  sd.u = u;
  sd.v = v;

  /* triangle_normal - geom/geom_triangle.h:26-41 */
  const uint4 tri_vindex = __ldg(&g_scene.tri_vindex[sd.prim]);
  const f3 v0 = mk3(__ldg(&g_scene.prim_tri_verts[tri_vindex.w + 0]));
  const f3 v1 = mk3(__ldg(&g_scene.prim_tri_verts[tri_vindex.w + 1]));
  const f3 v2 = mk3(__ldg(&g_scene.prim_tri_verts[tri_vindex.w + 2]));
  f3 Ng;
  if (sd.object_flag & CY_SD_OBJECT_NEGATIVE_SCALE_APPLIED)
    Ng = normalize(cross(v2 - v0, v1 - v0));
  else
    Ng = normalize(cross(v1 - v0, v2 - v0));

  sd.shader = (int)__ldg(&g_scene.tri_shader[sd.prim]);
  sd.P = triangle_refine(isect_prim, isect_object, t, rayP, rayD);
  sd.Ng = Ng;
  sd.N = Ng;

  if (sd.shader & CY_SHADER_SMOOTH_NORMAL) {
    /* triangle_smooth_normal - geom/geom_triangle.h:80-92 */
    f3 n0 = mk3(__ldg(&g_scene.tri_vnormal[tri_vindex.x]));
    f3 n1 = mk3(__ldg(&g_scene.tri_vnormal[tri_vindex.y]));
    f3 n2 = mk3(__ldg(&g_scene.tri_vnormal[tri_vindex.z]));
    f3 N = safe_normalize((1.0f - u - v) * n2 + u * n0 + v * n1);
    sd.N = is_zero(N) ? Ng : N;
  }

  /* triangle_dPdudv - geom/geom_triangle.h:96-110 */
  sd.dPdu = (v0 - v2);

  sd.I = -rayD;
  sd.flag |= shader_flags(sd.shader);

  if (isect_object != -1) {
    sd.N = object_normal_transform(sd.object, sd.N);
    sd.Ng = object_normal_transform(sd.object, sd.Ng);
    sd.dPdu = object_dir_transform(sd.object, sd.dPdu);
  }

  bool backfacing = (dot(sd.Ng, sd.I) < 0.0f);
  if (backfacing) {
    sd.flag |= CY_SD_BACKFACING;
    sd.Ng = -sd.Ng;
    sd.N = -sd.N;
    sd.dPdu = -sd.dPdu;
  }
}

/* ------------------------------------------------------------------ BSDFs */

#include "bsdf.cuh"

/* --------------------------------------------------------------------- SVM */

#define CY_NODE_GEOM_P 0
#define CY_NODE_GEOM_N 1
#define CY_NODE_GEOM_T 2
#define CY_NODE_GEOM_I 3
#define CY_NODE_GEOM_Ng 4
#define CY_NODE_GEOM_uv 5

#include "svm_closure.cuh"
#include "svm_nodes.cuh"
#include "svm_tex.cuh"
#include "svm_image.cuh"

/* The nodes beyond the closure and basic value set, behind ONE out-of-line call: with
 * their cases in the interpreter's own switch the common shaders (Principled, diffuse,
 * glossy) paid 3 % of a Cornell frame for code they never run, measured A/B on one
 * box. */
__device__ __noinline__ int svm_eval_extended_node(ShaderDataG &sd, float *stack, uint4 node,
                                                   int offset)
{
  /* the program counter comes and goes BY VALUE: taking its address in the interpreter
   * would move it from a register to local memory for every node of every shader */
  switch (node.x) {
    case CY_NODE_HSV:
      svm_node_hsv(stack, node);
      break;
    case CY_NODE_SEPARATE_HSV:
      svm_node_separate_hsv(stack, node, &offset);
      break;
    case CY_NODE_COMBINE_HSV:
      svm_node_combine_hsv(stack, node, &offset);
      break;
    case CY_NODE_MAP_RANGE:
      svm_node_map_range(stack, node, &offset);
      break;
    case CY_NODE_NORMAL:
      svm_node_normal(stack, node, &offset);
      break;
    case CY_NODE_VECTOR_ROTATE:
      svm_node_vector_rotate(stack, node);
      break;
    case CY_NODE_VECTOR_TRANSFORM:
      svm_node_vector_transform(sd, stack, node);
      break;
    case CY_NODE_OBJECT_INFO:
      svm_node_object_info(sd, stack, node);
      break;
    case CY_NODE_CAMERA:
      svm_node_camera(sd, stack, node);
      break;
    case CY_NODE_TEX_WHITE_NOISE:
      svm_node_tex_white_noise(stack, node);
      break;
    case CY_NODE_ATTR:
      svm_node_attr(sd, stack, node);
      break;
    case CY_NODE_TEX_COORD:
      svm_node_tex_coord(sd, stack, node, &offset);
      break;
    case CY_NODE_MAPPING:
      svm_node_mapping(stack, node);
      break;
    case CY_NODE_TEXTURE_MAPPING:
      svm_node_texture_mapping(stack, node, &offset);
      break;
    case CY_NODE_MIN_MAX:
      svm_node_min_max(stack, node, &offset);
      break;
    case CY_NODE_TEX_NOISE:
      svm_node_tex_noise(stack, node, &offset);
      break;
    case CY_NODE_TANGENT:
      svm_node_tangent(sd, stack, node);
      break;
    case CY_NODE_NORMAL_MAP:
      svm_node_normal_map(sd, stack, node);
      break;
    case CY_NODE_BLACKBODY:
      svm_node_blackbody(stack, node);
      break;
    case CY_NODE_WAVELENGTH:
      svm_node_wavelength(stack, node);
      break;
    case CY_NODE_TEX_MUSGRAVE:
      svm_node_tex_musgrave(stack, node, &offset);
      break;
    case CY_NODE_TEX_VORONOI:
      svm_node_tex_voronoi(stack, node, &offset);
      break;
    case CY_NODE_TEX_CHECKER:
      svm_node_tex_checker(stack, node);
      break;
    case CY_NODE_TEX_GRADIENT:
      svm_node_tex_gradient(stack, node);
      break;
    case CY_NODE_TEX_WAVE:
      svm_node_tex_wave(stack, node, &offset);
      break;
    case CY_NODE_TEX_MAGIC:
      svm_node_tex_magic(stack, node, &offset);
      break;
    case CY_NODE_TEX_BRICK:
      svm_node_tex_brick(stack, node, &offset);
      break;
    case CY_NODE_TEX_IMAGE:
      offset = svm_node_tex_image(stack, node, offset);
      break;
    case CY_NODE_TEX_IMAGE_BOX:
      svm_node_tex_image_box(sd, stack, node);
      break;
    case CY_NODE_TEX_ENVIRONMENT:
      svm_node_tex_environment(stack, node);
      break;
    default:
      return -1;
  }
  return offset;
}

/* svm/svm.h:220-300 for the supported opcodes.  max_closures = 0 evaluates only
 * emission / background weights (PATH_RAY_EMISSION / TERMINATE evaluation,
 * kernel_shader.h:1063-1075).
 *
 * Two instances.  FULL = false knows the closure nodes and the basic value nodes, calls
 * nothing (a leaf function, its program counter and temporaries all in registers) and
 * returns false the moment a shader needs more - an extended node, or a Principled BSDF
 * with sheen.  FULL = true runs everything; its extended nodes sit behind out-of-line
 * calls.  Measured A/B on one box with the Cornell workload (Principled + diffuse
 * shaders only): one interpreter carrying everything costs 4.5 % of the frame, the
 * split costs nothing. */
template<bool FULL, bool MS = FULL>
__device__ __noinline__ bool svm_eval_nodes_t(ShaderDataG &sd, LobeArena &arena,
                                              PathDepths depths, uint32_t path_flag,
                                              int max_closures)
{
  float stack[SVM_STACK_GPU];
  arena_reset(arena, max_closures);
  sd.transparent_at = -1;
  sd.svm_closure_weight = zero3();
  int offset = sd.shader & CY_SHADER_MASK;

  while (true) {
    const uint4 node = __ldg(&g_scene.svm_nodes[offset]);
    offset++;
    switch (node.x) {
      case CY_NODE_END:
        return true;
      case CY_NODE_SHADER_JUMP:
        offset = (int)node.y; /* SHADER_TYPE_SURFACE */
        break;
      case CY_NODE_CLOSURE_BSDF:
        if (!svm_node_closure_bsdf<FULL, MS>(sd, arena, stack, node, path_flag, &offset))
          return false;
        break;
      case CY_NODE_CLOSURE_EMISSION:
      case CY_NODE_CLOSURE_BACKGROUND: {
        /* svm_closure.h:1068-1100 */
        f3 weight = sd.svm_closure_weight;
        if (stack_valid(node.y)) {
          float mix_weight = stack[node.y];
          if (mix_weight == 0.0f)
            break;
          weight *= mix_weight;
        }
        emission_setup(sd, weight);
        break;
      }
      case CY_NODE_CLOSURE_SET_WEIGHT:
        sd.svm_closure_weight = mk3(__uint_as_float(node.y), __uint_as_float(node.z),
                                    __uint_as_float(node.w));
        break;
      case CY_NODE_CLOSURE_WEIGHT:
        sd.svm_closure_weight = stack_load_float3(stack, node.y);
        break;
      case CY_NODE_EMISSION_WEIGHT: {
        float strength = stack[node.z];
        sd.svm_closure_weight = stack_load_float3(stack, node.y) * strength;
        break;
      }
      case CY_NODE_MIX_CLOSURE: {
        uint32_t weight_offset = node.y & 0xff, in_weight_offset = (node.y >> 8) & 0xff;
        uint32_t weight1_offset = (node.y >> 16) & 0xff, weight2_offset = (node.y >> 24) & 0xff;
        float weight = saturate(stack[weight_offset]);
        float in_weight = stack_valid(in_weight_offset) ? stack[in_weight_offset] : 1.0f;
        if (stack_valid(weight1_offset))
          stack[weight1_offset] = in_weight * (1.0f - weight);
        if (stack_valid(weight2_offset))
          stack[weight2_offset] = in_weight * weight;
        break;
      }
      case CY_NODE_JUMP_IF_ZERO:
        if (stack[node.z] == 0.0f)
          offset += (int)node.y;
        break;
      case CY_NODE_JUMP_IF_ONE:
        if (stack[node.z] == 1.0f)
          offset += (int)node.y;
        break;
      case CY_NODE_GEOMETRY: {
        f3 data;
        switch (node.y) {
          case CY_NODE_GEOM_P:
            data = sd.P;
            break;
          case CY_NODE_GEOM_N:
            data = sd.N;
            break;
          case CY_NODE_GEOM_T:
            /* every Principled BSDF carries this node (its tangent input defaults to
             * it), so it belongs to the lean set too */
            data = primitive_tangent(sd);
            break;
          case CY_NODE_GEOM_I:
            data = sd.I;
            break;
          case CY_NODE_GEOM_Ng:
            data = sd.Ng;
            break;
          case CY_NODE_GEOM_uv:
            data = mk3(sd.u, sd.v, 0.0f);
            break;
          default:
            data = zero3();
        }
        stack_store_float3(stack, node.z, data);
        break;
      }
      case CY_NODE_VALUE_F:
        stack[node.z] = __uint_as_float(node.y);
        break;
      case CY_NODE_VALUE_V: {
        const uint4 node1 = __ldg(&g_scene.svm_nodes[offset]);
        offset++;
        stack_store_float3(stack, node.y,
                           mk3(__uint_as_float(node1.y), __uint_as_float(node1.z),
                               __uint_as_float(node1.w)));
        break;
      }
      case CY_NODE_CONVERT:
        svm_node_convert(stack, node.y, node.z, node.w);
        break;
      case CY_NODE_FRESNEL:
        svm_node_fresnel(sd, stack, node);
        break;
      case CY_NODE_LAYER_WEIGHT:
        svm_node_layer_weight(sd, stack, node);
        break;
      case CY_NODE_MATH:
        svm_node_math(stack, node);
        break;
      case CY_NODE_VECTOR_MATH:
        svm_node_vector_math(stack, node, &offset);
        break;
      case CY_NODE_MIX:
        svm_node_mix(stack, node, &offset);
        break;
      case CY_NODE_INVERT:
        svm_node_invert(stack, node);
        break;
      case CY_NODE_GAMMA:
        svm_node_gamma(stack, node);
        break;
      case CY_NODE_BRIGHTCONTRAST:
        svm_node_brightness(stack, node);
        break;
      case CY_NODE_SEPARATE_VECTOR: {
        /* svm_sepcomb_vector.h: (vector offset, component index, out offset) */
        const f3 vec = stack_load_float3(stack, node.y);
        if (stack_valid(node.w))
          stack[node.w] = (node.z == 0) ? vec.x : ((node.z == 1) ? vec.y : vec.z);
        break;
      }
      case CY_NODE_COMBINE_VECTOR:
        if (stack_valid(node.w))
          stack[node.w + node.z] = stack[node.y];
        break;
      case CY_NODE_CLAMP:
        svm_node_clamp(stack, node, &offset);
        break;
      case CY_NODE_LIGHT_PATH:
        svm_node_light_path(sd, depths, stack, node.y, node.z, path_flag);
        break;
      case CY_NODE_LIGHT_FALLOFF:
        svm_node_light_falloff(sd, stack, node);
        break;
      case CY_NODE_RGB_RAMP:
        svm_node_rgb_ramp(stack, node, &offset);
        break;
      case CY_NODE_RGB_CURVES:
      case CY_NODE_VECTOR_CURVES:
        svm_node_curves(stack, node, &offset);
        break;
      default:
        /* the texture / attribute / colour nodes; anything else was refused at bind
         * time by svm_validate() */
        if (!FULL)
          return false;
        offset = svm_eval_extended_node(sd, stack, node, offset);
        if (offset < 0)
          return true;
        break;
    }
  }
}

/* Set when the lean interpreter met something only the full one handles although the
 * host's scan of the program (svm_validate: SVM_USES_EXTENDED_NODES) promised there was
 * none; b200_render reports it as an internal error instead of returning a wrong frame. */
__device__ unsigned int g_svm_scope_miss;

/* EXT is a property of the bound program, decided on the host, and selects the kernel
 * instance: programs with extended nodes (or a Principled BSDF whose sheen is not a
 * constant zero, or a multi-scatter lobe) run the full interpreter, all others the lean
 * one. */
template<bool EXT, bool MS = EXT>
CY_DEV void svm_eval_nodes(ShaderDataG &sd, LobeArena &arena, PathDepths depths,
                           uint32_t path_flag, int max_closures)
{
  if (EXT) {
    svm_eval_nodes_t<true, true>(sd, arena, depths, path_flag, max_closures);
  }
  else if (!svm_eval_nodes_t<false, MS>(sd, arena, depths, path_flag, max_closures)) {
    g_svm_scope_miss = 1u;
    arena_reset(arena, 0);
    sd.flag &= ~(CY_SD_BSDF | CY_SD_EMISSION);
  }
}

/* kernel_shader.h:1057-1112: the surface shader of a hit.  `rng_seed` = rng_hash +
 * rng_offset + sample * 0xb4bc3953 of the path (lcg_state_init): the seed of the random
 * walks of multi-scatter lobes, drawn from only when the shader made one. */
template<bool EXT, bool MS = EXT>
CY_DEV void shader_eval_surface(ShaderDataG &sd, LobeArena &arena, PathDepths depths,
                                uint32_t path_flag, uint32_t rng_seed)
{
  int max_closures;
  if (path_flag & (CY_PATH_RAY_TERMINATE | CY_PATH_RAY_SHADOW | CY_PATH_RAY_EMISSION))
    max_closures = 0;
  else
    max_closures = min(kd_int(KD_INT_MAX_CLOSURES), MAX_CLOSURES_GPU);
  svm_eval_nodes<EXT, MS>(sd, arena, depths, path_flag, max_closures);
  if (MS && (sd.flag & CY_SD_BSDF_NEEDS_LCG))
    sd.lcg_state = lcg_seed(rng_seed);
}

/* A shader that can only emit (a lamp, the background, an emission-only evaluation of a
 * surface): no lobes, so no arena behind it. */
template<bool EXT>
CY_DEV void shader_eval_emission(ShaderDataG &sd, PathDepths depths, uint32_t path_flag)
{
  LobeArena none;
  none.q = nullptr;
  svm_eval_nodes<EXT>(sd, none, depths, path_flag, 0);
}

/* First hit of a path with several lobes: every lobe keeps at least an eighth of the
 * total sample weight, so that a lobe that matters after the bounce is not starved by
 * how it looks from the camera (shader_prepare_closures, kernel_shader.h:530-555);
 * EXT also decides whether the terminator factors are needed at this point. */
template<bool EXT>
CY_DEV void shader_prepare_lobes(ShaderDataG &sd, LobeArena &arena, bool first_hit)
{
  if (EXT)
    bsdf_terminator_terms_setup(sd, arena);
  if (first_hit && arena.n > 1) {
    float sum = 0.0f;
    int at = 0;
    for (int i = 0; i < arena.n; i++) {
      const uint32_t kind = lobe_kind_at(arena, at);
      if (lobe_is_sampled(kind))
        sum += lobe_sample_weight_at(arena, at);
      at += lobe_words(kind);
    }
    at = 0;
    for (int i = 0; i < arena.n; i++) {
      const uint32_t kind = lobe_kind_at(arena, at);
      if (lobe_is_sampled(kind))
        lobe_set_sample_weight_at(arena, at,
                                  fmaxf(lobe_sample_weight_at(arena, at), 0.125f * sum));
      at += lobe_words(kind);
    }
  }
}

/* filter_glossy (kernel_path.h:280-299): widen the roughness of the GGX lobes after a
 * blurry bounce */
CY_DEV void shader_blur_lobes(LobeArena &arena, float roughness)
{
  int at = 0;
  for (int i = 0; i < arena.n; i++) {
    lobe_blur_at(arena, at, roughness);
    at += lobe_words(lobe_kind_at(arena, at));
  }
}

/* power heuristic of multiple importance sampling (beta = 2) */
CY_DEV float power_heuristic(float a, float b)
{
  return (a * a) / (a * a + b * b);
}

/* Sum of all lobes but `skip` for direction omega_in (value weighted by the lobe's
 * colour, pdf by its sample weight), continuing the running sums of a sampled lobe
 * (_shader_bsdf_multi_eval, kernel_shader.h:556-582, no light passes) */
template<bool EXT, bool MS = EXT>
MULTI_EVAL_ATTR void shader_bsdf_multi_eval(ShaderDataG &sd, const LobeArena &arena, f3 omega_in,
                                   float *pdf, int skip, f3 *result_eval, float sum_pdf,
                                   float sum_sample_weight)
{
  int at = 0;
  for (int i = 0; i < arena.n; i++) {
    const uint32_t kind = lobe_kind_at(arena, at);
    if (i != skip && lobe_is_bsdf(kind)) {
      const Lobe l = lobe_fetch(arena, at);
      float bsdf_pdf = 0.0f;
      const f3 eval = bsdf_eval<EXT, MS>(sd, l, omega_in, &bsdf_pdf);
      if (bsdf_pdf != 0.0f) {
        *result_eval += eval * l.weight;
        sum_pdf += bsdf_pdf * l.sample_weight;
      }
      sum_sample_weight += l.sample_weight;
    }
    at += lobe_words(kind);
  }
  *pdf = (sum_sample_weight > 0.0f) ? sum_pdf / sum_sample_weight : 0.0f;
}

/* BSDF towards a light sample, MIS-weighted against BSDF sampling
 * (shader_bsdf_eval, kernel_shader.h:612-636, non-branched) */
template<bool EXT, bool MS = EXT>
CY_DEV f3 shader_bsdf_eval(ShaderDataG &sd, const LobeArena &arena, f3 omega_in, float light_pdf,
                           bool use_mis, f3 *sum_no_mis = nullptr)
{
  f3 eval = zero3();
  float pdf;
  shader_bsdf_multi_eval<EXT, MS>(sd, arena, omega_in, &pdf, -1, &eval, 0.0f, 0.0f);
  if (sum_no_mis) /* BsdfEval::sum_no_mis, kernel_accumulate.h:62-69 */
    *sum_no_mis = eval;
  if (use_mis)
    eval *= power_heuristic(light_pdf, pdf);
  return eval;
}

/* Picks a lobe in proportion to its sample weight, samples it, then adds the other lobes
 * for the sampled direction (shader_bsdf_pick + shader_bsdf_sample,
 * kernel_shader.h:638-680, 739-775) */
template<bool EXT, bool MS = EXT>
CY_DEV int shader_bsdf_sample(ShaderDataG &sd, const LobeArena &arena, float randu, float randv,
                              f3 *bsdf_eval_out, f3 *omega_in, float *pdf)
{
  int sampled = 0, sampled_at = 0;
  if (arena.n > 1) {
    float sum = 0.0f;
    int at = 0;
    for (int i = 0; i < arena.n; i++) {
      const uint32_t kind = lobe_kind_at(arena, at);
      if (lobe_is_sampled(kind))
        sum += lobe_sample_weight_at(arena, at);
      at += lobe_words(kind);
    }
    const float r = randu * sum;
    float partial_sum = 0.0f;
    at = 0;
    for (int i = 0; i < arena.n; i++) {
      const uint32_t kind = lobe_kind_at(arena, at);
      if (lobe_is_sampled(kind)) {
        const float sw = lobe_sample_weight_at(arena, at);
        const float next_sum = partial_sum + sw;
        if (r < next_sum) {
          sampled = i;
          sampled_at = at;
          /* rescale so the number can be reused for the lobe's own sampling */
          randu = (r - partial_sum) / sw;
          break;
        }
        partial_sum = next_sum;
      }
      at += lobe_words(kind);
    }
  }
  *pdf = 0.0f;
  if (!lobe_is_bsdf(lobe_kind_at(arena, sampled_at)))
    return CY_LABEL_NONE;
  const Lobe l = lobe_fetch(arena, sampled_at);
  f3 eval = zero3();
  const int label = bsdf_sample<EXT, MS>(sd, l, randu, randv, &eval, omega_in, pdf);
  if (*pdf != 0.0f) {
    *bsdf_eval_out = eval * l.weight;
    if (arena.n > 1) {
      const float sweight = l.sample_weight;
      shader_bsdf_multi_eval<EXT, MS>(sd, arena, *omega_in, pdf, sampled, bsdf_eval_out,
                                  *pdf * sweight, sweight);
    }
  }
  return label;
}

/* --------------------------------------------------------------- lights */

struct LightSampleG {
  f3 P, Ng, D;
  float t, u, v, pdf, eval_fac;
  int object, prim, shader, lamp, type;
};

CY_DEV const uint8_t *light_ptr(int lamp)
{
  return g_scene.lights + (size_t)lamp * SIZEOF_KERNEL_LIGHT;
}
CY_DEV float kl_float(const uint8_t *kl, int off)
{
  return __ldg((const float *)(kl + off));
}
CY_DEV int kl_int(const uint8_t *kl, int off)
{
  return __ldg((const int *)(kl + off));
}
CY_DEV f3 kl_float3(const uint8_t *kl, int off)
{
  return mk3(kl_float(kl, off), kl_float(kl, off + 4), kl_float(kl, off + 8));
}

/* kernel_montecarlo.h:39-46 */
CY_DEV void to_unit_disk(float *x, float *y)
{
  float phi = CY_2PI_F * (*x);
  float r = sqrtf(*y);
  *x = r * cosf(phi);
  *y = r * sinf(phi);
}

/* kernel_light_common.h:117-150 */
CY_DEV f3 ellipse_sample(f3 ru, f3 rv, float randu, float randv)
{
  to_unit_disk(&randu, &randv);
  return ru * randu + rv * randv;
}
CY_DEV f3 disk_light_sample(f3 v, float randu, float randv)
{
  f3 ru, rv;
  make_orthonormals(v, &ru, &rv);
  return ellipse_sample(ru, rv, randu, randv);
}
CY_DEV f3 distant_light_sample(f3 D, float radius, float randu, float randv)
{
  return normalize(D + disk_light_sample(D, randu, randv) * radius);
}
CY_DEV f3 sphere_light_sample(f3 P, f3 center, float radius, float randu, float randv)
{
  return disk_light_sample(normalize(P - center), randu, randv) * radius;
}
CY_DEV float smoothstepf(float f)
{
  float ff = f * f;
  return (3.0f * ff - 2.0f * ff * f);
}
CY_DEV float spot_light_attenuation(f3 dir, float spot_angle, float spot_smooth, f3 N)
{
  float attenuation = dot(dir, N);
  if (attenuation <= spot_angle) {
    attenuation = 0.0f;
  }
  else {
    float t = attenuation - spot_angle;
    if (t < spot_smooth && spot_smooth != 0.0f)
      attenuation *= smoothstepf(t / spot_smooth);
  }
  return attenuation;
}
CY_DEV float lamp_light_pdf(f3 Ng, f3 I, float t)
{
  float cos_pi = dot(Ng, I);
  if (cos_pi <= 0.0f)
    return 0.0f;
  return t * t / cos_pi;
}

/* kernel_light_common.h:31-115 (Urena et al. spherical rectangle) */
CY_DEV float rect_light_sample(f3 P, f3 *light_p, f3 axisu, f3 axisv, float randu, float randv,
                               bool sample_coord)
{
  f3 corner = *light_p - axisu * 0.5f - axisv * 0.5f;
  float axisu_len, axisv_len;
  f3 x = normalize_len(axisu, &axisu_len);
  f3 y = normalize_len(axisv, &axisv_len);
  f3 z = cross(x, y);
  f3 dir = corner - P;
  float z0 = dot(dir, z);
  if (z0 > 0.0f) {
    z *= -1.0f;
    z0 *= -1.0f;
  }
  float x0 = dot(dir, x);
  float y0 = dot(dir, y);
  float x1 = x0 + axisu_len;
  float y1 = y0 + axisv_len;
  /* float4 arithmetic component-wise, util_math_float4.h */
  float dfx = x0 - x1, dfy = y1 - y0, dfz = x1 - x0, dfw = y0 - y1;
  float nzx = y0 * dfx, nzy = x1 * dfy, nzz = y1 * dfz, nzw = x0 * dfw;
  nzx = nzx / sqrtf(z0 * z0 * dfx * dfx + nzx * nzx);
  nzy = nzy / sqrtf(z0 * z0 * dfy * dfy + nzy * nzy);
  nzz = nzz / sqrtf(z0 * z0 * dfz * dfz + nzz * nzz);
  nzw = nzw / sqrtf(z0 * z0 * dfw * dfw + nzw * nzw);
  float g0 = safe_acosf(-nzx * nzy);
  float g1 = safe_acosf(-nzy * nzz);
  float g2 = safe_acosf(-nzz * nzw);
  float g3 = safe_acosf(-nzw * nzx);
  float b0 = nzx;
  float b1 = nzz;
  float b0sq = b0 * b0;
  float k = CY_2PI_F - g2 - g3;
  float S = g0 + g1 - k;

  if (sample_coord) {
    float au = randu * S + k;
    float fu = (cosf(au) * b0 - b1) / sinf(au);
    float cu = 1.0f / sqrtf(fu * fu + b0sq) * (fu > 0.0f ? 1.0f : -1.0f);
    cu = clampf(cu, -1.0f, 1.0f);
    float xu = -(cu * z0) / fmaxf(sqrtf(1.0f - cu * cu), 1e-7f);
    xu = clampf(xu, x0, x1);
    float z0sq = z0 * z0;
    float y0sq = y0 * y0;
    float y1sq = y1 * y1;
    float d = sqrtf(xu * xu + z0sq);
    float h0 = y0 / sqrtf(d * d + y0sq);
    float h1 = y1 / sqrtf(d * d + y1sq);
    float hv = h0 + randv * (h1 - h0), hv2 = hv * hv;
    float yv = (hv2 < 1.0f - 1e-6f) ? (hv * d) / sqrtf(1.0f - hv2) : y1;
    *light_p = P + xu * x + yv * y + z0 * z;
  }
  if (S != 0.0f)
    return 1.0f / S;
  else
    return 0.0f;
}

#include "light_background.cuh"

/* kernel_light.h:40-168 (triangle lights: light_tri.cuh) */
CY_DEV bool lamp_light_sample(int lamp, float randu, float randv, f3 P, LightSampleG *ls)
{
  const uint8_t *kl = light_ptr(lamp);
  const int type = kl_int(kl, KL_TYPE);
  ls->type = type;
  ls->shader = kl_int(kl, KL_SHADER_ID);
  ls->object = CY_PRIM_NONE;
  ls->prim = CY_PRIM_NONE;
  ls->lamp = lamp;
  ls->u = randu;
  ls->v = randv;

  if (type == CY_LIGHT_DISTANT) {
    f3 lightD = kl_float3(kl, KL_CO);
    f3 D = lightD;
    float radius = kl_float(kl, KL_DISTANT_RADIUS);
    float invarea = kl_float(kl, KL_DISTANT_INVAREA);
    if (radius > 0.0f)
      D = distant_light_sample(D, radius, randu, randv);
    ls->P = D;
    ls->Ng = D;
    ls->D = -D;
    ls->t = FLT_MAX;
    float costheta = dot(lightD, D);
    ls->pdf = invarea / (costheta * costheta * costheta);
    ls->eval_fac = ls->pdf;
  }
  else if (type == CY_LIGHT_BACKGROUND) {
    /* the world as a light dome (kernel_light.h:72-83), importance-sampled by its map */
    const f3 D = -background_light_sample(randu, randv, &ls->pdf);
    ls->P = D;
    ls->Ng = D;
    ls->D = -D;
    ls->t = FLT_MAX;
    ls->eval_fac = 1.0f;
  }
  else {
    ls->P = kl_float3(kl, KL_CO);
    if (type == CY_LIGHT_POINT || type == CY_LIGHT_SPOT) {
      float radius = kl_float(kl, KL_SPOT_RADIUS);
      if (radius > 0.0f)
        ls->P += sphere_light_sample(P, ls->P, radius, randu, randv);
      ls->D = normalize_len(ls->P - P, &ls->t);
      ls->Ng = -ls->D;
      float invarea = kl_float(kl, KL_SPOT_INVAREA);
      ls->eval_fac = (0.25f * CY_1_PI_F) * invarea;
      ls->pdf = invarea;
      if (type == CY_LIGHT_SPOT) {
        f3 dir = kl_float3(kl, KL_SPOT_DIR);
        ls->eval_fac *= spot_light_attenuation(dir, kl_float(kl, KL_SPOT_SPOT_ANGLE),
                                               kl_float(kl, KL_SPOT_SPOT_SMOOTH), ls->Ng);
        if (ls->eval_fac == 0.0f)
          return false;
      }
      ls->pdf *= lamp_light_pdf(ls->Ng, -ls->D, ls->t);
    }
    else {
      f3 axisu = kl_float3(kl, KL_AREA_AXISU);
      f3 axisv = kl_float3(kl, KL_AREA_AXISV);
      f3 D = kl_float3(kl, KL_AREA_DIR);
      float invarea_raw = kl_float(kl, KL_AREA_INVAREA);
      float invarea = fabsf(invarea_raw);
      bool is_round = (invarea_raw < 0.0f);
      if (dot(ls->P - P, D) > 0.0f)
        return false;
      f3 inplane;
      if (is_round) {
        inplane = ellipse_sample(axisu * 0.5f, axisv * 0.5f, randu, randv);
        ls->P += inplane;
        ls->pdf = invarea;
      }
      else {
        inplane = ls->P;
        ls->pdf = rect_light_sample(P, &ls->P, axisu, axisv, randu, randv, true);
        inplane = ls->P - inplane;
      }
      ls->u = dot(inplane, axisu) * (1.0f / dot(axisu, axisu)) + 0.5f;
      ls->v = dot(inplane, axisv) * (1.0f / dot(axisv, axisv)) + 0.5f;
      ls->Ng = D;
      ls->D = normalize_len(ls->P - P, &ls->t);
      ls->eval_fac = 0.25f * invarea;
      if (is_round)
        ls->pdf *= lamp_light_pdf(D, -ls->D, ls->t);
    }
  }
  ls->pdf *= kd_float(KD_INT_PDF_LIGHTS);
  return (ls->pdf > 0.0f);
}

/* util_math_intersect.h:58-86 */
CY_DEV bool ray_aligned_disk_intersect(f3 ray_P, f3 ray_D, float ray_t, f3 disk_P,
                                       float disk_radius, f3 *isect_P, float *isect_t)
{
  float disk_t;
  const f3 disk_N = normalize_len(ray_P - disk_P, &disk_t);
  const float div = dot(ray_D, disk_N);
  if (div == 0.0f)
    return false;
  const float t = -disk_t / div;
  if (t < 0.0f || t > ray_t)
    return false;
  f3 P = ray_P + ray_D * t;
  if (len_squared(P - disk_P) > disk_radius * disk_radius)
    return false;
  *isect_P = P;
  *isect_t = t;
  return true;
}

/* util_math_intersect.h:202-245 */
CY_DEV bool ray_quad_intersect(f3 ray_P, f3 ray_D, float ray_mint, float ray_maxt, f3 quad_P,
                               f3 quad_u, f3 quad_v, f3 quad_n, f3 *isect_P, float *isect_t,
                               float *isect_u, float *isect_v, bool ellipse)
{
  float t = -(dot(ray_P, quad_n) - dot(quad_P, quad_n)) / dot(ray_D, quad_n);
  if (t < ray_mint || t > ray_maxt)
    return false;
  const f3 hit = ray_P + t * ray_D;
  const f3 inplane = hit - quad_P;
  const float u = dot(inplane, quad_u) / dot(quad_u, quad_u);
  if (u < -0.5f || u > 0.5f)
    return false;
  const float v = dot(inplane, quad_v) / dot(quad_v, quad_v);
  if (v < -0.5f || v > 0.5f)
    return false;
  if (ellipse && (u * u + v * v > 0.25f))
    return false;
  *isect_P = hit;
  *isect_t = t;
  *isect_u = u + 0.5f;
  *isect_v = v + 0.5f;
  return true;
}

/* kernel_light.h:170-300 */
CY_DEV bool lamp_light_eval(int lamp, f3 P, f3 D, float t, LightSampleG *ls)
{
  const uint8_t *kl = light_ptr(lamp);
  const int type = kl_int(kl, KL_TYPE);
  ls->type = type;
  ls->shader = kl_int(kl, KL_SHADER_ID);
  ls->object = CY_PRIM_NONE;
  ls->prim = CY_PRIM_NONE;
  ls->lamp = lamp;
  ls->u = 0.0f;
  ls->v = 0.0f;

  if (!(ls->shader & CY_SHADER_USE_MIS))
    return false;

  if (type == CY_LIGHT_DISTANT) {
    float radius = kl_float(kl, KL_DISTANT_RADIUS);
    if (radius == 0.0f)
      return false;
    if (t != FLT_MAX)
      return false;
    f3 lightD = kl_float3(kl, KL_CO);
    float costheta = dot(-lightD, D);
    float cosangle = kl_float(kl, KL_DISTANT_COSANGLE);
    if (costheta < cosangle)
      return false;
    ls->P = -D;
    ls->Ng = -D;
    ls->D = D;
    ls->t = FLT_MAX;
    float invarea = kl_float(kl, KL_DISTANT_INVAREA);
    ls->pdf = invarea / (costheta * costheta * costheta);
    ls->eval_fac = ls->pdf;
  }
  else if (type == CY_LIGHT_POINT || type == CY_LIGHT_SPOT) {
    f3 lightP = kl_float3(kl, KL_CO);
    float radius = kl_float(kl, KL_SPOT_RADIUS);
    if (radius == 0.0f)
      return false;
    if (!ray_aligned_disk_intersect(P, D, t, lightP, radius, &ls->P, &ls->t))
      return false;
    ls->Ng = -D;
    ls->D = D;
    float invarea = kl_float(kl, KL_SPOT_INVAREA);
    ls->eval_fac = (0.25f * CY_1_PI_F) * invarea;
    ls->pdf = invarea;
    if (type == CY_LIGHT_SPOT) {
      f3 dir = kl_float3(kl, KL_SPOT_DIR);
      ls->eval_fac *= spot_light_attenuation(dir, kl_float(kl, KL_SPOT_SPOT_ANGLE),
                                             kl_float(kl, KL_SPOT_SPOT_SMOOTH), ls->Ng);
      if (ls->eval_fac == 0.0f)
        return false;
    }
    if (ls->t != FLT_MAX)
      ls->pdf *= lamp_light_pdf(ls->Ng, -ls->D, ls->t);
  }
  else if (type == CY_LIGHT_AREA) {
    float invarea_raw = kl_float(kl, KL_AREA_INVAREA);
    float invarea = fabsf(invarea_raw);
    bool is_round = (invarea_raw < 0.0f);
    if (invarea == 0.0f)
      return false;
    f3 axisu = kl_float3(kl, KL_AREA_AXISU);
    f3 axisv = kl_float3(kl, KL_AREA_AXISV);
    f3 Ng = kl_float3(kl, KL_AREA_DIR);
    if (dot(D, Ng) >= 0.0f)
      return false;
    f3 light_P = kl_float3(kl, KL_CO);
    if (!ray_quad_intersect(P, D, 0.0f, t, light_P, axisu, axisv, Ng, &ls->P, &ls->t, &ls->u,
                            &ls->v, is_round))
      return false;
    ls->D = D;
    ls->Ng = Ng;
    if (is_round)
      ls->pdf = invarea * lamp_light_pdf(Ng, -D, ls->t);
    else
      ls->pdf = rect_light_sample(P, &light_P, axisu, axisv, 0, 0, false);
    ls->eval_fac = 0.25f * invarea;
  }
  else {
    return false;
  }
  ls->pdf *= kd_float(KD_INT_PDF_LIGHTS);
  return true;
}

/* kernel_light.h:582-614 */
CY_DEV int light_distribution_sample(float *randu)
{
  int first = 0;
  const int num_distribution = kd_int(KD_INT_NUM_DISTRIBUTION);
  int len = num_distribution + 1;
  float r = *randu;
  do {
    int half_len = len >> 1;
    int middle = first + half_len;
    float totarea = __ldg((const float *)(g_scene.light_distribution +
                                          (size_t)middle * SIZEOF_KERNEL_LIGHT_DISTRIBUTION +
                                          KLD_TOTAREA));
    if (r < totarea) {
      len = half_len;
    }
    else {
      first = middle + 1;
      len = len - half_len - 1;
    }
  } while (len > 0);
  int index = min(max(first - 1, 0), num_distribution - 1);
  float distr_min = __ldg((const float *)(g_scene.light_distribution +
                                          (size_t)index * SIZEOF_KERNEL_LIGHT_DISTRIBUTION));
  float distr_max = __ldg((const float *)(g_scene.light_distribution +
                                          (size_t)(index + 1) * SIZEOF_KERNEL_LIGHT_DISTRIBUTION));
  *randu = (r - distr_min) / (distr_max - distr_min);
  return index;
}

#include "light_tri.cuh"

/* kernel_light.h:624-660 (lamp < 0: pick from the distribution) */
LIGHT_SAMPLE_ATTR bool light_sample(float randu, float randv, f3 P, int bounce, LightSampleG *ls)
{
  int index = light_distribution_sample(&randu);
  const uint8_t *kd = g_scene.light_distribution + (size_t)index * SIZEOF_KERNEL_LIGHT_DISTRIBUTION;
  int prim = __ldg((const int *)(kd + KLD_PRIM));
  if (prim >= 0) {
    /* an emissive triangle */
    const int object = __ldg((const int *)(kd + KLD_MESH_OBJECT_ID));
    const int shader_flag = __ldg((const int *)(kd + KLD_MESH_SHADER_FLAG));
    triangle_light_sample(prim, object, randu, randv, ls, P);
    ls->shader |= shader_flag;
    return ls->pdf > 0.0f;
  }
  int lamp = -prim - 1;
  if ((float)bounce > kl_float(light_ptr(lamp), KL_MAX_BOUNCES))
    return false;
  return lamp_light_sample(lamp, randu, randv, P, ls);
}

/* ------------------------------------------------------------- emission */

/* kernel_shader.h:978-992 */
CY_DEV bool shader_constant_emission_eval(int shader, f3 *eval)
{
  const uint8_t *ks = g_scene.shaders + (size_t)(shader & CY_SHADER_MASK) * SIZEOF_KERNEL_SHADER;
  uint32_t flags = __ldg((const uint32_t *)(ks + KS_FLAGS));
  if (flags & CY_SD_HAS_CONSTANT_EMISSION) {
    *eval = mk3(__ldg((const float *)(ks + KS_CONSTANT_EMISSION)),
                __ldg((const float *)(ks + KS_CONSTANT_EMISSION + 4)),
                __ldg((const float *)(ks + KS_CONSTANT_EMISSION + 8)));
    return true;
  }
  return false;
}

/* kernel_emission.h:20-98.  `emission_sd` is scratch. */
template<bool EXT>
CY_DEV f3 direct_emissive_eval(ShaderDataG &emission_sd, PathDepths depths, LightSampleG *ls,
                               f3 I, float t)
{
  f3 eval = zero3();
  if (shader_constant_emission_eval(ls->shader, &eval)) {
    if ((ls->prim != CY_PRIM_NONE) && dot(ls->Ng, I) < 0.0f)
      ls->Ng = -ls->Ng;
  }
  else if (ls->type == CY_LIGHT_BACKGROUND) {
    /* kernel_emission.h:64-77: the world shader for the sampled direction
     * (shader_setup_from_background with ray.D = ls->D) */
    const f3 rayD = ls->D;
    emission_sd.P = rayD;
    emission_sd.N = -rayD;
    emission_sd.Ng = -rayD;
    emission_sd.I = -rayD;
    emission_sd.shader = kd_int(KD_BG_SURFACE_SHADER);
    emission_sd.flag = shader_flags(emission_sd.shader);
    emission_sd.object_flag = 0;
    emission_sd.ray_length = 0.0f;
    emission_sd.object = -1;
    emission_sd.prim = CY_PRIM_NONE;
    emission_sd.lamp = -1;
    emission_sd.type = 0;
    emission_sd.u = emission_sd.v = 0.0f;
    emission_sd.dPdu = zero3();
    shader_eval_emission<EXT>(emission_sd, depths, CY_PATH_RAY_EMISSION);
    if (emission_sd.flag & CY_SD_EMISSION)
      eval = emission_sd.closure_emission_background;
  }
  else {
    /* shader_setup_from_sample (kernel_shader.h:244-345): a lamp, or a point on an
     * emissive triangle */
    emission_sd.P = ls->P;
    emission_sd.N = ls->Ng;
    emission_sd.Ng = ls->Ng;
    emission_sd.I = I;
    emission_sd.shader = ls->shader;
    emission_sd.type = (ls->prim != CY_PRIM_NONE) ? (int)CY_PRIMITIVE_TRIANGLE : 0;
    emission_sd.object = ls->object;
    emission_sd.prim = ls->prim;
    emission_sd.lamp = (ls->prim != CY_PRIM_NONE) ? -1 : ls->lamp;
    emission_sd.u = ls->u;
    emission_sd.v = ls->v;
    emission_sd.ray_length = t;
    emission_sd.flag = shader_flags(ls->shader);
    emission_sd.object_flag = 0;
    emission_sd.dPdu = zero3();
    if (ls->prim != CY_PRIM_NONE) {
      emission_sd.object_flag = __ldg(&g_scene.object_flag[ls->object]);
      const bool applied = (emission_sd.object_flag & CY_SD_OBJECT_TRANSFORM_APPLIED) != 0;
      const uint4 tri_vindex = __ldg(&g_scene.tri_vindex[ls->prim]);
      if (ls->shader & CY_SHADER_SMOOTH_NORMAL) {
        /* triangle_smooth_normal - geom/geom_triangle.h:80-92 */
        f3 n0 = mk3(__ldg(&g_scene.tri_vnormal[tri_vindex.x]));
        f3 n1 = mk3(__ldg(&g_scene.tri_vnormal[tri_vindex.y]));
        f3 n2 = mk3(__ldg(&g_scene.tri_vnormal[tri_vindex.z]));
        f3 N = safe_normalize((1.0f - ls->u - ls->v) * n2 + ls->u * n0 + ls->v * n1);
        emission_sd.N = is_zero(N) ? ls->Ng : N;
        if (!applied)
          emission_sd.N = object_normal_transform(ls->object, emission_sd.N);
      }
      emission_sd.dPdu = mk3(__ldg(&g_scene.prim_tri_verts[tri_vindex.w + 0])) -
                         mk3(__ldg(&g_scene.prim_tri_verts[tri_vindex.w + 2]));
      if (!applied)
        emission_sd.dPdu = object_dir_transform(ls->object, emission_sd.dPdu);
      if (dot(emission_sd.Ng, emission_sd.I) < 0.0f) {
        emission_sd.flag |= CY_SD_BACKFACING;
        emission_sd.Ng = -emission_sd.Ng;
        emission_sd.N = -emission_sd.N;
        emission_sd.dPdu = -emission_sd.dPdu;
      }
    }
    ls->Ng = emission_sd.Ng;
    shader_eval_emission<EXT>(emission_sd, depths, CY_PATH_RAY_EMISSION);
    /* shader_emissive_eval: emissive_simple_eval(Ng, I) * weight */
    if (emission_sd.flag & CY_SD_EMISSION) {
      float cosNO = fabsf(dot(emission_sd.Ng, emission_sd.I));
      float res = (cosNO > 0.0f) ? 1.0f : 0.0f;
      eval = mk3(res, res, res) * emission_sd.closure_emission_background;
    }
  }
  eval *= ls->eval_fac;
  if (ls->lamp != CY_LAMP_NONE) {
    eval *= kl_float3(light_ptr(ls->lamp), KL_STRENGTH);
  }
  return eval;
}

/* kernel_emission.h:288-340 */
template<bool EXT>
CY_DEV f3 indirect_background(ShaderDataG &emission_sd, const PathStateG &state, f3 rayD)
{
  int shader = kd_int(KD_BG_SURFACE_SHADER);
  if (shader & CY_SHADER_EXCLUDE_ANY) {
    if (((shader & CY_SHADER_EXCLUDE_DIFFUSE) && (state.flag & CY_PATH_RAY_DIFFUSE)) ||
        ((shader & CY_SHADER_EXCLUDE_GLOSSY) &&
         ((state.flag & (CY_PATH_RAY_GLOSSY | CY_PATH_RAY_REFLECT)) ==
          (CY_PATH_RAY_GLOSSY | CY_PATH_RAY_REFLECT))) ||
        ((shader & CY_SHADER_EXCLUDE_TRANSMIT) && (state.flag & CY_PATH_RAY_TRANSMIT)) ||
        ((shader & CY_SHADER_EXCLUDE_CAMERA) && (state.flag & CY_PATH_RAY_CAMERA)) ||
        ((shader & CY_SHADER_EXCLUDE_SCATTER) && (state.flag & CY_PATH_RAY_VOLUME_SCATTER)))
      return zero3();
  }
  f3 L = zero3();
  if (!shader_constant_emission_eval(shader, &L)) {
    /* shader_setup_from_background (kernel_shader.h:395-430) */
    emission_sd.P = rayD;
    emission_sd.N = -rayD;
    emission_sd.Ng = -rayD;
    emission_sd.I = -rayD;
    emission_sd.shader = shader;
    emission_sd.flag = shader_flags(shader);
    emission_sd.object_flag = 0;
    emission_sd.ray_length = 0.0f;
    emission_sd.object = -1;
    emission_sd.prim = CY_PRIM_NONE;
    emission_sd.lamp = -1;
    emission_sd.type = 0;
    emission_sd.u = emission_sd.v = 0.0f;
    emission_sd.dPdu = zero3();
    shader_eval_emission<EXT>(emission_sd, path_depths(state), state.flag | CY_PATH_RAY_EMISSION);
    if (emission_sd.flag & CY_SD_EMISSION)
      L = emission_sd.closure_emission_background;
  }
  /* the world is also in the light distribution: weight the BSDF-sampled hit against the
   * pdf light sampling would have had for this direction (kernel_emission.h:325-337) */
  if (!(state.flag & CY_PATH_RAY_MIS_SKIP) && kd_int(KD_BG_USE_MIS)) {
    const float pdf = background_light_pdf(rayD);
    return L * power_heuristic(state.ray_pdf, pdf);
  }
  return L;
}

/* kernel_accumulate.h:279-290 */
CY_DEV void path_radiance_clamp(f3 *L, int bounce)
{
  float limit = (bounce > 0) ? kd_float(KD_INT_SAMPLE_CLAMP_INDIRECT) :
                               kd_float(KD_INT_SAMPLE_CLAMP_DIRECT);
  float sum = reduce_add(fabs3(*L));
  if (sum > limit)
    *L *= limit / sum;
}

#endif /* B200_SHADE_CUH */

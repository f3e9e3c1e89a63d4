/* cymath.cuh - scalar float3/float4 arithmetic, in two flavours.
 *
 * The plain operators (dot, cross, normalize ...) keep the operation ORDER of the
 * reference's generic (non-SSE) CPU math (intern/cycles/util/util_math_float3.h
 * :113-160,233-270,353-390, util_math_float4.h:243-250, util_transform.h:56-108); how
 * each operation rounds is the build's choice (Makefile FPFLAGS).  The shipped build
 * compiles them without FMA contraction and with IEEE division / square root, so every
 * expression rounds like the oracle built with -ffp-contract=off.
 *
 * The x-prefixed functions (xdot, xcross, xnormalize_len, xtransform_point ...) are EXACT
 * under any flags: every operation is an IEEE round-to-nearest intrinsic that the compiler
 * neither fuses nor approximates, in the reference's association order.  Everything a hit
 * id depends on goes through them - the ray / triangle test, the instance transforms of
 * the traversal - so u, v, t and the accept / reject decisions are the oracle's bits even
 * in a build that trades shading precision for speed.  Fused multiply-adds appear only
 * where written explicitly (fmaf): the BVH8 slab test, conservative by construction.
 */
#ifndef B200_CYMATH_CUH
#define B200_CYMATH_CUH

#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>

#ifndef CY_DEV /* tests/host_check compiles the node files for the host */
#define CY_DEV __device__ __forceinline__
#endif

#define CY_PI_F 3.1415926535897932f
#define CY_2PI_F 6.2831853071795864f
#define CY_1_PI_F 0.3183098861837067f
#define CY_1_2PI_F 0.1591549430918953f
#define CY_PI_2_F 1.5707963267948966f
#define CY_PI_4_F 0.7853981633974830f

struct f3 {
  float x, y, z;
};

#define CY_M_PI_F 3.1415926535897932f
#define CY_M_PI_2_F 1.5707963267948966f
#define CY_M_2PI_F 6.2831853071795864f
#define CY_M_4PI_F 12.566370614359172f
#define CY_M_1_PI_F 0.3183098861837067f
#define CY_M_PI_4_F 0.7853981633974483f

CY_DEV f3 mk3(float x, float y, float z)
{
  f3 r;
  r.x = x;
  r.y = y;
  r.z = z;
  return r;
}
CY_DEV f3 mk3(float4 a)
{
  return mk3(a.x, a.y, a.z);
}
CY_DEV f3 zero3()
{
  return mk3(0.0f, 0.0f, 0.0f);
}
CY_DEV f3 one3()
{
  return mk3(1.0f, 1.0f, 1.0f);
}
CY_DEV f3 operator+(f3 a, f3 b)
{
  return mk3(a.x + b.x, a.y + b.y, a.z + b.z);
}
CY_DEV f3 operator-(f3 a, f3 b)
{
  return mk3(a.x - b.x, a.y - b.y, a.z - b.z);
}
CY_DEV f3 operator-(f3 a)
{
  return mk3(-a.x, -a.y, -a.z);
}
CY_DEV f3 operator*(f3 a, f3 b)
{
  return mk3(a.x * b.x, a.y * b.y, a.z * b.z);
}
CY_DEV f3 operator*(f3 a, float f)
{
  return mk3(a.x * f, a.y * f, a.z * f);
}
CY_DEV f3 operator*(float f, f3 a)
{
  return mk3(a.x * f, a.y * f, a.z * f);
}
/* util_math_float3.h:140-144: division by a scalar is a multiply by the reciprocal */
CY_DEV f3 operator/(f3 a, float f)
{
  float invf = 1.0f / f;
  return a * invf;
}
CY_DEV f3 operator/(f3 a, f3 b)
{
  return mk3(a.x / b.x, a.y / b.y, a.z / b.z);
}
CY_DEV f3 &operator+=(f3 &a, f3 b)
{
  a = a + b;
  return a;
}
CY_DEV f3 &operator-=(f3 &a, f3 b)
{
  a = a - b;
  return a;
}
CY_DEV f3 &operator*=(f3 &a, f3 b)
{
  a = a * b;
  return a;
}
CY_DEV f3 &operator*=(f3 &a, float f)
{
  a = a * f;
  return a;
}
CY_DEV f3 &operator/=(f3 &a, float f)
{
  float invf = 1.0f / f;
  a = a * invf;
  return a;
}
CY_DEV float dot(f3 a, f3 b)
{
  return a.x * b.x + a.y * b.y + a.z * b.z;
}
CY_DEV float dot4(float4 a, float4 b)
{
  return (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w);
}
CY_DEV f3 cross(f3 a, f3 b)
{
  return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
CY_DEV float len(f3 a)
{
  return sqrtf(dot(a, a));
}
CY_DEV float len_squared(f3 a)
{
  return dot(a, a);
}
CY_DEV f3 normalize(f3 a)
{
  return a / len(a);
}
CY_DEV f3 normalize_len(f3 a, float *t)
{
  *t = len(a);
  float x = 1.0f / *t;
  return a * x;
}
CY_DEV f3 safe_normalize(f3 a)
{
  float t = len(a);
  return (t != 0.0f) ? a * (1.0f / t) : a;
}
CY_DEV f3 fabs3(f3 a)
{
  return mk3(fabsf(a.x), fabsf(a.y), fabsf(a.z));
}
CY_DEV float max3(f3 a)
{
  return fmaxf(fmaxf(a.x, a.y), a.z);
}
CY_DEV float reduce_add(f3 a)
{
  return (a.x + a.y + a.z);
}
CY_DEV float average(f3 a)
{
  return reduce_add(a) * (1.0f / 3.0f);
}
CY_DEV bool is_zero(f3 a)
{
  return (a.x == 0.0f && a.y == 0.0f && a.z == 0.0f);
}
CY_DEV bool isequal3(f3 a, f3 b)
{
  return a.x == b.x && a.y == b.y && a.z == b.z;
}
CY_DEV float saturate(float a)
{
  return fminf(fmaxf(a, 0.0f), 1.0f);
}
CY_DEV float clampf(float a, float mn, float mx)
{
  return fminf(fmaxf(a, mn), mx);
}
CY_DEV float sqr(float a)
{
  return a * a;
}
CY_DEV float safe_sqrtf(float f)
{
  return sqrtf(fmaxf(f, 0.0f));
}
CY_DEV float safe_acosf(float a)
{
  return acosf(clampf(a, -1.0f, 1.0f));
}
CY_DEV float safe_divide(float a, float b)
{
  return (b != 0.0f) ? a / b : 0.0f;
}
CY_DEV bool isfinite_safe(float f)
{
  /* util_math.h isfinite_safe: exponent bits not all ones */
  unsigned int x = __float_as_uint(f);
  return (f == f) && (x == 0 || x == (1u << 31) || (f != 2.0f * f)) && !((x << 1) > 0xff000000u);
}
CY_DEV float xor_signmask(float x, int y)
{
  return __int_as_float(__float_as_int(x) ^ y);
}

/* 3x4 transform, rows x,y,z (util_transform.h Transform) */
struct tfm34 {
  float4 x, y, z;
};
CY_DEV f3 transform_point(const tfm34 &t, f3 a)
{
  return mk3(a.x * t.x.x + a.y * t.x.y + a.z * t.x.z + t.x.w,
             a.x * t.y.x + a.y * t.y.y + a.z * t.y.z + t.y.w,
             a.x * t.z.x + a.y * t.z.y + a.z * t.z.z + t.z.w);
}
CY_DEV f3 transform_direction(const tfm34 &t, f3 a)
{
  return mk3(a.x * t.x.x + a.y * t.x.y + a.z * t.x.z, a.x * t.y.x + a.y * t.y.y + a.z * t.y.z,
             a.x * t.z.x + a.y * t.z.y + a.z * t.z.z);
}
CY_DEV f3 transform_direction_transposed(const tfm34 &t, f3 a)
{
  f3 x = mk3(t.x.x, t.y.x, t.z.x);
  f3 y = mk3(t.x.y, t.y.y, t.z.y);
  f3 z = mk3(t.x.z, t.y.z, t.z.z);
  return mk3(dot(x, a), dot(y, a), dot(z, a));
}

/* util_math.h:477-499 */
CY_DEV void make_orthonormals(f3 N, f3 *a, f3 *b)
{
  if (N.x != N.y || N.x != N.z)
    *a = mk3(N.z - N.y, N.x - N.z, N.y - N.x);
  else
    *a = mk3(N.z - N.y, N.x + N.z, -N.y - N.x);
  *a = normalize(*a);
  *b = cross(N, *a);
}

/* ----------------------------------------------------------- exact flavour */

#ifdef __CUDA_ARCH__
CY_DEV float xmul(float a, float b)
{
  return __fmul_rn(a, b);
}
CY_DEV float xadd(float a, float b)
{
  return __fadd_rn(a, b);
}
CY_DEV float xsub(float a, float b)
{
  return __fsub_rn(a, b);
}
CY_DEV float xdiv(float a, float b)
{
  return __fdiv_rn(a, b);
}
CY_DEV float xsqrt(float a)
{
  return __fsqrt_rn(a);
}
#else /* host builds are compiled with -ffp-contract=off */
CY_DEV float xmul(float a, float b)
{
  return a * b;
}
CY_DEV float xadd(float a, float b)
{
  return a + b;
}
CY_DEV float xsub(float a, float b)
{
  return a - b;
}
CY_DEV float xdiv(float a, float b)
{
  return a / b;
}
CY_DEV float xsqrt(float a)
{
  return sqrtf(a);
}
#endif

CY_DEV f3 xadd3(f3 a, f3 b)
{
  return mk3(xadd(a.x, b.x), xadd(a.y, b.y), xadd(a.z, b.z));
}
CY_DEV f3 xsub3(f3 a, f3 b)
{
  return mk3(xsub(a.x, b.x), xsub(a.y, b.y), xsub(a.z, b.z));
}
CY_DEV f3 xscale3(f3 a, float f)
{
  return mk3(xmul(a.x, f), xmul(a.y, f), xmul(a.z, f));
}
CY_DEV float xdot(f3 a, f3 b)
{
  return xadd(xadd(xmul(a.x, b.x), xmul(a.y, b.y)), xmul(a.z, b.z));
}
CY_DEV f3 xcross(f3 a, f3 b)
{
  return mk3(xsub(xmul(a.y, b.z), xmul(a.z, b.y)), xsub(xmul(a.z, b.x), xmul(a.x, b.z)),
             xsub(xmul(a.x, b.y), xmul(a.y, b.x)));
}
/* normalize_len: length, then a multiply by its reciprocal */
CY_DEV f3 xnormalize_len(f3 a, float *t)
{
  *t = xsqrt(xdot(a, a));
  return xscale3(a, xdiv(1.0f, *t));
}
CY_DEV f3 xtransform_point(const tfm34 &t, f3 a)
{
  return mk3(xadd(xadd(xadd(xmul(a.x, t.x.x), xmul(a.y, t.x.y)), xmul(a.z, t.x.z)), t.x.w),
             xadd(xadd(xadd(xmul(a.x, t.y.x), xmul(a.y, t.y.y)), xmul(a.z, t.y.z)), t.y.w),
             xadd(xadd(xadd(xmul(a.x, t.z.x), xmul(a.y, t.z.y)), xmul(a.z, t.z.z)), t.z.w));
}
CY_DEV f3 xtransform_direction(const tfm34 &t, f3 a)
{
  return mk3(xadd(xadd(xmul(a.x, t.x.x), xmul(a.y, t.x.y)), xmul(a.z, t.x.z)),
             xadd(xadd(xmul(a.x, t.y.x), xmul(a.y, t.y.y)), xmul(a.z, t.y.z)),
             xadd(xadd(xmul(a.x, t.z.x), xmul(a.y, t.z.y)), xmul(a.z, t.z.z)));
}

#endif /* B200_CYMATH_CUH */

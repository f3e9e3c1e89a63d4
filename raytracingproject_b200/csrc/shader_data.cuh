/* shader_data.cuh - the shading point record (ShaderData, kernel_types.h:1011-1110, cut
 * to what the supported closures and nodes read), the SVM stack
 * accessors (svm/svm.h:53-113) and the small object / projection transforms the node
 * implementations share.  Free of warp intrinsics, so that the node files
 * (svm_nodes.cuh, svm_tex.cuh) also compile for the host in tests/test_svm_host_cpu.py. */
#ifndef B200_SHADER_DATA_CUH
#define B200_SHADER_DATA_CUH

#include "device_scene.cuh"

#define CLOSURE_WEIGHT_CUTOFF 1e-5f
#ifndef MAX_CLOSURES_GPU
#  define MAX_CLOSURES_GPU 32 /* two to four mixed Principled BSDFs (8 each, graph.cpp:1147) */
#endif
#define SVM_STACK_GPU 256 /* SVM_STACK_SIZE 255 (svm_types.h): any offset the compiler can emit is in range */

/* util/util_projection.h:48-55 */
CY_DEV f3 transform_perspective(const float4 tx, const float4 ty, const float4 tz,
                                const float4 tw, f3 a)
{
  float4 b = make_float4(a.x, a.y, a.z, 1.0f);
  f3 c = mk3(dot4(tx, b), dot4(ty, b), dot4(tz, b));
  float w = dot4(tw, b);
  return (w != 0.0f) ? c / w : zero3();
}

/* ----------------------------------------------------------- ShaderData */

/* The shading point.  Its BSDF lobes are NOT part of it: they live in the shared-memory
 * arena of lobes.cuh, so this record is small enough to stay in registers and an
 * emission-only evaluation (a lamp or the background) costs no closure storage at all. */
struct ShaderDataG {
  f3 P, N, Ng, I;
  f3 dPdu;
  int shader;
  uint32_t flag, object_flag;
  int prim, type, object;
  int lamp; /* lamp index while its emission shader runs, else -1 (LAMP_NONE) */
  float u, v, ray_length;
  float terminator_freq; /* KernelObject::shadow_terminator_offset, fetched once per point */
  int terminator_terms;  /* some lobe needs the factors of bsdf_terminator_terms_setup() */
  f3 svm_closure_weight;
  f3 closure_emission_background;
  f3 closure_transparent_extinction; /* valid when flag & SD_TRANSPARENT */
  int transparent_at;  /* arena offset of the merged transparent lobe, -1 = none yet */
  uint32_t lcg_state;  /* random walk of the multi-scatter lobes (SD_BSDF_NEEDS_LCG) */
};

CY_DEV uint32_t shader_flags(int shader)
{
  return __ldg((const uint32_t *)(g_scene.shaders +
                                  (size_t)(shader & CY_SHADER_MASK) * SIZEOF_KERNEL_SHADER +
                                  KS_FLAGS));
}

/* KernelObject::shadow_terminator_offset (1 / (1 - offset), 1 = off) */
CY_DEV float object_shadow_terminator_offset(int object)
{
  return __ldg((const float *)(g_scene.objects + (size_t)object * SIZEOF_KERNEL_OBJECT +
                               KO_SHADOW_TERMINATOR_OFFSET));
}

/* geom/geom_object.h:166-186 */
CY_DEV f3 object_normal_transform(int object, f3 N)
{
  tfm34 itfm = object_itfm(object);
  return normalize(transform_direction_transposed(itfm, N));
}
CY_DEV f3 object_dir_transform(int object, f3 D)
{
  tfm34 tfm = object_tfm(object);
  return transform_direction(tfm, D);
}

/* bsdf_util.h:104-117 */
CY_DEV float fresnel_dielectric_cos(float cosi, float eta)
{
  float c = fabsf(cosi);
  float g = eta * eta - 1 + c * c;
  if (g > 0) {
    g = sqrtf(g);
    float A = (g - c) / (g + c);
    float B = (c * (g + c) - 1) / (c * (g - c) + 1);
    return 0.5f * A * A * (1 + B * B);
  }
  return 1.0f;
}

/* the bounce counters the Light Path node reads, passed BY VALUE into the (out-of-line)
 * SVM interpreter: handing it a pointer to the PathStateG would force the whole state of
 * k_shade_surface out of registers */
struct PathDepths {
  short bounce, diffuse, glossy, transparent, transmission;
};
/* util/util_math_fast.h:278-293 (used by the mesh-light pdf and the shadow terminator) */
CY_DEV float fast_acosf(float x)
{
  const float f = fabsf(x);
  const float m = (f < 1.0f) ? 1.0f - (1.0f - f) : 1.0f; /* clamp, crush denormals */
  const float a = sqrtf(1.0f - m) *
                  (1.5707963267f + m * (-0.213300989f + m * (0.077980478f + m * -0.02164095f)));
  return x < 0 ? CY_M_PI_F - a : a;
}

/* ------------------------------------------------------------ emission */

CY_DEV void emission_setup(ShaderDataG &sd, f3 weight)
{
  if (sd.flag & CY_SD_EMISSION) {
    sd.closure_emission_background += weight;
  }
  else {
    sd.flag |= CY_SD_EMISSION;
    sd.closure_emission_background = weight;
  }
}

/* ------------------------------------------------------------ SVM stack */

CY_DEV f3 stack_load_float3(const float *stack, uint32_t a)
{
  return mk3(stack[a + 0], stack[a + 1], stack[a + 2]);
}
CY_DEV void stack_store_float3(float *stack, uint32_t a, f3 f)
{
  stack[a + 0] = f.x;
  stack[a + 1] = f.y;
  stack[a + 2] = f.z;
}
CY_DEV bool stack_valid(uint32_t a)
{
  return a != (uint32_t)CY_SVM_STACK_INVALID;
}

#endif /* B200_SHADER_DATA_CUH */

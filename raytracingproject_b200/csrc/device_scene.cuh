/* device_scene.cuh - what the kernels see of the scene: the BVH8 built by
 * bvh8_build.cpp, the reference's own flat arrays bound by name
 * (kernel/kernel_textures.h:21-87) and a verbatim copy of KernelData ("__data",
 * render/scene.cpp:307) addressed through the generated offsets of
 * include/cycles_abi.h.  One instance per CUDA device, in __constant__ memory
 * (uniform across a warp -> constant-cache broadcast). */
#ifndef B200_DEVICE_SCENE_CUH
#define B200_DEVICE_SCENE_CUH

#include <stdint.h>

#include "../../include/cycles_abi.h"
#include "cymath.cuh"

struct DeviceScene {
  /* BVH8 */
  const uint4 *nodes;    /* 5 x uint4 per node */
  const float4 *records; /* 3 x float4 per leaf record */
  uint32_t bvh_root;
  uint32_t pad0;

  /* reference arrays (device pointers), named after kernel_textures.h */
  const float4 *prim_tri_verts;
  const uint32_t *prim_tri_index;
  const uint32_t *prim_index;
  const uint32_t *prim_object;
  const uint8_t *objects; /* KernelObject[], SIZEOF_KERNEL_OBJECT each */
  const uint32_t *object_flag;
  const uint32_t *tri_shader;
  const float4 *tri_vnormal;
  const uint4 *tri_vindex;
  const uint8_t *lights;             /* KernelLight[] */
  const uint8_t *light_distribution; /* KernelLightDistribution[] */
  const uint8_t *shaders;            /* KernelShader[] */
  const uint4 *svm_nodes;
  const float *lookup_table;
  const uint32_t *sample_pattern_lut;
  const uint4 *attributes_map;
  const float *attributes_float;
  const float2 *attributes_float2;
  const float4 *attributes_float3;
  const uchar4 *attributes_uchar4;
  /* world importance map (light_background.cuh): float2 (function value, CDF) entries */
  const float2 *light_background_marginal_cdf;
  const float2 *light_background_conditional_cdf;
  const uint8_t *texture_info; /* TextureInfo[], SIZEOF_TEXTURE_INFO each, data = device ptr */
  uint32_t num_textures;
  uint32_t pad1;

  /* KernelData, byte-for-byte */
  alignas(16) uint8_t kdata[SIZEOF_KERNEL_DATA];
};

/* All device code is one translation unit (b200_cycles.cu), so the symbol is
 * defined here rather than declared extern (no -rdc needed). */
__constant__ DeviceScene g_scene;

/* KernelData field access by generated offset */
CY_DEV int kd_int(int off)
{
  return *(const int *)(g_scene.kdata + off);
}
CY_DEV float kd_float(int off)
{
  return *(const float *)(g_scene.kdata + off);
}
CY_DEV float4 kd_float4(int off)
{
  return *(const float4 *)(g_scene.kdata + off);
}

/* KernelObject::itfm / tfm (kernel_types.h:1460-1490; object_fetch_transform,
 * geom/geom_object.h:36-47) */
CY_DEV tfm34 object_itfm(int object)
{
  const float4 *p = (const float4 *)(g_scene.objects + (size_t)object * SIZEOF_KERNEL_OBJECT +
                                     KO_ITFM);
  tfm34 t;
  t.x = __ldg(p + 0);
  t.y = __ldg(p + 1);
  t.z = __ldg(p + 2);
  return t;
}
CY_DEV tfm34 object_tfm(int object)
{
  const float4 *p = (const float4 *)(g_scene.objects + (size_t)object * SIZEOF_KERNEL_OBJECT +
                                     KO_TFM);
  tfm34 t;
  t.x = __ldg(p + 0);
  t.y = __ldg(p + 1);
  t.z = __ldg(p + 2);
  return t;
}

#endif /* B200_DEVICE_SCENE_CUH */

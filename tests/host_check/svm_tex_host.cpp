/* Host build of the device-side SVM texture / attribute nodes (csrc/svm_tex.cuh) for
 * tests/test_svm_host_cpu.py: the SAME source the CUDA kernels include, compiled by g++
 * with the handful of CUDA built-ins it uses mapped to their host meaning.  The test
 * feeds identical node words, stacks and shading points to this library and to the
 * reference's own svm_node_* functions (oracle/ref_probe.cpp) and compares the stacks.
 * TEST INFRASTRUCTURE - never loaded by the product path. */
#include <cuda_runtime.h> /* vector types and make_float4 & co: plain host headers */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define CY_DEV static inline
#define SVM_TEX_FN static inline
#define SVM_NODES_INLINE 1
#undef __device__
#define __device__
#undef __noinline__
#define __noinline__
#undef __forceinline__
#define __forceinline__ inline
#undef __constant__
#define __constant__
template<class T> static inline T __ldg(const T *p) { return *p; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }

#include "../../raytracingproject_b200/csrc/shader_data.cuh"
#include "../../raytracingproject_b200/csrc/bsdf.cuh"
#include "../../raytracingproject_b200/csrc/svm_closure.cuh"
#include "../../raytracingproject_b200/csrc/svm_nodes.cuh"
#include "../../raytracingproject_b200/csrc/svm_tex.cuh"
#include "../../raytracingproject_b200/csrc/svm_image.cuh"

struct HostShadingPoint {
  float P[3], N[3], I[3], dPdu[3];
  float u, v;
  int object, prim, lamp, shader, backfacing;
};

struct HostSceneArrays {
  const void *svm_nodes, *objects, *tri_vindex, *lights, *shaders, *attributes_map, *attributes_float,
      *attributes_float2, *attributes_float3, *attributes_uchar4, *kernel_data;
  /* image textures: TextureInfo records whose `data` are host addresses */
  const void *texture_info;
  uint64_t num_textures;
};

extern "C" __attribute__((visibility("default"))) void host_svm_bind(const HostSceneArrays *a)
{
  memset(&g_scene, 0, sizeof(g_scene));
  g_scene.svm_nodes = (const uint4 *)a->svm_nodes;
  g_scene.objects = (const uint8_t *)a->objects;
  g_scene.tri_vindex = (const uint4 *)a->tri_vindex;
  g_scene.lights = (const uint8_t *)a->lights;
  g_scene.shaders = (const uint8_t *)a->shaders;
  g_scene.attributes_map = (const uint4 *)a->attributes_map;
  g_scene.attributes_float = (const float *)a->attributes_float;
  g_scene.attributes_float2 = (const float2 *)a->attributes_float2;
  g_scene.attributes_float3 = (const float4 *)a->attributes_float3;
  g_scene.attributes_uchar4 = (const uchar4 *)a->attributes_uchar4;
  g_scene.texture_info = (const uint8_t *)a->texture_info;
  g_scene.num_textures = (uint32_t)a->num_textures;
  if (a->kernel_data)
    memcpy(g_scene.kdata, a->kernel_data, SIZEOF_KERNEL_DATA);
}

/* Runs the one node at `offset` of the bound program; returns the offset after it, or
 * -1 for an opcode svm_tex.cuh does not implement. */
extern "C" __attribute__((visibility("default"))) int host_svm_node(int offset, float *stack, const HostShadingPoint *p)
{
  ShaderDataG sd;
  memset(&sd, 0, sizeof(sd));
  sd.P = mk3(p->P[0], p->P[1], p->P[2]);
  sd.N = mk3(p->N[0], p->N[1], p->N[2]);
  sd.Ng = sd.N;
  sd.I = mk3(p->I[0], p->I[1], p->I[2]);
  sd.dPdu = mk3(p->dPdu[0], p->dPdu[1], p->dPdu[2]);
  sd.u = p->u;
  sd.v = p->v;
  sd.object = p->object;
  sd.prim = p->prim;
  sd.lamp = p->lamp;
  sd.shader = p->shader;
  sd.flag = p->backfacing ? CY_SD_BACKFACING : 0;
  sd.type = (p->prim != -1) ? (int)CY_PRIMITIVE_TRIANGLE : 0;
  const uint4 node = g_scene.svm_nodes[offset];
  offset++;
  switch (node.x) {
    case CY_NODE_ATTR:
      svm_node_attr(sd, stack, node);
      break;
    case CY_NODE_GEOMETRY: /* only the tangent lives in svm_tex.cuh */
      if (node.y != 2)
        return -1;
      stack_store_float3(stack, node.z, primitive_tangent(sd));
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
    case CY_NODE_TEX_WHITE_NOISE:
      svm_node_tex_white_noise(stack, node);
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
    case CY_NODE_OBJECT_INFO:
      svm_node_object_info(sd, stack, node);
      break;
    case CY_NODE_CAMERA:
      svm_node_camera(sd, stack, node);
      break;
    case CY_NODE_VECTOR_TRANSFORM:
      svm_node_vector_transform(sd, stack, node);
      break;
    case CY_NODE_VECTOR_ROTATE:
      svm_node_vector_rotate(stack, node);
      break;
    case CY_NODE_NORMAL:
      svm_node_normal(stack, node, &offset);
      break;
    case CY_NODE_MAP_RANGE:
      svm_node_map_range(stack, node, &offset);
      break;
    case CY_NODE_HSV:
      svm_node_hsv(stack, node);
      break;
    case CY_NODE_SEPARATE_HSV:
      svm_node_separate_hsv(stack, node, &offset);
      break;
    case CY_NODE_COMBINE_HSV:
      svm_node_combine_hsv(stack, node, &offset);
      break;
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
    case CY_NODE_RGB_RAMP:
      svm_node_rgb_ramp(stack, node, &offset);
      break;
    case CY_NODE_RGB_CURVES:
    case CY_NODE_VECTOR_CURVES:
      svm_node_curves(stack, node, &offset);
      break;
    case CY_NODE_GAMMA:
      svm_node_gamma(stack, node);
      break;
    case CY_NODE_BRIGHTCONTRAST:
      svm_node_brightness(stack, node);
      break;
    case CY_NODE_INVERT:
      svm_node_invert(stack, node);
      break;
    case CY_NODE_MIX:
      svm_node_mix(stack, node, &offset);
      break;
    case CY_NODE_CLAMP:
      svm_node_clamp(stack, node, &offset);
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

/* NODE_CLOSURE_BSDF at `offset` -> lobes in an arena (csrc/svm_closure.cuh, the full
 * variant), then bsdf_eval for `omega_in` and bsdf_sample for (randu, randv) on each of
 * them (csrc/bsdf.cuh, microfacet.cuh, microfacet_multi.cuh).  out: [0] = number of lobes,
 * then 20 floats per lobe: type, weight xyz, sample_weight, eval xyz, pdf, label, sampled
 * eval xyz, omega xyz, pdf.  The LCG of the multi-scatter lobes starts at 0 and runs on
 * from lobe to lobe, like the reference probe's zeroed ShaderData. */
extern "C" __attribute__((visibility("default"))) int host_svm_closure(
    int offset, float *stack, const HostShadingPoint *p, const float *closure_weight,
    unsigned int path_flag, const float *omega_in, float randu, float randv, float *out)
{
  static ShaderDataG sd;
  static float4 arena_words[ARENA_QUADS];
  memset(&sd, 0, sizeof(sd));
  sd.P = mk3(p->P[0], p->P[1], p->P[2]);
  sd.N = mk3(p->N[0], p->N[1], p->N[2]);
  sd.Ng = sd.N;
  sd.I = mk3(p->I[0], p->I[1], p->I[2]);
  sd.flag = p->backfacing ? CY_SD_BACKFACING : 0;
  sd.object = p->object;
  sd.prim = p->prim;
  sd.lamp = p->lamp;
  sd.terminator_freq = object_shadow_terminator_offset(sd.object);
  sd.transparent_at = -1;
  sd.lcg_state = 0;
  LobeArena arena;
  arena.q = arena_words;
  arena_reset(arena, MAX_CLOSURES_GPU);
  sd.svm_closure_weight = mk3(closure_weight[0], closure_weight[1], closure_weight[2]);
  const uint4 node = g_scene.svm_nodes[offset];
  offset++;
  svm_node_closure_bsdf<true>(sd, arena, stack, node, path_flag, &offset);
  bsdf_terminator_terms_setup(sd, arena);
  out[0] = (float)arena.n;
  const f3 wi = mk3(omega_in[0], omega_in[1], omega_in[2]);
  int at = 0;
  for (int i = 0; i < arena.n; i++) {
    const Lobe l = lobe_fetch(arena, at);
    at += lobe_words(l.kind);
    float *o = out + 1 + 20 * i;
    o[0] = (float)lobe_id(l.kind);
    o[1] = l.weight.x, o[2] = l.weight.y, o[3] = l.weight.z;
    o[4] = l.sample_weight;
    float pdf = 0.0f;
    const f3 ev = bsdf_eval<true>(sd, l, wi, &pdf);
    o[5] = ev.x, o[6] = ev.y, o[7] = ev.z, o[8] = pdf;
    f3 sev = zero3(), swi = zero3();
    float spdf = 0.0f;
    const int label = bsdf_sample<true>(sd, l, randu, randv, &sev, &swi, &spdf);
    o[9] = (float)label;
    o[10] = sev.x, o[11] = sev.y, o[12] = sev.z;
    o[13] = swi.x, o[14] = swi.y, o[15] = swi.z;
    o[16] = spdf;
  }
  return offset;
}

/* Kernel-side probe: instantiates the REFERENCE kernel headers (generic scalar
 * variant, same flags as kernels/cpu/kernel.cpp in oracle/Makefile) and exposes
 * the individual hot-path functions on batches, so tests can compare the CUDA
 * path stage by stage:
 *   - scene_intersect            (kernel/bvh/bvh.h:154-237)      closest hit
 *   - scene_intersect any-hit    (PATH_RAY_SHADOW_OPAQUE, bvh_traversal.h:144-147)
 *   - kernel_path_trace_setup    (kernel/kernel_path_common.h:21-46) camera rays
 *   - the first-bounce light connection (kernel_path_surface.h:214-268) to dump
 *     the shadow rays the reference would trace.
 * TEST INFRASTRUCTURE ONLY - part of oracle/_ref/libcycles_ref.so. */

#if defined(__x86_64__) || defined(_M_X64)
#  define __KERNEL_SSE2__
#endif

#include "kernel/kernel.h"
#define KERNEL_ARCH cpu_probe

// clang-format off
#include "kernel/kernel_compat_cpu.h"
#include "kernel/kernel_math.h"
#include "kernel/kernel_types.h"
#include "kernel/split/kernel_split_data.h"
#include "kernel/kernel_globals.h"
#include "kernel/kernel_color.h"
#include "kernel/kernels/cpu/kernel_cpu_image.h"
#include "kernel/kernel_film.h"
#include "kernel/kernel_path.h"
// clang-format on

#include "ref_probe.h"

CCL_NAMESPACE_BEGIN

static inline void ray_from_probe(const RefProbeRay &in, Ray *ray)
{
  ray->P = make_float3(in.P[0], in.P[1], in.P[2]);
  ray->D = make_float3(in.D[0], in.D[1], in.D[2]);
  ray->t = in.t;
  ray->time = 0.0f;
  ray->dP = differential3_zero();
  ray->dD = differential3_zero();
}

void ref_probe_intersect(KernelGlobals *kg, const RefProbeRay *rays, RefProbeHit *hits, size_t n)
{
  for (size_t i = 0; i < n; i++) {
    Ray ray;
    ray_from_probe(rays[i], &ray);
    Intersection isect;
    isect.t = 0.0f;
    isect.u = isect.v = 0.0f;
    isect.prim = PRIM_NONE;
    isect.object = OBJECT_NONE;
    isect.type = PRIMITIVE_NONE;
    bool hit = (ray.t != 0.0f) && scene_intersect(kg, &ray, rays[i].visibility, &isect);
    RefProbeHit &h = hits[i];
    if (hit) {
      h.t = isect.t;
      h.u = isect.u;
      h.v = isect.v;
      h.prim = isect.prim;
      h.object = isect.object;
      h.type = isect.type;
    }
    else {
      h.t = rays[i].t;
      h.u = h.v = 0.0f;
      h.prim = -1;
      h.object = -1;
      h.type = 0;
    }
  }
}

void ref_probe_camera_rays(
    KernelGlobals *kg, int sample, int x0, int y0, int w, int h, RefProbeRay *rays, uint *rng_hash)
{
  for (int y = 0; y < h; y++) {
    for (int x = 0; x < w; x++) {
      Ray ray;
      uint hash;
      kernel_path_trace_setup(kg, sample, x0 + x, y0 + y, &hash, &ray);
      RefProbeRay &r = rays[(size_t)y * w + x];
      r.P[0] = ray.P.x;
      r.P[1] = ray.P.y;
      r.P[2] = ray.P.z;
      r.t = ray.t;
      r.D[0] = ray.D.x;
      r.D[1] = ray.D.y;
      r.D[2] = ray.D.z;
      r.visibility = PATH_RAY_CAMERA | PATH_RAY_ALL_VISIBILITY; /* refined below */
      if (rng_hash)
        rng_hash[(size_t)y * w + x] = hash;
    }
  }
  /* visibility exactly as path_state_ray_visibility() would compute for the
   * first segment (kernel_path_state.h:190-203). */
  {
    Ray ray;
    uint hash;
    kernel_path_trace_setup(kg, sample, x0, y0, &hash, &ray);
    ShaderDataTinyStorage sd_storage;
    PathState state;
    path_state_init(kg, AS_SHADER_DATA(&sd_storage), &state, hash, sample, &ray);
    uint vis = path_state_ray_visibility(kg, &state);
    for (size_t i = 0; i < (size_t)w * h; i++)
      rays[i].visibility = vis;
  }
}

/* First-bounce shadow rays: run the reference path up to the light connection
 * (kernel_path_integrate, kernel_path.h:509-641, first iteration only) and
 * record the light ray that shadow_blocked() would be given
 * (kernel_path_surface.h:236-262).  Rays that the reference would not trace
 * (miss, no light sample, zero contribution) get t = 0. */
void ref_probe_shadow_rays(
    KernelGlobals *kg, int sample, int x0, int y0, int w, int h, RefProbeRay *rays)
{
  for (int y = 0; y < h; y++) {
    for (int x = 0; x < w; x++) {
      RefProbeRay &out = rays[(size_t)y * w + x];
      memset(&out, 0, sizeof(out));
      out.visibility = PATH_RAY_SHADOW_OPAQUE;

      Ray ray;
      uint rng_hash;
      kernel_path_trace_setup(kg, sample, x0 + x, y0 + y, &rng_hash, &ray);
      if (ray.t == 0.0f)
        continue;

      PathRadiance L;
      path_radiance_init(kg, &L);
      ShaderDataTinyStorage emission_sd_storage;
      ShaderData *emission_sd = AS_SHADER_DATA(&emission_sd_storage);
      PathState state;
      path_state_init(kg, emission_sd, &state, rng_hash, sample, &ray);

      Intersection isect;
      if (!kernel_path_scene_intersect(kg, &state, &ray, &isect, &L))
        continue;

      ShaderData sd;
      shader_setup_from_ray(kg, &sd, &isect, &ray);
      shader_eval_surface(kg, &sd, &state, NULL, state.flag);
      shader_prepare_closures(&sd, &state);

      if (!(kernel_data.integrator.use_direct_light && (sd.flag & SD_BSDF_HAS_EVAL)))
        continue;

      float light_u, light_v;
      path_state_rng_2D(kg, &state, PRNG_LIGHT_U, &light_u, &light_v);

      Ray light_ray;
      BsdfEval L_light;
      bool is_lamp;
      light_ray.time = sd.time;

      LightSample ls;
      if (light_sample(kg, -1, light_u, light_v, sd.time, sd.P, state.bounce, &ls)) {
        float terminate = path_state_rng_light_termination(kg, &state);
        if (direct_emission(
                kg, &sd, emission_sd, &ls, &state, &light_ray, &L_light, &is_lamp, terminate)) {
          out.P[0] = light_ray.P.x;
          out.P[1] = light_ray.P.y;
          out.P[2] = light_ray.P.z;
          out.t = light_ray.t;
          out.D[0] = light_ray.D.x;
          out.D[1] = light_ray.D.y;
          out.D[2] = light_ray.D.z;
        }
      }
    }
  }
}

/* Ray census of the reference path loop: follows kernel_path_integrate
 * (kernel_path.h:509-641, surface-only branches) with the reference's own inline
 * functions and counts scene_intersect calls for camera rays, bounce rays and
 * shadow rays (light rays with t != 0, kernel_shadow.h:398-400).  Used to turn a
 * timed CPU render into Mrays/s without a __KERNEL_DEBUG__ rebuild. */
void ref_probe_count_rays(KernelGlobals *kg, int sample, int x0, int y0, int w, int h,
                          unsigned long long counts[3])
{
  for (int y = 0; y < h; y++) {
    for (int x = 0; x < w; x++) {
      Ray ray;
      uint rng_hash;
      kernel_path_trace_setup(kg, sample, x0 + x, y0 + y, &rng_hash, &ray);
      if (ray.t == 0.0f)
        continue;
      float3 throughput = make_float3(1.0f, 1.0f, 1.0f);
      PathRadiance L;
      path_radiance_init(kg, &L);
      ShaderDataTinyStorage emission_sd_storage;
      ShaderData *emission_sd = AS_SHADER_DATA(&emission_sd_storage);
      PathState state;
      path_state_init(kg, emission_sd, &state, rng_hash, sample, &ray);
      ShaderData sd;
      for (;;) {
        Intersection isect;
        counts[(state.flag & PATH_RAY_CAMERA) ? 0 : 1]++;
        bool hit = kernel_path_scene_intersect(kg, &state, &ray, &isect, &L);
        kernel_path_lamp_emission(kg, &state, &ray, throughput, &isect, &sd, &L);
        if (!hit) {
          kernel_path_background(kg, &state, &ray, throughput, &sd, NULL, &L);
          break;
        }
        else if (path_state_ao_bounce(kg, &state)) {
          break;
        }
        shader_setup_from_ray(kg, &sd, &isect, &ray);
        shader_eval_surface(kg, &sd, &state, NULL, state.flag);
        shader_prepare_closures(&sd, &state);
        if (!kernel_path_shader_apply(kg, &sd, &state, &ray, throughput, emission_sd, &L, NULL))
          break;
        float probability = path_state_continuation_probability(kg, &state, throughput);
        if (probability == 0.0f) {
          break;
        }
        else if (probability != 1.0f) {
          float terminate = path_state_rng_1D(kg, &state, PRNG_TERMINATE);
          if (terminate >= probability)
            break;
          throughput /= probability;
        }
        /* kernel_branched_path_surface_connect_light with one light (kernel_path_surface.h:22-125) */
        if (kernel_data.integrator.use_direct_light && (sd.flag & SD_BSDF_HAS_EVAL)) {
          float light_u, light_v;
          path_state_rng_2D(kg, &state, PRNG_LIGHT_U, &light_u, &light_v);
          float terminate = path_state_rng_light_termination(kg, &state);
          Ray light_ray;
          light_ray.t = 0.0f;
          light_ray.time = sd.time;
          BsdfEval L_light;
          bool is_lamp = false;
          LightSample ls;
          if (light_sample(kg, -1, light_u, light_v, sd.time, sd.P, state.bounce, &ls)) {
            if (direct_emission(
                    kg, &sd, emission_sd, &ls, &state, &light_ray, &L_light, &is_lamp, terminate)) {
              if (light_ray.t != 0.0f)
                counts[2]++;
            }
          }
        }
        if (!kernel_path_surface_bounce(kg, &sd, &throughput, &state, &L.state, &ray))
          break;
      }
    }
  }
}

/* Debugging aid: the same 32 floats per bounce the B200 device records with option
 * "debug_slot" (wavefront.cuh), taken from the reference path loop. */
void ref_probe_path_dump(KernelGlobals *kg, int sample, int x, int y, float *out)
{
  Ray ray;
  uint rng_hash;
  kernel_path_trace_setup(kg, sample, x, y, &rng_hash, &ray);
  if (ray.t == 0.0f)
    return;
  float3 throughput = make_float3(1.0f, 1.0f, 1.0f);
  PathRadiance L;
  path_radiance_init(kg, &L);
  ShaderDataTinyStorage emission_sd_storage;
  ShaderData *emission_sd = AS_SHADER_DATA(&emission_sd_storage);
  PathState state;
  path_state_init(kg, emission_sd, &state, rng_hash, sample, &ray);
  ShaderData sd;
  for (;;) {
    Intersection isect;
    bool hit = kernel_path_scene_intersect(kg, &state, &ray, &isect, &L);
    kernel_path_lamp_emission(kg, &state, &ray, throughput, &isect, &sd, &L);
    if (!hit) {
      kernel_path_background(kg, &state, &ray, throughput, &sd, NULL, &L);
      break;
    }
    else if (path_state_ao_bounce(kg, &state)) {
      break;
    }
    shader_setup_from_ray(kg, &sd, &isect, &ray);
    shader_eval_surface(kg, &sd, &state, NULL, state.flag);
    shader_prepare_closures(&sd, &state);
    if (!kernel_path_shader_apply(kg, &sd, &state, &ray, throughput, emission_sd, &L, NULL))
      break;
    float probability = path_state_continuation_probability(kg, &state, throughput);
    if (probability == 0.0f) {
      break;
    }
    else if (probability != 1.0f) {
      float terminate = path_state_rng_1D(kg, &state, PRNG_TERMINATE);
      if (terminate >= probability)
        break;
      throughput /= probability;
    }
    kernel_path_surface_connect_light(kg, &sd, emission_sd, throughput, &state, &L);
    /* kernel_path_surface_bounce, opened up to record the sample */
    if (!(sd.flag & SD_BSDF))
      break;
    float bsdf_pdf;
    BsdfEval bsdf_eval;
    float3 bsdf_omega_in;
    differential3 bsdf_domega_in;
    float bsdf_u, bsdf_v;
    path_state_rng_2D(kg, &state, PRNG_BSDF_U, &bsdf_u, &bsdf_v);
    Ray in_ray = ray;
    int label = shader_bsdf_sample(
        kg, &sd, bsdf_u, bsdf_v, &bsdf_eval, &bsdf_omega_in, &bsdf_domega_in, &bsdf_pdf);
    if (bsdf_pdf == 0.0f || bsdf_eval_is_zero(&bsdf_eval))
      break;
    path_radiance_bsdf_bounce(kg, &L.state, &throughput, &bsdf_eval, bsdf_pdf, state.bounce, label);
    if (!(label & LABEL_TRANSPARENT)) {
      state.ray_pdf = bsdf_pdf;
      state.ray_t = 0.0f;
      state.min_ray_pdf = fminf(bsdf_pdf, state.min_ray_pdf);
    }
    path_state_next(kg, &state, label);
    ray.P = ray_offset(sd.P, (label & LABEL_TRANSMIT) ? -sd.Ng : sd.Ng);
    ray.D = normalize(bsdf_omega_in);
    if (state.bounce == 0)
      ray.t -= sd.ray_length;
    else
      ray.t = FLT_MAX;
    ray.dP = sd.dP;
    ray.dD = bsdf_domega_in;
    if (state.bounce < 16) {
      float *dbg = out + 32 * (state.bounce + state.transparent_bounce - 1);
      dbg[0] = in_ray.P.x, dbg[1] = in_ray.P.y, dbg[2] = in_ray.P.z, dbg[3] = in_ray.t;
      dbg[4] = in_ray.D.x, dbg[5] = in_ray.D.y, dbg[6] = in_ray.D.z, dbg[7] = isect.t;
      dbg[8] = (float)isect.prim, dbg[9] = (float)isect.object, dbg[10] = sd.P.x;
      dbg[11] = sd.P.y, dbg[12] = sd.P.z, dbg[13] = sd.N.x, dbg[14] = sd.N.y;
      dbg[15] = sd.N.z, dbg[16] = (float)(sd.flag & 0xffff), dbg[17] = (float)sd.num_closure;
      dbg[18] = (float)sd.closure[0].type, dbg[19] = sd.closure[0].sample_weight;
      dbg[20] = (float)sd.closure[1].type, dbg[21] = sd.closure[1].sample_weight;
      dbg[22] = (float)label, dbg[23] = bsdf_pdf, dbg[24] = bsdf_omega_in.x;
      dbg[25] = bsdf_omega_in.y, dbg[26] = bsdf_omega_in.z, dbg[27] = throughput.x;
      dbg[28] = throughput.y, dbg[29] = throughput.z, dbg[30] = bsdf_u, dbg[31] = bsdf_v;
    }
  }
}

void ref_probe_path_trace(
    KernelGlobals *kg, float *buffer, int sample, int x, int y, int offset, int stride)
{
  kernel_path_trace(kg, buffer, sample, x, y, offset, stride);
}

/* Runs ONE node of a caller-supplied SVM program through the reference's own
 * svm_node_* function (the dispatch of svm/svm.h:220-500 for the texture, attribute and
 * mapping opcodes) on a caller-supplied shading point; returns the offset after the
 * node or -1.  kg's scene arrays (objects, attributes, lights, camera) are the bound
 * scene's; only __svm_nodes is swapped for the call. */
int ref_probe_svm_node(KernelGlobals *kg, const void *nodes, int offset, float *stack,
                       const RefShadingPoint *p)
{
  uint4 *saved = kg->__svm_nodes.data;
  kg->__svm_nodes.data = (uint4 *)nodes;
  ShaderData sd_storage;
  ShaderData *sd = &sd_storage;
  memset((void *)sd, 0, sizeof(ShaderData));
  sd->P = make_float3(p->P[0], p->P[1], p->P[2]);
  sd->N = make_float3(p->N[0], p->N[1], p->N[2]);
  sd->Ng = sd->N;
  sd->I = make_float3(p->I[0], p->I[1], p->I[2]);
  sd->dPdu = make_float3(p->dPdu[0], p->dPdu[1], p->dPdu[2]);
  sd->u = p->u;
  sd->v = p->v;
  sd->object = p->object;
  sd->prim = p->prim;
  sd->lamp = p->lamp;
  sd->shader = p->shader;
  sd->flag = p->backfacing ? SD_BACKFACING : 0;
  sd->type = (p->prim != PRIM_NONE) ? PRIMITIVE_TRIANGLE :
                                      ((p->lamp != LAMP_NONE) ? PRIMITIVE_LAMP : PRIMITIVE_NONE);
  if (sd->object != OBJECT_NONE) {
    sd->ob_tfm = object_fetch_transform(kg, sd->object, OBJECT_TRANSFORM);
    sd->ob_itfm = object_fetch_transform(kg, sd->object, OBJECT_INVERSE_TRANSFORM);
  }
  else if (sd->lamp != LAMP_NONE) {
    sd->ob_tfm = lamp_fetch_transform(kg, sd->lamp, false);
    sd->ob_itfm = lamp_fetch_transform(kg, sd->lamp, true);
  }
  const int path_flag = 0;
  uint4 node = read_node(kg, &offset);
  switch (node.x) {
    case NODE_ATTR:
      svm_node_attr(kg, sd, stack, node);
      break;
    case NODE_GEOMETRY:
      svm_node_geometry(kg, sd, stack, node.y, node.z);
      break;
    case NODE_TEX_COORD:
      svm_node_tex_coord(kg, sd, path_flag, stack, node, &offset);
      break;
    case NODE_MAPPING:
      svm_node_mapping(kg, sd, stack, node.y, node.z, node.w, &offset);
      break;
    case NODE_TEXTURE_MAPPING:
      svm_node_texture_mapping(kg, sd, stack, node.y, node.z, &offset);
      break;
    case NODE_MIN_MAX:
      svm_node_min_max(kg, sd, stack, node.y, node.z, &offset);
      break;
    case NODE_TEX_NOISE:
      svm_node_tex_noise(kg, sd, stack, node.y, node.z, node.w, &offset);
      break;
    case NODE_TEX_CHECKER:
      svm_node_tex_checker(kg, sd, stack, node);
      break;
    case NODE_TEX_GRADIENT:
      svm_node_tex_gradient(sd, stack, node);
      break;
    case NODE_TEX_WAVE:
      svm_node_tex_wave(kg, sd, stack, node, &offset);
      break;
    case NODE_TEX_MAGIC:
      svm_node_tex_magic(kg, sd, stack, node, &offset);
      break;
    case NODE_TEX_BRICK:
      svm_node_tex_brick(kg, sd, stack, node, &offset);
      break;
    case NODE_TEX_WHITE_NOISE:
      svm_node_tex_white_noise(kg, sd, stack, node.y, node.z, node.w, &offset);
      break;
    case NODE_TANGENT:
      svm_node_tangent(kg, sd, stack, node);
      break;
    case NODE_NORMAL_MAP:
      svm_node_normal_map(kg, sd, stack, node);
      break;
    case NODE_BLACKBODY:
      svm_node_blackbody(kg, sd, stack, node.y, node.z);
      break;
    case NODE_WAVELENGTH:
      svm_node_wavelength(kg, sd, stack, node.y, node.z);
      break;
    case NODE_TEX_MUSGRAVE:
      svm_node_tex_musgrave(kg, sd, stack, node.y, node.z, node.w, &offset);
      break;
    case NODE_TEX_VORONOI:
      svm_node_tex_voronoi(kg, sd, stack, node.y, node.z, node.w, &offset);
      break;
    case NODE_OBJECT_INFO:
      svm_node_object_info(kg, sd, stack, node.y, node.z);
      break;
    case NODE_CAMERA:
      svm_node_camera(kg, sd, stack, node.y, node.z, node.w);
      break;
    case NODE_VECTOR_TRANSFORM:
      svm_node_vector_transform(kg, sd, stack, node);
      break;
    case NODE_VECTOR_ROTATE:
      svm_node_vector_rotate(sd, stack, node.y, node.z, node.w);
      break;
    case NODE_NORMAL:
      svm_node_normal(kg, sd, stack, node.y, node.z, node.w, &offset);
      break;
    case NODE_MAP_RANGE:
      svm_node_map_range(kg, sd, stack, node.y, node.z, node.w, &offset);
      break;
    case NODE_HSV:
      svm_node_hsv(kg, sd, stack, node, &offset);
      break;
    case NODE_SEPARATE_HSV:
      svm_node_separate_hsv(kg, sd, stack, node.y, node.z, node.w, &offset);
      break;
    case NODE_COMBINE_HSV:
      svm_node_combine_hsv(kg, sd, stack, node.y, node.z, node.w, &offset);
      break;
    case NODE_CONVERT:
      svm_node_convert(kg, sd, stack, node.y, node.z, node.w);
      break;
    case NODE_FRESNEL:
      svm_node_fresnel(sd, stack, node.y, node.z, node.w);
      break;
    case NODE_LAYER_WEIGHT:
      svm_node_layer_weight(sd, stack, node);
      break;
    case NODE_MATH:
      svm_node_math(kg, sd, stack, node.y, node.z, node.w, &offset);
      break;
    case NODE_VECTOR_MATH:
      svm_node_vector_math(kg, sd, stack, node.y, node.z, node.w, &offset);
      break;
    case NODE_RGB_RAMP:
      svm_node_rgb_ramp(kg, sd, stack, node, &offset);
      break;
    case NODE_RGB_CURVES:
    case NODE_VECTOR_CURVES:
      svm_node_curves(kg, sd, stack, node, &offset);
      break;
    case NODE_GAMMA:
      svm_node_gamma(sd, stack, node.y, node.z, node.w);
      break;
    case NODE_BRIGHTCONTRAST:
      svm_node_brightness(sd, stack, node.y, node.z, node.w);
      break;
    case NODE_INVERT:
      svm_node_invert(sd, stack, node.y, node.z, node.w);
      break;
    case NODE_MIX:
      svm_node_mix(kg, sd, stack, node.y, node.z, node.w, &offset);
      break;
    case NODE_CLAMP:
      svm_node_clamp(kg, sd, stack, node.y, node.z, node.w, &offset);
      break;
    case NODE_TEX_IMAGE:
      svm_node_tex_image(kg, sd, stack, node, &offset);
      break;
    case NODE_TEX_IMAGE_BOX:
      svm_node_tex_image_box(kg, sd, stack, node);
      break;
    case NODE_TEX_ENVIRONMENT:
      svm_node_tex_environment(kg, sd, stack, node);
      break;
    default:
      offset = -1;
  }
  kg->__svm_nodes.data = saved;
  return offset;
}

/* NODE_CLOSURE_BSDF at `offset` through the reference's svm_node_closure_bsdf, then the
 * reference's bsdf_eval / bsdf_sample on every closure it made.  Output layout as
 * host_svm_closure (tests/host_check/svm_tex_host.cpp). */
int ref_probe_svm_closure(KernelGlobals *kg, const void *nodes, int offset, float *stack,
                          const RefShadingPoint *p, const float *closure_weight,
                          unsigned int path_flag, const float *omega_in, float randu,
                          float randv, float *out)
{
  uint4 *saved = kg->__svm_nodes.data;
  kg->__svm_nodes.data = (uint4 *)nodes;
  ShaderData sd_storage;
  ShaderData *sd = &sd_storage;
  memset((void *)sd, 0, sizeof(ShaderData));
  sd->P = make_float3(p->P[0], p->P[1], p->P[2]);
  sd->N = make_float3(p->N[0], p->N[1], p->N[2]);
  sd->Ng = sd->N;
  sd->I = make_float3(p->I[0], p->I[1], p->I[2]);
  sd->flag = p->backfacing ? SD_BACKFACING : 0;
  sd->object = p->object;
  sd->prim = p->prim;
  sd->lamp = p->lamp;
  sd->type = PRIMITIVE_TRIANGLE;
  sd->num_closure = 0;
  sd->num_closure_left = 32;
  sd->svm_closure_weight = make_float3(closure_weight[0], closure_weight[1], closure_weight[2]);
  uint4 node = read_node(kg, &offset);
  svm_node_closure_bsdf(kg, sd, stack, node, SHADER_TYPE_SURFACE, (int)path_flag, &offset);
  out[0] = (float)sd->num_closure;
  const float3 wi = make_float3(omega_in[0], omega_in[1], omega_in[2]);
  for (int i = 0; i < sd->num_closure; i++) {
    const ShaderClosure *sc = &sd->closure[i];
    float *o = out + 1 + 20 * i;
    o[0] = (float)sc->type;
    o[1] = sc->weight.x, o[2] = sc->weight.y, o[3] = sc->weight.z;
    o[4] = sc->sample_weight;
    float pdf = 0.0f;
    const float3 ev = bsdf_eval(kg, sd, sc, wi, &pdf);
    o[5] = ev.x, o[6] = ev.y, o[7] = ev.z, o[8] = pdf;
    float3 sev = make_float3(0.0f, 0.0f, 0.0f), swi = make_float3(0.0f, 0.0f, 0.0f);
    differential3 dwi;
    float spdf = 0.0f;
    const int label = bsdf_sample(kg, sd, sc, randu, randv, &sev, &swi, &dwi, &spdf);
    o[9] = (float)label;
    o[10] = sev.x, o[11] = sev.y, o[12] = sev.z;
    o[13] = swi.x, o[14] = swi.y, o[15] = swi.z;
    o[16] = spdf;
  }
  kg->__svm_nodes.data = saved;
  return offset;
}

CCL_NAMESPACE_END



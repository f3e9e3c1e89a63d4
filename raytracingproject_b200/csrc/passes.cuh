/* passes.cuh - render passes next to the combined one: the light passes (direct /
 * indirect / colour per BSDF class, emission, background, shadow, mist) and the data
 * passes (depth, normal, UV, object and material id).
 *
 * Semantics to match (reference = blender/intern/cycles/kernel):
 *   kernel_accumulate.h:25-166    BsdfEval split per BSDF class
 *   kernel_accumulate.h:235-268   path_radiance_bsdf_bounce: the throughput of the FIRST
 *                                 bounce is kept per class (PathRadianceState), later
 *                                 bounces carry one colour
 *   kernel_accumulate.h:302-515   where emission / light / background contributions go,
 *                                 by bounce
 *   kernel_accumulate.h:537-560   path_radiance_sum_indirect: the indirect light is
 *                                 divided by the first bounce's total and re-multiplied by
 *                                 each class
 *   kernel_accumulate.h:640-700   path_radiance_clamp_and_sum: with light passes the
 *                                 combined colour is the sum of the classes
 *   kernel_passes.h:174-389       kernel_write_data_passes, kernel_write_light_passes
 * `film.use_light_pass` is set by any light pass - and by a lamp that is invisible to
 * diffuse / glossy / transmission rays (render/light.cpp:352-364), which zeroes that
 * class of the light's contribution (kernel_emission.h:146-157).
 *
 * On this device a path's accumulators live in a per-path block of PASS_WORDS floats
 * (PathSoA::pass), carved only when the film asks for more than the combined pass; the
 * shading kernels compiled with PASSES = true add to it, k_film_accumulate_passes folds a
 * pixel's samples into the film in sample order like the combined-only kernel.  The
 * emission of directly visible surfaces stays in PathSoA::L (it IS the emission pass). */
#ifndef B200_PASSES_CUH
#define B200_PASSES_CUH

/* per-path block, word offsets */
enum {
  PB_STATE_DIFFUSE = 0,        /* PathRadianceState: first-bounce throughput per class */
  PB_STATE_GLOSSY = 3,
  PB_STATE_TRANSMISSION = 6,
  PB_STATE_DIRECT = 9,
  PB_DIRECT_DIFFUSE = 12,
  PB_DIRECT_GLOSSY = 15,
  PB_DIRECT_TRANSMISSION = 18,
  PB_INDIRECT = 21,
  PB_DIRECT_EMISSION = 24,
  PB_BACKGROUND = 27,
  PB_COLOR_DIFFUSE = 30,
  PB_COLOR_GLOSSY = 33,
  PB_COLOR_TRANSMISSION = 36,
  PB_SHADOW = 39,
  PB_NORMAL = 42,
  PB_UV = 45,
  PB_MIST = 48,
  PB_DEPTH = 49,
  PB_OBJECT_ID = 50,
  PB_MATERIAL_ID = 51,
  PB_HAS_DATA = 52, /* != 0: the path wrote its data passes (PATH_RAY_SINGLE_PASS_DONE) */
  PB_UNTRACED = 53, /* != 0: no camera ray for this pixel sample, nothing is written */
  PB_AO = 56,       /* the ambient-occlusion pass (path_radiance_accum_ao) */
  /* denoising features (film.pass_denoising_data != 0), see denoising_update_features */
  PB_DN_WEIGHT = 60,       /* PathState::denoising_feature_weight */
  PB_DN_THROUGHPUT = 61,   /* PathState::denoising_feature_throughput */
  PB_DN_NORMAL = 64,       /* PathRadiance::denoising_normal / _albedo / _depth */
  PB_DN_ALBEDO = 67,
  PB_DN_DEPTH = 70,
  PB_PATH_TOTAL = 71,        /* light that could have arrived while the feature is open */
  PB_PATH_TOTAL_SHADED = 74, /* light that did (the "shadowing" feature is their ratio) */
  PASS_WORDS = 80
};

CY_DEV bool film_has_denoising()
{
  return kd_int(KD_FILM_PASS_DENOISING_DATA) != 0;
}

CY_DEV f3 pb_get3(const float *pb, int off)
{
  return mk3(pb[off], pb[off + 1], pb[off + 2]);
}
CY_DEV void pb_set3(float *pb, int off, f3 v)
{
  pb[off] = v.x;
  pb[off + 1] = v.y;
  pb[off + 2] = v.z;
}
CY_DEV void pb_add3(float *pb, int off, f3 v)
{
  pb[off] += v.x;
  pb[off + 1] += v.y;
  pb[off + 2] += v.z;
}

/* the light ray and the AO ray of one path may arrive in the same launch */
CY_DEV void pb_atomic_add3(float *pb, int off, f3 v)
{
  atomicAdd(pb + off, v.x);
  atomicAdd(pb + off + 1, v.y);
  atomicAdd(pb + off + 2, v.z);
}

CY_DEV bool light_pass_on(int pass_type) /* PassType 32..63 */
{
  return (kd_int(KD_FILM_LIGHT_PASS_FLAG) & (1 << (pass_type % 32))) != 0;
}
CY_DEV bool data_pass_on(int pass_type) /* PassType < 32 */
{
  return (kd_int(KD_FILM_PASS_FLAG) & (1 << (pass_type % 32))) != 0;
}

/* BsdfEval with use_light_pass: one colour per BSDF class (no volumes on this device) */
struct EvalSplit {
  f3 diffuse, glossy, transmission, transparent;
};

CY_DEV void eval_split_zero(EvalSplit &e)
{
  e.diffuse = e.glossy = e.transmission = e.transparent = zero3();
}
/* bsdf_eval_init / bsdf_eval_accum: which class a closure id belongs to
 * (svm_types.h:587-596) */
CY_DEV void eval_split_add(EvalSplit &e, uint32_t kind, f3 v)
{
  const int id = lobe_id(kind);
  if (id == CY_CLOSURE_BSDF_TRANSPARENT_ID)
    e.transparent += v;
  else if (id <= CY_CLOSURE_BSDF_TRANSLUCENT_ID)
    e.diffuse += v;
  else if (id < CY_CLOSURE_BSDF_REFRACTION_ID)
    e.glossy += v;
  else
    e.transmission += v;
}
/* bsdf_eval_sum: transparent is not part of it */
CY_DEV f3 eval_split_sum(const EvalSplit &e)
{
  return e.diffuse + e.glossy + e.transmission;
}
CY_DEV void eval_split_mul(EvalSplit &e, float f)
{
  e.diffuse *= f;
  e.glossy *= f;
  e.transmission *= f;
}
CY_DEV void eval_split_mul3(EvalSplit &e, f3 f)
{
  e.diffuse *= f;
  e.glossy *= f;
  e.transmission *= f;
}
CY_DEV bool eval_split_is_zero(const EvalSplit &e)
{
  return is_zero(e.diffuse) && is_zero(e.glossy) && is_zero(e.transmission) &&
         is_zero(e.transparent);
}

/* _shader_bsdf_multi_eval with a split result */
template<bool EXT, bool MS>
CY_DEV void shader_bsdf_multi_eval_split(ShaderDataG &sd, const LobeArena &arena, f3 omega_in,
                                         float *pdf, int skip, EvalSplit &result, float sum_pdf,
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
        eval_split_add(result, kind, eval * l.weight);
        sum_pdf += bsdf_pdf * l.sample_weight;
      }
      sum_sample_weight += l.sample_weight;
    }
    at += lobe_words(kind);
  }
  *pdf = (sum_sample_weight > 0.0f) ? sum_pdf / sum_sample_weight : 0.0f;
}

/* shader_bsdf_eval towards a light sample, MIS-weighted */
template<bool EXT, bool MS>
CY_DEV void shader_bsdf_eval_split(ShaderDataG &sd, const LobeArena &arena, f3 omega_in,
                                   float light_pdf, bool use_mis, EvalSplit &eval,
                                   f3 *sum_no_mis = nullptr)
{
  eval_split_zero(eval);
  float pdf;
  shader_bsdf_multi_eval_split<EXT, MS>(sd, arena, omega_in, &pdf, -1, eval, 0.0f, 0.0f);
  if (sum_no_mis)
    *sum_no_mis = eval_split_sum(eval);
  if (use_mis)
    eval_split_mul(eval, power_heuristic(light_pdf, pdf));
}

/* shader_bsdf_pick + shader_bsdf_sample with a split result */
template<bool EXT, bool MS>
CY_DEV int shader_bsdf_sample_split(ShaderDataG &sd, const LobeArena &arena, float randu,
                                    float randv, EvalSplit &out, f3 *omega_in, float *pdf)
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
          randu = (r - partial_sum) / sw;
          break;
        }
        partial_sum = next_sum;
      }
      at += lobe_words(kind);
    }
  }
  *pdf = 0.0f;
  eval_split_zero(out);
  const uint32_t kind = lobe_kind_at(arena, sampled_at);
  if (!lobe_is_bsdf(kind))
    return CY_LABEL_NONE;
  const Lobe l = lobe_fetch(arena, sampled_at);
  f3 eval = zero3();
  const int label = bsdf_sample<EXT, MS>(sd, l, randu, randv, &eval, omega_in, pdf);
  if (*pdf != 0.0f) {
    eval_split_add(out, kind, eval * l.weight);
    if (arena.n > 1) {
      const float sweight = l.sample_weight;
      shader_bsdf_multi_eval_split<EXT, MS>(sd, arena, *omega_in, pdf, sampled, out,
                                            *pdf * sweight, sweight);
    }
  }
  return label;
}

/* path_radiance_accum_emission / _background: which accumulator a contribution of a path
 * at `bounce` goes to - the emission pass itself (PathSoA::L, returned as NULL), the
 * light seen after exactly one bounce, or everything later */
CY_DEV int emission_bucket(int bounce)
{
  return (bounce == 0) ? -1 : (bounce == 1 ? PB_DIRECT_EMISSION : PB_INDIRECT);
}

/* kernel_camera.h:435-460 */
CY_DEV float pass_camera_distance(f3 P)
{
  const float4 cx = kd_float4(KD_CAM_CAMERATOWORLD), cy = kd_float4(KD_CAM_CAMERATOWORLD + 16),
               cz = kd_float4(KD_CAM_CAMERATOWORLD + 32);
  const f3 camP = mk3(cx.w, cy.w, cz.w);
  if (kd_int(KD_CAM_TYPE) == CY_CAMERA_ORTHOGRAPHIC)
    return fabsf(dot(P - camP, mk3(cx.z, cy.z, cz.z)));
  return len(P - camP);
}
CY_DEV float pass_camera_z_depth(f3 P)
{
  if (kd_int(KD_CAM_TYPE) != CY_CAMERA_PANORAMA) {
    tfm34 w2c;
    w2c.x = kd_float4(KD_CAM_WORLDTOCAMERA);
    w2c.y = kd_float4(KD_CAM_WORLDTOCAMERA + 16);
    w2c.z = kd_float4(KD_CAM_WORLDTOCAMERA + 32);
    return transform_point(w2c, P).z;
  }
  const f3 camP = mk3(kd_float4(KD_CAM_CAMERATOWORLD).w, kd_float4(KD_CAM_CAMERATOWORLD + 16).w,
                      kd_float4(KD_CAM_CAMERATOWORLD + 32).w);
  return len(P - camP);
}

/* kernel_write_data_passes (kernel_passes.h:174-282) for one shading point of a camera
 * path: data passes once per path (on the first surface that is opaque enough), the
 * colour passes and the mist on every surface the camera ray crosses */
CY_DEV void pass_write_data(const ShaderDataG &sd, const LobeArena &arena, PathStateG &st,
                            f3 throughput, float *pb)
{
  if (!(st.flag & CY_PATH_RAY_CAMERA))
    return;
  const int flag = kd_int(KD_FILM_PASS_FLAG), light_flag = kd_int(KD_FILM_LIGHT_PASS_FLAG);
  /* shader_bsdf_alpha, kernel_shader.h:886-894 */
  const f3 transparency = (sd.flag & CY_SD_TRANSPARENT) ? sd.closure_transparent_extinction :
                                                          zero3();
  f3 alpha = mk3(1.0f, 1.0f, 1.0f) - transparency;
  alpha = mk3(fminf(fmaxf(alpha.x, 0.0f), 1.0f), fminf(fmaxf(alpha.y, 0.0f), 1.0f),
              fminf(fmaxf(alpha.z, 0.0f), 1.0f));

  if (!(st.flag & CY_PATH_RAY_SINGLE_PASS_DONE)) {
    const float threshold = kd_float(KD_FILM_PASS_ALPHA_THRESHOLD);
    if (!(sd.flag & CY_SD_TRANSPARENT) || threshold == 0.0f || average(alpha) >= threshold) {
      pb[PB_DEPTH] = pass_camera_z_depth(sd.P);
      pb[PB_OBJECT_ID] = (sd.object == -1) ?
                             0.0f :
                             __ldg((const float *)(g_scene.objects +
                                                   (size_t)sd.object * SIZEOF_KERNEL_OBJECT +
                                                   KO_PASS_ID));
      pb[PB_MATERIAL_ID] = __ldg((const float *)(g_scene.shaders +
                                                 (size_t)(sd.shader & CY_SHADER_MASK) *
                                                     SIZEOF_KERNEL_SHADER +
                                                 KS_PASS_ID));
      if (flag & (1 << CY_PASS_NORMAL)) {
        /* shader_bsdf_average_normal, kernel_shader.h:934-947 */
        f3 N = zero3();
        int at = 0;
        for (int i = 0; i < arena.n; i++) {
          const uint32_t kind = lobe_kind_at(arena, at);
          if (lobe_is_sampled(kind))
            N += lobe_normal_at(arena, at) * fabsf(average(lobe_weight_at(arena, at)));
          at += lobe_words(kind);
        }
        pb_set3(pb, PB_NORMAL, is_zero(N) ? sd.N : normalize(N));
      }
      if (flag & (1 << CY_PASS_UV)) {
        /* primitive_uv, geom/geom_primitive.h */
        f3 uv = zero3();
        const AttrDesc d = find_attribute(sd, CY_ATTR_STD_UV);
        if (d.offset != CY_ATTR_STD_NOT_FOUND) {
          const float2 t = attribute_float2(sd, d);
          uv = mk3(t.x, t.y, 1.0f);
        }
        pb_set3(pb, PB_UV, uv);
      }
      pb[PB_HAS_DATA] = 1.0f;
      st.flag |= CY_PATH_RAY_SINGLE_PASS_DONE;
    }
  }

  /* shader_bsdf_diffuse / _glossy / _transmission: summed lobe weights per class */
  const bool want_d = light_flag & ((1 << (CY_PASS_DIFFUSE_DIRECT % 32)) |
                                    (1 << (CY_PASS_DIFFUSE_INDIRECT % 32)) |
                                    (1 << (CY_PASS_DIFFUSE_COLOR % 32)));
  const bool want_g = light_flag & ((1 << (CY_PASS_GLOSSY_DIRECT % 32)) |
                                    (1 << (CY_PASS_GLOSSY_INDIRECT % 32)) |
                                    (1 << (CY_PASS_GLOSSY_COLOR % 32)));
  const bool want_t = light_flag & ((1 << (CY_PASS_TRANSMISSION_DIRECT % 32)) |
                                    (1 << (CY_PASS_TRANSMISSION_INDIRECT % 32)) |
                                    (1 << (CY_PASS_TRANSMISSION_COLOR % 32)));
  if (want_d || want_g || want_t) {
    EvalSplit w;
    eval_split_zero(w);
    int at = 0;
    for (int i = 0; i < arena.n; i++) {
      const uint32_t kind = lobe_kind_at(arena, at);
      if (lobe_is_sampled(kind) && lobe_id(kind) != CY_CLOSURE_BSDF_TRANSPARENT_ID)
        eval_split_add(w, kind, lobe_weight_at(arena, at));
      at += lobe_words(kind);
    }
    if (want_d)
      pb_add3(pb, PB_COLOR_DIFFUSE, w.diffuse * throughput);
    if (want_g)
      pb_add3(pb, PB_COLOR_GLOSSY, w.glossy * throughput);
    if (want_t)
      pb_add3(pb, PB_COLOR_TRANSMISSION, w.transmission * throughput);
  }

  if (light_flag & (1 << (CY_PASS_MIST % 32))) {
    const float depth = pass_camera_distance(sd.P);
    float mist = saturate((depth - kd_float(KD_FILM_MIST_START)) * kd_float(KD_FILM_MIST_INV_DEPTH));
    const float falloff = kd_float(KD_FILM_MIST_FALLOFF);
    if (falloff == 1.0f)
      ;
    else if (falloff == 2.0f)
      mist = mist * mist;
    else if (falloff == 0.5f)
      mist = sqrtf(mist);
    else
      mist = powf(mist, falloff);
    pb[PB_MIST] += (1.0f - mist) * average(throughput * alpha);
  }
}

CY_DEV float ensure_finite(float v)
{
  return isfinite_safe(v) ? v : 0.0f;
}
CY_DEV f3 ensure_finite3(f3 v)
{
  return mk3(ensure_finite(v.x), ensure_finite(v.y), ensure_finite(v.z));
}

/* kernel_update_denoising_features (kernel_passes.h:46-122): the denoiser's guide images
 * are the normal, albedo and depth of the first surface that is not (mostly) specular -
 * a path crossing glass or bouncing off a mirror keeps the feature open, weighted by the
 * specular albedo so far.  Needs what the lobes know: the Fresnel tint a Principled-style
 * lobe was weighted by (recomputed from its ior and cspec0, svm_closure.cuh
 * lobe_weigh_by_fresnel), the sheen lobe's average value, the roughness pair. */
CY_DEV void denoising_update_features(const ShaderDataG &sd, const LobeArena &arena, float *pb)
{
  const float fw = pb[PB_DN_WEIGHT];
  if (fw == 0.0f)
    return;
  pb[PB_DN_DEPTH] += ensure_finite(fw * sd.ray_length);

  f3 normal = zero3(), diffuse_albedo = zero3(), specular_albedo = zero3();
  float sum_weight = 0.0f, sum_nonspecular_weight = 0.0f;
  int at = 0;
  for (int i = 0; i < arena.n; i++) {
    const uint32_t kind = lobe_kind_at(arena, at);
    if (lobe_is_sampled(kind)) {
      const Lobe l = lobe_fetch(arena, at);
      const int id = lobe_id(kind);
      normal += l.N * l.sample_weight;
      sum_weight += l.sample_weight;
      f3 albedo = l.weight;
      float roughness2 = 1.0f; /* bsdf_get_specular_roughness_squared, closure/bsdf.h:43-55 */
      const bool fresnel = id == CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_FRESNEL_ID ||
                           id == CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_GLASS_FRESNEL_ID ||
                           id == CY_CLOSURE_BSDF_MICROFACET_GGX_FRESNEL_ID ||
                           id == CY_CLOSURE_BSDF_MICROFACET_GGX_CLEARCOAT_ID;
      if (fresnel) {
        const float F0 = fresnel_dielectric_cos(1.0f, l.ior);
        f3 tint = fresnel_tint(sd.I, l.N, l.ior, F0, l.cspec0);
        if (id == CY_CLOSURE_BSDF_MICROFACET_GGX_CLEARCOAT_ID)
          tint *= 0.25f * l.aux;
        albedo *= tint;
      }
      else if (id == CY_CLOSURE_BSDF_PRINCIPLED_SHEEN_ID) {
        const float NdotI = dot(l.N, sd.I);
        albedo *= (NdotI < 0.0f) ? 0.0f : schlick_weight(NdotI) * NdotI;
      }
      if (id == CY_CLOSURE_BSDF_REFLECTION_ID || id == CY_CLOSURE_BSDF_REFRACTION_ID ||
          id == CY_CLOSURE_BSDF_TRANSPARENT_ID)
        roughness2 = 0.0f;
      else if ((id >= CY_CLOSURE_BSDF_MICROFACET_GGX_ID &&
                id <= CY_CLOSURE_BSDF_ASHIKHMIN_SHIRLEY_ID) ||
               (id >= CY_CLOSURE_BSDF_MICROFACET_BECKMANN_REFRACTION_ID &&
                id <= CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_GLASS_ID) ||
               id == CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_GLASS_FRESNEL_ID)
        roughness2 = l.ax * l.ay;
      if (roughness2 > 0.075f * 0.075f) {
        diffuse_albedo += albedo;
        sum_nonspecular_weight += l.sample_weight;
      }
      else {
        specular_albedo += albedo;
      }
    }
    at += lobe_words(kind);
  }

  /* wait for the next bounce while 75 % or more of the sample weight is specular */
  if (sum_weight == 0.0f || sum_nonspecular_weight * 4.0f > sum_weight) {
    if (sum_weight != 0.0f)
      normal = normal / sum_weight;
    tfm34 w2c;
    w2c.x = kd_float4(KD_CAM_WORLDTOCAMERA);
    w2c.y = kd_float4(KD_CAM_WORLDTOCAMERA + 16);
    w2c.z = kd_float4(KD_CAM_WORLDTOCAMERA + 32);
    normal = transform_direction(w2c, normal);
    pb_add3(pb, PB_DN_NORMAL, ensure_finite3(normal * fw));
    pb_add3(pb, PB_DN_ALBEDO,
            ensure_finite3(pb_get3(pb, PB_DN_THROUGHPUT) * fw * diffuse_albedo));
    pb[PB_DN_WEIGHT] = 0.0f;
  }
  else {
    pb_set3(pb, PB_DN_THROUGHPUT, pb_get3(pb, PB_DN_THROUGHPUT) * specular_albedo);
  }
}

CY_DEV f3 safe_divide_color(f3 a, f3 b)
{
  return mk3((b.x != 0.0f) ? a.x / b.x : 0.0f, (b.y != 0.0f) ? a.y / b.y : 0.0f,
             (b.z != 0.0f) ? a.z / b.z : 0.0f);
}

#endif /* B200_PASSES_CUH */

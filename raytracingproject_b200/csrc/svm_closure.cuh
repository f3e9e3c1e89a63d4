/* svm_closure.cuh - NODE_CLOSURE_BSDF: turns a BSDF node of the compiled shader into
 * lobes in the shading point's arena (lobes.cuh).
 *
 * Semantics to match: kernel/svm/svm_closure.h:60-1000 for the BSDFs in scope - Diffuse /
 * Oren-Nayar, Translucent, Transparent, Glossy and the Anisotropic BSDF (sharp, GGX,
 * multi-scatter GGX), Glass and Refraction (sharp, GGX, multi-scatter GGX), and the
 * Principled BSDF with the GGX or the Multiscatter-GGX distribution (no subsurface) - and
 * closure/alloc.h:19-68 for the closure budget (a lobe whose weight is under the cutoff
 * still occupies a slot; a Fresnel lobe costs two).  The node's word layout is the SVM
 * bytecode the host compiler emits (render/nodes.cpp, BsdfNode::compile /
 * PrincipledBsdfNode::compile), i.e. part of the binary interface.
 *
 * Shape: one builder per BSDF layer, each filling a `Lobe` in registers and committing it
 * with lobe_store(); the Principled node is the sum of its five layers.
 * Included by shade.cuh; host-compilable. */
#ifndef B200_SVM_CLOSURE_CUH
#define B200_SVM_CLOSURE_CUH

CY_DEV f3 saturate3(f3 a)
{
  return mk3(saturate(a.x), saturate(a.y), saturate(a.z));
}

/* Rodrigues rotation of p about the unit axis by `angle` (matrix form, as
 * util_math.h:563-582 evaluates it) */
CY_DEV f3 rotate_around_axis(f3 p, f3 axis, float angle)
{
  const float c = cosf(angle), s = sinf(angle), ic = 1 - c;
  f3 r;
  r.x = ((c + ic * axis.x * axis.x) * p.x) + ((ic * axis.x * axis.y - axis.z * s) * p.y) +
        ((ic * axis.x * axis.z + axis.y * s) * p.z);
  r.y = ((ic * axis.x * axis.y + axis.z * s) * p.x) + ((c + ic * axis.y * axis.y) * p.y) +
        ((ic * axis.y * axis.z - axis.x * s) * p.z);
  r.z = ((ic * axis.x * axis.z - axis.y * s) * p.x) + ((ic * axis.y * axis.z + axis.x * s) * p.y) +
        ((c + ic * axis.z * axis.z) * p.z);
  return r;
}

/* --------------------------------------------------------- closure budget */

/* Starts a lobe of `weight`.  False = nothing to set up: the budget is spent, or the
 * weight is under the cutoff - then a placeholder (kind NONE) takes the slot, because the
 * reference's bsdf_alloc has already counted the closure when it rejects it, and the
 * sampling sums see its sample weight.  `fresnel_extra`: the lobe needs the second slot
 * the reference spends on MicrofacetExtra; without room for it the lobe is dropped. */
CY_DEV bool lobe_open(LobeArena &arena, Lobe &l, f3 weight, bool fresnel_extra)
{
  if (arena.left == 0)
    return false;
  l.weight = weight;
  l.sample_weight = fabsf(average(weight));
  l.kind = CY_CLOSURE_NONE_ID;
  l.N = zero3();
  if (!(l.sample_weight >= CLOSURE_WEIGHT_CUTOFF)) {
    lobe_store(arena, l);
    arena.left -= 1;
    return false;
  }
  if (fresnel_extra && arena.left < 2)
    return false;
  arena.left -= fresnel_extra ? 2 : 1;
  l.ax = l.ay = l.ior = l.aux = 0.0f;
  l.cspec0 = l.color = l.T = zero3();
  return true;
}

/* Principled-style lobes are sampled in proportion to how much they reflect towards the
 * viewer: sample weight *= average Fresnel tint at the shading normal */
CY_DEV void lobe_weigh_by_fresnel(const ShaderDataG &sd, Lobe &l, float clearcoat_scale)
{
  const float F0 = fresnel_dielectric_cos(1.0f, l.ior);
  f3 tint = fresnel_tint(sd.I, l.N, l.ior, F0, l.cspec0);
  if (clearcoat_scale >= 0.0f)
    tint *= 0.25f * clearcoat_scale;
  l.sample_weight *= average(tint);
}

/* ------------------------------------------------------------ GGX layers */

/* plain GGX reflection or refraction; isotropic unless ax != ay was set by the caller */
CY_DEV void commit_ggx(ShaderDataG &sd, LobeArena &arena, Lobe &l, bool refraction)
{
  l.ax = saturate(l.ax);
  l.ay = refraction ? l.ax : saturate(l.ay);
  l.kind = (l.kind & ~LOBE_ID_MASK) | (refraction ? CY_CLOSURE_BSDF_MICROFACET_GGX_REFRACTION_ID :
                                                    CY_CLOSURE_BSDF_MICROFACET_GGX_ID);
  if (l.ax == l.ay)
    l.kind &= ~LOBE_HAS_TANGENT;
  lobe_store(arena, l);
  sd.flag |= CY_SD_BSDF | CY_SD_BSDF_HAS_EVAL;
}

CY_DEV void commit_ggx_fresnel(ShaderDataG &sd, LobeArena &arena, Lobe &l)
{
  l.cspec0 = saturate3(l.cspec0);
  l.ax = saturate(l.ax);
  l.ay = saturate(l.ay);
  l.kind = (l.kind & ~LOBE_ID_MASK) | CY_CLOSURE_BSDF_MICROFACET_GGX_FRESNEL_ID;
  if (l.ax == l.ay)
    l.kind &= ~LOBE_HAS_TANGENT;
  lobe_weigh_by_fresnel(sd, l, -1.0f);
  lobe_store(arena, l);
  sd.flag |= CY_SD_BSDF | CY_SD_BSDF_HAS_EVAL;
}

/* multi-scatter reflection (mirror microfacets), plain or Fresnel tinted */
CY_DEV void commit_multi_ggx(ShaderDataG &sd, LobeArena &arena, Lobe &l, bool fresnel)
{
  if (is_zero(l.T))
    l.T = mk3(1.0f, 0.0f, 0.0f);
  if (fresnel)
    lobe_weigh_by_fresnel(sd, l, -1.0f); /* before the clamps, like the reference */
  l.ax = clampf(l.ax, 1e-4f, 1.0f);
  l.ay = clampf(l.ay, 1e-4f, 1.0f);
  l.color = saturate3(l.color);
  l.cspec0 = saturate3(l.cspec0);
  l.kind = fresnel ? CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_FRESNEL_ID :
                     CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_ID;
  if (l.ax != l.ay)
    l.kind |= LOBE_HAS_TANGENT;
  lobe_store(arena, l);
  sd.flag |= CY_SD_BSDF | CY_SD_BSDF_HAS_EVAL | CY_SD_BSDF_NEEDS_LCG;
}

/* multi-scatter dielectric interface */
CY_DEV void commit_multi_glass(ShaderDataG &sd, LobeArena &arena, Lobe &l, bool fresnel)
{
  l.ax = clampf(l.ax, 1e-4f, 1.0f);
  l.ay = l.ax;
  l.ior = fmaxf(0.0f, l.ior);
  l.color = saturate3(l.color);
  if (fresnel) {
    l.cspec0 = saturate3(l.cspec0);
    lobe_weigh_by_fresnel(sd, l, -1.0f);
  }
  l.kind = fresnel ? CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_GLASS_FRESNEL_ID :
                     CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_GLASS_ID;
  lobe_store(arena, l);
  sd.flag |= CY_SD_BSDF | CY_SD_BSDF_HAS_EVAL | CY_SD_BSDF_NEEDS_LCG;
}

/* ------------------------------------------------------ Principled BSDF */

struct PrincipledInputs {
  f3 N, T, clearcoat_normal, base_color, weight;
  float metallic, specular, roughness, specular_tint, anisotropic, sheen, sheen_tint;
  float clearcoat, clearcoat_roughness, transmission, transmission_roughness, ior;
  float dielectric_fresnel; /* at the shading normal, for the glass split */
  bool multiscatter;
  bool may_reflect, may_refract; /* caustics switches for this path */
};

CY_DEV f3 hue_of(f3 color, float luminance, f3 fallback)
{
  return luminance > 0.0f ? color / luminance : fallback;
}

CY_DEV void principled_diffuse_layer(ShaderDataG &sd, LobeArena &arena,
                                     const PrincipledInputs &in, float diffuse_weight)
{
  Lobe l;
  if (lobe_open(arena, l, in.weight * in.base_color * diffuse_weight, false)) {
    l.N = in.N;
    l.aux = in.roughness;
    l.kind = CY_CLOSURE_BSDF_PRINCIPLED_DIFFUSE_ID;
    lobe_store(arena, l);
    sd.flag |= CY_SD_BSDF | CY_SD_BSDF_HAS_EVAL;
  }
}

/* Sheen: full interpreter only.  The host routes every program whose Principled sheen
 * input is not a constant zero there (svm_validate, SVM_USES_EXTENDED_NODES), so the lean
 * one never sees sheen - and a test for it in the lean node cost those kernels 2 %. */
CY_DEV void principled_sheen_layer(ShaderDataG &sd, LobeArena &arena,
                                   const PrincipledInputs &in, float diffuse_weight)
{
  if (!(in.sheen > CLOSURE_WEIGHT_CUTOFF))
    return;
  const float lum = dot(in.base_color, mk3(kd_float(KD_FILM_RGB_TO_Y),
                                           kd_float(KD_FILM_RGB_TO_Y + 4),
                                           kd_float(KD_FILM_RGB_TO_Y + 8)));
  const f3 tint = hue_of(in.base_color, lum, one3());
  const f3 sheen_color = one3() * (1.0f - in.sheen_tint) + tint * in.sheen_tint;
  Lobe l;
  if (lobe_open(arena, l, in.weight * in.sheen * sheen_color * diffuse_weight, false)) {
    l.N = in.N;
    l.kind = CY_CLOSURE_BSDF_PRINCIPLED_SHEEN_ID;
    const float NdotI = dot(in.N, sd.I);
    l.sample_weight *= (NdotI < 0.0f) ? 0.0f : schlick_weight(NdotI) * NdotI;
    lobe_store(arena, l);
    sd.flag |= CY_SD_BSDF | CY_SD_BSDF_HAS_EVAL;
  }
}

template<bool MS>
CY_DEV bool principled_specular_layer(ShaderDataG &sd, LobeArena &arena,
                                      const PrincipledInputs &in, float specular_weight)
{
  if (!in.may_reflect || !(specular_weight > CLOSURE_WEIGHT_CUTOFF) ||
      !(in.specular > CLOSURE_WEIGHT_CUTOFF || in.metallic > CLOSURE_WEIGHT_CUTOFF))
    return true;
  Lobe l;
  if (!lobe_open(arena, l, in.weight * specular_weight, true))
    return true;
  l.N = in.N;
  l.ior = (2.0f / (1.0f - safe_sqrtf(0.08f * in.specular))) - 1.0f;
  const float aspect = safe_sqrtf(1.0f - in.anisotropic * 0.9f);
  const float r2 = in.roughness * in.roughness;
  l.ax = r2 / aspect;
  l.ay = r2 * aspect;
  const float lum = 0.3f * in.base_color.x + 0.6f * in.base_color.y + 0.1f * in.base_color.z;
  const f3 tint = hue_of(in.base_color, lum, zero3());
  const f3 dielectric_col = one3() * (1.0f - in.specular_tint) + tint * in.specular_tint;
  l.cspec0 = (in.specular * 0.08f * dielectric_col) * (1.0f - in.metallic) +
             in.base_color * in.metallic;
  l.color = in.base_color;
  l.T = in.T;
  l.kind = LOBE_HAS_TANGENT;
  /* smooth surfaces scatter once: single-scatter GGX below roughness 0.075 */
  if (!in.multiscatter || in.roughness <= 0.075f) {
    commit_ggx_fresnel(sd, arena, l);
  }
  else {
    if (!MS)
      return false; /* this kernel does not carry the random walk */
    commit_multi_ggx(sd, arena, l, true);
  }
  return true;
}

template<bool MS>
CY_DEV bool principled_transmission_layers(ShaderDataG &sd, LobeArena &arena,
                                           const PrincipledInputs &in, float final_transmission)
{
  if (!(in.may_reflect || in.may_refract) || !(final_transmission > CLOSURE_WEIGHT_CUTOFF))
    return true;
  const f3 glass_weight = in.weight * final_transmission;
  const f3 cspec0 = in.base_color * in.specular_tint + one3() * (1.0f - in.specular_tint);
  Lobe l;
  if (in.roughness <= 5e-2f || !in.multiscatter) {
    /* a reflection and a refraction lobe, split by the dielectric Fresnel term */
    if (in.may_reflect && lobe_open(arena, l, glass_weight * in.dielectric_fresnel, true)) {
      l.N = in.N;
      l.ax = l.ay = in.roughness * in.roughness;
      l.ior = in.ior;
      l.cspec0 = cspec0;
      commit_ggx_fresnel(sd, arena, l);
    }
    if (in.may_refract &&
        lobe_open(arena, l, in.base_color * glass_weight * (1.0f - in.dielectric_fresnel),
                  false)) {
      l.N = in.N;
      /* the GGX distribution has its own transmission roughness on top of the surface's */
      const float tr = in.multiscatter ?
                           in.roughness :
                           1.0f - (1.0f - in.roughness) * (1.0f - in.transmission_roughness);
      l.ax = l.ay = tr * tr;
      l.ior = in.ior;
      commit_ggx(sd, arena, l, true);
    }
    return true;
  }
  if (arena.left == 0)
    return true;
  if (!MS)
    return false;
  if (lobe_open(arena, l, glass_weight, true)) {
    l.N = in.N;
    l.ax = l.ay = in.roughness * in.roughness;
    l.ior = in.ior;
    l.color = in.base_color;
    l.cspec0 = cspec0;
    commit_multi_glass(sd, arena, l, true);
  }
  return true;
}

CY_DEV void principled_clearcoat_layer(ShaderDataG &sd, LobeArena &arena,
                                       const PrincipledInputs &in)
{
  Lobe l;
  if (!in.may_reflect || !(in.clearcoat > CLOSURE_WEIGHT_CUTOFF) ||
      !lobe_open(arena, l, in.weight, true))
    return;
  l.N = in.clearcoat_normal;
  l.ior = 1.5f;
  l.cspec0 = mk3(0.04f, 0.04f, 0.04f);
  l.ax = l.ay = saturate(in.clearcoat_roughness * in.clearcoat_roughness);
  l.aux = in.clearcoat;
  l.kind = CY_CLOSURE_BSDF_MICROFACET_GGX_CLEARCOAT_ID;
  lobe_weigh_by_fresnel(sd, l, in.clearcoat);
  lobe_store(arena, l);
  sd.flag |= CY_SD_BSDF | CY_SD_BSDF_HAS_EVAL;
}

/* ------------------------------------------------------------ the node */

/* FULL = false is the interpreter for the common shaders (see svm_eval_nodes); false is
 * returned when the shader needs more than this kernel carries (a bent normal -> the full
 * interpreter; a multi-scatter lobe -> a kernel with MS). */
template<bool FULL, bool MS = FULL>
CY_DEV bool svm_node_closure_bsdf(ShaderDataG &sd, LobeArena &arena, float *stack, uint4 node,
                                  uint32_t path_flag, int *offset)
{
  const uint32_t type = node.y & 0xff, param1_offset = (node.y >> 8) & 0xff;
  const uint32_t param2_offset = (node.y >> 16) & 0xff, mix_weight_offset = (node.y >> 24) & 0xff;
  const float mix_weight = stack_valid(mix_weight_offset) ? stack[mix_weight_offset] : 1.0f;

  const uint4 data_node = __ldg(&g_scene.svm_nodes[*offset]);
  (*offset)++;

  if (mix_weight == 0.0f) {
    if (type == CY_CLOSURE_BSDF_PRINCIPLED_ID)
      (*offset) += 4; /* the Principled node carries four more data nodes */
    return true;
  }

  const f3 N = stack_valid(data_node.x) ? stack_load_float3(stack, data_node.x) : sd.N;
  /* a linked normal that is not the shading normal needs the terminator factors of
   * bsdf_eval / bsdf_sample, which only the full kernels carry */
  if (!FULL && stack_valid(data_node.x) && !isequal3(N, sd.N))
    return false;
  const float param1 = stack_valid(param1_offset) ? stack[param1_offset] :
                                                    __uint_as_float(node.z);
  const float param2 = stack_valid(param2_offset) ? stack[param2_offset] :
                                                    __uint_as_float(node.w);
  const f3 weight = sd.svm_closure_weight * mix_weight;
  const bool on_diffuse_path = (path_flag & CY_PATH_RAY_DIFFUSE) != 0;
  const bool may_reflect = kd_int(KD_INT_CAUSTICS_REFLECTIVE) || !on_diffuse_path;
  const bool may_refract = kd_int(KD_INT_CAUSTICS_REFRACTIVE) || !on_diffuse_path;
  Lobe l;

  switch (type) {
    case CY_CLOSURE_BSDF_PRINCIPLED_ID: {
      const uint4 data_node2 = __ldg(&g_scene.svm_nodes[*offset]);
      const uint4 data_base_color = __ldg(&g_scene.svm_nodes[*offset + 1]);
      const uint4 data_cn_ssr = __ldg(&g_scene.svm_nodes[*offset + 2]);
      const uint4 data_subsurface_color = __ldg(&g_scene.svm_nodes[*offset + 3]);
      (*offset) += 4;

      PrincipledInputs in;
      in.N = N;
      in.weight = weight;
      in.T = stack_load_float3(stack, data_node.y);
      in.metallic = param1;
      float subsurface = param2;
      in.specular = stack[data_node.z & 0xff];
      in.roughness = stack[(data_node.z >> 8) & 0xff];
      in.specular_tint = stack[(data_node.z >> 16) & 0xff];
      in.anisotropic = stack[(data_node.z >> 24) & 0xff];
      in.sheen = stack[data_node.w & 0xff];
      in.sheen_tint = stack[(data_node.w >> 8) & 0xff];
      in.clearcoat = stack[(data_node.w >> 16) & 0xff];
      in.clearcoat_roughness = stack[(data_node.w >> 24) & 0xff];
      const float eta = fmaxf(stack[data_node2.x & 0xff], 1e-5f);
      in.transmission = stack[(data_node2.x >> 8) & 0xff];
      const float anisotropic_rotation = stack[(data_node2.x >> 16) & 0xff];
      in.transmission_roughness = stack[(data_node2.x >> 24) & 0xff];
      in.multiscatter = (int)data_node2.y != CY_CLOSURE_BSDF_MICROFACET_GGX_GLASS_ID;
      if (anisotropic_rotation != 0.0f)
        in.T = rotate_around_axis(in.T, N, anisotropic_rotation * CY_M_2PI_F);
      in.ior = (sd.flag & CY_SD_BACKFACING) ? 1.0f / eta : eta;
      in.dielectric_fresnel = fresnel_dielectric_cos(dot(N, sd.I), in.ior);
      in.may_reflect = may_reflect;
      in.may_refract = may_refract;

      in.base_color = stack_valid(data_base_color.x) ?
                          stack_load_float3(stack, data_base_color.x) :
                          mk3(__uint_as_float(data_base_color.y),
                              __uint_as_float(data_base_color.z),
                              __uint_as_float(data_base_color.w));
      in.clearcoat_normal = stack_valid(data_cn_ssr.x) ? stack_load_float3(stack, data_cn_ssr.x) :
                                                         sd.N;
      const f3 subsurface_color = stack_valid(data_subsurface_color.x) ?
                                      stack_load_float3(stack, data_subsurface_color.x) :
                                      mk3(__uint_as_float(data_subsurface_color.y),
                                          __uint_as_float(data_subsurface_color.z),
                                          __uint_as_float(data_subsurface_color.w));

      const float diffuse_weight = (1.0f - saturate(in.metallic)) *
                                   (1.0f - saturate(in.transmission));
      const float final_transmission = saturate(in.transmission) * (1.0f - saturate(in.metallic));
      const float specular_weight = 1.0f - final_transmission;

      /* Subsurface is out of scope (svm_validate accepts the node only with a provably
       * zero subsurface input), but the reference's bookkeeping around it still decides
       * whether the diffuse lobe exists: it is dropped when the blended colour is black,
       * and on a path with a diffuse ancestor the blended colour replaces the base. */
      const f3 blended = subsurface_color * subsurface + in.base_color * (1.0f - subsurface);
      if (path_flag & CY_PATH_RAY_DIFFUSE_ANCESTOR) {
        subsurface = 0.0f;
        in.base_color = blended;
      }
      if (diffuse_weight > CLOSURE_WEIGHT_CUTOFF) {
        if (fabsf(average(blended)) > CLOSURE_WEIGHT_CUTOFF && subsurface <= CLOSURE_WEIGHT_CUTOFF)
          principled_diffuse_layer(sd, arena, in, diffuse_weight);
        if (FULL)
          principled_sheen_layer(sd, arena, in, diffuse_weight);
      }
      if (!principled_specular_layer<MS>(sd, arena, in, specular_weight))
        return false;
      if (!principled_transmission_layers<MS>(sd, arena, in, final_transmission))
        return false;
      principled_clearcoat_layer(sd, arena, in);
      break;
    }
    case CY_CLOSURE_BSDF_DIFFUSE_ID:
      /* Lambert, or Oren-Nayar when the node has roughness */
      if (lobe_open(arena, l, weight, false)) {
        l.N = N;
        if (param1 == 0.0f) {
          l.kind = CY_CLOSURE_BSDF_DIFFUSE_ID;
        }
        else {
          l.kind = CY_CLOSURE_BSDF_OREN_NAYAR_ID;
          oren_nayar_coefficients(param1, &l.ax, &l.aux);
        }
        lobe_store(arena, l);
        sd.flag |= CY_SD_BSDF | CY_SD_BSDF_HAS_EVAL;
      }
      break;
    case CY_CLOSURE_BSDF_TRANSLUCENT_ID:
      if (lobe_open(arena, l, weight, false)) {
        l.N = N;
        l.kind = CY_CLOSURE_BSDF_TRANSLUCENT_ID;
        lobe_store(arena, l);
        sd.flag |= CY_SD_BSDF | CY_SD_BSDF_HAS_EVAL;
      }
      break;
    case CY_CLOSURE_BSDF_TRANSPARENT_ID: {
      /* all transparent closures of a shader merge into one lobe; the summed weight is
       * what shadow rays multiply by (closure/bsdf_transparent.h:38-85) */
      const float sample_weight = fabsf(average(weight));
      if (!(sample_weight >= CLOSURE_WEIGHT_CUTOFF))
        break;
      if (sd.flag & CY_SD_TRANSPARENT) {
        sd.closure_transparent_extinction += weight;
        if (sd.transparent_at >= 0) {
          lobe_add_weight_at(arena, sd.transparent_at, weight, sample_weight);
        }
      }
      else {
        sd.flag |= CY_SD_BSDF | CY_SD_TRANSPARENT;
        sd.closure_transparent_extinction = weight;
        /* a terminating path evaluates no closures, but still has to pass through */
        const bool terminating = (path_flag & CY_PATH_RAY_TERMINATE) != 0 && arena.q != nullptr;
        if (terminating)
          arena.left = 1;
        if (arena.left != 0) {
          l.weight = weight;
          l.sample_weight = sample_weight;
          l.N = sd.N;
          l.kind = CY_CLOSURE_BSDF_TRANSPARENT_ID;
          sd.transparent_at = lobe_store(arena, l);
          arena.left -= 1;
        }
        else if (terminating) {
          arena.left = 0;
        }
      }
      break;
    }
    case CY_CLOSURE_BSDF_REFRACTION_ID:
    case CY_CLOSURE_BSDF_MICROFACET_GGX_REFRACTION_ID:
      /* Refraction BSDF node, sharp or GGX */
      if (may_refract && lobe_open(arena, l, weight, false)) {
        l.N = N;
        const float eta = fmaxf(param2, 1e-5f);
        l.ior = (sd.flag & CY_SD_BACKFACING) ? 1.0f / eta : eta;
        if (type == CY_CLOSURE_BSDF_REFRACTION_ID) {
          l.kind = CY_CLOSURE_BSDF_REFRACTION_ID;
          lobe_store(arena, l);
          sd.flag |= CY_SD_BSDF;
        }
        else {
          l.ax = l.ay = sqr(param1);
          commit_ggx(sd, arena, l, true);
        }
      }
      break;
    case CY_CLOSURE_BSDF_SHARP_GLASS_ID:
    case CY_CLOSURE_BSDF_MICROFACET_GGX_GLASS_ID: {
      /* Glass BSDF node: a reflection and a refraction lobe weighted by the dielectric
       * Fresnel term at the shading normal */
      if (!may_reflect && !may_refract)
        break;
      float eta = fmaxf(param2, 1e-5f);
      eta = (sd.flag & CY_SD_BACKFACING) ? 1.0f / eta : eta;
      const float fresnel = fresnel_dielectric_cos(dot(N, sd.I), eta);
      const float alpha = sqr(param1);
      const bool sharp = (type == CY_CLOSURE_BSDF_SHARP_GLASS_ID);
      if (may_reflect && lobe_open(arena, l, weight * fresnel, false)) {
        l.N = N;
        if (sharp) {
          l.kind = CY_CLOSURE_BSDF_REFLECTION_ID;
          lobe_store(arena, l);
          sd.flag |= CY_SD_BSDF;
        }
        else {
          l.ax = l.ay = alpha;
          l.ior = eta;
          commit_ggx(sd, arena, l, false);
        }
      }
      if (may_refract && lobe_open(arena, l, weight * (1.0f - fresnel), false)) {
        l.N = N;
        l.ior = eta;
        if (sharp) {
          l.kind = CY_CLOSURE_BSDF_REFRACTION_ID;
          lobe_store(arena, l);
          sd.flag |= CY_SD_BSDF;
        }
        else {
          l.ax = l.ay = alpha;
          commit_ggx(sd, arena, l, true);
        }
      }
      break;
    }
    case CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_GLASS_ID:
      /* Glass BSDF node, Multiscatter GGX: one lobe for both sides of the interface */
      if (!may_reflect && !may_refract)
        break;
      if (arena.left == 0)
        break;
      if (!MS)
        return false;
      if (lobe_open(arena, l, weight, true)) {
        l.N = N;
        l.ax = l.ay = sqr(param1);
        const float eta = fmaxf(param2, 1e-5f);
        l.ior = (sd.flag & CY_SD_BACKFACING) ? 1.0f / eta : eta;
        l.color = stack_load_float3(stack, data_node.z);
        commit_multi_glass(sd, arena, l, false);
      }
      break;
    case CY_CLOSURE_BSDF_REFLECTION_ID:
    case CY_CLOSURE_BSDF_MICROFACET_GGX_ID:
    case CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_ID: {
      /* Glossy BSDF node (sharp / GGX / Multiscatter GGX); with a tangent input it is the
       * Anisotropic BSDF node */
      if (!may_reflect)
        break;
      const bool multi = (type == CY_CLOSURE_BSDF_MICROFACET_MULTI_GGX_ID);
      if (multi && !MS) {
        if (arena.left == 0)
          break;
        return false;
      }
      /* the multi-scatter variant takes its MicrofacetExtra slot AFTER the closure's own:
       * same budget arithmetic as asking for both up front */
      if (!lobe_open(arena, l, weight, multi))
        break;
      const float alpha = sqr(param1);
      l.N = N;
      l.ax = l.ay = alpha;
      /* tangent, rotation and anisotropy: full interpreter only - the host routes
       * programs with a tangent input there (svm_validate) */
      if (FULL && stack_valid(data_node.y)) {
        f3 T = stack_load_float3(stack, data_node.y);
        const float rotation = stack[data_node.z];
        if (rotation != 0.0f)
          T = rotate_around_axis(T, N, rotation * CY_M_2PI_F);
        const float anisotropy = clampf(param2, -0.99f, 0.99f);
        if (anisotropy < 0.0f) {
          l.ax = alpha / (1.0f + anisotropy);
          l.ay = alpha * (1.0f + anisotropy);
        }
        else {
          l.ax = alpha * (1.0f - anisotropy);
          l.ay = alpha / (1.0f - anisotropy);
        }
        l.T = T;
        l.kind = LOBE_HAS_TANGENT;
      }
      if (type == CY_CLOSURE_BSDF_REFLECTION_ID) {
        l.kind = CY_CLOSURE_BSDF_REFLECTION_ID;
        lobe_store(arena, l);
        sd.flag |= CY_SD_BSDF;
      }
      else if (multi) {
        l.color = stack_load_float3(stack, data_node.w);
        commit_multi_ggx(sd, arena, l, false);
      }
      else {
        commit_ggx(sd, arena, l, false);
      }
      break;
    }
    default:
      break; /* refused at bind time */
  }
  return true;
}

#endif

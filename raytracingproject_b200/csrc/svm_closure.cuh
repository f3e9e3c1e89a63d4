/* svm_closure.cuh - NODE_CLOSURE_BSDF (kernel/svm/svm_closure.h:60-1000) for the
 * closures in scope: Diffuse / Oren-Nayar, Translucent, Principled (single-scatter GGX
 * distribution, no subsurface), Glossy, Glass and Refraction (GGX or sharp).
 * Included by shade.cuh. */
#ifndef B200_SVM_CLOSURE_CUH
#define B200_SVM_CLOSURE_CUH

CY_DEV f3 saturate3(f3 a)
{
  return mk3(saturate(a.x), saturate(a.y), saturate(a.z));
}

/* bsdf_microfacet.h:275-288 */
CY_DEV void bsdf_microfacet_fresnel_color(const ShaderDataG &sd, Closure *bsdf)
{
  float F0 = fresnel_dielectric_cos(1.0f, bsdf->ior);
  bsdf->fresnel_color = interpolate_fresnel_color(sd.I, bsdf->N, bsdf->ior, F0, bsdf->cspec0);
  if (bsdf->type == CY_CLOSURE_BSDF_MICROFACET_GGX_CLEARCOAT_ID)
    bsdf->fresnel_color *= 0.25f * bsdf->clearcoat;
  bsdf->sample_weight *= average(bsdf->fresnel_color);
}
CY_DEV uint32_t bsdf_microfacet_ggx_setup(Closure *bsdf)
{
  bsdf->alpha_x = saturate(bsdf->alpha_x);
  bsdf->alpha_y = saturate(bsdf->alpha_y);
  bsdf->type = CY_CLOSURE_BSDF_MICROFACET_GGX_ID;
  return CY_SD_BSDF | CY_SD_BSDF_HAS_EVAL;
}
CY_DEV uint32_t bsdf_microfacet_ggx_fresnel_setup(Closure *bsdf, const ShaderDataG &sd)
{
  bsdf->cspec0 = saturate3(bsdf->cspec0);
  bsdf->alpha_x = saturate(bsdf->alpha_x);
  bsdf->alpha_y = saturate(bsdf->alpha_y);
  bsdf->type = CY_CLOSURE_BSDF_MICROFACET_GGX_FRESNEL_ID;
  bsdf_microfacet_fresnel_color(sd, bsdf);
  return CY_SD_BSDF | CY_SD_BSDF_HAS_EVAL;
}
/* util_math.h:563-582 */
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

CY_DEV uint32_t bsdf_microfacet_ggx_clearcoat_setup(Closure *bsdf, const ShaderDataG &sd)
{
  bsdf->cspec0 = saturate3(bsdf->cspec0);
  bsdf->alpha_x = saturate(bsdf->alpha_x);
  bsdf->alpha_y = bsdf->alpha_x;
  bsdf->type = CY_CLOSURE_BSDF_MICROFACET_GGX_CLEARCOAT_ID;
  bsdf_microfacet_fresnel_color(sd, bsdf);
  return CY_SD_BSDF | CY_SD_BSDF_HAS_EVAL;
}
CY_DEV uint32_t bsdf_microfacet_ggx_refraction_setup(Closure *bsdf)
{
  bsdf->alpha_x = saturate(bsdf->alpha_x);
  bsdf->alpha_y = bsdf->alpha_x;
  bsdf->type = CY_CLOSURE_BSDF_MICROFACET_GGX_REFRACTION_ID;
  return CY_SD_BSDF | CY_SD_BSDF_HAS_EVAL;
}

/* MicrofacetBsdf + MicrofacetExtra allocation (svm_closure.h:287-292) */
CY_DEV Closure *microfacet_alloc(ShaderDataG &sd, f3 weight, bool with_extra)
{
  Closure *bsdf = bsdf_alloc(sd, weight);
  if (bsdf && with_extra) {
    if (!closure_alloc_extra(sd))
      return NULL;
  }
  return bsdf;
}

/* FULL = false is the interpreter for the common shaders (see svm_eval_nodes); the
 * return value is reserved for "this shader needs the full one". */
template<bool FULL>
CY_DEV bool svm_node_closure_bsdf(ShaderDataG &sd, float *stack, uint4 node, uint32_t path_flag,
                                  int *offset)
{
  const uint32_t type = node.y & 0xff, param1_offset = (node.y >> 8) & 0xff;
  const uint32_t param2_offset = (node.y >> 16) & 0xff, mix_weight_offset = (node.y >> 24) & 0xff;
  float mix_weight = stack_valid(mix_weight_offset) ? stack[mix_weight_offset] : 1.0f;

  const uint4 data_node = __ldg(&g_scene.svm_nodes[*offset]);
  (*offset)++;

  if (mix_weight == 0.0f) {
    if (type == CY_CLOSURE_BSDF_PRINCIPLED_ID)
      (*offset) += 4;
    return true;
  }

  f3 N = stack_valid(data_node.x) ? stack_load_float3(stack, data_node.x) : sd.N;
  /* a linked normal that is not the shading normal needs the terminator terms of
   * bsdf_eval / bsdf_sample, which only the full kernels carry */
  if (!FULL && stack_valid(data_node.x) && !isequal3(N, sd.N))
    return false;
  float param1 = stack_valid(param1_offset) ? stack[param1_offset] : __uint_as_float(node.z);
  float param2 = stack_valid(param2_offset) ? stack[param2_offset] : __uint_as_float(node.w);

  switch (type) {
    case CY_CLOSURE_BSDF_PRINCIPLED_ID: {
      /* svm_closure.h:100-465 */
      const uint4 data_node2 = __ldg(&g_scene.svm_nodes[*offset]);
      (*offset)++;
      f3 T = stack_load_float3(stack, data_node.y);
      const uint32_t specular_offset = data_node.z & 0xff,
                     roughness_offset = (data_node.z >> 8) & 0xff,
                     specular_tint_offset = (data_node.z >> 16) & 0xff,
                     anisotropic_offset = (data_node.z >> 24) & 0xff;
      const uint32_t sheen_offset = data_node.w & 0xff, sheen_tint_offset = (data_node.w >> 8) & 0xff,
                     clearcoat_offset = (data_node.w >> 16) & 0xff,
                     clearcoat_roughness_offset = (data_node.w >> 24) & 0xff;
      const uint32_t eta_offset = data_node2.x & 0xff, transmission_offset = (data_node2.x >> 8) & 0xff,
                     anisotropic_rotation_offset = (data_node2.x >> 16) & 0xff,
                     transmission_roughness_offset = (data_node2.x >> 24) & 0xff;

      float metallic = param1;
      float subsurface = param2;
      float specular = stack[specular_offset];
      float roughness = stack[roughness_offset];
      float specular_tint = stack[specular_tint_offset];
      float anisotropic = stack[anisotropic_offset];
      float sheen = stack[sheen_offset];
      float sheen_tint = stack[sheen_tint_offset];
      float clearcoat = stack[clearcoat_offset];
      float clearcoat_roughness = stack[clearcoat_roughness_offset];
      float transmission = stack[transmission_offset];
      float anisotropic_rotation = stack[anisotropic_rotation_offset];
      float transmission_roughness = stack[transmission_roughness_offset];
      float eta = fmaxf(stack[eta_offset], 1e-5f);
      const int distribution = (int)data_node2.y;
      if (anisotropic_rotation != 0.0f)
        T = rotate_around_axis(T, N, anisotropic_rotation * CY_M_2PI_F);

      float ior = (sd.flag & CY_SD_BACKFACING) ? 1.0f / eta : eta;
      float cosNO = dot(N, sd.I);
      float fresnel = fresnel_dielectric_cos(cosNO, ior);

      float diffuse_weight = (1.0f - saturate(metallic)) * (1.0f - saturate(transmission));
      float final_transmission = saturate(transmission) * (1.0f - saturate(metallic));
      float specular_weight = (1.0f - final_transmission);

      const uint4 data_base_color = __ldg(&g_scene.svm_nodes[*offset]);
      (*offset)++;
      f3 base_color = stack_valid(data_base_color.x) ?
                          stack_load_float3(stack, data_base_color.x) :
                          mk3(__uint_as_float(data_base_color.y), __uint_as_float(data_base_color.z),
                              __uint_as_float(data_base_color.w));
      const uint4 data_cn_ssr = __ldg(&g_scene.svm_nodes[*offset]);
      (*offset)++;
      f3 clearcoat_normal = stack_valid(data_cn_ssr.x) ? stack_load_float3(stack, data_cn_ssr.x) :
                                                         sd.N;
      const uint4 data_subsurface_color = __ldg(&g_scene.svm_nodes[*offset]);
      (*offset)++;
      f3 subsurface_color = stack_valid(data_subsurface_color.x) ?
                                stack_load_float3(stack, data_subsurface_color.x) :
                                mk3(__uint_as_float(data_subsurface_color.y),
                                    __uint_as_float(data_subsurface_color.z),
                                    __uint_as_float(data_subsurface_color.w));

      f3 weight = sd.svm_closure_weight * mix_weight;

      /* __SUBSURFACE__ branch of the reference; subsurface > cutoff is out of scope and
       * falls back to nothing being allocated for the diffuse lobe */
      f3 mixed_ss_base_color = subsurface_color * subsurface + base_color * (1.0f - subsurface);
      if (path_flag & CY_PATH_RAY_DIFFUSE_ANCESTOR) {
        subsurface = 0.0f;
        base_color = mixed_ss_base_color;
      }
      if (fabsf(average(mixed_ss_base_color)) > CLOSURE_WEIGHT_CUTOFF) {
        if (subsurface <= CLOSURE_WEIGHT_CUTOFF && diffuse_weight > CLOSURE_WEIGHT_CUTOFF) {
          f3 diff_weight = weight * base_color * diffuse_weight;
          Closure *bsdf = bsdf_alloc(sd, diff_weight);
          if (bsdf) {
            bsdf->N = N;
            bsdf->roughness = roughness;
            bsdf->type = CY_CLOSURE_BSDF_PRINCIPLED_DIFFUSE_ID;
            sd.flag |= CY_SD_BSDF | CY_SD_BSDF_HAS_EVAL;
          }
        }
      }

      /* sheen (svm_closure.h:256-279).  Full interpreter only: the host routes every
       * program whose Principled sheen input is not a constant zero to it (svm_validate,
       * SVM_USES_EXTENDED_NODES), so the lean one never sees sheen - and a test for it
       * here, of all places, cost the lean kernels 2 % (register allocation). */
      if (FULL && diffuse_weight > CLOSURE_WEIGHT_CUTOFF && sheen > CLOSURE_WEIGHT_CUTOFF) {
        const float m_cdlum = dot(base_color, mk3(kd_float(KD_FILM_RGB_TO_Y),
                                                  kd_float(KD_FILM_RGB_TO_Y + 4),
                                                  kd_float(KD_FILM_RGB_TO_Y + 8)));
        const f3 m_ctint = m_cdlum > 0.0f ? base_color / m_cdlum : one3();
        const f3 sheen_color = one3() * (1.0f - sheen_tint) + m_ctint * sheen_tint;
        Closure *bsdf = bsdf_alloc(sd, weight * sheen * sheen_color * diffuse_weight);
        if (bsdf) {
          bsdf->N = N;
          /* bsdf_principled_sheen_setup (closure/bsdf_principled_sheen.h:67-73) */
          bsdf->type = CY_CLOSURE_BSDF_PRINCIPLED_SHEEN_ID;
          const float NdotI = dot(N, sd.I);
          bsdf->sample_weight *= (NdotI < 0.0f) ? 0.0f : schlick_fresnel(NdotI) * NdotI;
          sd.flag |= CY_SD_BSDF | CY_SD_BSDF_HAS_EVAL;
        }
      }

      /* specular reflection */
      if (kd_int(KD_INT_CAUSTICS_REFLECTIVE) || (path_flag & CY_PATH_RAY_DIFFUSE) == 0) {
        if (specular_weight > CLOSURE_WEIGHT_CUTOFF &&
            (specular > CLOSURE_WEIGHT_CUTOFF || metallic > CLOSURE_WEIGHT_CUTOFF)) {
          f3 spec_weight = weight * specular_weight;
          Closure *bsdf = microfacet_alloc(sd, spec_weight, true);
          if (bsdf) {
            bsdf->N = N;
            bsdf->ior = (2.0f / (1.0f - safe_sqrtf(0.08f * specular))) - 1.0f;
            bsdf->T = T;
            float aspect = safe_sqrtf(1.0f - anisotropic * 0.9f);
            float r2 = roughness * roughness;
            bsdf->alpha_x = r2 / aspect;
            bsdf->alpha_y = r2 * aspect;
            float m_cdlum = 0.3f * base_color.x + 0.6f * base_color.y + 0.1f * base_color.z;
            f3 m_ctint = m_cdlum > 0.0f ? base_color / m_cdlum : zero3();
            f3 tmp_col = one3() * (1.0f - specular_tint) + m_ctint * specular_tint;
            bsdf->cspec0 = (specular * 0.08f * tmp_col) * (1.0f - metallic) + base_color * metallic;
            bsdf->color = base_color;
            bsdf->clearcoat = 0.0f;
            /* distribution is GGX (svm_validate refuses multiscatter unless roughness
             * takes the single-scatter branch) */
            sd.flag |= bsdf_microfacet_ggx_fresnel_setup(bsdf, sd);
          }
        }
      }

      /* transmission */
      if (kd_int(KD_INT_CAUSTICS_REFLECTIVE) || kd_int(KD_INT_CAUSTICS_REFRACTIVE) ||
          (path_flag & CY_PATH_RAY_DIFFUSE) == 0) {
        if (final_transmission > CLOSURE_WEIGHT_CUTOFF) {
          f3 glass_weight = weight * final_transmission;
          f3 cspec0 = base_color * specular_tint + one3() * (1.0f - specular_tint);
          float refl_roughness = roughness;
          if (kd_int(KD_INT_CAUSTICS_REFLECTIVE) || (path_flag & CY_PATH_RAY_DIFFUSE) == 0) {
            Closure *bsdf = microfacet_alloc(sd, glass_weight * fresnel, true);
            if (bsdf) {
              bsdf->N = N;
              bsdf->T = zero3();
              bsdf->alpha_x = refl_roughness * refl_roughness;
              bsdf->alpha_y = refl_roughness * refl_roughness;
              bsdf->ior = ior;
              bsdf->color = base_color;
              bsdf->cspec0 = cspec0;
              bsdf->clearcoat = 0.0f;
              sd.flag |= bsdf_microfacet_ggx_fresnel_setup(bsdf, sd);
            }
          }
          if (kd_int(KD_INT_CAUSTICS_REFRACTIVE) || (path_flag & CY_PATH_RAY_DIFFUSE) == 0) {
            Closure *bsdf = bsdf_alloc(sd, base_color * glass_weight * (1.0f - fresnel));
            if (bsdf) {
              bsdf->N = N;
              bsdf->T = zero3();
              if (distribution == CY_CLOSURE_BSDF_MICROFACET_GGX_GLASS_ID)
                transmission_roughness = 1.0f - (1.0f - refl_roughness) *
                                                    (1.0f - transmission_roughness);
              else
                transmission_roughness = refl_roughness;
              bsdf->alpha_x = transmission_roughness * transmission_roughness;
              bsdf->alpha_y = transmission_roughness * transmission_roughness;
              bsdf->ior = ior;
              sd.flag |= bsdf_microfacet_ggx_refraction_setup(bsdf);
            }
          }
        }
      }

      /* clearcoat */
      if (kd_int(KD_INT_CAUSTICS_REFLECTIVE) || (path_flag & CY_PATH_RAY_DIFFUSE) == 0) {
        if (clearcoat > CLOSURE_WEIGHT_CUTOFF) {
          Closure *bsdf = microfacet_alloc(sd, weight, true);
          if (bsdf) {
            bsdf->N = clearcoat_normal;
            bsdf->T = zero3();
            bsdf->ior = 1.5f;
            bsdf->alpha_x = clearcoat_roughness * clearcoat_roughness;
            bsdf->alpha_y = clearcoat_roughness * clearcoat_roughness;
            bsdf->color = zero3();
            bsdf->cspec0 = mk3(0.04f, 0.04f, 0.04f);
            bsdf->clearcoat = clearcoat;
            sd.flag |= bsdf_microfacet_ggx_clearcoat_setup(bsdf, sd);
          }
        }
      }
      break;
    }
    case CY_CLOSURE_BSDF_DIFFUSE_ID: {
      /* svm_closure.h:465-483: Lambert, or Oren-Nayar when the node has roughness */
      f3 weight = sd.svm_closure_weight * mix_weight;
      Closure *bsdf = bsdf_alloc(sd, weight);
      if (bsdf) {
        bsdf->N = N;
        if (param1 == 0.0f) {
          bsdf->type = CY_CLOSURE_BSDF_DIFFUSE_ID;
          sd.flag |= CY_SD_BSDF | CY_SD_BSDF_HAS_EVAL;
        }
        else {
          sd.flag |= bsdf_oren_nayar_setup(bsdf, param1);
        }
      }
      break;
    }
    case CY_CLOSURE_BSDF_TRANSLUCENT_ID: {
      /* svm_closure.h:484-493 */
      f3 weight = sd.svm_closure_weight * mix_weight;
      Closure *bsdf = bsdf_alloc(sd, weight);
      if (bsdf) {
        bsdf->N = N;
        bsdf->type = CY_CLOSURE_BSDF_TRANSLUCENT_ID;
        sd.flag |= CY_SD_BSDF | CY_SD_BSDF_HAS_EVAL;
      }
      break;
    }
    case CY_CLOSURE_BSDF_TRANSPARENT_ID: {
      /* bsdf_transparent_setup - closure/bsdf_transparent.h:38-85: all transparent
       * closures of a shader merge into one; the summed weight is what shadow rays
       * multiply by (shader_bsdf_transparency) */
      const f3 weight = sd.svm_closure_weight * mix_weight;
      const float sample_weight = fabsf(average(weight));
      if (!(sample_weight >= CLOSURE_WEIGHT_CUTOFF))
        break;
      if (sd.flag & CY_SD_TRANSPARENT) {
        sd.closure_transparent_extinction += weight;
        for (int i = 0; i < sd.num_closure; i++) {
          Closure &sc = sd.closure[i];
          if (sc.type == CY_CLOSURE_BSDF_TRANSPARENT_ID) {
            sc.weight += weight;
            sc.sample_weight += sample_weight;
            break;
          }
        }
      }
      else {
        sd.flag |= CY_SD_BSDF | CY_SD_TRANSPARENT;
        sd.closure_transparent_extinction = weight;
        /* a terminating path evaluates no closures, but still has to pass through */
        if (path_flag & CY_PATH_RAY_TERMINATE)
          sd.num_closure_left = 1;
        Closure *bsdf = closure_alloc(sd, weight);
        if (bsdf) {
          bsdf->type = CY_CLOSURE_BSDF_TRANSPARENT_ID;
          bsdf->sample_weight = sample_weight;
          bsdf->N = sd.N;
        }
        else if (path_flag & CY_PATH_RAY_TERMINATE) {
          sd.num_closure_left = 0;
        }
      }
      break;
    }
    case CY_CLOSURE_BSDF_REFRACTION_ID:
    case CY_CLOSURE_BSDF_MICROFACET_GGX_REFRACTION_ID: {
      /* svm_closure.h:571-608: Refraction BSDF node, sharp or GGX */
      if (!kd_int(KD_INT_CAUSTICS_REFRACTIVE) && (path_flag & CY_PATH_RAY_DIFFUSE))
        break;
      f3 weight = sd.svm_closure_weight * mix_weight;
      Closure *bsdf = bsdf_alloc(sd, weight);
      if (bsdf) {
        bsdf->N = N;
        bsdf->T = zero3();
        float eta = fmaxf(param2, 1e-5f);
        eta = (sd.flag & CY_SD_BACKFACING) ? 1.0f / eta : eta;
        if (type == CY_CLOSURE_BSDF_REFRACTION_ID) {
          bsdf->alpha_x = 0.0f;
          bsdf->alpha_y = 0.0f;
          bsdf->ior = eta;
          bsdf->type = CY_CLOSURE_BSDF_REFRACTION_ID;
          sd.flag |= CY_SD_BSDF;
        }
        else {
          float roughness = sqr(param1);
          bsdf->alpha_x = roughness;
          bsdf->alpha_y = roughness;
          bsdf->ior = eta;
          sd.flag |= bsdf_microfacet_ggx_refraction_setup(bsdf);
        }
      }
      break;
    }
    case CY_CLOSURE_BSDF_SHARP_GLASS_ID:
    case CY_CLOSURE_BSDF_MICROFACET_GGX_GLASS_ID: {
      /* svm_closure.h:609-660 + svm_node_glass_setup :25-58: Glass BSDF node = a
       * reflection and a refraction closure weighted by the dielectric Fresnel term */
      const bool refl_ok = kd_int(KD_INT_CAUSTICS_REFLECTIVE) != 0;
      const bool refr_ok = kd_int(KD_INT_CAUSTICS_REFRACTIVE) != 0;
      if (!refl_ok && !refr_ok && (path_flag & CY_PATH_RAY_DIFFUSE))
        break;
      f3 weight = sd.svm_closure_weight * mix_weight;
      float eta = fmaxf(param2, 1e-5f);
      eta = (sd.flag & CY_SD_BACKFACING) ? 1.0f / eta : eta;
      float cosNO = dot(N, sd.I);
      float fresnel = fresnel_dielectric_cos(cosNO, eta);
      float roughness = sqr(param1);
      const bool sharp = (type == CY_CLOSURE_BSDF_SHARP_GLASS_ID);
      if (refl_ok || (path_flag & CY_PATH_RAY_DIFFUSE) == 0) {
        Closure *bsdf = bsdf_alloc(sd, weight * fresnel);
        if (bsdf) {
          bsdf->N = N;
          bsdf->T = zero3();
          if (sharp) {
            bsdf->alpha_x = bsdf->alpha_y = 0.0f;
            bsdf->ior = 0.0f;
            bsdf->type = CY_CLOSURE_BSDF_REFLECTION_ID;
            sd.flag |= CY_SD_BSDF;
          }
          else {
            bsdf->alpha_x = bsdf->alpha_y = roughness;
            bsdf->ior = eta;
            sd.flag |= bsdf_microfacet_ggx_setup(bsdf);
          }
        }
      }
      if (refr_ok || (path_flag & CY_PATH_RAY_DIFFUSE) == 0) {
        Closure *bsdf = bsdf_alloc(sd, weight * (1.0f - fresnel));
        if (bsdf) {
          bsdf->N = N;
          bsdf->T = zero3();
          if (sharp) {
            bsdf->alpha_x = bsdf->alpha_y = 0.0f;
            bsdf->ior = eta;
            bsdf->type = CY_CLOSURE_BSDF_REFRACTION_ID;
            sd.flag |= CY_SD_BSDF;
          }
          else {
            bsdf->alpha_x = bsdf->alpha_y = roughness;
            bsdf->ior = eta;
            sd.flag |= bsdf_microfacet_ggx_refraction_setup(bsdf);
          }
        }
      }
      break;
    }
    case CY_CLOSURE_BSDF_REFLECTION_ID:
    case CY_CLOSURE_BSDF_MICROFACET_GGX_ID: {
      /* svm_closure.h:500-560: Glossy BSDF node, sharp or GGX (isotropic) */
      if (!kd_int(KD_INT_CAUSTICS_REFLECTIVE) && (path_flag & CY_PATH_RAY_DIFFUSE))
        break;
      f3 weight = sd.svm_closure_weight * mix_weight;
      Closure *bsdf = bsdf_alloc(sd, weight);
      if (bsdf) {
        float roughness = sqr(param1);
        bsdf->N = N;
        bsdf->ior = 0.0f;
        bsdf->alpha_x = roughness;
        bsdf->alpha_y = roughness;
        bsdf->T = zero3();
        /* the Anisotropic BSDF node: tangent, rotation and anisotropy (svm_closure.h:
         * 526-545).  Full interpreter only - the host routes programs with a tangent
         * input here (svm_validate), the lean one never sees them. */
        if (FULL && stack_valid(data_node.y)) {
          bsdf->T = stack_load_float3(stack, data_node.y);
          const float rotation = stack[data_node.z];
          if (rotation != 0.0f)
            bsdf->T = rotate_around_axis(bsdf->T, bsdf->N, rotation * CY_M_2PI_F);
          const float anisotropy = clampf(param2, -0.99f, 0.99f);
          if (anisotropy < 0.0f) {
            bsdf->alpha_x = roughness / (1.0f + anisotropy);
            bsdf->alpha_y = roughness * (1.0f + anisotropy);
          }
          else {
            bsdf->alpha_x = roughness * (1.0f - anisotropy);
            bsdf->alpha_y = roughness / (1.0f - anisotropy);
          }
        }
        if (type == CY_CLOSURE_BSDF_REFLECTION_ID) {
          bsdf->type = CY_CLOSURE_BSDF_REFLECTION_ID;
          sd.flag |= CY_SD_BSDF;
        }
        else {
          sd.flag |= bsdf_microfacet_ggx_setup(bsdf);
        }
      }
      break;
    }
    default:
      break; /* refused at bind time */
  }
  return true;
}

#endif

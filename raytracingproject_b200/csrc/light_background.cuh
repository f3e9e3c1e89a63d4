/* light_background.cuh - the world as a light: importance sampling of the background by
 * the luminance map the host builds (LightManager::device_update_background,
 * render/light.cpp:560-720, from one DeviceTask::SHADER evaluation of the world shader -
 * k_background_evaluate in wavefront.cuh), and the pdf of a direction under that sampling
 * for the MIS weight of background hits.
 *
 * Semantics to match: kernel/kernel_light_background.h
 *   :25-105   background_map_sample   (marginal CDF over rows, conditional CDF in a row)
 *   :107-138  background_map_pdf
 *   :302-398  background_light_sample (strategy mix by weight)
 *   :400-445  background_light_pdf
 * Of the three strategies the map is in scope; portals and the Nishita sun disc are
 * refused at bind time (check_scope), so their weights are zero here.
 * The two CDF arrays arrive by name (__light_background_marginal_cdf, _conditional_cdf),
 * float2 entries (function value, CDF), the last entry of a row holding the row total. */
#ifndef B200_LIGHT_BACKGROUND_CUH
#define B200_LIGHT_BACKGROUND_CUH

/* kernel_projection.h:56-83, with this tree's range (-2 pi, pi, -pi, pi) */
CY_DEV f3 equirectangular_to_direction(float u, float v)
{
  const float phi = -CY_M_2PI_F * u + CY_M_PI_F;
  const float theta = -CY_M_PI_F * v + CY_M_PI_F;
  const float sin_theta = sinf(theta);
  return mk3(sin_theta * cosf(phi), sin_theta * sinf(phi), cosf(theta));
}
CY_DEV float2 direction_to_equirectangular(f3 dir)
{
  if (is_zero(dir))
    return make_float2(0.0f, 0.0f);
  return make_float2((atan2f(dir.y, dir.x) - CY_M_PI_F) / -CY_M_2PI_F,
                     (acosf(dir.z / len(dir)) - CY_M_PI_F) / -CY_M_PI_F);
}

/* std::lower_bound over the .y (CDF) member of `count` float2 entries */
CY_DEV int cdf_lower_bound(const float2 *cdf, int count, float value)
{
  int first = 0;
  while (count > 0) {
    const int step = count >> 1;
    const int middle = first + step;
    if (__ldg(&cdf[middle]).y < value) {
      first = middle + 1;
      count -= step + 1;
    }
    else {
      count = step;
    }
  }
  return first;
}

CY_DEV f3 background_map_sample(float randu, float randv, float *pdf)
{
  const int res_x = kd_int(KD_BG_MAP_RES_X), res_y = kd_int(KD_BG_MAP_RES_Y);
  const int cdf_width = res_x + 1;
  const float2 *marg = g_scene.light_background_marginal_cdf;
  const float2 *cond = g_scene.light_background_conditional_cdf;

  const int index_v = max(0, cdf_lower_bound(marg, res_y, randv) - 1);
  const float2 cdf_v = __ldg(&marg[index_v]);
  const float2 cdf_next_v = __ldg(&marg[index_v + 1]);
  const float2 cdf_last_v = __ldg(&marg[res_y]);
  const float dv = (randv - cdf_v.y) / (cdf_next_v.y - cdf_v.y);
  const float v = ((float)index_v + dv) / (float)res_y;

  const float2 *row = cond + (size_t)index_v * cdf_width;
  const int index_u = max(0, cdf_lower_bound(row, res_x, randu) - 1);
  const float2 cdf_u = __ldg(&row[index_u]);
  const float2 cdf_next_u = __ldg(&row[index_u + 1]);
  const float2 cdf_last_u = __ldg(&row[res_x]);
  const float du = (randu - cdf_u.y) / (cdf_next_u.y - cdf_u.y);
  const float u = ((float)index_u + du) / (float)res_x;

  const float sin_theta = sinf(CY_M_PI_F * v);
  const float denom = (CY_M_2PI_F * CY_M_PI_F * sin_theta) * cdf_last_u.x * cdf_last_v.x;
  *pdf = (sin_theta == 0.0f || denom == 0.0f) ? 0.0f : (cdf_u.x * cdf_v.x) / denom;
  return equirectangular_to_direction(u, v);
}

CY_DEV float background_map_pdf(f3 direction)
{
  const float2 uv = direction_to_equirectangular(direction);
  const int res_x = kd_int(KD_BG_MAP_RES_X), res_y = kd_int(KD_BG_MAP_RES_Y);
  const int cdf_width = res_x + 1;
  const float sin_theta = sinf(uv.y * CY_M_PI_F);
  if (sin_theta == 0.0f)
    return 0.0f;
  const int index_u = min(max((int)(uv.x * (float)res_x), 0), res_x - 1);
  const int index_v = min(max((int)(uv.y * (float)res_y), 0), res_y - 1);
  const float2 *marg = g_scene.light_background_marginal_cdf;
  const float2 *row = g_scene.light_background_conditional_cdf + (size_t)index_v * cdf_width;
  const float2 cdf_last_u = __ldg(&row[res_x]);
  const float2 cdf_last_v = __ldg(&marg[res_y]);
  const float denom = (CY_M_2PI_F * CY_M_PI_F * sin_theta) * cdf_last_u.x * cdf_last_v.x;
  if (denom == 0.0f)
    return 0.0f;
  const float2 cdf_u = __ldg(&row[index_u]);
  const float2 cdf_v = __ldg(&marg[index_v]);
  return (cdf_u.x * cdf_v.x) / denom;
}

/* kernel_montecarlo.h sample_uniform_sphere */
CY_DEV f3 bg_sample_uniform_sphere(float u1, float u2)
{
  const float z = 1.0f - 2.0f * u1;
  const float r = sqrtf(fmaxf(0.0f, 1.0f - z * z));
  const float phi = CY_M_2PI_F * u2;
  return mk3(r * cosf(phi), r * sinf(phi), z);
}

/* background_light_sample with the map as the only strategy that can be active */
CY_DEV f3 background_light_sample(float randu, float randv, float *pdf)
{
  const float map_method_pdf = kd_float(KD_BG_MAP_WEIGHT);
  if (map_method_pdf == 0.0f) {
    *pdf = 1.0f / CY_M_4PI_F;
    return bg_sample_uniform_sphere(randu, randv);
  }
  /* the weights are normalised to 1: the map is sampled alone, no MIS between strategies */
  return background_map_sample(randu, randv, pdf);
}

CY_DEV float background_light_pdf(f3 direction)
{
  const float map_method_pdf = kd_float(KD_BG_MAP_WEIGHT);
  if (map_method_pdf == 0.0f)
    return kd_float(KD_INT_PDF_LIGHTS) / CY_M_4PI_F;
  /* pdf_fac = 1 / map_weight, map_method_pdf * pdf_fac: the reference's own arithmetic */
  const float pdf_fac = 1.0f / map_method_pdf;
  const float w = map_method_pdf * pdf_fac;
  const float pdf = background_map_pdf(direction) * w;
  return pdf * kd_float(KD_INT_PDF_LIGHTS);
}

#endif /* B200_LIGHT_BACKGROUND_CUH */

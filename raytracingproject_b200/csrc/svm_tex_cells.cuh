/* svm_tex_cells.cuh - the Musgrave and Voronoi textures (svm/svm_musgrave.h,
 * svm/svm_voronoi.h), included at the end of svm_tex.cuh (uses its Perlin noise and
 * hashes).  The reference spells every function out four times, once per dimension
 * count; here a point is a float[4] with `dims` live components and each algorithm is
 * written once.  What must not change is kept: the order the neighbour cells are
 * visited in (x fastest, w slowest - it decides ties and the smooth-F1 accumulation),
 * the per-dimension hash recipes of util_hash.h:165-216 and the operation order of the
 * distance metrics. */
#ifndef B200_SVM_TEX_CELLS_CUH
#define B200_SVM_TEX_CELLS_CUH

/* --------------------------------------------------------------- Musgrave */

/* noise_musgrave_{fBm, multi_fractal, hetero_terrain, hybrid_multi_fractal,
 * ridged_multi_fractal}_{1..4}d - svm_musgrave.h:28-700 */
__device__ __noinline__ float musgrave_nd(uint32_t type, const float *co, int dims, float H,
                                          float lacunarity, float octaves, float offset,
                                          float gain)
{
  float p[4];
  for (int d = 0; d < dims; d++)
    p[d] = co[d];
  const float pwHL = powf(lacunarity, -H);
  const int n = (int)octaves;
  const float rmd = octaves - floorf(octaves);
#define MUSGRAVE_NEXT_OCTAVE \
  for (int d = 0; d < dims; d++) \
  p[d] *= lacunarity

  switch (type) {
    case CY_NODE_MUSGRAVE_FBM: {
      float value = 0.0f, pwr = 1.0f;
      for (int i = 0; i < n; i++) {
        value += snoise_nd(p, dims) * pwr;
        pwr *= pwHL;
        MUSGRAVE_NEXT_OCTAVE;
      }
      if (rmd != 0.0f)
        value += rmd * snoise_nd(p, dims) * pwr;
      return value;
    }
    case CY_NODE_MUSGRAVE_MULTIFRACTAL: {
      float value = 1.0f, pwr = 1.0f;
      for (int i = 0; i < n; i++) {
        value *= (pwr * snoise_nd(p, dims) + 1.0f);
        pwr *= pwHL;
        MUSGRAVE_NEXT_OCTAVE;
      }
      if (rmd != 0.0f)
        value *= (rmd * pwr * snoise_nd(p, dims) + 1.0f);
      return value;
    }
    case CY_NODE_MUSGRAVE_HETERO_TERRAIN: {
      float pwr = pwHL;
      float value = offset + snoise_nd(p, dims); /* first octave unscaled */
      MUSGRAVE_NEXT_OCTAVE;
      for (int i = 1; i < n; i++) {
        const float increment = (snoise_nd(p, dims) + offset) * pwr * value;
        value += increment;
        pwr *= pwHL;
        MUSGRAVE_NEXT_OCTAVE;
      }
      if (rmd != 0.0f) {
        const float increment = (snoise_nd(p, dims) + offset) * pwr * value;
        value += rmd * increment;
      }
      return value;
    }
    case CY_NODE_MUSGRAVE_HYBRID_MULTIFRACTAL: {
      float pwr = pwHL;
      float value = snoise_nd(p, dims) + offset;
      float weight = gain * value;
      MUSGRAVE_NEXT_OCTAVE;
      for (int i = 1; (weight > 0.001f) && (i < n); i++) {
        if (weight > 1.0f)
          weight = 1.0f;
        const float signal = (snoise_nd(p, dims) + offset) * pwr;
        pwr *= pwHL;
        value += weight * signal;
        weight *= gain * signal;
        MUSGRAVE_NEXT_OCTAVE;
      }
      if (rmd != 0.0f)
        value += rmd * ((snoise_nd(p, dims) + offset) * pwr);
      return value;
    }
    case CY_NODE_MUSGRAVE_RIDGED_MULTIFRACTAL: {
      float pwr = pwHL;
      float signal = offset - fabsf(snoise_nd(p, dims));
      signal *= signal;
      float value = signal;
      for (int i = 1; i < n; i++) {
        MUSGRAVE_NEXT_OCTAVE;
        const float weight = saturate(signal * gain);
        signal = offset - fabsf(snoise_nd(p, dims));
        signal *= signal;
        signal *= weight;
        value += signal * pwr;
        pwr *= pwHL;
      }
      return value;
    }
    default:
      return 0.0f;
  }
#undef MUSGRAVE_NEXT_OCTAVE
}

/* svm_musgrave.h:703-847 */
SVM_TEX_FN void svm_node_tex_musgrave(float *stack, uint4 node, int *offset)
{
  const uint32_t type = node.y & 0xff, dims = (node.y >> 8) & 0xff,
                 co_off = (node.y >> 16) & 0xff, w_off = (node.y >> 24) & 0xff;
  const uint32_t scale_off = node.z & 0xff, detail_off = (node.z >> 8) & 0xff,
                 dimension_off = (node.z >> 16) & 0xff, lacunarity_off = (node.z >> 24) & 0xff;
  const uint32_t offset_off = node.w & 0xff, gain_off = (node.w >> 8) & 0xff,
                 fac_off = (node.w >> 16) & 0xff;
  const uint4 defaults1 = __ldg(&g_scene.svm_nodes[*offset]);
  const uint4 defaults2 = __ldg(&g_scene.svm_nodes[*offset + 1]);
  *offset += 2;

  const f3 co = stack_load_float3(stack, co_off);
  const float w = stack_load_default(stack, w_off, defaults1.x);
  const float scale = stack_load_default(stack, scale_off, defaults1.y);
  float detail = stack_load_default(stack, detail_off, defaults1.z);
  float dimension = stack_load_default(stack, dimension_off, defaults1.w);
  float lacunarity = stack_load_default(stack, lacunarity_off, defaults2.x);
  const float foffset = stack_load_default(stack, offset_off, defaults2.y);
  const float gain = stack_load_default(stack, gain_off, defaults2.z);
  dimension = fmaxf(dimension, 1e-5f);
  detail = clampf(detail, 0.0f, 16.0f);
  lacunarity = fmaxf(lacunarity, 1e-5f);

  float fac = 0.0f;
  if (dims >= 1 && dims <= 4) {
    float p[4] = {co.x * scale, co.y * scale, co.z * scale, w * scale};
    if (dims == 1)
      p[0] = w * scale;
    fac = musgrave_nd(type, p, (int)dims, dimension, lacunarity, detail, foffset, gain);
  }
  stack[fac_off] = fac;
}

/* ---------------------------------------------------------------- Voronoi */

/* hash_float_to_float, hash_float2_to_float2, hash_float3_to_float3,
 * hash_float4_to_float4: the random point of a cell */
CY_DEV void voronoi_hash_point(const float *k, int dims, float *out)
{
  uint32_t b[4];
  for (int d = 0; d < dims; d++)
    b[d] = __float_as_uint(k[d]);
  const uint32_t one = __float_as_uint(1.0f), two = __float_as_uint(2.0f);
  if (dims == 1) {
    out[0] = hash_to_unit_float(hash_uint_n(b, 1));
  }
  else if (dims == 2) {
    const uint32_t k1[3] = {b[0], b[1], one};
    out[0] = hash_to_unit_float(hash_uint_n(b, 2));
    out[1] = hash_to_unit_float(hash_uint_n(k1, 3));
  }
  else if (dims == 3) {
    const uint32_t k1[4] = {b[0], b[1], b[2], one}, k2[4] = {b[0], b[1], b[2], two};
    out[0] = hash_to_unit_float(hash_uint_n(b, 3));
    out[1] = hash_to_unit_float(hash_uint_n(k1, 4));
    out[2] = hash_to_unit_float(hash_uint_n(k2, 4));
  }
  else {
    const uint32_t k1[4] = {b[3], b[0], b[1], b[2]}, k2[4] = {b[2], b[3], b[0], b[1]},
                   k3[4] = {b[1], b[2], b[3], b[0]};
    out[0] = hash_to_unit_float(hash_uint_n(b, 4));
    out[1] = hash_to_unit_float(hash_uint_n(k1, 4));
    out[2] = hash_to_unit_float(hash_uint_n(k2, 4));
    out[3] = hash_to_unit_float(hash_uint_n(k3, 4));
  }
}

/* hash_float_to_float3 .. hash_float4_to_float3: the colour of a cell */
CY_DEV f3 voronoi_hash_color(const float *k, int dims)
{
  uint32_t b[4];
  for (int d = 0; d < dims; d++)
    b[d] = __float_as_uint(k[d]);
  const uint32_t one = __float_as_uint(1.0f), two = __float_as_uint(2.0f);
  uint32_t k1[4], k2[4];
  int n12 = dims + 1;
  if (dims == 4) {
    k1[0] = b[2], k1[1] = b[0], k1[2] = b[3], k1[3] = b[1];
    k2[0] = b[3], k2[1] = b[2], k2[2] = b[1], k2[3] = b[0];
    n12 = 4;
  }
  else {
    for (int d = 0; d < dims; d++)
      k1[d] = k2[d] = b[d];
    k1[dims] = one;
    k2[dims] = two;
  }
  return mk3(hash_to_unit_float(hash_uint_n(b, dims)), hash_to_unit_float(hash_uint_n(k1, n12)),
             hash_to_unit_float(hash_uint_n(k2, n12)));
}

CY_DEV float voronoi_dot(const float *a, const float *b, int dims)
{
  float s = a[0] * b[0];
  for (int d = 1; d < dims; d++)
    s += a[d] * b[d];
  return s;
}

/* voronoi_distance_{1..4}d */
CY_DEV float voronoi_distance(const float *a, const float *b, int dims, uint32_t metric,
                              float exponent)
{
  float diff[4];
  for (int d = 0; d < dims; d++)
    diff[d] = a[d] - b[d];
  if (dims == 1)
    return fabsf(b[0] - a[0]);
  if (metric == CY_NODE_VORONOI_EUCLIDEAN)
    return sqrtf(voronoi_dot(diff, diff, dims));
  if (metric == CY_NODE_VORONOI_MANHATTAN) {
    float s = fabsf(diff[0]);
    for (int d = 1; d < dims; d++)
      s += fabsf(diff[d]);
    return s;
  }
  if (metric == CY_NODE_VORONOI_CHEBYCHEV) {
    /* max(x, max(y, max(z, w))) */
    float m = fabsf(diff[dims - 1]);
    for (int d = dims - 2; d >= 0; d--)
      m = fmaxf(fabsf(diff[d]), m);
    return m;
  }
  if (metric == CY_NODE_VORONOI_MINKOWSKI) {
    float s = powf(fabsf(diff[0]), exponent);
    for (int d = 1; d < dims; d++)
      s += powf(fabsf(diff[d]), exponent);
    return powf(s, 1.0f / exponent);
  }
  return 0.0f;
}

/* The neighbourhood of a cell: visit `index` of (2 r + 1)^dims, x fastest.  Writes the
 * integer offset of the neighbour as floats. */
CY_DEV void voronoi_cell_offset(int index, int radius, int dims, float *off)
{
  const int side = 2 * radius + 1;
  for (int d = 0; d < dims; d++) {
    off[d] = (float)(index % side - radius);
    index /= side;
  }
}
CY_DEV int voronoi_cell_count(int radius, int dims)
{
  const int side = 2 * radius + 1;
  int n = side;
  for (int d = 1; d < dims; d++)
    n *= side;
  return n;
}

/* offset + hash(cell + offset) * randomness */
CY_DEV void voronoi_point(const float *cell, const float *off, int dims, float randomness,
                          float *point)
{
  float key[4], h[4];
  for (int d = 0; d < dims; d++)
    key[d] = cell[d] + off[d];
  voronoi_hash_point(key, dims, h);
  for (int d = 0; d < dims; d++)
    point[d] = off[d] + h[d] * randomness;
}

struct VoronoiOut {
  float distance, radius;
  f3 color;
  float position[4];
};

/* voronoi_f1 / voronoi_f2 / voronoi_smooth_f1 / voronoi_distance_to_edge /
 * voronoi_n_sphere_radius, any dimension */
__device__ __noinline__ void voronoi_nd(const float *coord, int dims, uint32_t feature,
                                        uint32_t metric, float smoothness, float exponent,
                                        float randomness, VoronoiOut *out)
{
  float cell[4], local[4];
  for (int d = 0; d < dims; d++) {
    cell[d] = floorf(coord[d]);
    local[d] = coord[d] - cell[d];
  }
  out->distance = 0.0f;
  out->radius = 0.0f;
  out->color = zero3();
  for (int d = 0; d < 4; d++)
    out->position[d] = 0.0f;
  float off[4], point[4], key[4];

  if (feature == CY_NODE_VORONOI_F1 || feature == CY_NODE_VORONOI_F2) {
    float d1 = 8.0f, d2 = 8.0f;
    float off1[4] = {0, 0, 0, 0}, off2[4] = {0, 0, 0, 0};
    float pos1[4] = {0, 0, 0, 0}, pos2[4] = {0, 0, 0, 0};
    const int n = voronoi_cell_count(1, dims);
    for (int c = 0; c < n; c++) {
      voronoi_cell_offset(c, 1, dims, off);
      voronoi_point(cell, off, dims, randomness, point);
      const float dist = voronoi_distance(point, local, dims, metric, exponent);
      if (dist < d1) {
        d2 = d1;
        d1 = dist;
        for (int d = 0; d < dims; d++) {
          off2[d] = off1[d];
          off1[d] = off[d];
          pos2[d] = pos1[d];
          pos1[d] = point[d];
        }
      }
      else if (feature == CY_NODE_VORONOI_F2 && dist < d2) {
        d2 = dist;
        for (int d = 0; d < dims; d++) {
          off2[d] = off[d];
          pos2[d] = point[d];
        }
      }
    }
    const bool f2 = (feature == CY_NODE_VORONOI_F2);
    out->distance = f2 ? d2 : d1;
    for (int d = 0; d < dims; d++) {
      key[d] = cell[d] + (f2 ? off2[d] : off1[d]);
      out->position[d] = (f2 ? pos2[d] : pos1[d]) + cell[d];
    }
    out->color = voronoi_hash_color(key, dims);
  }
  else if (feature == CY_NODE_VORONOI_SMOOTH_F1) {
    float sd = 8.0f;
    f3 sc = zero3();
    float sp[4] = {0, 0, 0, 0};
    const int n = voronoi_cell_count(2, dims);
    for (int c = 0; c < n; c++) {
      voronoi_cell_offset(c, 2, dims, off);
      voronoi_point(cell, off, dims, randomness, point);
      const float dist = voronoi_distance(point, local, dims, metric, exponent);
      /* smoothstep(0, 1, x) */
      const float x = 0.5f + 0.5f * (sd - dist) / smoothness;
      float h;
      if (x < 0.0f)
        h = 0.0f;
      else if (x >= 1.0f)
        h = 1.0f;
      else {
        const float t = (x - 0.0f) / (1.0f - 0.0f);
        h = (3.0f - 2.0f * t) * (t * t);
      }
      float correction = smoothness * h * (1.0f - h);
      sd = (sd + h * (dist - sd)) - correction;
      correction /= 1.0f + 3.0f * smoothness;
      for (int d = 0; d < dims; d++)
        key[d] = cell[d] + off[d];
      const f3 cc = voronoi_hash_color(key, dims);
      sc = mk3((sc.x + h * (cc.x - sc.x)) - correction, (sc.y + h * (cc.y - sc.y)) - correction,
               (sc.z + h * (cc.z - sc.z)) - correction);
      for (int d = 0; d < dims; d++)
        sp[d] = (sp[d] + h * (point[d] - sp[d])) - correction;
    }
    out->distance = sd;
    out->color = sc;
    for (int d = 0; d < dims; d++)
      out->position[d] = cell[d] + sp[d];
  }
  else if (feature == CY_NODE_VORONOI_DISTANCE_TO_EDGE) {
    const int n = voronoi_cell_count(1, dims);
    float min_dist = 8.0f;
    if (dims == 1) {
      for (int c = 0; c < n; c++) {
        voronoi_cell_offset(c, 1, dims, off);
        voronoi_point(cell, off, dims, randomness, point);
        min_dist = fminf(fabsf(point[0] - local[0]), min_dist);
      }
      out->distance = min_dist;
      return;
    }
    float to_closest[4] = {0, 0, 0, 0}, to_point[4];
    for (int c = 0; c < n; c++) {
      voronoi_cell_offset(c, 1, dims, off);
      voronoi_point(cell, off, dims, randomness, point);
      for (int d = 0; d < dims; d++)
        to_point[d] = point[d] - local[d];
      const float dist = voronoi_dot(to_point, to_point, dims);
      if (dist < min_dist) {
        min_dist = dist;
        for (int d = 0; d < dims; d++)
          to_closest[d] = to_point[d];
      }
    }
    min_dist = 8.0f;
    for (int c = 0; c < n; c++) {
      voronoi_cell_offset(c, 1, dims, off);
      voronoi_point(cell, off, dims, randomness, point);
      float perp[4], mid[4];
      for (int d = 0; d < dims; d++) {
        to_point[d] = point[d] - local[d];
        perp[d] = to_point[d] - to_closest[d];
      }
      const float pp = voronoi_dot(perp, perp, dims);
      if (pp > 0.0001f) {
        const float inv_len = 1.0f / sqrtf(pp);
        for (int d = 0; d < dims; d++) {
          mid[d] = (to_closest[d] + to_point[d]) * (1.0f / 2.0f);
          perp[d] *= inv_len;
        }
        min_dist = fminf(min_dist, voronoi_dot(mid, perp, dims));
      }
    }
    out->distance = min_dist;
  }
  else if (feature == CY_NODE_VORONOI_N_SPHERE_RADIUS) {
    const int n = voronoi_cell_count(1, dims);
    float closest[4] = {0, 0, 0, 0}, closest_off[4] = {0, 0, 0, 0};
    float min_dist = 8.0f;
    for (int c = 0; c < n; c++) {
      voronoi_cell_offset(c, 1, dims, off);
      voronoi_point(cell, off, dims, randomness, point);
      const float dist = voronoi_distance(point, local, dims, CY_NODE_VORONOI_EUCLIDEAN, 0.0f);
      if (dist < min_dist) {
        min_dist = dist;
        for (int d = 0; d < dims; d++) {
          closest[d] = point[d];
          closest_off[d] = off[d];
        }
      }
    }
    min_dist = 8.0f;
    float second[4] = {0, 0, 0, 0};
    for (int c = 0; c < n; c++) {
      if (c == n / 2)
        continue; /* the zero offset */
      voronoi_cell_offset(c, 1, dims, off);
      for (int d = 0; d < dims; d++)
        off[d] += closest_off[d];
      voronoi_point(cell, off, dims, randomness, point);
      const float dist = voronoi_distance(closest, point, dims, CY_NODE_VORONOI_EUCLIDEAN, 0.0f);
      if (dist < min_dist) {
        min_dist = dist;
        for (int d = 0; d < dims; d++)
          second[d] = point[d];
      }
    }
    out->radius = voronoi_distance(second, closest, dims, CY_NODE_VORONOI_EUCLIDEAN, 0.0f) / 2.0f;
  }
}

/* svm_voronoi.h:901-1137 */
SVM_TEX_FN void svm_node_tex_voronoi(float *stack, uint4 node, int *offset)
{
  const uint32_t dims = node.y, feature = node.z, metric = node.w;
  const uint4 so = __ldg(&g_scene.svm_nodes[*offset]);
  const uint4 defaults = __ldg(&g_scene.svm_nodes[*offset + 1]);
  *offset += 2;
  const uint32_t coord_off = so.x & 0xff, w_off = (so.x >> 8) & 0xff,
                 scale_off = (so.x >> 16) & 0xff, smooth_off = (so.x >> 24) & 0xff;
  const uint32_t exponent_off = so.y & 0xff, random_off = (so.y >> 8) & 0xff,
                 distance_out = (so.y >> 16) & 0xff, color_out = (so.y >> 24) & 0xff;
  const uint32_t position_out = so.z & 0xff, w_out = (so.z >> 8) & 0xff,
                 radius_out = (so.z >> 16) & 0xff;

  f3 coord = stack_load_float3(stack, coord_off);
  float w = stack_load_default(stack, w_off, so.w);
  const float scale = stack_load_default(stack, scale_off, defaults.x);
  float smoothness = stack_load_default(stack, smooth_off, defaults.y);
  const float exponent = stack_load_default(stack, exponent_off, defaults.z);
  float randomness = stack_load_default(stack, random_off, defaults.w);
  randomness = clampf(randomness, 0.0f, 1.0f);
  smoothness = clampf(smoothness / 2.0f, 0.0f, 0.5f);
  w *= scale;
  coord *= scale;

  VoronoiOut r;
  r.distance = r.radius = 0.0f;
  r.color = zero3();
  r.position[0] = r.position[1] = r.position[2] = r.position[3] = 0.0f;
  f3 position = zero3();
  float w_result = 0.0f;
  if (dims >= 1 && dims <= 4) {
    float p[4] = {coord.x, coord.y, coord.z, w};
    if (dims == 1)
      p[0] = w;
    voronoi_nd(p, (int)dims, feature, metric, smoothness, exponent, randomness, &r);
    /* safe_divide of the position by the scale; a / b of a vector is a * (1 / b) */
    const float inv = (scale != 0.0f) ? 1.0f / scale : 0.0f;
    if (dims == 1) {
      w_result = (scale != 0.0f) ? r.position[0] / scale : 0.0f;
    }
    else {
      position = mk3(r.position[0] * inv, r.position[1] * inv,
                     (dims >= 3) ? r.position[2] * inv : 0.0f);
      if (dims == 4)
        w_result = r.position[3] * inv;
    }
  }
  if (stack_valid(distance_out))
    stack[distance_out] = r.distance;
  if (stack_valid(color_out))
    stack_store_float3(stack, color_out, r.color);
  if (stack_valid(position_out))
    stack_store_float3(stack, position_out, position);
  if (stack_valid(w_out))
    stack[w_out] = w_result;
  if (stack_valid(radius_out))
    stack[radius_out] = r.radius;
}

/* --------------------------------------------------------------- Blackbody */

/* svm_blackbody.h + svm_math_blackbody_color (svm_math_util.h:183-250): Rec.709 colour
 * of a black body, piecewise rational fit in six temperature bands.  The coefficients
 * are the reference's fit data: per band r = a/t + b t + c, g likewise, b a cubic. */
SVM_TEX_FN void svm_node_blackbody(float *stack, uint4 node)
{
  const float band_from[5] = {1167.0f, 1449.0f, 1902.0f, 3315.0f, 6365.0f};
  const float fit[6][10] = {
      {2.52432244e+03f, -1.06185848e-03f, 3.11067539e+00f, -7.50343014e+02f, 3.15679613e-04f,
       4.73464526e-01f, 0.0f, 0.0f, 0.0f, 0.0f},
      {3.37763626e+03f, -4.34581697e-04f, 1.64843306e+00f, -1.00402363e+03f, 1.29189794e-04f,
       9.08181524e-01f, 0.0f, 0.0f, 0.0f, 0.0f},
      {4.10671449e+03f, -8.61949938e-05f, 6.41423749e-01f, -1.22075471e+03f, 2.56245413e-05f,
       1.20753416e+00f, 0.0f, 0.0f, 0.0f, 0.0f},
      {4.66849800e+03f, 2.85655028e-05f, 1.29075375e-01f, -1.42546105e+03f, -4.01730887e-05f,
       1.44002695e+00f, -2.02524603e-11f, 1.79435860e-07f, -2.60561875e-04f, -1.41761141e-02f},
      {4.60124770e+03f, 2.89727618e-05f, 1.48001316e-01f, -1.18134453e+03f, -2.18913373e-05f,
       1.30656109e+00f, -2.22463426e-13f, -1.55078698e-08f, 3.81675160e-04f, -7.30646033e-01f},
      {3.78765709e+03f, 9.36026367e-06f, 3.98995841e-01f, -5.00279505e+02f, -4.59745390e-06f,
       1.09090465e+00f, 6.72595954e-13f, -2.73059993e-08f, 4.24068546e-04f, -7.52204323e-01f},
  };
  const float t = stack[node.y];
  f3 rgb;
  if (t >= 12000.0f) {
    rgb = mk3(0.826270103f, 0.994478524f, 1.56626022f);
  }
  else if (t < 965.0f) {
    rgb = mk3(4.70366907f, 0.0f, 0.0f);
  }
  else {
    int band = 0;
    for (int k = 0; k < 5; k++)
      band += (t >= band_from[k]) ? 1 : 0;
    const float *c = fit[band];
    const float t_inv = 1.0f / t;
    rgb = mk3(c[0] * t_inv + c[1] * t + c[2], c[3] * t_inv + c[4] * t + c[5],
              ((c[6] * t + c[7]) * t + c[8]) * t + c[9]);
  }
  stack_store_float3(stack, node.z, rgb);
}

/* -------------------------------------------------------------- Wavelength */

/* svm_wavelength.h:76-96: CIE 1931 2-degree standard observer (x, y, z bar, 380-780 nm in
 * 5 nm steps - the published colour matching functions, as the reference tabulates
 * them), interpolated, taken to scene-linear RGB with the film's XYZ matrix. */
SVM_TEX_FN void svm_node_wavelength(float *stack, uint4 node)
{
  const float cmf[81][3] = {
      {0.0014f, 0.0000f, 0.0065f}, {0.0022f, 0.0001f, 0.0105f}, {0.0042f, 0.0001f, 0.0201f},
      {0.0076f, 0.0002f, 0.0362f}, {0.0143f, 0.0004f, 0.0679f}, {0.0232f, 0.0006f, 0.1102f},
      {0.0435f, 0.0012f, 0.2074f}, {0.0776f, 0.0022f, 0.3713f}, {0.1344f, 0.0040f, 0.6456f},
      {0.2148f, 0.0073f, 1.0391f}, {0.2839f, 0.0116f, 1.3856f}, {0.3285f, 0.0168f, 1.6230f},
      {0.3483f, 0.0230f, 1.7471f}, {0.3481f, 0.0298f, 1.7826f}, {0.3362f, 0.0380f, 1.7721f},
      {0.3187f, 0.0480f, 1.7441f}, {0.2908f, 0.0600f, 1.6692f}, {0.2511f, 0.0739f, 1.5281f},
      {0.1954f, 0.0910f, 1.2876f}, {0.1421f, 0.1126f, 1.0419f}, {0.0956f, 0.1390f, 0.8130f},
      {0.0580f, 0.1693f, 0.6162f}, {0.0320f, 0.2080f, 0.4652f}, {0.0147f, 0.2586f, 0.3533f},
      {0.0049f, 0.3230f, 0.2720f}, {0.0024f, 0.4073f, 0.2123f}, {0.0093f, 0.5030f, 0.1582f},
      {0.0291f, 0.6082f, 0.1117f}, {0.0633f, 0.7100f, 0.0782f}, {0.1096f, 0.7932f, 0.0573f},
      {0.1655f, 0.8620f, 0.0422f}, {0.2257f, 0.9149f, 0.0298f}, {0.2904f, 0.9540f, 0.0203f},
      {0.3597f, 0.9803f, 0.0134f}, {0.4334f, 0.9950f, 0.0087f}, {0.5121f, 1.0000f, 0.0057f},
      {0.5945f, 0.9950f, 0.0039f}, {0.6784f, 0.9786f, 0.0027f}, {0.7621f, 0.9520f, 0.0021f},
      {0.8425f, 0.9154f, 0.0018f}, {0.9163f, 0.8700f, 0.0017f}, {0.9786f, 0.8163f, 0.0014f},
      {1.0263f, 0.7570f, 0.0011f}, {1.0567f, 0.6949f, 0.0010f}, {1.0622f, 0.6310f, 0.0008f},
      {1.0456f, 0.5668f, 0.0006f}, {1.0026f, 0.5030f, 0.0003f}, {0.9384f, 0.4412f, 0.0002f},
      {0.8544f, 0.3810f, 0.0002f}, {0.7514f, 0.3210f, 0.0001f}, {0.6424f, 0.2650f, 0.0000f},
      {0.5419f, 0.2170f, 0.0000f}, {0.4479f, 0.1750f, 0.0000f}, {0.3608f, 0.1382f, 0.0000f},
      {0.2835f, 0.1070f, 0.0000f}, {0.2187f, 0.0816f, 0.0000f}, {0.1649f, 0.0610f, 0.0000f},
      {0.1212f, 0.0446f, 0.0000f}, {0.0874f, 0.0320f, 0.0000f}, {0.0636f, 0.0232f, 0.0000f},
      {0.0468f, 0.0170f, 0.0000f}, {0.0329f, 0.0119f, 0.0000f}, {0.0227f, 0.0082f, 0.0000f},
      {0.0158f, 0.0057f, 0.0000f}, {0.0114f, 0.0041f, 0.0000f}, {0.0081f, 0.0029f, 0.0000f},
      {0.0058f, 0.0021f, 0.0000f}, {0.0041f, 0.0015f, 0.0000f}, {0.0029f, 0.0010f, 0.0000f},
      {0.0020f, 0.0007f, 0.0000f}, {0.0014f, 0.0005f, 0.0000f}, {0.0010f, 0.0004f, 0.0000f},
      {0.0007f, 0.0002f, 0.0000f}, {0.0005f, 0.0002f, 0.0000f}, {0.0003f, 0.0001f, 0.0000f},
      {0.0002f, 0.0001f, 0.0000f}, {0.0002f, 0.0001f, 0.0000f}, {0.0001f, 0.0000f, 0.0000f},
      {0.0001f, 0.0000f, 0.0000f}, {0.0001f, 0.0000f, 0.0000f}, {0.0000f, 0.0000f, 0.0000f},
  };
  const float lambda_nm = stack[node.y];
  float ii = (lambda_nm - 380.0f) * (1.0f / 5.0f);
  const int i = (int)ii;
  f3 xyz = zero3();
  if (i >= 0 && i < 80) {
    ii -= (float)i;
    const f3 a = mk3(cmf[i][0], cmf[i][1], cmf[i][2]);
    const f3 b = mk3(cmf[i + 1][0], cmf[i + 1][1], cmf[i + 1][2]);
    xyz = a + ii * (b - a);
  }
  f3 rgb = mk3(dot(mk3(kd_float4(KD_FILM_XYZ_TO_R)), xyz), dot(mk3(kd_float4(KD_FILM_XYZ_TO_G)), xyz),
               dot(mk3(kd_float4(KD_FILM_XYZ_TO_B)), xyz));
  rgb *= 1.0f / 2.52f; /* the reference's empirical scale: all components <= 1 */
  stack_store_float3(stack, node.z, mk3(fmaxf(rgb.x, 0.0f), fmaxf(rgb.y, 0.0f), fmaxf(rgb.z, 0.0f)));
}

#endif /* B200_SVM_TEX_CELLS_CUH */

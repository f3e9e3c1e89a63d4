/* svm_tex.cuh - mesh attributes, texture coordinates, mapping and the procedural
 * textures of the SVM interpreter.  Included by shade.cuh (needs ShaderDataG and the
 * stack helpers).  What each block restates:
 *
 *   attributes      kernel/geom/geom_attribute.h:48-86 (find_attribute),
 *                   geom/geom_triangle.h:110-356 (per-element fetch), svm/svm_attribute.h:21-90
 *   tangent         geom/geom_primitive.h:292-320
 *   tex coord       svm/svm_tex_coord.h:20-92 (object, normal, camera, window, reflection)
 *   mapping         svm/svm_mapping.h, svm_mapping_util.h, util_transform.h:151-175
 *   hashes          util/util_hash.h:27-165 (Jenkins lookup3 "final"/"mix")
 *   Perlin noise    svm/svm_noise.h:36-268 (the scalar formulation; the CPU reference
 *                   runs the SSE one, same lattice and gradients, sums in another order)
 *   fractal noise   svm/svm_fractal_noise.h, noise texture svm/svm_noisetex.h
 *   checker, gradient, wave, magic, brick   svm/svm_checker.h, svm_gradient.h, svm_wave.h,
 *                   svm_magic.h, svm_brick.h
 *
 * Subdivision patches (ATTR_PRIM_SUBD), curves and voxel attributes are outside the
 * hot-path scope; scenes carrying them are refused on the host (check_scope). */
#ifndef B200_SVM_TEX_CUH
#define B200_SVM_TEX_CUH

/* node entry points stay out of line: they are rare next to the closure nodes and
 * inlining them into the interpreter's switch costs it registers */
#ifndef SVM_TEX_FN
#  define SVM_TEX_FN __device__ __noinline__
#endif

/* ------------------------------------------------------------ attributes */

struct AttrDesc {
  uint32_t element; /* CY_ATTR_ELEMENT_* */
  uint32_t type;    /* CY_NODE_ATTR_* */
  uint32_t offset;  /* CY_ATTR_STD_NOT_FOUND when absent */
};

CY_DEV AttrDesc attribute_not_found()
{
  AttrDesc d;
  d.element = CY_ATTR_ELEMENT_NONE;
  d.type = 0;
  d.offset = CY_ATTR_STD_NOT_FOUND;
  return d;
}

CY_DEV AttrDesc find_attribute(const ShaderDataG &sd, uint32_t id)
{
  if (sd.object == -1)
    return attribute_not_found();
  uint32_t at = __ldg((const uint32_t *)(g_scene.objects +
                                         (size_t)sd.object * SIZEOF_KERNEL_OBJECT +
                                         KO_ATTRIBUTE_MAP_OFFSET));
  /* + ATTR_PRIM_GEOMETRY (0): no subdivision patches in scope */
  uint4 m = __ldg(&g_scene.attributes_map[at]);
  while (m.x != id) {
    if (m.x == CY_ATTR_STD_NONE)
      return attribute_not_found();
    at += CY_ATTR_PRIM_TYPES;
    m = __ldg(&g_scene.attributes_map[at]);
  }
  AttrDesc d;
  d.element = m.y;
  if (sd.prim == CY_PRIM_NONE && d.element != CY_ATTR_ELEMENT_MESH &&
      d.element != CY_ATTR_ELEMENT_OBJECT)
    return attribute_not_found();
  d.offset = (m.y == CY_ATTR_ELEMENT_NONE) ? CY_ATTR_STD_NOT_FOUND : m.z;
  d.type = m.w & 0xff;
  return d;
}

/* Which three array slots a triangle interpolates for this element; false = the
 * attribute is constant over the triangle (slot i0) or absent. */
CY_DEV bool attr_triangle_slots(const ShaderDataG &sd, const AttrDesc &d, uint32_t *i0,
                                uint32_t *i1, uint32_t *i2, bool *present)
{
  *present = true;
  if (d.element == CY_ATTR_ELEMENT_VERTEX || d.element == CY_ATTR_ELEMENT_VERTEX_MOTION) {
    const uint4 vi = __ldg(&g_scene.tri_vindex[sd.prim]);
    *i0 = d.offset + vi.x;
    *i1 = d.offset + vi.y;
    *i2 = d.offset + vi.z;
    return true;
  }
  if (d.element == CY_ATTR_ELEMENT_CORNER) {
    *i0 = d.offset + (uint32_t)sd.prim * 3u;
    *i1 = *i0 + 1;
    *i2 = *i0 + 2;
    return true;
  }
  if (d.element == CY_ATTR_ELEMENT_FACE)
    *i0 = d.offset + (uint32_t)sd.prim;
  else if (d.element == CY_ATTR_ELEMENT_OBJECT || d.element == CY_ATTR_ELEMENT_MESH)
    *i0 = d.offset;
  else
    *present = false;
  return false;
}

CY_DEV float attribute_float(const ShaderDataG &sd, const AttrDesc &d)
{
  uint32_t i0, i1, i2;
  bool present;
  if (attr_triangle_slots(sd, d, &i0, &i1, &i2, &present)) {
    const float f0 = __ldg(&g_scene.attributes_float[i0]);
    const float f1 = __ldg(&g_scene.attributes_float[i1]);
    const float f2 = __ldg(&g_scene.attributes_float[i2]);
    return sd.u * f0 + sd.v * f1 + (1.0f - sd.u - sd.v) * f2;
  }
  return present ? __ldg(&g_scene.attributes_float[i0]) : 0.0f;
}

CY_DEV float2 attribute_float2(const ShaderDataG &sd, const AttrDesc &d)
{
  uint32_t i0, i1, i2;
  bool present;
  if (attr_triangle_slots(sd, d, &i0, &i1, &i2, &present)) {
    const float2 f0 = __ldg(&g_scene.attributes_float2[i0]);
    const float2 f1 = __ldg(&g_scene.attributes_float2[i1]);
    const float2 f2 = __ldg(&g_scene.attributes_float2[i2]);
    const float w = 1.0f - sd.u - sd.v;
    return make_float2(sd.u * f0.x + sd.v * f1.x + w * f2.x, sd.u * f0.y + sd.v * f1.y + w * f2.y);
  }
  return present ? __ldg(&g_scene.attributes_float2[i0]) : make_float2(0.0f, 0.0f);
}

CY_DEV f3 attribute_float3(const ShaderDataG &sd, const AttrDesc &d)
{
  uint32_t i0, i1, i2;
  bool present;
  if (attr_triangle_slots(sd, d, &i0, &i1, &i2, &present)) {
    const f3 f0 = mk3(__ldg(&g_scene.attributes_float3[i0]));
    const f3 f1 = mk3(__ldg(&g_scene.attributes_float3[i1]));
    const f3 f2 = mk3(__ldg(&g_scene.attributes_float3[i2]));
    return sd.u * f0 + sd.v * f1 + (1.0f - sd.u - sd.v) * f2;
  }
  return present ? mk3(__ldg(&g_scene.attributes_float3[i0])) : zero3();
}

CY_DEV float4 uchar4_to_float4(uchar4 c)
{
  return make_float4(c.x * (1.0f / 255.0f), c.y * (1.0f / 255.0f), c.z * (1.0f / 255.0f),
                     c.w * (1.0f / 255.0f));
}

/* triangle_attribute_float4: byte corners or float4 vertices interpolate, a mesh /
 * object constant comes from the byte array, anything else is zero */
CY_DEV f3 attribute_rgba_rgb(const ShaderDataG &sd, const AttrDesc &d)
{
  float4 f0, f1, f2;
  if (d.element == CY_ATTR_ELEMENT_CORNER_BYTE) {
    const uint32_t tri = d.offset + (uint32_t)sd.prim * 3u;
    f0 = uchar4_to_float4(__ldg(&g_scene.attributes_uchar4[tri + 0]));
    f1 = uchar4_to_float4(__ldg(&g_scene.attributes_uchar4[tri + 1]));
    f2 = uchar4_to_float4(__ldg(&g_scene.attributes_uchar4[tri + 2]));
  }
  else if (d.element == CY_ATTR_ELEMENT_VERTEX) {
    const uint4 vi = __ldg(&g_scene.tri_vindex[sd.prim]);
    f0 = __ldg(&g_scene.attributes_float3[d.offset + vi.x]);
    f1 = __ldg(&g_scene.attributes_float3[d.offset + vi.y]);
    f2 = __ldg(&g_scene.attributes_float3[d.offset + vi.z]);
  }
  else if (d.element == CY_ATTR_ELEMENT_OBJECT || d.element == CY_ATTR_ELEMENT_MESH) {
    return mk3(uchar4_to_float4(__ldg(&g_scene.attributes_uchar4[d.offset])));
  }
  else {
    return zero3();
  }
  return sd.u * mk3(f0) + sd.v * mk3(f1) + (1.0f - sd.u - sd.v) * mk3(f2);
}

SVM_TEX_FN void svm_node_attr(const ShaderDataG &sd, float *stack, uint4 node)
{
  const uint32_t out = node.z, want = node.w;
  AttrDesc d = find_attribute(sd, node.y);
  if (d.offset == CY_ATTR_STD_NOT_FOUND) {
    d = attribute_not_found();
    d.offset = 0;
    d.type = want;
  }
  /* a descriptor that was not found keeps element NONE: every fetch returns zero */
  f3 v;
  float f;
  if (d.type == CY_NODE_ATTR_FLOAT) {
    f = attribute_float(sd, d);
    v = mk3(f, f, f);
  }
  else if (d.type == CY_NODE_ATTR_FLOAT2) {
    const float2 t = attribute_float2(sd, d);
    f = t.x;
    v = mk3(t.x, t.y, 0.0f);
  }
  else if (d.type == CY_NODE_ATTR_RGBA) {
    v = attribute_rgba_rgb(sd, d);
    f = average(v);
  }
  else {
    v = attribute_float3(sd, d);
    f = average(v);
  }
  if (want == CY_NODE_ATTR_FLOAT)
    stack[out] = f;
  else
    stack_store_float3(stack, out, v);
}

/* primitive_tangent: spherical tangent from the generated coordinates when the mesh
 * carries them, else the surface derivative */
CY_DEV f3 primitive_tangent(const ShaderDataG &sd)
{
  const AttrDesc d = find_attribute(sd, CY_ATTR_STD_GENERATED);
  if (d.offset != CY_ATTR_STD_NOT_FOUND) {
    f3 data = attribute_float3(sd, d);
    data = mk3(-(data.y - 0.5f), (data.x - 0.5f), 0.0f);
    data = object_normal_transform(sd.object, data);
    return cross(sd.N, normalize(cross(data, sd.N)));
  }
  return normalize(sd.dPdu);
}

/* -------------------------------------------------------- tex coordinates */

CY_DEV tfm34 node_transform(int *offset)
{
  tfm34 t;
  t.x = __ldg((const float4 *)&g_scene.svm_nodes[*offset + 0]);
  t.y = __ldg((const float4 *)&g_scene.svm_nodes[*offset + 1]);
  t.z = __ldg((const float4 *)&g_scene.svm_nodes[*offset + 2]);
  *offset += 3;
  return t;
}

CY_DEV tfm34 kd_transform(int off)
{
  tfm34 t;
  t.x = kd_float4(off);
  t.y = kd_float4(off + 16);
  t.z = kd_float4(off + 32);
  return t;
}

CY_DEV f3 camera_position()
{
  const tfm34 c2w = kd_transform(KD_CAM_CAMERATOWORLD);
  return mk3(c2w.x.w, c2w.y.w, c2w.z.w);
}

SVM_TEX_FN void svm_node_tex_coord(const ShaderDataG &sd, float *stack, uint4 node, int *offset)
{
  f3 data = zero3();
  switch (node.y) {
    case CY_NODE_TEXCO_OBJECT:
      data = sd.P;
      if (node.w == 0) {
        if (sd.object != -1)
          data = transform_point(object_itfm(sd.object), data);
      }
      else {
        data = transform_point(node_transform(offset), data);
      }
      break;
    case CY_NODE_TEXCO_NORMAL:
      data = sd.N;
      if (sd.object != -1) {
        data = normalize(transform_direction_transposed(object_tfm(sd.object), data));
      }
      else if (sd.lamp != -1) {
        /* PRIMITIVE_LAMP: sd->ob_tfm is the lamp's (kernel_shader.h:286-290) */
        const float4 *p = (const float4 *)(g_scene.lights +
                                           (size_t)sd.lamp * SIZEOF_KERNEL_LIGHT + KL_TFM);
        tfm34 t;
        t.x = __ldg(p + 0);
        t.y = __ldg(p + 1);
        t.z = __ldg(p + 2);
        data = normalize(transform_direction_transposed(t, data));
      }
      break;
    case CY_NODE_TEXCO_CAMERA: {
      const tfm34 w2c = kd_transform(KD_CAM_WORLDTOCAMERA);
      data = (sd.object != -1) ? transform_point(w2c, sd.P) :
                                 transform_point(w2c, sd.P + camera_position());
      break;
    }
    case CY_NODE_TEXCO_WINDOW: {
      /* camera_world_to_ndc (kernel_camera.h:476-485) for the perspective camera; with
       * an orthographic or panoramic camera the node is refused (check_scope) */
      f3 P = sd.P;
      if (sd.object == -1 && kd_int(KD_CAM_TYPE) == CY_CAMERA_PERSPECTIVE)
        P += camera_position();
      data = transform_perspective(kd_float4(KD_CAM_WORLDTONDC), kd_float4(KD_CAM_WORLDTONDC + 16),
                                   kd_float4(KD_CAM_WORLDTONDC + 32),
                                   kd_float4(KD_CAM_WORLDTONDC + 48), P);
      data.z = 0.0f;
      break;
    }
    case CY_NODE_TEXCO_REFLECTION:
      data = (sd.object != -1) ? 2.0f * dot(sd.N, sd.I) * sd.N - sd.I : sd.I;
      break;
    default:
      break; /* dupli / volume coordinates: refused by svm_validate */
  }
  stack_store_float3(stack, node.z, data);
}

/* ---------------------------------------------------------------- mapping */

CY_DEV tfm34 euler_to_transform(f3 e)
{
  const float cx = cosf(e.x), cy = cosf(e.y), cz = cosf(e.z);
  const float sx = sinf(e.x), sy = sinf(e.y), sz = sinf(e.z);
  tfm34 t;
  t.x = make_float4(cy * cz, sy * sx * cz - cx * sz, sy * cx * cz + sx * sz, 0.0f);
  t.y = make_float4(cy * sz, sy * sx * sz + cx * cz, sy * cx * sz - sx * cz, 0.0f);
  t.z = make_float4(-sy, cy * sx, cy * cx, 0.0f);
  return t;
}

CY_DEV f3 tex_safe_divide3(f3 a, f3 b)
{
  return mk3((b.x != 0.0f) ? a.x / b.x : 0.0f, (b.y != 0.0f) ? a.y / b.y : 0.0f,
             (b.z != 0.0f) ? a.z / b.z : 0.0f);
}

SVM_TEX_FN void svm_node_mapping(float *stack, uint4 node)
{
  const f3 vector = stack_load_float3(stack, node.z & 0xff);
  const f3 location = stack_load_float3(stack, (node.z >> 8) & 0xff);
  const f3 rotation = stack_load_float3(stack, (node.z >> 16) & 0xff);
  const f3 scale = stack_load_float3(stack, (node.z >> 24) & 0xff);
  const tfm34 rot = euler_to_transform(rotation);
  f3 r;
  switch (node.y) {
    case CY_NODE_MAPPING_TYPE_POINT:
      r = transform_direction(rot, vector * scale) + location;
      break;
    case CY_NODE_MAPPING_TYPE_TEXTURE:
      r = tex_safe_divide3(transform_direction_transposed(rot, vector - location), scale);
      break;
    case CY_NODE_MAPPING_TYPE_VECTOR:
      r = transform_direction(rot, vector * scale);
      break;
    case CY_NODE_MAPPING_TYPE_NORMAL:
      r = safe_normalize(transform_direction(rot, tex_safe_divide3(vector, scale)));
      break;
    default:
      r = zero3();
  }
  stack_store_float3(stack, node.w, r);
}

SVM_TEX_FN void svm_node_texture_mapping(float *stack, uint4 node, int *offset)
{
  const f3 v = stack_load_float3(stack, node.y);
  stack_store_float3(stack, node.z, transform_point(node_transform(offset), v));
}

SVM_TEX_FN void svm_node_min_max(float *stack, uint4 node, int *offset)
{
  const f3 v = stack_load_float3(stack, node.y);
  const f3 mn = mk3(__ldg((const float4 *)&g_scene.svm_nodes[*offset]));
  const f3 mx = mk3(__ldg((const float4 *)&g_scene.svm_nodes[*offset + 1]));
  *offset += 2;
  stack_store_float3(stack, node.z,
                     mk3(fminf(fmaxf(mn.x, v.x), mx.x), fminf(fmaxf(mn.y, v.y), mx.y),
                         fminf(fmaxf(mn.z, v.z), mx.z)));
}

CY_DEV float stack_load_default(const float *stack, uint32_t a, uint32_t bits)
{
  return stack_valid(a) ? stack[a] : __uint_as_float(bits);
}

/* ----------------------------------------------------------------- hashes */

CY_DEV uint32_t rotl32(uint32_t x, int k)
{
  return (x << k) | (x >> (32 - k));
}

/* lookup3 final() */
CY_DEV uint32_t jenkins_final(uint32_t a, uint32_t b, uint32_t c)
{
  c ^= b; c -= rotl32(b, 14);
  a ^= c; a -= rotl32(c, 11);
  b ^= a; b -= rotl32(a, 25);
  c ^= b; c -= rotl32(b, 16);
  a ^= c; a -= rotl32(c, 4);
  b ^= a; b -= rotl32(a, 14);
  c ^= b; c -= rotl32(b, 24);
  return c;
}

/* hash_uint / hash_uint2 / hash_uint3 / hash_uint4 as one function of the number of
 * keys: init 0xdeadbeef + (n << 2) + 13, keys added to a, b, c (a fourth key after one
 * lookup3 mix() round) */
CY_DEV uint32_t hash_uint_n(const uint32_t *k, int n)
{
  uint32_t a, b, c;
  a = b = c = 0xdeadbeefu + ((uint32_t)n << 2) + 13u;
  a += k[0];
  if (n > 1)
    b += k[1];
  if (n > 2)
    c += k[2];
  if (n > 3) {
    a -= c; a ^= rotl32(c, 4); c += b;
    b -= a; b ^= rotl32(a, 6); a += c;
    c -= b; c ^= rotl32(b, 8); b += a;
    a -= c; a ^= rotl32(c, 16); c += b;
    b -= a; b ^= rotl32(a, 19); a += c;
    c -= b; c ^= rotl32(b, 4); b += a;
    a += k[3];
  }
  return jenkins_final(a, b, c);
}

CY_DEV float hash_to_unit_float(uint32_t h)
{
  return (float)h / (float)0xFFFFFFFFu;
}

/* ------------------------------------------------------------ Perlin noise */

CY_DEV float perlin_fade(float t)
{
  return t * t * t * (t * (t * 6.0f - 15.0f) + 10.0f);
}

CY_DEV float perlin_neg(float v, uint32_t cond)
{
  return cond ? -v : v;
}

/* grad1..grad4 of svm_noise.h:47-52, 174-205; f = offset of the point from the corner */
CY_DEV float perlin_grad(uint32_t hash, const float *f, int dims)
{
  if (dims == 1) {
    const uint32_t h = hash & 15;
    const float g = (float)(1 + (h & 7));
    return perlin_neg(g, h & 8) * f[0];
  }
  if (dims == 2) {
    const uint32_t h = hash & 7;
    const float u = h < 4 ? f[0] : f[1];
    const float v = 2.0f * (h < 4 ? f[1] : f[0]);
    return perlin_neg(u, h & 1) + perlin_neg(v, h & 2);
  }
  if (dims == 3) {
    const uint32_t h = hash & 15;
    const float u = h < 8 ? f[0] : f[1];
    const float vt = ((h == 12) || (h == 14)) ? f[0] : f[2];
    const float v = h < 4 ? f[1] : vt;
    return perlin_neg(u, h & 1) + perlin_neg(v, h & 2);
  }
  const uint32_t h = hash & 31;
  const float u = h < 24 ? f[0] : f[1];
  const float v = h < 16 ? f[1] : f[2];
  const float s = h < 8 ? f[2] : f[3];
  return perlin_neg(u, h & 1) + perlin_neg(v, h & 2) + perlin_neg(s, h & 4);
}

/* perlin_1d .. perlin_4d: bit d of a corner index is the +1 step along axis d.  The
 * reduction runs x first, then y, z, w, with the same formulas as mix / bi_mix /
 * tri_mix / quad_mix. */
__device__ __noinline__ float perlin_nd(const float *p, int dims)
{
  int cell[4];
  float frac[4], fade[4];
  for (int d = 0; d < dims; d++) {
    const int i = (int)p[d] - ((p[d] < 0.0f) ? 1 : 0); /* quick_floor_to_int */
    cell[d] = i;
    frac[d] = p[d] - (float)i;
    fade[d] = perlin_fade(frac[d]);
  }
  float g[16];
  const int corners = 1 << dims;
  for (int c = 0; c < corners; c++) {
    uint32_t key[4];
    float off[4];
    for (int d = 0; d < dims; d++) {
      const int step = (c >> d) & 1;
      key[d] = (uint32_t)(cell[d] + step);
      off[d] = step ? frac[d] - 1.0f : frac[d];
    }
    g[c] = perlin_grad(hash_uint_n(key, dims), off, dims);
  }
  if (dims == 1)
    return g[0] + fade[0] * (g[1] - g[0]); /* mix() */
  const int lerp_dims = (dims == 4) ? 3 : dims;
  int n = corners;
  for (int d = 0; d < lerp_dims; d++) {
    n >>= 1;
    const float t = fade[d], t1 = 1.0f - fade[d];
    for (int i = 0; i < n; i++)
      g[i] = t1 * g[2 * i] + t * g[2 * i + 1];
  }
  if (dims == 4)
    return g[0] + fade[3] * (g[1] - g[0]); /* quad_mix: mix() of two tri_mix */
  return g[0];
}

/* snoise_*: Perlin scaled to [-1, 1] (svm_noise.h:676-741); noise_* = 0.5 s + 0.5 */
CY_DEV float snoise_nd(const float *p, int dims)
{
  float r = perlin_nd(p, dims);
  r = isfinite_safe(r) ? r : 0.0f;
  const float scale = (dims == 1) ? 0.2500f : ((dims == 2) ? 0.6616f :
                                               ((dims == 3) ? 0.9820f : 0.8344f));
  return scale * r;
}
CY_DEV float noise_nd(const float *p, int dims)
{
  return 0.5f * snoise_nd(p, dims) + 0.5f;
}

/* fractal_noise_1d .. 4d */
__device__ __noinline__ float fractal_noise_nd(const float *p, int dims, float octaves,
                                               float roughness)
{
  float fscale = 1.0f, amp = 1.0f, maxamp = 0.0f, sum = 0.0f;
  octaves = clampf(octaves, 0.0f, 16.0f);
  const int n = (int)octaves;
  float q[4];
  for (int i = 0; i <= n; i++) {
    for (int d = 0; d < dims; d++)
      q[d] = fscale * p[d];
    const float t = noise_nd(q, dims);
    sum += t * amp;
    maxamp += amp;
    amp *= clampf(roughness, 0.0f, 1.0f);
    fscale *= 2.0f;
  }
  const float rmd = octaves - floorf(octaves);
  if (rmd != 0.0f) {
    for (int d = 0; d < dims; d++)
      q[d] = fscale * p[d];
    const float t = noise_nd(q, dims);
    float sum2 = sum + t * amp;
    sum /= maxamp;
    sum2 /= maxamp + amp;
    return (1.0f - rmd) * sum + rmd * sum2;
  }
  return sum / maxamp;
}

/* random_float*_offset (svm_noisetex.h:28-56): component i of the offset for a seed */
CY_DEV float noise_seed_offset(float seed, int component, int dims)
{
  uint32_t key[2] = {__float_as_uint(seed), __float_as_uint((float)component)};
  return 100.0f + hash_to_unit_float(hash_uint_n(key, dims == 1 ? 1 : 2)) * 100.0f;
}

SVM_TEX_FN void svm_node_tex_noise(float *stack, uint4 node, int *offset)
{
  const int dims = (int)node.y;
  const uint32_t vector_off = node.z & 0xff, w_off = (node.z >> 8) & 0xff,
                 scale_off = (node.z >> 16) & 0xff, detail_off = (node.z >> 24) & 0xff;
  const uint32_t rough_off = node.w & 0xff, distortion_off = (node.w >> 8) & 0xff,
                 value_off = (node.w >> 16) & 0xff, color_off = (node.w >> 24) & 0xff;
  const uint4 defaults1 = __ldg(&g_scene.svm_nodes[*offset]);
  const uint4 defaults2 = __ldg(&g_scene.svm_nodes[*offset + 1]);
  *offset += 2;

  f3 vector = stack_load_float3(stack, vector_off);
  float w = stack_load_default(stack, w_off, defaults1.x);
  const float scale = stack_load_default(stack, scale_off, defaults1.y);
  const float detail = stack_load_default(stack, detail_off, defaults1.z);
  const float roughness = stack_load_default(stack, rough_off, defaults1.w);
  const float distortion = stack_load_default(stack, distortion_off, defaults2.x);
  vector *= scale;
  w *= scale;

  /* 1D noise runs on w alone, 4D appends it */
  float p[4] = {vector.x, vector.y, vector.z, w};
  if (dims == 1)
    p[0] = w;

  /* the distortion of component d is signed noise at p + offset(seed d) */
  if (distortion != 0.0f) {
    float shift[4], q[4];
    for (int d = 0; d < dims; d++) {
      for (int e = 0; e < dims; e++)
        q[e] = p[e] + noise_seed_offset((float)d, e, dims);
      shift[d] = snoise_nd(q, dims) * distortion;
    }
    for (int d = 0; d < dims; d++)
      p[d] += shift[d];
  }

  const float value = fractal_noise_nd(p, dims, detail, roughness);
  if (stack_valid(value_off))
    stack[value_off] = value;
  if (stack_valid(color_off)) {
    /* colour seeds continue after the distortion seeds: dims, dims + 1 */
    f3 color = mk3(value, 0.0f, 0.0f);
    float q[4];
    for (int k = 0; k < 2; k++) {
      const float seed = (float)(dims + k);
      for (int e = 0; e < dims; e++)
        q[e] = p[e] + noise_seed_offset(seed, e, dims);
      const float c = fractal_noise_nd(q, dims, detail, roughness);
      if (k == 0)
        color.y = c;
      else
        color.z = c;
    }
    stack_store_float3(stack, color_off, color);
  }
}

/* ------------------------------------------------ checker, gradient, wave */

SVM_TEX_FN void svm_node_tex_checker(float *stack, uint4 node)
{
  const uint32_t co_off = node.y & 0xff, color1_off = (node.y >> 8) & 0xff,
                 color2_off = (node.y >> 16) & 0xff, scale_off = (node.y >> 24) & 0xff;
  const uint32_t color_off = node.z & 0xff, fac_off = (node.z >> 8) & 0xff;
  const f3 color1 = stack_load_float3(stack, color1_off);
  const f3 color2 = stack_load_float3(stack, color2_off);
  f3 p = stack_load_float3(stack, co_off) * stack_load_default(stack, scale_off, node.w);
  /* nudge off the unit lattice */
  p.x = (p.x + 0.000001f) * 0.999999f;
  p.y = (p.y + 0.000001f) * 0.999999f;
  p.z = (p.z + 0.000001f) * 0.999999f;
  const int xi = abs((int)floorf(p.x)), yi = abs((int)floorf(p.y)), zi = abs((int)floorf(p.z));
  const bool on = ((xi % 2 == yi % 2) == (zi % 2));
  if (stack_valid(color_off))
    stack_store_float3(stack, color_off, on ? color1 : color2);
  if (stack_valid(fac_off))
    stack[fac_off] = on ? 1.0f : 0.0f;
}

SVM_TEX_FN void svm_node_tex_gradient(float *stack, uint4 node)
{
  const uint32_t type = node.y & 0xff, co_off = (node.y >> 8) & 0xff,
                 fac_off = (node.y >> 16) & 0xff, color_off = (node.y >> 24) & 0xff;
  const f3 p = stack_load_float3(stack, co_off);
  float f = 0.0f;
  switch (type) {
    case CY_NODE_BLEND_LINEAR:
      f = p.x;
      break;
    case CY_NODE_BLEND_QUADRATIC: {
      const float r = fmaxf(p.x, 0.0f);
      f = r * r;
      break;
    }
    case CY_NODE_BLEND_EASING: {
      const float r = fminf(fmaxf(p.x, 0.0f), 1.0f);
      const float t = r * r;
      f = 3.0f * t - 2.0f * t * r;
      break;
    }
    case CY_NODE_BLEND_DIAGONAL:
      f = (p.x + p.y) * 0.5f;
      break;
    case CY_NODE_BLEND_RADIAL:
      f = atan2f(p.y, p.x) / CY_M_2PI_F + 0.5f;
      break;
    case CY_NODE_BLEND_QUADRATIC_SPHERE:
    case CY_NODE_BLEND_SPHERICAL: {
      /* the bias keeps a unit-length p at exactly zero */
      const float r = fmaxf(0.999999f - sqrtf(p.x * p.x + p.y * p.y + p.z * p.z), 0.0f);
      f = (type == CY_NODE_BLEND_QUADRATIC_SPHERE) ? r * r : r;
      break;
    }
    default:
      break;
  }
  f = saturate(f);
  if (stack_valid(fac_off))
    stack[fac_off] = f;
  if (stack_valid(color_off))
    stack_store_float3(stack, color_off, mk3(f, f, f));
}

SVM_TEX_FN void svm_node_tex_wave(float *stack, uint4 node, int *offset)
{
  const uint4 node2 = __ldg(&g_scene.svm_nodes[*offset]);
  const uint4 node3 = __ldg(&g_scene.svm_nodes[*offset + 1]);
  *offset += 2;
  const uint32_t type = node.y & 0xff, bands_dir = (node.y >> 8) & 0xff,
                 rings_dir = (node.y >> 16) & 0xff, profile = (node.y >> 24) & 0xff;
  const uint32_t co_off = node.z & 0xff, scale_off = (node.z >> 8) & 0xff,
                 distortion_off = (node.z >> 16) & 0xff;
  const uint32_t detail_off = node.w & 0xff, dscale_off = (node.w >> 8) & 0xff,
                 drough_off = (node.w >> 16) & 0xff, phase_off = (node.w >> 24) & 0xff;
  const uint32_t color_off = node2.x & 0xff, fac_off = (node2.x >> 8) & 0xff;

  const float scale = stack_load_default(stack, scale_off, node2.y);
  const float distortion = stack_load_default(stack, distortion_off, node2.z);
  const float detail = stack_load_default(stack, detail_off, node2.w);
  const float dscale = stack_load_default(stack, dscale_off, node3.x);
  const float droughness = stack_load_default(stack, drough_off, node3.y);
  const float phase = stack_load_default(stack, phase_off, node3.z);

  f3 p = stack_load_float3(stack, co_off) * scale;
  p = mk3((p.x + 0.000001f) * 0.999999f, (p.y + 0.000001f) * 0.999999f,
          (p.z + 0.000001f) * 0.999999f);

  float n;
  if (type == CY_NODE_WAVE_BANDS) {
    if (bands_dir == CY_NODE_WAVE_BANDS_DIRECTION_X)
      n = p.x * 20.0f;
    else if (bands_dir == CY_NODE_WAVE_BANDS_DIRECTION_Y)
      n = p.y * 20.0f;
    else if (bands_dir == CY_NODE_WAVE_BANDS_DIRECTION_Z)
      n = p.z * 20.0f;
    else
      n = (p.x + p.y + p.z) * 10.0f;
  }
  else {
    f3 rp = p;
    if (rings_dir == CY_NODE_WAVE_RINGS_DIRECTION_X)
      rp = rp * mk3(0.0f, 1.0f, 1.0f);
    else if (rings_dir == CY_NODE_WAVE_RINGS_DIRECTION_Y)
      rp = rp * mk3(1.0f, 0.0f, 1.0f);
    else if (rings_dir == CY_NODE_WAVE_RINGS_DIRECTION_Z)
      rp = rp * mk3(1.0f, 1.0f, 0.0f);
    n = len(rp) * 20.0f;
  }
  n += phase;
  if (distortion != 0.0f) {
    const float q[3] = {p.x * dscale, p.y * dscale, p.z * dscale};
    n += distortion * (fractal_noise_nd(q, 3, detail, droughness) * 2.0f - 1.0f);
  }
  float f;
  if (profile == CY_NODE_WAVE_PROFILE_SIN) {
    f = 0.5f + 0.5f * sinf(n - CY_M_PI_2_F);
  }
  else if (profile == CY_NODE_WAVE_PROFILE_SAW) {
    n /= CY_M_2PI_F;
    f = n - floorf(n);
  }
  else {
    n /= CY_M_2PI_F;
    f = fabsf(n - floorf(n + 0.5f)) * 2.0f;
  }
  if (stack_valid(fac_off))
    stack[fac_off] = f;
  if (stack_valid(color_off))
    stack_store_float3(stack, color_off, mk3(f, f, f));
}

/* ----------------------------------------------------------- magic, brick */

SVM_TEX_FN void svm_node_tex_magic(float *stack, uint4 node, int *offset)
{
  const int depth = (int)(node.y & 0xff);
  const uint32_t color_off = (node.y >> 8) & 0xff, fac_off = (node.y >> 16) & 0xff;
  const uint32_t co_off = node.z & 0xff, scale_off = (node.z >> 8) & 0xff,
                 distortion_off = (node.z >> 16) & 0xff;
  const uint4 node2 = __ldg(&g_scene.svm_nodes[*offset]);
  (*offset)++;
  const f3 p = stack_load_float3(stack, co_off) * stack_load_default(stack, scale_off, node2.x);
  float dist = stack_load_default(stack, distortion_off, node2.y);

  float x = sinf((p.x + p.y + p.z) * 5.0f);
  float y = cosf((-p.x + p.y - p.z) * 5.0f);
  float z = -cosf((-p.x - p.y + p.z) * 5.0f);
  /* each further depth level rewrites one channel from the three (svm_magic.h:27-83) */
  if (depth > 0) {
    x *= dist;
    y *= dist;
    z *= dist;
    y = -cosf(x - y + z) * dist;
  }
  if (depth > 1)
    x = cosf(x - y - z) * dist;
  if (depth > 2)
    z = sinf(-x - y - z) * dist;
  if (depth > 3)
    x = -cosf(-x + y - z) * dist;
  if (depth > 4)
    y = -sinf(-x + y + z) * dist;
  if (depth > 5)
    y = -cosf(-x + y + z) * dist;
  if (depth > 6)
    x = cosf(x + y + z) * dist;
  if (depth > 7)
    z = sinf(x + y - z) * dist;
  if (depth > 8)
    x = -cosf(-x - y + z) * dist;
  if (depth > 9)
    y = -sinf(x - y + z) * dist;
  if (dist != 0.0f) {
    dist *= 2.0f;
    x /= dist;
    y /= dist;
    z /= dist;
  }
  const f3 color = mk3(0.5f - x, 0.5f - y, 0.5f - z);
  if (stack_valid(fac_off))
    stack[fac_off] = average(color);
  if (stack_valid(color_off))
    stack_store_float3(stack, color_off, color);
}

CY_DEV float brick_noise(uint32_t n)
{
  n = (n + 1013) & 0x7fffffff;
  n = (n >> 13) ^ n;
  const uint32_t nn = (n * (n * n * 60493 + 19990303) + 1376312589) & 0x7fffffff;
  return 0.5f * ((float)nn / 1073741824.0f);
}

SVM_TEX_FN void svm_node_tex_brick(float *stack, uint4 node, int *offset)
{
  const uint4 node2 = __ldg(&g_scene.svm_nodes[*offset]);
  const uint4 node3 = __ldg(&g_scene.svm_nodes[*offset + 1]);
  const uint4 node4 = __ldg(&g_scene.svm_nodes[*offset + 2]);
  *offset += 3;
  const uint32_t co_off = node.y & 0xff, color1_off = (node.y >> 8) & 0xff,
                 color2_off = (node.y >> 16) & 0xff, mortar_off = (node.y >> 24) & 0xff;
  const uint32_t scale_off = node.z & 0xff, mortar_size_off = (node.z >> 8) & 0xff,
                 bias_off = (node.z >> 16) & 0xff, width_off = (node.z >> 24) & 0xff;
  const uint32_t height_off = node.w & 0xff, color_off = (node.w >> 8) & 0xff,
                 fac_off = (node.w >> 16) & 0xff, smooth_off = (node.w >> 24) & 0xff;
  const int offset_frequency = (int)(node2.x & 0xff), squash_frequency = (int)((node2.x >> 8) & 0xff);

  f3 color1 = stack_load_float3(stack, color1_off);
  const f3 color2 = stack_load_float3(stack, color2_off);
  const f3 mortar_color = stack_load_float3(stack, mortar_off);
  const float scale = stack_load_default(stack, scale_off, node2.y);
  const float mortar_size = stack_load_default(stack, mortar_size_off, node2.z);
  const float mortar_smooth = stack_load_default(stack, smooth_off, node4.x);
  const float bias = stack_load_default(stack, bias_off, node2.w);
  float brick_width = stack_load_default(stack, width_off, node3.x);
  const float row_height = stack_load_default(stack, height_off, node3.y);
  const float offset_amount = __uint_as_float(node3.z);
  const float squash_amount = __uint_as_float(node3.w);
  const f3 p = stack_load_float3(stack, co_off) * scale;

  const int rownum = (int)floorf(p.y / row_height);
  float shift = 0.0f;
  if (offset_frequency && squash_frequency) {
    brick_width *= (rownum % squash_frequency) ? 1.0f : squash_amount;
    shift = (rownum % offset_frequency) ? 0.0f : (brick_width * offset_amount);
  }
  const int bricknum = (int)floorf((p.x + shift) / brick_width);
  const float x = (p.x + shift) - brick_width * bricknum;
  const float y = p.y - row_height * rownum;
  const float tint = saturate(brick_noise(((uint32_t)rownum << 16) + ((uint32_t)bricknum & 0xFFFF)) +
                              bias);
  float min_dist = fminf(fminf(x, y), fminf(brick_width - x, row_height - y));
  float f;
  if (min_dist >= mortar_size) {
    f = 0.0f;
  }
  else if (mortar_smooth == 0.0f) {
    f = 1.0f;
  }
  else {
    min_dist = 1.0f - min_dist / mortar_size;
    if (min_dist < mortar_smooth) {
      const float s = min_dist / mortar_smooth, ss = s * s;
      f = 3.0f * ss - 2.0f * ss * s;
    }
    else {
      f = 1.0f;
    }
  }
  if (f != 1.0f)
    color1 = (1.0f - tint) * color1 + tint * color2;
  if (stack_valid(color_off))
    stack_store_float3(stack, color_off, color1 * (1.0f - f) + mortar_color * f);
  if (stack_valid(fac_off))
    stack[fac_off] = f;
}

/* --------------------------------- object info, camera, vector transform */

CY_DEV float ko_float(int object, int off)
{
  return __ldg((const float *)(g_scene.objects + (size_t)object * SIZEOF_KERNEL_OBJECT + off));
}

/* svm_geometry.h:102-139 */
SVM_TEX_FN void svm_node_object_info(const ShaderDataG &sd, float *stack, uint4 node)
{
  float data = 0.0f;
  switch (node.y) {
    case CY_NODE_INFO_OB_LOCATION: {
      f3 loc = zero3();
      if (sd.object != -1) {
        const tfm34 t = object_tfm(sd.object);
        loc = mk3(t.x.w, t.y.w, t.z.w);
      }
      stack_store_float3(stack, node.z, loc);
      return;
    }
    case CY_NODE_INFO_OB_COLOR:
      stack_store_float3(stack, node.z,
                         (sd.object == -1) ? zero3() :
                                             mk3(ko_float(sd.object, KO_COLOR),
                                                 ko_float(sd.object, KO_COLOR + 4),
                                                 ko_float(sd.object, KO_COLOR + 8)));
      return;
    case CY_NODE_INFO_OB_INDEX:
      data = (sd.object == -1) ? 0.0f : ko_float(sd.object, KO_PASS_ID);
      break;
    case CY_NODE_INFO_MAT_INDEX:
      data = (float)__ldg((const int *)(g_scene.shaders +
                                        (size_t)(sd.shader & CY_SHADER_MASK) * SIZEOF_KERNEL_SHADER +
                                        KS_PASS_ID));
      break;
    case CY_NODE_INFO_OB_RANDOM:
      if (sd.lamp != -1)
        data = __ldg((const float *)(g_scene.lights + (size_t)sd.lamp * SIZEOF_KERNEL_LIGHT +
                                     KL_RANDOM));
      else
        data = (sd.object == -1) ? 0.0f : ko_float(sd.object, KO_RANDOM_NUMBER);
      break;
    default:
      break;
  }
  stack[node.z] = data;
}

/* svm_camera.h:19-43 */
SVM_TEX_FN void svm_node_camera(const ShaderDataG &sd, float *stack, uint4 node)
{
  const f3 vector = transform_point(kd_transform(KD_CAM_WORLDTOCAMERA), sd.P);
  if (stack_valid(node.y))
    stack_store_float3(stack, node.y, normalize(vector));
  if (stack_valid(node.z))
    stack[node.z] = vector.z;
  if (stack_valid(node.w))
    stack[node.w] = len(vector);
}

/* svm_vector_transform.h:21-105 */
SVM_TEX_FN void svm_node_vector_transform(const ShaderDataG &sd, float *stack, uint4 node)
{
  const uint32_t type = node.y & 0xff, from = (node.y >> 8) & 0xff, to = (node.y >> 16) & 0xff;
  const uint32_t vector_in = node.z & 0xff, vector_out = (node.z >> 8) & 0xff;
  f3 in = stack_load_float3(stack, vector_in);
  const bool is_object = (sd.object != -1);
  const bool is_direction = (type == CY_NODE_VECTOR_TRANSFORM_TYPE_VECTOR ||
                             type == CY_NODE_VECTOR_TRANSFORM_TYPE_NORMAL);
  const uint32_t WORLD = CY_NODE_VECTOR_TRANSFORM_CONVERT_SPACE_WORLD,
                 OBJECT = CY_NODE_VECTOR_TRANSFORM_CONVERT_SPACE_OBJECT,
                 CAMERA = CY_NODE_VECTOR_TRANSFORM_CONVERT_SPACE_CAMERA;
  /* world <-> camera through the camera matrices, world <-> object through the object's */
  const bool to_world_first = (from == CAMERA && (to == WORLD || to == OBJECT));
  const bool from_object_first = (from == OBJECT && (to == WORLD || to == CAMERA) && is_object);
  if (to_world_first) {
    const tfm34 t = kd_transform(KD_CAM_CAMERATOWORLD);
    in = is_direction ? transform_direction(t, in) : transform_point(t, in);
  }
  if (from_object_first) {
    const tfm34 t = object_tfm(sd.object);
    in = is_direction ? transform_direction(t, in) : transform_point(t, in);
  }
  if ((from == WORLD || from == CAMERA) && to == OBJECT && is_object) {
    const tfm34 t = object_itfm(sd.object);
    in = is_direction ? transform_direction(t, in) : transform_point(t, in);
  }
  if ((from == WORLD || from == OBJECT) && to == CAMERA) {
    const tfm34 t = kd_transform(KD_CAM_WORLDTOCAMERA);
    in = is_direction ? transform_direction(t, in) : transform_point(t, in);
  }
  if (type == CY_NODE_VECTOR_TRANSFORM_TYPE_NORMAL)
    in = normalize(in);
  if (stack_valid(vector_out))
    stack_store_float3(stack, vector_out, in);
}

/* svm_white_noise.h:19-79; hash_float*_to_float3 of util_hash.h:180-216 */
SVM_TEX_FN void svm_node_tex_white_noise(float *stack, uint4 node)
{
  const uint32_t dims = node.y;
  const uint32_t vector_offset = node.z & 0xff, w_offset = (node.z >> 8) & 0xff;
  const uint32_t value_offset = node.w & 0xff, color_offset = (node.w >> 8) & 0xff;
  const f3 v = stack_load_float3(stack, vector_offset);
  const float w = stack[w_offset];
  const uint32_t x = __float_as_uint(v.x), y = __float_as_uint(v.y), z = __float_as_uint(v.z),
                 ww = __float_as_uint(w);
  const uint32_t one = __float_as_uint(1.0f), two = __float_as_uint(2.0f);
  uint32_t k0[4], k1[4], k2[4];
  int n0 = (int)dims, n12 = (int)dims + 1; /* the 2nd and 3rd channel append 1.0 / 2.0 */
  switch (dims) {
    case 1:
      k0[0] = k1[0] = k2[0] = ww;
      k1[1] = one;
      k2[1] = two;
      break;
    case 2:
      k0[0] = k1[0] = k2[0] = x;
      k0[1] = k1[1] = k2[1] = y;
      k1[2] = one;
      k2[2] = two;
      break;
    case 3:
      k0[0] = k1[0] = k2[0] = x;
      k0[1] = k1[1] = k2[1] = y;
      k0[2] = k1[2] = k2[2] = z;
      k1[3] = one;
      k2[3] = two;
      break;
    default: /* 4D: the other channels hash permutations of the key */
      k0[0] = x, k0[1] = y, k0[2] = z, k0[3] = ww;
      k1[0] = z, k1[1] = x, k1[2] = ww, k1[3] = y;
      k2[0] = ww, k2[1] = z, k2[2] = y, k2[3] = x;
      n0 = n12 = 4;
      break;
  }
  const float value = hash_to_unit_float(hash_uint_n(k0, n0));
  if (stack_valid(color_offset))
    stack_store_float3(stack, color_offset,
                       mk3(value, hash_to_unit_float(hash_uint_n(k1, n12)),
                           hash_to_unit_float(hash_uint_n(k2, n12))));
  if (stack_valid(value_offset))
    stack[value_offset] = value;
}

/* -------------------------------------------------- tangent, normal map */

/* svm_tex_coord.h:363-411 (Tangent node): the UV-map tangent attribute, or a radial
 * tangent from the generated coordinates (the position where the mesh has none) */
SVM_TEX_FN void svm_node_tangent(const ShaderDataG &sd, float *stack, uint4 node)
{
  uint32_t tangent_offset, direction_type, axis;
  unpack_uchar3(node.y, &tangent_offset, &direction_type, &axis);
  const AttrDesc desc = find_attribute(sd, node.z);
  const bool found = desc.offset != CY_ATTR_STD_NOT_FOUND;
  f3 value = zero3();
  if (found) {
    if (desc.type == CY_NODE_ATTR_FLOAT2) {
      const float2 v = attribute_float2(sd, desc);
      value = mk3(v.x, v.y, 0.0f);
    }
    else {
      value = attribute_float3(sd, desc);
    }
  }
  f3 tangent;
  if (direction_type == CY_NODE_TANGENT_UVMAP) {
    tangent = found ? value : zero3();
  }
  else {
    const f3 g = found ? value : sd.P;
    if (axis == CY_NODE_TANGENT_AXIS_X)
      tangent = mk3(0.0f, -(g.z - 0.5f), (g.y - 0.5f));
    else if (axis == CY_NODE_TANGENT_AXIS_Y)
      tangent = mk3(-(g.z - 0.5f), 0.0f, (g.x - 0.5f));
    else
      tangent = mk3(-(g.y - 0.5f), (g.x - 0.5f), 0.0f);
  }
  if (sd.object != -1) /* surfaces only: the reference reads an unset transform otherwise */
    tangent = object_normal_transform(sd.object, tangent);
  tangent = cross(sd.N, normalize(cross(tangent, sd.N)));
  stack_store_float3(stack, tangent_offset, tangent);
}

/* kernel_montecarlo.h:196-299: bend a shading normal back until the reflection of I
 * about it leaves the geometric surface */
CY_DEV f3 ensure_valid_reflection(f3 Ng, f3 I, f3 N)
{
  const f3 R = 2 * dot(N, I) * N - I;
  const float threshold = fminf(0.9f * dot(Ng, I), 0.01f);
  if (dot(Ng, R) >= threshold)
    return N;
  const float NdotNg = dot(N, Ng);
  const f3 X = normalize(N - NdotNg * Ng);
  const float Ix = dot(I, X), Iz = dot(I, Ng);
  const float Ix2 = sqr(Ix), Iz2 = sqr(Iz);
  const float a = Ix2 + Iz2;
  const float b = safe_sqrtf(Ix2 * (a - sqr(threshold)));
  const float c = Iz * threshold + a;
  const float fac = 0.5f / a;
  const float N1_z2 = fac * (b + c), N2_z2 = fac * (-b + c);
  bool valid1 = (N1_z2 > 1e-5f) && (N1_z2 <= (1.0f + 1e-5f));
  bool valid2 = (N2_z2 > 1e-5f) && (N2_z2 <= (1.0f + 1e-5f));
  float nx, ny;
  if (valid1 && valid2) {
    const float n1x = safe_sqrtf(1.0f - N1_z2), n1y = safe_sqrtf(N1_z2);
    const float n2x = safe_sqrtf(1.0f - N2_z2), n2y = safe_sqrtf(N2_z2);
    const float R1 = 2 * (n1x * Ix + n1y * Iz) * n1y - Iz;
    const float R2 = 2 * (n2x * Ix + n2y * Iz) * n2y - Iz;
    valid1 = (R1 >= 1e-5f);
    valid2 = (R2 >= 1e-5f);
    /* both valid: the shallower reflection; else the positive one */
    const bool first = (valid1 && valid2) ? (R1 < R2) : (R1 > R2);
    nx = first ? n1x : n2x;
    ny = first ? n1y : n2y;
  }
  else if (valid1 || valid2) {
    const float Nz2 = valid1 ? N1_z2 : N2_z2;
    nx = safe_sqrtf(1.0f - Nz2);
    ny = safe_sqrtf(Nz2);
  }
  else {
    return Ng;
  }
  return nx * X + ny * Ng;
}

/* svm_tex_coord.h:264-361 (Normal Map node): tangent space with the host's tangent and
 * sign attributes, object / world space and their Blender-convention variants */
SVM_TEX_FN void svm_node_normal_map(const ShaderDataG &sd, float *stack, uint4 node)
{
  const uint32_t color_offset = node.y & 0xff, strength_offset = (node.y >> 8) & 0xff,
                 normal_offset = (node.y >> 16) & 0xff, space = (node.y >> 24) & 0xff;
  f3 color = stack_load_float3(stack, color_offset);
  color = 2.0f * mk3(color.x - 0.5f, color.y - 0.5f, color.z - 0.5f);
  const bool is_backfacing = (sd.flag & CY_SD_BACKFACING) != 0;
  f3 N;
  if (space == CY_NODE_NORMAL_MAP_TANGENT) {
    if (sd.object == -1) {
      stack_store_float3(stack, normal_offset, zero3());
      return;
    }
    const AttrDesc attr = find_attribute(sd, node.z);
    const AttrDesc attr_sign = find_attribute(sd, node.w);
    const AttrDesc attr_normal = find_attribute(sd, CY_ATTR_STD_VERTEX_NORMAL);
    if (attr.offset == CY_ATTR_STD_NOT_FOUND || attr_sign.offset == CY_ATTR_STD_NOT_FOUND ||
        attr_normal.offset == CY_ATTR_STD_NOT_FOUND) {
      stack_store_float3(stack, normal_offset, zero3());
      return;
    }
    const f3 tangent = attribute_float3(sd, attr);
    const float sign = attribute_float(sd, attr_sign);
    f3 normal;
    if (sd.shader & CY_SHADER_SMOOTH_NORMAL) {
      normal = attribute_float3(sd, attr_normal);
    }
    else {
      normal = is_backfacing ? -sd.Ng : sd.Ng;
      normal = normalize(transform_direction_transposed(object_tfm(sd.object), normal));
    }
    const f3 B = sign * cross(normal, tangent);
    N = safe_normalize(color.x * tangent + color.y * B + color.z * normal);
    N = object_normal_transform(sd.object, N);
  }
  else {
    if (space == CY_NODE_NORMAL_MAP_BLENDER_OBJECT || space == CY_NODE_NORMAL_MAP_BLENDER_WORLD) {
      color.y = -color.y;
      color.z = -color.z;
    }
    N = color;
    if ((space == CY_NODE_NORMAL_MAP_OBJECT || space == CY_NODE_NORMAL_MAP_BLENDER_OBJECT) &&
        sd.object != -1)
      N = object_normal_transform(sd.object, N);
    else
      N = safe_normalize(N);
  }
  if (is_backfacing)
    N = -N;
  float strength = stack[strength_offset];
  if (strength != 1.0f) {
    strength = fmaxf(strength, 0.0f);
    N = safe_normalize(sd.N + (N - sd.N) * strength);
  }
  N = ensure_valid_reflection(sd.Ng, sd.I, N);
  if (is_zero(N))
    N = sd.N;
  stack_store_float3(stack, normal_offset, N);
}

#include "svm_tex_cells.cuh"

#endif /* B200_SVM_TEX_CUH */

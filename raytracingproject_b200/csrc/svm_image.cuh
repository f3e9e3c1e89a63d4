/* svm_image.cuh - image textures: the Image Texture node (flat, sphere, tube and box
 * projections, UDIM tiles) and the Environment Texture node (equirectangular, mirror
 * ball), sampling images that live in device memory as plain pixel arrays.
 *
 * Semantics to match (reference = blender/intern/cycles):
 *   kernel/svm/svm_image.h                     the three nodes, alpha un-association, the
 *                                              sRGB decompression of 8-bit colour images
 *   kernel/kernels/cpu/kernel_cpu_image.h      TextureInterpolator: closest / linear / cubic
 *                                              B-spline lookups with repeat / extend / clip
 *                                              extension, every pixel format
 *   util/util_math.h map_to_sphere / map_to_tube, kernel_projection.h direction_to_*
 * The CPU device's arithmetic is followed rather than a hardware texture unit's (9-bit
 * fixed-point weights, its own wrap rules): the CPU path is the oracle, and a bilinear
 * lookup is four 128-bit loads - cheap next to the shading around it.  3D (volume) images
 * are out of scope: the binder refuses them.
 *
 * Images are bound per slot through the C ABI (b200_texture_set): the kernels see an array
 * of the reference's TextureInfo records (cycles_abi.h TI_* offsets) whose `data` is the
 * device address of the pixels - what CUDADevice::tex_alloc keeps for its non-bindless
 * path (device/cuda/device_cuda_impl.cpp:1105-1304).
 * Included by svm_tex.cuh users (shade.cuh); host-compilable (tests/host_check). */
#ifndef B200_SVM_IMAGE_CUH
#define B200_SVM_IMAGE_CUH

#ifndef SVM_TEX_FN
#  define SVM_TEX_FN __device__ __noinline__
#endif

#define CY_TEX_IMAGE_MISSING make_float4(1.0f, 0.0f, 1.0f, 1.0f) /* kernel_types.h:76-79 */

struct ImageView {
  const uint8_t *data;
  uint32_t type, interpolation, extension;
  int width, height;
};

CY_DEV ImageView image_view(int id)
{
  const uint8_t *ti = g_scene.texture_info + (size_t)id * SIZEOF_TEXTURE_INFO;
  ImageView v;
  v.data = (const uint8_t *)(uintptr_t)__ldg((const unsigned long long *)(ti + TI_DATA));
  v.type = __ldg((const uint32_t *)(ti + TI_DATA_TYPE));
  v.interpolation = __ldg((const uint32_t *)(ti + TI_INTERPOLATION));
  v.extension = __ldg((const uint32_t *)(ti + TI_EXTENSION));
  v.width = (int)__ldg((const uint32_t *)(ti + TI_WIDTH));
  v.height = (int)__ldg((const uint32_t *)(ti + TI_HEIGHT));
  return v;
}

CY_DEV float4 f4_scale(float4 a, float s)
{
  return make_float4(a.x * s, a.y * s, a.z * s, a.w * s);
}
CY_DEV float4 f4_add(float4 a, float4 b)
{
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

CY_DEV float half_bits_to_float(unsigned short h)
{
  /* util_half.h:120-127 half_to_float, scalar path: sign, rebiased exponent, mantissa -
   * bit for bit, including its reading of half zero as 2^-15 */
  const uint32_t u = (uint32_t)h;
  return __uint_as_float(((u & 0x8000u) << 16) | (((u & 0x7c00u) + 0x1C000u) << 13) |
                         ((u & 0x03FFu) << 13));
}

/* one texel as float4; single-channel formats replicate into rgb with alpha 1 */
CY_DEV float4 image_texel(const ImageView &im, int x, int y)
{
  const size_t i = (size_t)y * (size_t)im.width + (size_t)x;
  switch (im.type) {
    case CY_IMAGE_DATA_TYPE_FLOAT4:
      return __ldg((const float4 *)im.data + i);
    case CY_IMAGE_DATA_TYPE_BYTE4: {
      const uchar4 c = __ldg((const uchar4 *)im.data + i);
      const float f = 1.0f / 255.0f;
      return make_float4(c.x * f, c.y * f, c.z * f, c.w * f);
    }
    case CY_IMAGE_DATA_TYPE_FLOAT: {
      const float f = __ldg((const float *)im.data + i);
      return make_float4(f, f, f, 1.0f);
    }
    case CY_IMAGE_DATA_TYPE_BYTE: {
      const float f = __ldg(im.data + i) * (1.0f / 255.0f);
      return make_float4(f, f, f, 1.0f);
    }
    case CY_IMAGE_DATA_TYPE_USHORT4: {
      const ushort4 c = __ldg((const ushort4 *)im.data + i);
      const float f = 1.0f / 65535.0f;
      return make_float4(c.x * f, c.y * f, c.z * f, c.w * f);
    }
    case CY_IMAGE_DATA_TYPE_USHORT: {
      const float f = __ldg((const unsigned short *)im.data + i) * (1.0f / 65535.0f);
      return make_float4(f, f, f, 1.0f);
    }
    case CY_IMAGE_DATA_TYPE_HALF4: {
      const ushort4 c = __ldg((const ushort4 *)im.data + i);
      return make_float4(half_bits_to_float(c.x), half_bits_to_float(c.y),
                         half_bits_to_float(c.z), half_bits_to_float(c.w));
    }
    case CY_IMAGE_DATA_TYPE_HALF: {
      const float f = half_bits_to_float(__ldg((const unsigned short *)im.data + i));
      return make_float4(f, f, f, 1.0f);
    }
    default:
      return CY_TEX_IMAGE_MISSING;
  }
}

/* texel, or transparent black outside the image (what the clip extension reads) */
CY_DEV float4 image_texel_or_zero(const ImageView &im, int x, int y)
{
  if (x < 0 || y < 0 || x >= im.width || y >= im.height)
    return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  return image_texel(im, x, y);
}

CY_DEV int image_wrap_periodic(int x, int n)
{
  x %= n;
  return (x < 0) ? x + n : x;
}
CY_DEV int image_wrap_clamp(int x, int n)
{
  return min(max(x, 0), n - 1);
}
/* floor and fraction the way the CPU device takes them (truncate, step down for negatives) */
CY_DEV float image_frac(float x, int *ix)
{
  const int i = (int)x - ((x < 0.0f) ? 1 : 0);
  *ix = i;
  return x - (float)i;
}

/* the coordinate of tap `k` (0 = the texel at or below the sample point) under the
 * image's extension; the clip extension keeps out-of-range coordinates, which then read
 * as transparent black */
CY_DEV int image_tap(const ImageView &im, int base, int k, int n)
{
  switch (im.extension) {
    case CY_EXTENSION_REPEAT:
      return image_wrap_periodic(image_wrap_periodic(base, n) + k, n);
    case CY_EXTENSION_EXTEND:
      return image_wrap_clamp(base + k, n);
    default:
      return base + k;
  }
}

CY_DEV void cubic_bspline_weights(float t, float w[4])
{
  w[0] = (((-1.0f / 6.0f) * t + 0.5f) * t - 0.5f) * t + (1.0f / 6.0f);
  w[1] = ((0.5f * t - 1.0f) * t) * t + (2.0f / 3.0f);
  w[2] = ((-0.5f * t + 0.5f) * t + 0.5f) * t + (1.0f / 6.0f);
  w[3] = (1.0f / 6.0f) * t * t * t;
}

/* kernel_tex_image_interp: the image of slot `id` at (x, y) in [0, 1]^2 */
CY_DEV float4 image_sample(int id, float x, float y)
{
  const ImageView im = image_view(id);
  if (!im.data)
    return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  if (im.interpolation == CY_INTERPOLATION_CLOSEST) {
    int ix, iy;
    image_frac(x * (float)im.width, &ix);
    image_frac(y * (float)im.height, &iy);
    if (im.extension == CY_EXTENSION_REPEAT) {
      ix = image_wrap_periodic(ix, im.width);
      iy = image_wrap_periodic(iy, im.height);
    }
    else {
      if (im.extension == CY_EXTENSION_CLIP && (x < 0.0f || y < 0.0f || x > 1.0f || y > 1.0f))
        return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      ix = image_wrap_clamp(ix, im.width);
      iy = image_wrap_clamp(iy, im.height);
    }
    return image_texel(im, ix, iy);
  }
  int ix, iy;
  const float tx = image_frac(x * (float)im.width - 0.5f, &ix);
  const float ty = image_frac(y * (float)im.height - 0.5f, &iy);
  if (im.interpolation == CY_INTERPOLATION_LINEAR) {
    const int x0 = image_tap(im, ix, 0, im.width), x1 = image_tap(im, ix, 1, im.width);
    const int y0 = image_tap(im, iy, 0, im.height), y1 = image_tap(im, iy, 1, im.height);
    float4 r = f4_scale(image_texel_or_zero(im, x0, y0), (1.0f - ty) * (1.0f - tx));
    r = f4_add(r, f4_scale(image_texel_or_zero(im, x1, y0), (1.0f - ty) * tx));
    r = f4_add(r, f4_scale(image_texel_or_zero(im, x0, y1), ty * (1.0f - tx)));
    r = f4_add(r, f4_scale(image_texel_or_zero(im, x1, y1), ty * tx));
    return r;
  }
  /* cubic B-spline over the 4 x 4 neighbourhood, rows summed in the reference's order */
  float u[4], v[4];
  cubic_bspline_weights(tx, u);
  cubic_bspline_weights(ty, v);
  int xc[4], yc[4];
  for (int k = 0; k < 4; k++) {
    xc[k] = image_tap(im, ix, k - 1, im.width);
    yc[k] = image_tap(im, iy, k - 1, im.height);
  }
  float4 r = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  for (int row = 0; row < 4; row++) {
    float4 s = f4_scale(image_texel_or_zero(im, xc[0], yc[row]), u[0]);
    s = f4_add(s, f4_scale(image_texel_or_zero(im, xc[1], yc[row]), u[1]));
    s = f4_add(s, f4_scale(image_texel_or_zero(im, xc[2], yc[row]), u[2]));
    s = f4_add(s, f4_scale(image_texel_or_zero(im, xc[3], yc[row]), u[3]));
    r = (row == 0) ? f4_scale(s, v[row]) : f4_add(r, f4_scale(s, v[row]));
  }
  return r;
}

CY_DEV float srgb_to_linear(float c)
{
  if (c < 0.04045f)
    return (c < 0.0f) ? 0.0f : c * (1.0f / 12.92f);
  return powf((c + 0.055f) * (1.0f / 1.055f), 2.4f);
}

/* svm_image_texture: lookup + the node's colour handling */
CY_DEV float4 image_texture_lookup(int id, float x, float y, uint32_t flags)
{
  if (id < 0 || (uint32_t)id >= g_scene.num_textures)
    return CY_TEX_IMAGE_MISSING; /* -1 = the image failed to load; beyond the bound slots */
  float4 r = image_sample(id, x, y);
  const float alpha = r.w;
  if ((flags & CY_NODE_IMAGE_ALPHA_UNASSOCIATE) && alpha != 1.0f && alpha != 0.0f) {
    const float inv = 1.0f / alpha; /* float4 / scalar: one reciprocal, four products */
    r = make_float4(r.x * inv, r.y * inv, r.z * inv, alpha);
  }
  if (flags & CY_NODE_IMAGE_COMPRESS_AS_SRGB)
    r = make_float4(srgb_to_linear(r.x), srgb_to_linear(r.y), srgb_to_linear(r.z), r.w);
  return r;
}

CY_DEV void image_store_result(float *stack, uint32_t out_offset, uint32_t alpha_offset, float4 f)
{
  if (stack_valid(out_offset))
    stack_store_float3(stack, out_offset, mk3(f.x, f.y, f.z));
  if (stack_valid(alpha_offset))
    stack[alpha_offset] = f.w;
}

/* unit-cube coordinates (0..1) around the centre, then onto a sphere / a tube */
CY_DEV float2 image_map_sphere(f3 co)
{
  co = (co - mk3(0.5f, 0.5f, 0.5f)) * 2.0f;
  const float l = len(co);
  if (!(l > 0.0f))
    return make_float2(0.0f, 0.0f);
  const float u = (co.x == 0.0f && co.y == 0.0f) ? 0.0f :
                                                   (1.0f - atan2f(co.x, co.y) / CY_M_PI_F) / 2.0f;
  return make_float2(u, 1.0f - safe_acosf(co.z / l) / CY_M_PI_F);
}
CY_DEV float2 image_map_tube(f3 co)
{
  co = (co - mk3(0.5f, 0.5f, 0.5f)) * 2.0f;
  const float l = sqrtf(co.x * co.x + co.y * co.y);
  if (!(l > 0.0f))
    return make_float2(0.0f, 0.0f);
  return make_float2((1.0f - (atan2f(co.x / l, co.y / l) / CY_M_PI_F)) * 0.5f,
                     (co.z + 1.0f) * 0.5f);
}

SVM_TEX_FN int svm_node_tex_image(float *stack, uint4 node, int offset)
{
  const uint32_t co_offset = node.z & 0xff, out_offset = (node.z >> 8) & 0xff;
  const uint32_t alpha_offset = (node.z >> 16) & 0xff, flags = (node.z >> 24) & 0xff;
  const f3 co = stack_load_float3(stack, co_offset);
  float2 uv;
  if (node.w == CY_NODE_IMAGE_PROJ_SPHERE)
    uv = image_map_sphere(co);
  else if (node.w == CY_NODE_IMAGE_PROJ_TUBE)
    uv = image_map_tube(co);
  else
    uv = make_float2(co.x, co.y);

  int id = -1;
  const int num_tile_nodes = (int)node.y;
  if (num_tile_nodes > 0) {
    /* UDIM: tile 1001 + 10 v + u holds the unit square at (u, v); two tiles per node */
    const int tx = (int)uv.x, ty = (int)uv.y;
    if (tx >= 0 && ty >= 0 && tx < 10) {
      const uint32_t tile = (uint32_t)(1001 + 10 * ty + tx);
      for (int i = 0; i < num_tile_nodes && id == -1; i++) {
        const uint4 t = __ldg(&g_scene.svm_nodes[offset + i]);
        if (t.x == tile)
          id = (int)t.y;
        else if (t.z == tile)
          id = (int)t.w;
      }
      if (id != -1) {
        uv.x -= (float)tx;
        uv.y -= (float)ty;
      }
    }
    offset += num_tile_nodes;
  }
  else {
    id = -num_tile_nodes;
  }
  image_store_result(stack, out_offset, alpha_offset,
                     image_texture_lookup(id, uv.x, uv.y, flags));
  return offset;
}

/* Box projection: the three axis-aligned projections blended by the object-space normal
 * (seven zones of the barycentric triangle of |N|, svm_image.h:117-215) */
SVM_TEX_FN void svm_node_tex_image_box(const ShaderDataG &sd, float *stack, uint4 node)
{
  f3 N = sd.N;
  if (sd.object != -1) {
    /* object_inverse_normal_transform: back to object space through the forward matrix */
    N = normalize(transform_direction_transposed(object_tfm(sd.object), N));
  }
  const f3 signed_N = N;
  N = fabs3(N);
  N /= (N.x + N.y + N.z);

  f3 weight = zero3();
  const float blend = __uint_as_float(node.w);
  const float limit = 0.5f * (1.0f + blend);
  if (N.x > limit * (N.x + N.y) && N.x > limit * (N.x + N.z)) {
    weight.x = 1.0f;
  }
  else if (N.y > limit * (N.x + N.y) && N.y > limit * (N.y + N.z)) {
    weight.y = 1.0f;
  }
  else if (N.z > limit * (N.x + N.z) && N.z > limit * (N.y + N.z)) {
    weight.z = 1.0f;
  }
  else if (blend > 0.0f) {
    if (N.z < (1.0f - limit) * (N.y + N.x)) {
      weight.x = N.x / (N.x + N.y);
      weight.x = saturate((weight.x - 0.5f * (1.0f - blend)) / blend);
      weight.y = 1.0f - weight.x;
    }
    else if (N.x < (1.0f - limit) * (N.y + N.z)) {
      weight.y = N.y / (N.y + N.z);
      weight.y = saturate((weight.y - 0.5f * (1.0f - blend)) / blend);
      weight.z = 1.0f - weight.y;
    }
    else if (N.y < (1.0f - limit) * (N.x + N.z)) {
      weight.x = N.x / (N.x + N.z);
      weight.x = saturate((weight.x - 0.5f * (1.0f - blend)) / blend);
      weight.z = 1.0f - weight.x;
    }
    else {
      const float d = 2.0f * limit - 1.0f;
      weight.x = ((2.0f - limit) * N.x + (limit - 1.0f)) / d;
      weight.y = ((2.0f - limit) * N.y + (limit - 1.0f)) / d;
      weight.z = ((2.0f - limit) * N.z + (limit - 1.0f)) / d;
    }
  }
  else {
    weight.x = 1.0f;
  }

  const uint32_t co_offset = node.z & 0xff, out_offset = (node.z >> 8) & 0xff;
  const uint32_t alpha_offset = (node.z >> 16) & 0xff, flags = (node.z >> 24) & 0xff;
  const f3 co = stack_load_float3(stack, co_offset);
  const int id = (int)node.y;
  float4 f = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  if (weight.x > 0.0f)
    f = f4_add(f, f4_scale(image_texture_lookup(id, (signed_N.x < 0.0f) ? 1.0f - co.y : co.y,
                                                co.z, flags),
                           weight.x));
  if (weight.y > 0.0f)
    f = f4_add(f, f4_scale(image_texture_lookup(id, (signed_N.y > 0.0f) ? 1.0f - co.x : co.x,
                                                co.z, flags),
                           weight.y));
  if (weight.z > 0.0f)
    f = f4_add(f, f4_scale(image_texture_lookup(id, (signed_N.z > 0.0f) ? 1.0f - co.y : co.y,
                                                co.x, flags),
                           weight.z));
  image_store_result(stack, out_offset, alpha_offset, f);
}

/* Environment texture: a direction looked up in an equirectangular or a mirror-ball image */
SVM_TEX_FN void svm_node_tex_environment(float *stack, uint4 node)
{
  const uint32_t co_offset = node.z & 0xff, out_offset = (node.z >> 8) & 0xff;
  const uint32_t alpha_offset = (node.z >> 16) & 0xff, flags = (node.z >> 24) & 0xff;
  f3 dir = safe_normalize(stack_load_float3(stack, co_offset));
  float2 uv;
  if (node.w == CY_NODE_ENVIRONMENT_EQUIRECTANGULAR) {
    if (is_zero(dir)) {
      uv = make_float2(0.0f, 0.0f);
    }
    else {
      /* direction_to_equirectangular with this fork's range (-2 pi, pi, -pi, pi):
       * u runs against the longitude, v = 1 at the +Z pole */
      uv.x = (atan2f(dir.y, dir.x) - CY_M_PI_F) / -CY_M_2PI_F;
      uv.y = (acosf(dir.z / len(dir)) - CY_M_PI_F) / -CY_M_PI_F;
    }
  }
  else {
    dir.y -= 1.0f;
    const float div = 2.0f * sqrtf(fmaxf(-0.5f * dir.y, 0.0f));
    if (div > 0.0f)
      dir /= div;
    uv = make_float2(0.5f * (dir.x + 1.0f), 0.5f * (dir.z + 1.0f));
  }
  image_store_result(stack, out_offset, alpha_offset,
                     image_texture_lookup((int)node.y, uv.x, uv.y, flags));
}

#endif /* B200_SVM_IMAGE_CUH */

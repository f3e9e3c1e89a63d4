/* svm_nodes.cuh - value nodes of the SVM interpreter (the nodes that compute numbers
 * on the stack, as opposed to the closure nodes of svm_closure.cuh).
 *
 * Restated from blender/intern/cycles/kernel/svm (same stack encodings, same float
 * operation order - the file is compiled without FMA contraction):
 *   convert                         svm_convert.h:22-74
 *   fresnel, layer weight           svm_fresnel.h:21-74
 *   math, vector math               svm_math.h:19-77, svm_math_util.h:19-196
 *   mix (colour blend modes)        svm_mix.h:21-38, svm_color_util.h:19-300
 *   invert, gamma, bright/contrast  svm_invert.h, svm_gamma.h, svm_brightness.h
 *   separate / combine vector       svm_sepcomb_vector.h
 *   clamp                           svm_clamp.h
 * Helper semantics (safe_divide, wrapf, pingpongf, smoothminf ...) follow
 * util/util_math.h:345-660.
 *
 * Blend modes that need RGB<->HSV (hue,
 * saturation, value, colour) and dodge / burn are refused by svm_validate.
 */
#ifndef B200_SVM_NODES_CUH
#define B200_SVM_NODES_CUH

/* measured on the shading-bound Cornell workload: inlined -0.4 %, out of line -1.3 %
 * against a build without these nodes */
#ifndef SVM_NODES_INLINE
#  define SVM_NODES_INLINE 1
#endif
#if SVM_NODES_INLINE
#  define SVM_NODES_FN __device__ __forceinline__
#else
#  define SVM_NODES_FN __device__ __noinline__
#endif

CY_DEV void unpack_uchar2(uint32_t i, uint32_t *x, uint32_t *y)
{
  *x = i & 0xffu;
  *y = (i >> 8) & 0xffu;
}
CY_DEV void unpack_uchar3(uint32_t i, uint32_t *x, uint32_t *y, uint32_t *z)
{
  *x = i & 0xffu;
  *y = (i >> 8) & 0xffu;
  *z = (i >> 16) & 0xffu;
}
CY_DEV float stack_load_float_default(const float *stack, uint32_t a, uint32_t value)
{
  return (a == (uint32_t)CY_SVM_STACK_INVALID) ? __uint_as_float(value) : stack[a];
}

/* ---- scalar helpers ---- */

CY_DEV float nodes_safe_divide(float a, float b)
{
  return (b != 0.0f) ? a / b : 0.0f;
}
CY_DEV float nodes_safe_modulo(float a, float b)
{
  return (b != 0.0f) ? fmodf(a, b) : 0.0f;
}
CY_DEV float nodes_wrapf(float value, float max, float min)
{
  const float range = max - min;
  return (range != 0.0f) ? value - (range * floorf((value - min) / range)) : min;
}
CY_DEV float nodes_fractf(float x)
{
  return x - floorf(x);
}
CY_DEV float nodes_pingpongf(float a, float b)
{
  return (b != 0.0f) ? fabsf(nodes_fractf((a - b) / (b * 2.0f)) * b * 2.0f - b) : 0.0f;
}
CY_DEV float nodes_smoothminf(float a, float b, float k)
{
  if (k != 0.0f) {
    const float h = fmaxf(k - fabsf(a - b), 0.0f) / k;
    return fminf(a, b) - h * h * h * k * (1.0f / 6.0f);
  }
  return fminf(a, b);
}
/* the CPU flavour of compatible_powf is plain powf */
CY_DEV float nodes_safe_powf(float a, float b)
{
  if (a < 0.0f && b != (float)(int)b)
    return 0.0f;
  return powf(a, b);
}
CY_DEV float nodes_safe_logf(float a, float b)
{
  if (a <= 0.0f || b <= 0.0f)
    return 0.0f;
  return nodes_safe_divide(logf(a), logf(b));
}
CY_DEV float nodes_clamp(float v, float lo, float hi)
{
  return fminf(fmaxf(v, lo), hi);
}

SVM_NODES_FN float svm_math(uint32_t type, float a, float b, float c)
{
  switch (type) {
    case CY_NODE_MATH_ADD:
      return a + b;
    case CY_NODE_MATH_SUBTRACT:
      return a - b;
    case CY_NODE_MATH_MULTIPLY:
      return a * b;
    case CY_NODE_MATH_DIVIDE:
      return nodes_safe_divide(a, b);
    case CY_NODE_MATH_POWER:
      return nodes_safe_powf(a, b);
    case CY_NODE_MATH_LOGARITHM:
      return nodes_safe_logf(a, b);
    case CY_NODE_MATH_SQRT:
      return sqrtf(fmaxf(a, 0.0f));
    case CY_NODE_MATH_INV_SQRT:
      return (a > 0.0f) ? 1.0f / sqrtf(a) : 0.0f;
    case CY_NODE_MATH_ABSOLUTE:
      return fabsf(a);
    case CY_NODE_MATH_RADIANS:
      return a * (CY_M_PI_F / 180.0f);
    case CY_NODE_MATH_DEGREES:
      return a * (180.0f / CY_M_PI_F);
    case CY_NODE_MATH_MINIMUM:
      return fminf(a, b);
    case CY_NODE_MATH_MAXIMUM:
      return fmaxf(a, b);
    case CY_NODE_MATH_LESS_THAN:
      return (a < b) ? 1.0f : 0.0f;
    case CY_NODE_MATH_GREATER_THAN:
      return (a > b) ? 1.0f : 0.0f;
    case CY_NODE_MATH_ROUND:
      return floorf(a + 0.5f);
    case CY_NODE_MATH_FLOOR:
      return floorf(a);
    case CY_NODE_MATH_CEIL:
      return ceilf(a);
    case CY_NODE_MATH_FRACTION:
      return a - floorf(a);
    case CY_NODE_MATH_MODULO:
      return nodes_safe_modulo(a, b);
    case CY_NODE_MATH_TRUNC:
      return a >= 0.0f ? floorf(a) : ceilf(a);
    case CY_NODE_MATH_SNAP:
      return floorf(nodes_safe_divide(a, b)) * b;
    case CY_NODE_MATH_WRAP:
      return nodes_wrapf(a, b, c);
    case CY_NODE_MATH_PINGPONG:
      return nodes_pingpongf(a, b);
    case CY_NODE_MATH_SINE:
      return sinf(a);
    case CY_NODE_MATH_COSINE:
      return cosf(a);
    case CY_NODE_MATH_TANGENT:
      return tanf(a);
    case CY_NODE_MATH_SINH:
      return sinhf(a);
    case CY_NODE_MATH_COSH:
      return coshf(a);
    case CY_NODE_MATH_TANH:
      return tanhf(a);
    case CY_NODE_MATH_ARCSINE:
      return asinf(nodes_clamp(a, -1.0f, 1.0f));
    case CY_NODE_MATH_ARCCOSINE:
      return acosf(nodes_clamp(a, -1.0f, 1.0f));
    case CY_NODE_MATH_ARCTANGENT:
      return atanf(a);
    case CY_NODE_MATH_ARCTAN2:
      return atan2f(a, b);
    case CY_NODE_MATH_SIGN:
      return (a == 0.0f) ? 0.0f : ((a < 0.0f) ? -1.0f : 1.0f);
    case CY_NODE_MATH_EXPONENT:
      return expf(a);
    case CY_NODE_MATH_COMPARE:
      return ((a == b) || (fabsf(a - b) <= fmaxf(c, FLT_EPSILON))) ? 1.0f : 0.0f;
    case CY_NODE_MATH_MULTIPLY_ADD:
      return a * b + c;
    case CY_NODE_MATH_SMOOTH_MIN:
      return nodes_smoothminf(a, b, c);
    case CY_NODE_MATH_SMOOTH_MAX:
      return -nodes_smoothminf(-a, -b, c);
    default:
      return 0.0f;
  }
}

CY_DEV f3 nodes_safe_divide3(f3 a, f3 b)
{
  return mk3((b.x != 0.0f) ? a.x / b.x : 0.0f, (b.y != 0.0f) ? a.y / b.y : 0.0f,
             (b.z != 0.0f) ? a.z / b.z : 0.0f);
}
CY_DEV f3 nodes_floor3(f3 a)
{
  return mk3(floorf(a.x), floorf(a.y), floorf(a.z));
}

SVM_NODES_FN void svm_vector_math(
    float *value, f3 *vector, uint32_t type, f3 a, f3 b, f3 c, float scale)
{
  switch (type) {
    case CY_NODE_VECTOR_MATH_ADD:
      *vector = a + b;
      break;
    case CY_NODE_VECTOR_MATH_SUBTRACT:
      *vector = a - b;
      break;
    case CY_NODE_VECTOR_MATH_MULTIPLY:
      *vector = a * b;
      break;
    case CY_NODE_VECTOR_MATH_DIVIDE:
      *vector = nodes_safe_divide3(a, b);
      break;
    case CY_NODE_VECTOR_MATH_CROSS_PRODUCT:
      *vector = cross(a, b);
      break;
    case CY_NODE_VECTOR_MATH_PROJECT: {
      const float l2 = dot(b, b);
      *vector = (l2 != 0.0f) ? (dot(a, b) / l2) * b : zero3();
      break;
    }
    case CY_NODE_VECTOR_MATH_REFLECT: {
      const f3 n = normalize(b);
      *vector = a - 2.0f * n * dot(a, n);
      break;
    }
    case CY_NODE_VECTOR_MATH_DOT_PRODUCT:
      *value = dot(a, b);
      break;
    case CY_NODE_VECTOR_MATH_DISTANCE:
      *value = len(a - b);
      break;
    case CY_NODE_VECTOR_MATH_LENGTH:
      *value = len(a);
      break;
    case CY_NODE_VECTOR_MATH_SCALE:
      *vector = a * scale;
      break;
    case CY_NODE_VECTOR_MATH_NORMALIZE:
      *vector = safe_normalize(a);
      break;
    case CY_NODE_VECTOR_MATH_SNAP:
      *vector = nodes_floor3(nodes_safe_divide3(a, b)) * b;
      break;
    case CY_NODE_VECTOR_MATH_FLOOR:
      *vector = nodes_floor3(a);
      break;
    case CY_NODE_VECTOR_MATH_CEIL:
      *vector = mk3(ceilf(a.x), ceilf(a.y), ceilf(a.z));
      break;
    case CY_NODE_VECTOR_MATH_MODULO:
      *vector = mk3(nodes_safe_modulo(a.x, b.x), nodes_safe_modulo(a.y, b.y),
                    nodes_safe_modulo(a.z, b.z));
      break;
    case CY_NODE_VECTOR_MATH_WRAP:
      *vector = mk3(nodes_wrapf(a.x, b.x, c.x), nodes_wrapf(a.y, b.y, c.y),
                    nodes_wrapf(a.z, b.z, c.z));
      break;
    case CY_NODE_VECTOR_MATH_FRACTION:
      *vector = a - nodes_floor3(a);
      break;
    case CY_NODE_VECTOR_MATH_ABSOLUTE:
      *vector = fabs3(a);
      break;
    case CY_NODE_VECTOR_MATH_MINIMUM:
      *vector = mk3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z));
      break;
    case CY_NODE_VECTOR_MATH_MAXIMUM:
      *vector = mk3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z));
      break;
    case CY_NODE_VECTOR_MATH_SINE:
      *vector = mk3(sinf(a.x), sinf(a.y), sinf(a.z));
      break;
    case CY_NODE_VECTOR_MATH_COSINE:
      *vector = mk3(cosf(a.x), cosf(a.y), cosf(a.z));
      break;
    case CY_NODE_VECTOR_MATH_TANGENT:
      *vector = mk3(tanf(a.x), tanf(a.y), tanf(a.z));
      break;
    default:
      *vector = zero3();
      *value = 0.0f;
  }
}

/* ---- colour mixing ---- */

CY_DEV f3 nodes_interp(f3 a, f3 b, float t)
{
  return a + t * (b - a);
}
CY_DEV float mix_overlay_1(float c1, float c2, float t, float tm)
{
  if (c1 < 0.5f)
    return c1 * (tm + 2.0f * t * c2);
  return 1.0f - (tm + 2.0f * t * (1.0f - c2)) * (1.0f - c1);
}

/* ---- HSV (util/util_color.h:85-166) ---- */

CY_DEV f3 rgb_to_hsv(f3 rgb)
{
  const float cmax = fmaxf(rgb.x, fmaxf(rgb.y, rgb.z));
  const float cmin = fminf(rgb.x, fminf(rgb.y, rgb.z));
  const float cdelta = cmax - cmin;
  float h = 0.0f;
  const float s = (cmax != 0.0f) ? cdelta / cmax : 0.0f;
  if (s != 0.0f) {
    const f3 c = (mk3(cmax, cmax, cmax) - rgb) / cdelta;
    if (rgb.x == cmax)
      h = c.z - c.y;
    else if (rgb.y == cmax)
      h = 2.0f + c.x - c.z;
    else
      h = 4.0f + c.y - c.x;
    h /= 6.0f;
    if (h < 0.0f)
      h += 1.0f;
  }
  return mk3(h, s, cmax);
}

CY_DEV f3 hsv_to_rgb(f3 hsv)
{
  float h = hsv.x;
  const float s = hsv.y, v = hsv.z;
  if (s == 0.0f)
    return mk3(v, v, v);
  if (h == 1.0f)
    h = 0.0f;
  h *= 6.0f;
  const float i = floorf(h);
  const float f = h - i;
  const float p = v * (1.0f - s);
  const float q = v * (1.0f - (s * f));
  const float t = v * (1.0f - (s * (1.0f - f)));
  if (i == 0.0f)
    return mk3(v, t, p);
  if (i == 1.0f)
    return mk3(q, v, p);
  if (i == 2.0f)
    return mk3(p, v, t);
  if (i == 3.0f)
    return mk3(p, q, v);
  if (i == 4.0f)
    return mk3(t, p, v);
  return mk3(v, p, q);
}

/* one channel of svm_mix_dodge / svm_mix_burn (svm_color_util.h:103-175) */
CY_DEV float mix_dodge_1(float c1, float c2, float t)
{
  if (c1 == 0.0f)
    return c1;
  float tmp = 1.0f - t * c2;
  if (tmp <= 0.0f)
    return 1.0f;
  tmp = c1 / tmp;
  return (tmp > 1.0f) ? 1.0f : tmp;
}
CY_DEV float mix_burn_1(float c1, float c2, float t, float tm)
{
  float tmp = tm + t * c2;
  if (tmp <= 0.0f)
    return 0.0f;
  tmp = 1.0f - (1.0f - c1) / tmp;
  if (tmp < 0.0f)
    return 0.0f;
  return (tmp > 1.0f) ? 1.0f : tmp;
}

SVM_NODES_FN f3 svm_mix(uint32_t type, float fac, f3 c1, f3 c2)
{
  const float t = saturate(fac);
  const float tm = 1.0f - t;
  const f3 one = one3();
  switch (type) {
    case CY_NODE_MIX_BLEND:
      return nodes_interp(c1, c2, t);
    case CY_NODE_MIX_ADD:
      return nodes_interp(c1, c1 + c2, t);
    case CY_NODE_MIX_MUL:
      return nodes_interp(c1, c1 * c2, t);
    case CY_NODE_MIX_SCREEN:
      return one - (mk3(tm, tm, tm) + t * (one - c2)) * (one - c1);
    case CY_NODE_MIX_OVERLAY:
      return mk3(mix_overlay_1(c1.x, c2.x, t, tm), mix_overlay_1(c1.y, c2.y, t, tm),
                 mix_overlay_1(c1.z, c2.z, t, tm));
    case CY_NODE_MIX_SUB:
      return nodes_interp(c1, c1 - c2, t);
    case CY_NODE_MIX_DIV: {
      f3 out = c1;
      if (c2.x != 0.0f)
        out.x = tm * out.x + t * out.x / c2.x;
      if (c2.y != 0.0f)
        out.y = tm * out.y + t * out.y / c2.y;
      if (c2.z != 0.0f)
        out.z = tm * out.z + t * out.z / c2.z;
      return out;
    }
    case CY_NODE_MIX_DIFF:
      return nodes_interp(c1, fabs3(c1 - c2), t);
    case CY_NODE_MIX_DARK:
      return nodes_interp(c1, mk3(fminf(c1.x, c2.x), fminf(c1.y, c2.y), fminf(c1.z, c2.z)), t);
    case CY_NODE_MIX_LIGHT:
      return nodes_interp(c1, mk3(fmaxf(c1.x, c2.x), fmaxf(c1.y, c2.y), fmaxf(c1.z, c2.z)), t);
    case CY_NODE_MIX_DODGE:
      return mk3(mix_dodge_1(c1.x, c2.x, t), mix_dodge_1(c1.y, c2.y, t),
                 mix_dodge_1(c1.z, c2.z, t));
    case CY_NODE_MIX_BURN:
      return mk3(mix_burn_1(c1.x, c2.x, t, tm), mix_burn_1(c1.y, c2.y, t, tm),
                 mix_burn_1(c1.z, c2.z, t, tm));
    case CY_NODE_MIX_HUE:
    case CY_NODE_MIX_COLOR: {
      /* svm_mix_hue / svm_mix_color: take hue (and saturation) of the second colour */
      const f3 hsv2 = rgb_to_hsv(c2);
      if (hsv2.y == 0.0f)
        return c1;
      f3 hsv = rgb_to_hsv(c1);
      hsv.x = hsv2.x;
      if (type == CY_NODE_MIX_COLOR)
        hsv.y = hsv2.y;
      return nodes_interp(c1, hsv_to_rgb(hsv), t);
    }
    case CY_NODE_MIX_SAT: {
      f3 hsv = rgb_to_hsv(c1);
      if (hsv.y == 0.0f)
        return c1;
      const f3 hsv2 = rgb_to_hsv(c2);
      hsv.y = tm * hsv.y + t * hsv2.y;
      return hsv_to_rgb(hsv);
    }
    case CY_NODE_MIX_VAL: {
      f3 hsv = rgb_to_hsv(c1);
      const f3 hsv2 = rgb_to_hsv(c2);
      hsv.z = tm * hsv.z + t * hsv2.z;
      return hsv_to_rgb(hsv);
    }
    case CY_NODE_MIX_SOFT: {
      const f3 scr = one - (one - c2) * (one - c1);
      return tm * c1 + t * ((one - c1) * c2 * c1 + c1 * scr);
    }
    case CY_NODE_MIX_LINEAR:
      return c1 + t * (2.0f * c2 + mk3(-1.0f, -1.0f, -1.0f));
    case CY_NODE_MIX_CLAMP:
      return mk3(saturate(c1.x), saturate(c1.y), saturate(c1.z));
    default:
      return zero3();
  }
}

/* ---- node bodies; `node` is the instruction, `offset` the program counter ---- */

CY_DEV void svm_node_convert(float *stack, uint32_t type, uint32_t from, uint32_t to)
{
  switch (type) {
    case CY_NODE_CONVERT_FI:
      stack[to] = __int_as_float((int)stack[from]);
      break;
    case CY_NODE_CONVERT_FV: {
      const float f = stack[from];
      stack_store_float3(stack, to, mk3(f, f, f));
      break;
    }
    case CY_NODE_CONVERT_CF:
    case CY_NODE_CONVERT_CI: {
      /* linear_rgb_to_gray - kernel_color.h:31-34 */
      const f3 c = stack_load_float3(stack, from);
      const float g = dot(c, mk3(kd_float(KD_FILM_RGB_TO_Y), kd_float(KD_FILM_RGB_TO_Y + 4),
                                 kd_float(KD_FILM_RGB_TO_Y + 8)));
      stack[to] = (type == CY_NODE_CONVERT_CF) ? g : __int_as_float((int)g);
      break;
    }
    case CY_NODE_CONVERT_VF:
    case CY_NODE_CONVERT_VI: {
      const float g = average(stack_load_float3(stack, from));
      stack[to] = (type == CY_NODE_CONVERT_VF) ? g : __int_as_float((int)g);
      break;
    }
    case CY_NODE_CONVERT_IF:
      stack[to] = (float)__float_as_int(stack[from]);
      break;
    case CY_NODE_CONVERT_IV: {
      const float f = (float)__float_as_int(stack[from]);
      stack_store_float3(stack, to, mk3(f, f, f));
      break;
    }
  }
}

CY_DEV void svm_node_fresnel(const ShaderDataG &sd, float *stack, uint4 node)
{
  uint32_t normal_offset, out_offset;
  unpack_uchar2(node.w, &normal_offset, &out_offset);
  float eta = stack_valid(node.y) ? stack[node.y] : __uint_as_float(node.z);
  const f3 normal_in = stack_valid(normal_offset) ? stack_load_float3(stack, normal_offset) : sd.N;
  eta = fmaxf(eta, 1e-5f);
  eta = (sd.flag & CY_SD_BACKFACING) ? 1.0f / eta : eta;
  stack[out_offset] = fresnel_dielectric_cos(dot(sd.I, normal_in), eta);
}

CY_DEV void svm_node_layer_weight(const ShaderDataG &sd, float *stack, uint4 node)
{
  uint32_t type, normal_offset, out_offset;
  unpack_uchar3(node.w, &type, &normal_offset, &out_offset);
  float blend = stack_valid(node.y) ? stack[node.y] : __uint_as_float(node.z);
  const f3 normal_in = stack_valid(normal_offset) ? stack_load_float3(stack, normal_offset) : sd.N;
  float f;
  if (type == CY_NODE_LAYER_WEIGHT_FRESNEL) {
    float eta = fmaxf(1.0f - blend, 1e-5f);
    eta = (sd.flag & CY_SD_BACKFACING) ? eta : 1.0f / eta;
    f = fresnel_dielectric_cos(dot(sd.I, normal_in), eta);
  }
  else {
    f = fabsf(dot(sd.I, normal_in));
    if (blend != 0.5f) {
      blend = nodes_clamp(blend, 0.0f, 1.0f - 1e-5f);
      blend = (blend < 0.5f) ? 2.0f * blend : 0.5f / (1.0f - blend);
      f = powf(f, blend);
    }
    f = 1.0f - f;
  }
  stack[out_offset] = f;
}

CY_DEV void svm_node_math(float *stack, uint4 node)
{
  uint32_t a, b, c;
  unpack_uchar3(node.z, &a, &b, &c);
  stack[node.w] = svm_math(node.y, stack[a], stack[b], stack[c]);
}

CY_DEV void svm_node_vector_math(float *stack, uint4 node, int *offset)
{
  uint32_t a_off, b_off, scale_off, value_off, vector_off;
  unpack_uchar3(node.z, &a_off, &b_off, &scale_off);
  unpack_uchar2(node.w, &value_off, &vector_off);
  const f3 a = stack_load_float3(stack, a_off);
  const f3 b = stack_load_float3(stack, b_off);
  f3 c = zero3();
  const float scale = stack[scale_off];
  if (node.y == CY_NODE_VECTOR_MATH_WRAP) {
    const uint4 extra = __ldg(&g_scene.svm_nodes[*offset]);
    (*offset)++;
    c = stack_load_float3(stack, extra.x);
  }
  float value = 0.0f;
  f3 vector = zero3();
  svm_vector_math(&value, &vector, node.y, a, b, c, scale);
  if (stack_valid(value_off))
    stack[value_off] = value;
  if (stack_valid(vector_off))
    stack_store_float3(stack, vector_off, vector);
}

CY_DEV void svm_node_mix(float *stack, uint4 node, int *offset)
{
  const uint4 node1 = __ldg(&g_scene.svm_nodes[*offset]);
  (*offset)++;
  const float fac = stack[node.y];
  const f3 c1 = stack_load_float3(stack, node.z);
  const f3 c2 = stack_load_float3(stack, node.w);
  stack_store_float3(stack, node1.z, svm_mix(node1.y, fac, c1, c2));
}

CY_DEV void svm_node_invert(float *stack, uint4 node)
{
  const float factor = stack[node.y];
  f3 color = stack_load_float3(stack, node.z);
  color.x = factor * (1.0f - color.x) + (1.0f - factor) * color.x;
  color.y = factor * (1.0f - color.y) + (1.0f - factor) * color.y;
  color.z = factor * (1.0f - color.z) + (1.0f - factor) * color.z;
  if (stack_valid(node.w))
    stack_store_float3(stack, node.w, color);
}

CY_DEV void svm_node_gamma(float *stack, uint4 node)
{
  f3 color = stack_load_float3(stack, node.z);
  const float gamma = stack[node.y];
  if (gamma == 0.0f) {
    color = one3();
  }
  else {
    if (color.x > 0.0f)
      color.x = powf(color.x, gamma);
    if (color.y > 0.0f)
      color.y = powf(color.y, gamma);
    if (color.z > 0.0f)
      color.z = powf(color.z, gamma);
  }
  if (stack_valid(node.w))
    stack_store_float3(stack, node.w, color);
}

CY_DEV void svm_node_brightness(float *stack, uint4 node)
{
  uint32_t bright_offset, contrast_offset;
  unpack_uchar2(node.w, &bright_offset, &contrast_offset);
  f3 color = stack_load_float3(stack, node.y);
  const float brightness = stack[bright_offset];
  const float contrast = stack[contrast_offset];
  const float a = 1.0f + contrast;
  const float b = brightness - contrast * 0.5f;
  color.x = fmaxf(a * color.x + b, 0.0f);
  color.y = fmaxf(a * color.y + b, 0.0f);
  color.z = fmaxf(a * color.z + b, 0.0f);
  if (stack_valid(node.z))
    stack_store_float3(stack, node.z, color);
}

CY_DEV void svm_node_clamp(float *stack, uint4 node, int *offset)
{
  uint32_t min_off, max_off, type;
  unpack_uchar3(node.z, &min_off, &max_off, &type);
  const uint4 defaults = __ldg(&g_scene.svm_nodes[*offset]);
  (*offset)++;
  const float value = stack[node.y];
  const float lo = stack_load_float_default(stack, min_off, defaults.x);
  const float hi = stack_load_float_default(stack, max_off, defaults.y);
  if (type == CY_NODE_CLAMP_RANGE && (lo > hi))
    stack[node.w] = nodes_clamp(value, hi, lo);
  else
    stack[node.w] = nodes_clamp(value, lo, hi);
}

/* svm_light_path.h:21-75 */
CY_DEV void svm_node_light_path(const ShaderDataG &sd, PathDepths depths, float *stack,
                                uint32_t type, uint32_t out_offset, uint32_t path_flag)
{
  float info = 0.0f;
  switch (type) {
    case CY_NODE_LP_camera:
      info = (path_flag & CY_PATH_RAY_CAMERA) ? 1.0f : 0.0f;
      break;
    case CY_NODE_LP_shadow:
      info = (path_flag & CY_PATH_RAY_SHADOW) ? 1.0f : 0.0f;
      break;
    case CY_NODE_LP_diffuse:
      info = (path_flag & CY_PATH_RAY_DIFFUSE) ? 1.0f : 0.0f;
      break;
    case CY_NODE_LP_glossy:
      info = (path_flag & CY_PATH_RAY_GLOSSY) ? 1.0f : 0.0f;
      break;
    case CY_NODE_LP_singular:
      info = (path_flag & CY_PATH_RAY_SINGULAR) ? 1.0f : 0.0f;
      break;
    case CY_NODE_LP_reflection:
      info = (path_flag & CY_PATH_RAY_REFLECT) ? 1.0f : 0.0f;
      break;
    case CY_NODE_LP_transmission:
      info = (path_flag & CY_PATH_RAY_TRANSMIT) ? 1.0f : 0.0f;
      break;
    case CY_NODE_LP_volume_scatter:
      info = (path_flag & CY_PATH_RAY_VOLUME_SCATTER) ? 1.0f : 0.0f;
      break;
    case CY_NODE_LP_backfacing:
      info = (sd.flag & CY_SD_BACKFACING) ? 1.0f : 0.0f;
      break;
    case CY_NODE_LP_ray_length:
      info = sd.ray_length;
      break;
    case CY_NODE_LP_ray_depth:
      info = (float)depths.bounce;
      break;
    case CY_NODE_LP_ray_diffuse:
      info = (float)depths.diffuse;
      break;
    case CY_NODE_LP_ray_glossy:
      info = (float)depths.glossy;
      break;
    case CY_NODE_LP_ray_transparent:
      info = (float)depths.transparent;
      break;
    case CY_NODE_LP_ray_transmission:
      info = (float)depths.transmission;
      break;
  }
  stack[out_offset] = info;
}

/* svm_light_path.h:79-112 */
CY_DEV void svm_node_light_falloff(const ShaderDataG &sd, float *stack, uint4 node)
{
  uint32_t strength_offset, smooth_offset, out_offset;
  unpack_uchar3(node.z, &strength_offset, &smooth_offset, &out_offset);
  float strength = stack[strength_offset];
  switch (node.y) {
    case CY_NODE_LIGHT_FALLOFF_QUADRATIC:
      break;
    case CY_NODE_LIGHT_FALLOFF_LINEAR:
      strength *= sd.ray_length;
      break;
    case CY_NODE_LIGHT_FALLOFF_CONSTANT:
      strength *= sd.ray_length * sd.ray_length;
      break;
  }
  const float smooth = stack[smooth_offset];
  if (smooth > 0.0f) {
    const float squared = sd.ray_length * sd.ray_length;
    /* distant lamps have ray_length FLT_MAX: the square overflows */
    if (isfinite_safe(squared))
      strength *= squared / (smooth + squared);
  }
  stack[out_offset] = strength;
}

/* svm_ramp.h:24-110: ColorRamp and RGB / vector curves; the table is stored in the SVM
 * program right behind the instruction, one float4 per entry */
CY_DEV float4 ramp_fetch(int offset)
{
  const uint4 n = __ldg(&g_scene.svm_nodes[offset]);
  return make_float4(__uint_as_float(n.x), __uint_as_float(n.y), __uint_as_float(n.z),
                     __uint_as_float(n.w));
}
CY_DEV float4 f4_lerp_terms(float wa, float4 a, float wb, float4 b)
{
  return make_float4(wa * a.x + wb * b.x, wa * a.y + wb * b.y, wa * a.z + wb * b.z,
                     wa * a.w + wb * b.w);
}
CY_DEV float4 rgb_ramp_lookup(int offset, float f, bool interpolate, bool extrapolate,
                              int table_size)
{
  if ((f < 0.0f || f > 1.0f) && extrapolate) {
    float4 t0, t1;
    if (f < 0.0f) {
      t0 = ramp_fetch(offset);
      t1 = ramp_fetch(offset + 1);
      f = -f;
    }
    else {
      t0 = ramp_fetch(offset + table_size - 1);
      t1 = ramp_fetch(offset + table_size - 2);
      f = f - 1.0f;
    }
    /* t0 + (t0 - t1) * f * (table_size - 1), evaluated left to right */
    const float n1 = (float)(table_size - 1);
    return make_float4(t0.x + (t0.x - t1.x) * f * n1, t0.y + (t0.y - t1.y) * f * n1,
                       t0.z + (t0.z - t1.z) * f * n1, t0.w + (t0.w - t1.w) * f * n1);
  }
  f = saturate(f) * (table_size - 1);
  const int i = min(max((int)f, 0), table_size - 1);
  const float t = f - (float)i;
  float4 a = ramp_fetch(offset + i);
  if (interpolate && t > 0.0f)
    a = f4_lerp_terms(1.0f - t, a, t, ramp_fetch(offset + i + 1));
  return a;
}
CY_DEV void svm_node_rgb_ramp(float *stack, uint4 node, int *offset)
{
  uint32_t fac_offset, color_offset, alpha_offset;
  unpack_uchar3(node.y, &fac_offset, &color_offset, &alpha_offset);
  const int table_size = (int)__ldg(&g_scene.svm_nodes[*offset]).x;
  (*offset)++;
  const float4 color = rgb_ramp_lookup(*offset, stack[fac_offset], node.z != 0, false, table_size);
  if (stack_valid(color_offset))
    stack_store_float3(stack, color_offset, mk3(color.x, color.y, color.z));
  if (stack_valid(alpha_offset))
    stack[alpha_offset] = color.w;
  *offset += table_size;
}
CY_DEV void svm_node_curves(float *stack, uint4 node, int *offset)
{
  uint32_t fac_offset, color_offset, out_offset;
  unpack_uchar3(node.y, &fac_offset, &color_offset, &out_offset);
  const int table_size = (int)__ldg(&g_scene.svm_nodes[*offset]).x;
  (*offset)++;
  const float fac = stack[fac_offset];
  f3 color = stack_load_float3(stack, color_offset);
  const float min_x = __uint_as_float(node.z), max_x = __uint_as_float(node.w);
  const float range_x = max_x - min_x;
  const f3 relpos = (color - mk3(min_x, min_x, min_x)) / range_x;
  const float r = rgb_ramp_lookup(*offset, relpos.x, true, true, table_size).x;
  const float g = rgb_ramp_lookup(*offset, relpos.y, true, true, table_size).y;
  const float b = rgb_ramp_lookup(*offset, relpos.z, true, true, table_size).z;
  color = (1.0f - fac) * color + fac * mk3(r, g, b);
  stack_store_float3(stack, out_offset, color);
  *offset += table_size;
}

/* ---- HSV, map range, normal, vector rotate (svm_hsv.h, svm_sepcomb_hsv.h,
 * svm_map_range.h, svm_normal.h, svm_vector_rotate.h) ---- */

__device__ __noinline__ void svm_node_hsv(float *stack, uint4 node)
{
  uint32_t in_color_offset, fac_offset, out_color_offset, hue_offset, sat_offset, val_offset;
  unpack_uchar3(node.y, &in_color_offset, &fac_offset, &out_color_offset);
  unpack_uchar3(node.z, &hue_offset, &sat_offset, &val_offset);
  const float fac = stack[fac_offset];
  const f3 in_color = stack_load_float3(stack, in_color_offset);
  f3 color = rgb_to_hsv(in_color);
  color.x = fmodf(color.x + stack[hue_offset] + 0.5f, 1.0f);
  color.y = saturate(color.y * stack[sat_offset]);
  color.z *= stack[val_offset];
  color = hsv_to_rgb(color);
  color.x = fmaxf(fac * color.x + (1.0f - fac) * in_color.x, 0.0f);
  color.y = fmaxf(fac * color.y + (1.0f - fac) * in_color.y, 0.0f);
  color.z = fmaxf(fac * color.z + (1.0f - fac) * in_color.z, 0.0f);
  if (stack_valid(out_color_offset))
    stack_store_float3(stack, out_color_offset, color);
}

__device__ __noinline__ void svm_node_combine_hsv(float *stack, uint4 node, int *offset)
{
  const uint32_t color_out = __ldg(&g_scene.svm_nodes[*offset]).y;
  (*offset)++;
  const f3 color = hsv_to_rgb(mk3(stack[node.y], stack[node.z], stack[node.w]));
  if (stack_valid(color_out))
    stack_store_float3(stack, color_out, color);
}

__device__ __noinline__ void svm_node_separate_hsv(float *stack, uint4 node, int *offset)
{
  const uint32_t value_out = __ldg(&g_scene.svm_nodes[*offset]).y;
  (*offset)++;
  const f3 hsv = rgb_to_hsv(stack_load_float3(stack, node.y));
  if (stack_valid(node.z))
    stack[node.z] = hsv.x;
  if (stack_valid(node.w))
    stack[node.w] = hsv.y;
  if (stack_valid(value_out))
    stack[value_out] = hsv.z;
}

CY_DEV float nodes_smoothstep(float edge0, float edge1, float x)
{
  if (x < edge0)
    return 0.0f;
  if (x >= edge1)
    return 1.0f;
  const float t = (x - edge0) / (edge1 - edge0);
  return (3.0f - 2.0f * t) * (t * t);
}
CY_DEV float nodes_smootherstep(float edge0, float edge1, float x)
{
  x = clampf(nodes_safe_divide(x - edge0, edge1 - edge0), 0.0f, 1.0f);
  return x * x * x * (x * (x * 6.0f - 15.0f) + 10.0f);
}

__device__ __noinline__ void svm_node_map_range(float *stack, uint4 node, int *offset)
{
  const uint32_t from_min_offset = node.z & 0xff, from_max_offset = (node.z >> 8) & 0xff,
                 to_min_offset = (node.z >> 16) & 0xff, to_max_offset = (node.z >> 24) & 0xff;
  uint32_t type, steps_offset, result_offset;
  unpack_uchar3(node.w, &type, &steps_offset, &result_offset);
  const uint4 defaults = __ldg(&g_scene.svm_nodes[*offset]);
  const uint4 defaults2 = __ldg(&g_scene.svm_nodes[*offset + 1]);
  *offset += 2;
  const float value = stack[node.y];
  const float from_min = stack_load_float_default(stack, from_min_offset, defaults.x);
  const float from_max = stack_load_float_default(stack, from_max_offset, defaults.y);
  const float to_min = stack_load_float_default(stack, to_min_offset, defaults.z);
  const float to_max = stack_load_float_default(stack, to_max_offset, defaults.w);
  const float steps = stack_load_float_default(stack, steps_offset, defaults2.x);
  float result = 0.0f;
  if (from_max != from_min) {
    float factor = value;
    switch (type) {
      default:
      case CY_NODE_MAP_RANGE_LINEAR:
        factor = (value - from_min) / (from_max - from_min);
        break;
      case CY_NODE_MAP_RANGE_STEPPED:
        factor = (value - from_min) / (from_max - from_min);
        factor = (steps > 0.0f) ? floorf(factor * (steps + 1.0f)) / steps : 0.0f;
        break;
      case CY_NODE_MAP_RANGE_SMOOTHSTEP:
        factor = (from_min > from_max) ? 1.0f - nodes_smoothstep(from_max, from_min, factor) :
                                         nodes_smoothstep(from_min, from_max, factor);
        break;
      case CY_NODE_MAP_RANGE_SMOOTHERSTEP:
        factor = (from_min > from_max) ? 1.0f - nodes_smootherstep(from_max, from_min, factor) :
                                         nodes_smootherstep(from_min, from_max, factor);
        break;
    }
    result = to_min + factor * (to_max - to_min);
  }
  stack[result_offset] = result;
}

__device__ __noinline__ void svm_node_normal(float *stack, uint4 node, int *offset)
{
  const uint4 node1 = __ldg(&g_scene.svm_nodes[*offset]);
  (*offset)++;
  const f3 normal = stack_load_float3(stack, node.y);
  const f3 direction = normalize(mk3(__uint_as_float(node1.x), __uint_as_float(node1.y),
                                     __uint_as_float(node1.z)));
  if (stack_valid(node.z))
    stack_store_float3(stack, node.z, direction);
  if (stack_valid(node.w))
    stack[node.w] = dot(direction, normalize(normal));
}

/* util_math.h:563-582 */
CY_DEV f3 nodes_rotate_around_axis(f3 p, f3 axis, float angle)
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

CY_DEV tfm34 nodes_euler_to_transform(f3 e)
{
  const float cx = cosf(e.x), cy = cosf(e.y), cz = cosf(e.z);
  const float sx = sinf(e.x), sy = sinf(e.y), sz = sinf(e.z);
  tfm34 t;
  t.x = make_float4(cy * cz, sy * sx * cz - cx * sz, sy * cx * cz + sx * sz, 0.0f);
  t.y = make_float4(cy * sz, sy * sx * sz + cx * cz, sy * cx * sz - sx * cz, 0.0f);
  t.z = make_float4(-sy, cy * sx, cy * cx, 0.0f);
  return t;
}

__device__ __noinline__ void svm_node_vector_rotate(float *stack, uint4 node)
{
  const uint32_t type = node.y & 0xff, vector_offset = (node.y >> 8) & 0xff,
                 rotation_offset = (node.y >> 16) & 0xff, invert = (node.y >> 24) & 0xff;
  uint32_t center_offset, axis_offset, angle_offset;
  unpack_uchar3(node.z, &center_offset, &axis_offset, &angle_offset);
  if (!stack_valid(node.w))
    return;
  const f3 vector = stack_load_float3(stack, vector_offset);
  const f3 center = stack_load_float3(stack, center_offset);
  f3 result;
  if (type == CY_NODE_VECTOR_ROTATE_TYPE_EULER_XYZ) {
    const tfm34 rot = nodes_euler_to_transform(stack_load_float3(stack, rotation_offset));
    result = (invert ? transform_direction_transposed(rot, vector - center) :
                       transform_direction(rot, vector - center)) +
             center;
  }
  else {
    f3 axis;
    switch (type) {
      case CY_NODE_VECTOR_ROTATE_TYPE_AXIS_X:
        axis = mk3(1.0f, 0.0f, 0.0f);
        break;
      case CY_NODE_VECTOR_ROTATE_TYPE_AXIS_Y:
        axis = mk3(0.0f, 1.0f, 0.0f);
        break;
      case CY_NODE_VECTOR_ROTATE_TYPE_AXIS_Z:
        axis = mk3(0.0f, 0.0f, 1.0f);
        break;
      default:
        axis = normalize(stack_load_float3(stack, axis_offset));
        break;
    }
    float angle = stack[angle_offset];
    angle = invert ? -angle : angle;
    result = (len_squared(axis) != 0.0f) ?
                 nodes_rotate_around_axis(vector - center, axis, angle) + center :
                 vector;
  }
  stack_store_float3(stack, node.w, result);
}

#endif /* B200_SVM_NODES_CUH */

/* cycles_port.c - plain-C RESTATEMENT of the reference's scene_intersect for the
 * hot path (BVH2, static triangles, two-level instancing, closest hit and the
 * opaque-shadow early-out).  TEST INFRASTRUCTURE ONLY: the portable checker that
 * needs nothing from /root/reference at run time.  It is pinned against the
 * reference itself: tests/test_oracle_cpu.py compares it bit for bit with the
 * golden vectors dumped from oracle/_ref (tests/golden/make_golden.py).
 *
 * Functions cite the reference code they restate (paths under
 * /root/reference/blender/intern/cycles/).  Build: make -C oracle -f Makefile.port
 * (gcc -O2 -ffp-contract=off: no FMA contraction, like the generic CPU kernel).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct {
  float x, y, z;
} v3;

typedef struct {
  const float *bvh_nodes;      /* __bvh_nodes, float4 units */
  const float *bvh_leaf_nodes; /* __bvh_leaf_nodes */
  const float *prim_tri_verts; /* __prim_tri_verts */
  const uint32_t *prim_tri_index;
  const uint32_t *prim_visibility;
  const uint32_t *prim_object;
  const int32_t *object_node;
  const uint8_t *objects; /* KernelObject[] */
  uint32_t object_stride, object_itfm_offset;
  int32_t root; /* KernelData.bvh.root */
} port_scene;

typedef struct {
  float P[3], t, D[3];
  uint32_t visibility;
} port_ray;

typedef struct {
  float t, u, v;
  int32_t prim, object, type;
} port_hit;

#define ENTRYPOINT_SENTINEL 0x76543210 /* kernel/bvh/bvh_types.h:30 */
#define BVH_STACK_SIZE 192             /* bvh_types.h:33 */
#define OBJECT_NONE (-1)
#define PRIM_NONE (-1)
#define PRIMITIVE_TRIANGLE 1u
#define PATH_RAY_SHADOW_OPAQUE 0x180u

static inline int as_int(float f)
{
  int i;
  memcpy(&i, &f, 4);
  return i;
}
static inline float as_float(int i)
{
  float f;
  memcpy(&f, &i, 4);
  return f;
}
static inline v3 mk(float x, float y, float z)
{
  v3 r = {x, y, z};
  return r;
}
static inline v3 sub(v3 a, v3 b)
{
  return mk(a.x - b.x, a.y - b.y, a.z - b.z);
}
static inline v3 add(v3 a, v3 b)
{
  return mk(a.x + b.x, a.y + b.y, a.z + b.z);
}
/* util/util_math_float3.h:233-255 */
static inline float dot(v3 a, v3 b)
{
  return a.x * b.x + a.y * b.y + a.z * b.z;
}
static inline v3 cross(v3 a, v3 b)
{
  return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

/* geom/geom_object.h:414-425 */
static inline v3 bvh_clamp_direction(v3 dir)
{
  const float ooeps = 8.271806E-25f;
  return mk((fabsf(dir.x) > ooeps) ? dir.x : copysignf(ooeps, dir.x),
            (fabsf(dir.y) > ooeps) ? dir.y : copysignf(ooeps, dir.y),
            (fabsf(dir.z) > ooeps) ? dir.z : copysignf(ooeps, dir.z));
}
static inline v3 bvh_inverse_direction(v3 dir)
{
  return mk(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
}

static inline float max4(float a, float b, float c, float d)
{
  return fmaxf(fmaxf(a, b), fmaxf(c, d));
}
static inline float min4(float a, float b, float c, float d)
{
  return fminf(fminf(a, b), fminf(c, d));
}

/* kernel/bvh/bvh_nodes.h:31-77 (bvh_aligned_node_intersect with visibility) */
static int node_intersect(const port_scene *s, v3 P, v3 idir, float t, int node_addr,
                          uint32_t visibility, float dist[2])
{
  const float *cn = s->bvh_nodes + 4 * (size_t)node_addr;
  const float *n0 = cn + 4, *n1 = cn + 8, *n2 = cn + 12;
  float c0lox = (n0[0] - P.x) * idir.x, c0hix = (n0[2] - P.x) * idir.x;
  float c0loy = (n1[0] - P.y) * idir.y, c0hiy = (n1[2] - P.y) * idir.y;
  float c0loz = (n2[0] - P.z) * idir.z, c0hiz = (n2[2] - P.z) * idir.z;
  float c0min = max4(0.0f, fminf(c0lox, c0hix), fminf(c0loy, c0hiy), fminf(c0loz, c0hiz));
  float c0max = min4(t, fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy), fmaxf(c0loz, c0hiz));
  float c1lox = (n0[1] - P.x) * idir.x, c1hix = (n0[3] - P.x) * idir.x;
  float c1loy = (n1[1] - P.y) * idir.y, c1hiy = (n1[3] - P.y) * idir.y;
  float c1loz = (n2[1] - P.z) * idir.z, c1hiz = (n2[3] - P.z) * idir.z;
  float c1min = max4(0.0f, fminf(c1lox, c1hix), fminf(c1loy, c1hiy), fminf(c1loz, c1hiz));
  float c1max = min4(t, fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy), fmaxf(c1loz, c1hiz));
  dist[0] = c0min;
  dist[1] = c1min;
  return (((c0max >= c0min) && ((uint32_t)as_int(cn[0]) & visibility)) ? 1 : 0) |
         (((c1max >= c1min) && ((uint32_t)as_int(cn[1]) & visibility)) ? 2 : 0);
}

/* util/util_math_intersect.h:88-195, scalar branch */
static int ray_triangle_intersect(v3 P, v3 dir, float ray_t, v3 tri_a, v3 tri_b, v3 tri_c,
                                  float *isect_u, float *isect_v, float *isect_t)
{
  const v3 v0 = sub(tri_c, P), v1 = sub(tri_a, P), v2 = sub(tri_b, P);
  const v3 e0 = sub(v2, v0), e1 = sub(v0, v1), e2 = sub(v1, v2);
  const float U = dot(cross(add(v2, v0), e0), dir);
  const float V = dot(cross(add(v0, v1), e1), dir);
  const float W = dot(cross(add(v1, v2), e2), dir);
  const float minUVW = fminf(U, fminf(V, W));
  const float maxUVW = fmaxf(U, fmaxf(V, W));
  if (minUVW < 0.0f && maxUVW > 0.0f)
    return 0;
  const v3 Ng1 = cross(e1, e0);
  const v3 Ng = add(Ng1, Ng1);
  const float den = dot(Ng, dir);
  if (den == 0.0f)
    return 0;
  const float T = dot(v0, Ng);
  const int sign_den = (as_int(den) & 0x80000000);
  const float sign_T = as_float(as_int(T) ^ sign_den);
  if ((sign_T < 0.0f) || (sign_T > ray_t * as_float(as_int(den) ^ sign_den)))
    return 0;
  const float inv_den = 1.0f / den;
  *isect_u = U * inv_den;
  *isect_v = V * inv_den;
  *isect_t = T * inv_den;
  return 1;
}

/* geom/geom_triangle_intersect.h:25-72 */
static int triangle_intersect(const port_scene *s, port_hit *isect, v3 P, v3 dir,
                              uint32_t visibility, int object, int prim_addr)
{
  const uint32_t vi = s->prim_tri_index[prim_addr];
  const float *a = s->prim_tri_verts + 4 * (size_t)vi, *b = a + 4, *c = a + 8;
  float t, u, v;
  if (ray_triangle_intersect(P, dir, isect->t, mk(a[0], a[1], a[2]), mk(b[0], b[1], b[2]),
                             mk(c[0], c[1], c[2]), &u, &v, &t)) {
    if (s->prim_visibility[prim_addr] & visibility) {
      isect->prim = prim_addr;
      isect->object = object;
      isect->type = PRIMITIVE_TRIANGLE;
      isect->u = u;
      isect->v = v;
      isect->t = t;
      return 1;
    }
  }
  return 0;
}

/* util/util_transform.h:56-108 on KernelObject::itfm */
static inline v3 tfm_point(const float *m, v3 a)
{
  return mk(a.x * m[0] + a.y * m[1] + a.z * m[2] + m[3], a.x * m[4] + a.y * m[5] + a.z * m[6] + m[7],
            a.x * m[8] + a.y * m[9] + a.z * m[10] + m[11]);
}
static inline v3 tfm_direction(const float *m, v3 a)
{
  return mk(a.x * m[0] + a.y * m[1] + a.z * m[2], a.x * m[4] + a.y * m[5] + a.z * m[6],
            a.x * m[8] + a.y * m[9] + a.z * m[10]);
}
static inline const float *object_itfm(const port_scene *s, int object)
{
  return (const float *)(s->objects + (size_t)object * s->object_stride + s->object_itfm_offset);
}

/* kernel/bvh/bvh_traversal.h:34-227 (BVH_FUNCTION_FEATURES = 0) with
 * geom/geom_object.h:427-460 for the instance push / pop */
static int bvh_intersect(const port_scene *s, const port_ray *ray, port_hit *isect,
                         uint32_t visibility)
{
  int traversal_stack[BVH_STACK_SIZE];
  traversal_stack[0] = ENTRYPOINT_SENTINEL;
  int stack_ptr = 0;
  int node_addr = s->root;
  const v3 rayP = mk(ray->P[0], ray->P[1], ray->P[2]), rayD = mk(ray->D[0], ray->D[1], ray->D[2]);
  v3 P = rayP;
  v3 dir = bvh_clamp_direction(rayD);
  v3 idir = bvh_inverse_direction(dir);
  int object = OBJECT_NONE;

  isect->t = ray->t;
  isect->u = 0.0f;
  isect->v = 0.0f;
  isect->prim = PRIM_NONE;
  isect->object = OBJECT_NONE;
  isect->type = 0;

  do {
    do {
      while (node_addr >= 0 && node_addr != ENTRYPOINT_SENTINEL) {
        int node_addr_child1, traverse_mask;
        float dist[2];
        const float *cnodes = s->bvh_nodes + 4 * (size_t)node_addr;
        traverse_mask = node_intersect(s, P, idir, isect->t, node_addr, visibility, dist);
        node_addr = as_int(cnodes[2]);
        node_addr_child1 = as_int(cnodes[3]);
        if (traverse_mask == 3) {
          int is_closest_child1 = (dist[1] < dist[0]);
          if (is_closest_child1) {
            int tmp = node_addr;
            node_addr = node_addr_child1;
            node_addr_child1 = tmp;
          }
          ++stack_ptr;
          traversal_stack[stack_ptr] = node_addr_child1;
        }
        else {
          if (traverse_mask == 2) {
            node_addr = node_addr_child1;
          }
          else if (traverse_mask == 0) {
            node_addr = traversal_stack[stack_ptr];
            --stack_ptr;
          }
        }
      }
      if (node_addr < 0) {
        const float *leaf = s->bvh_leaf_nodes + 4 * (size_t)(-node_addr - 1);
        int prim_addr = as_int(leaf[0]);
        if (prim_addr >= 0) {
          const int prim_addr2 = as_int(leaf[1]);
          node_addr = traversal_stack[stack_ptr];
          --stack_ptr;
          for (; prim_addr < prim_addr2; prim_addr++) {
            if (triangle_intersect(s, isect, P, dir, visibility, object, prim_addr)) {
              if (visibility & PATH_RAY_SHADOW_OPAQUE)
                return 1;
            }
          }
        }
        else {
          /* instance push */
          object = (int)s->prim_object[-prim_addr - 1];
          {
            const float *itfm = object_itfm(s, object);
            P = tfm_point(itfm, rayP);
            v3 d = tfm_direction(itfm, rayD);
            float len = sqrtf(dot(d, d));
            float x = 1.0f / len;
            dir = bvh_clamp_direction(mk(d.x * x, d.y * x, d.z * x));
            idir = bvh_inverse_direction(dir);
            if (isect->t != __FLT_MAX__)
              isect->t *= len;
          }
          ++stack_ptr;
          traversal_stack[stack_ptr] = ENTRYPOINT_SENTINEL;
          node_addr = s->object_node[object];
        }
      }
    } while (node_addr != ENTRYPOINT_SENTINEL);

    if (stack_ptr >= 0) {
      /* instance pop */
      if (isect->t != __FLT_MAX__) {
        const float *itfm = object_itfm(s, object);
        v3 d = tfm_direction(itfm, rayD);
        isect->t /= sqrtf(dot(d, d));
      }
      P = rayP;
      dir = bvh_clamp_direction(rayD);
      idir = bvh_inverse_direction(dir);
      object = OBJECT_NONE;
      node_addr = traversal_stack[stack_ptr];
      --stack_ptr;
    }
  } while (node_addr != ENTRYPOINT_SENTINEL);

  return (isect->prim != PRIM_NONE);
}

/* kernel/bvh/bvh.h:154-237 for a batch (ray.t == 0 marks an inactive ray) */
void port_scene_intersect(const port_scene *s, const port_ray *rays, port_hit *hits, uint64_t n)
{
  for (uint64_t i = 0; i < n; i++) {
    port_hit h;
    int hit = 0;
    if (rays[i].t != 0.0f)
      hit = bvh_intersect(s, &rays[i], &h, rays[i].visibility);
    if (!hit) {
      h.t = rays[i].t;
      h.u = h.v = 0.0f;
      h.prim = -1;
      h.object = -1;
      h.type = 0;
    }
    hits[i] = h;
  }
}

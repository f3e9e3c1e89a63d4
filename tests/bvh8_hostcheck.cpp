/* tests/bvh8_hostcheck.cpp - TEST INFRASTRUCTURE (never shipped, never timed).
 *
 * Links the product's host BVH8 builder (raytracingproject_b200/csrc/bvh8_build.cpp)
 * and checks its output on the CPU so that `pytest -m "not gpu"` covers the host
 * logic:  (1) structural invariants - every triangle appears in exactly one leaf
 * record, quantised child boxes contain the true boxes, inner children are laid
 * out contiguously in slot order;  (2) a plain recursive walk of the BVH8 with
 * the reference's scalar triangle test (util/util_math_intersect.h:88-195) whose
 * closest-hit prim ids are compared with the oracle's BVH2 traversal.
 * Built by tests/conftest.py with -ffp-contract=off. */
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../raytracingproject_b200/csrc/bvh8_build.h"
#include "../include/cycles_abi.h"

using namespace b200;

struct Check {
  BVH8Output out;
  const uint8_t *objects;
  std::string error;
};

static float as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static uint32_t as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

struct V3 { float x, y, z; };
static V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

static bool tri_test(V3 P, V3 dir, float ray_t, V3 a, V3 b, V3 c, float *u, float *v, float *t)
{
  V3 v0 = sub(c, P), v1 = sub(a, P), v2 = sub(b, P);
  V3 e0 = sub(v2, v0), e1 = sub(v0, v1), e2 = sub(v1, v2);
  float U = dot(cross(add(v2, v0), e0), dir);
  float V = dot(cross(add(v0, v1), e1), dir);
  float W = dot(cross(add(v1, v2), e2), dir);
  float mn = fminf(U, fminf(V, W)), mx = fmaxf(U, fmaxf(V, W));
  if (mn < 0.0f && mx > 0.0f) return false;
  V3 Ng1 = cross(e1, e0);
  V3 Ng = add(Ng1, Ng1);
  float den = dot(Ng, dir);
  if (den == 0.0f) return false;
  float T = dot(v0, Ng);
  uint32_t sign = as_uint(den) & 0x80000000u;
  float sT = as_float(as_uint(T) ^ sign);
  if (sT < 0.0f || sT > ray_t * as_float(as_uint(den) ^ sign)) return false;
  float inv = 1.0f / den;
  *u = U * inv; *v = V * inv; *t = T * inv;
  return true;
}

static V3 clamp_dir(V3 d)
{
  const float e = 8.271806E-25f;
  return {fabsf(d.x) > e ? d.x : copysignf(e, d.x), fabsf(d.y) > e ? d.y : copysignf(e, d.y),
          fabsf(d.z) > e ? d.z : copysignf(e, d.z)};
}

struct Hit { float t, u, v; int prim, object; };

static void walk(const Check &ck, uint32_t node_index, V3 P, V3 D, V3 wP, V3 wD, int object,
                 uint32_t vis, float &tmax, Hit &hit, bool &inst_hit)
{
  const BVH8Node &n = ck.out.nodes[node_index];
  V3 dir = clamp_dir(D);
  uint32_t inner_i = 0;
  for (int s = 0; s < 8; s++) {
    uint8_t meta = n.meta[s];
    bool inner = (n.imask >> s) & 1;
    uint32_t child = n.child_base + inner_i;
    if (inner) inner_i++;
    if (meta == 0) continue;
    /* dequantised slab test in double, generous */
    double t0 = 0.0, t1 = tmax;
    const float pd[3] = {dir.x, dir.y, dir.z}, pp[3] = {P.x, P.y, P.z};
    bool ok = true;
    for (int k = 0; k < 3 && ok; k++) {
      double sc = std::ldexp(1.0, (int)n.e[k] - 127);
      double lo = (double)n.origin[k] + n.qlo[k][s] * sc, hi = (double)n.origin[k] + n.qhi[k][s] * sc;
      double a = (lo - pp[k]) / pd[k], b = (hi - pp[k]) / pd[k];
      if (a > b) std::swap(a, b);
      t0 = std::max(t0, a * (1 - 1e-6) - 1e-30); t1 = std::min(t1, b * (1 + 1e-6) + 1e-30);
      if (t0 > t1) ok = false;
    }
    if (!ok) continue;
    if (inner) { walk(ck, child, P, D, wP, wD, object, vis, tmax, hit, inst_hit); continue; }
    int count = (meta >> 5) == 1 ? 1 : (meta >> 5) == 3 ? 2 : 3;
    uint32_t off = meta & 31;
    for (int r = 0; r < count; r++) {
      const float *rec = &ck.out.records[12 * (size_t)(n.prim_base + off + r)];
      int tag = (int)as_uint(rec[3]);
      if (tag >= 0) {
        float u, v, t;
        if (tri_test(P, dir, tmax, {rec[0], rec[1], rec[2]}, {rec[4], rec[5], rec[6]},
                     {rec[8], rec[9], rec[10]}, &u, &v, &t) && (as_uint(rec[7]) & vis)) {
          hit = {t, u, v, tag, object}; tmax = t; if (object >= 0) inst_hit = true;
        }
      }
      else if (as_uint(rec[1]) & vis) {
        int ob = ~tag;
        const float *itfm = (const float *)(ck.objects + (size_t)ob * SIZEOF_KERNEL_OBJECT + KO_ITFM);
        V3 oP = {wP.x * itfm[0] + wP.y * itfm[1] + wP.z * itfm[2] + itfm[3],
                 wP.x * itfm[4] + wP.y * itfm[5] + wP.z * itfm[6] + itfm[7],
                 wP.x * itfm[8] + wP.y * itfm[9] + wP.z * itfm[10] + itfm[11]};
        V3 oD = {wD.x * itfm[0] + wD.y * itfm[1] + wD.z * itfm[2],
                 wD.x * itfm[4] + wD.y * itfm[5] + wD.z * itfm[6],
                 wD.x * itfm[8] + wD.y * itfm[9] + wD.z * itfm[10]};
        float len = sqrtf(dot(oD, oD)); float x = 1.0f / len; oD = {oD.x * x, oD.y * x, oD.z * x};
        float ot = (tmax != __FLT_MAX__) ? tmax * len : tmax;
        bool ih = false;
        walk(ck, as_uint(rec[0]), oP, oD, wP, wD, ob, vis, ot, hit, ih);
        if (ih) { tmax = ot / len; hit.t = tmax; }
      }
    }
  }
}

extern "C" {

void *hostcheck_build(const float *nodes, size_t n_nodes_f4, const float *leaves, size_t n_leaves_f4,
                      const float *verts, const uint32_t *tri_index, const uint32_t *prim_vis,
                      const uint32_t *prim_object, size_t num_prims, const int32_t *object_node,
                      const uint8_t *objects, size_t num_objects, int32_t root, char *err, size_t errlen)
{
  BVH2Input in; memset(&in, 0, sizeof(in));
  in.nodes = nodes; in.num_nodes_f4 = n_nodes_f4; in.leaf_nodes = leaves; in.num_leaf_nodes_f4 = n_leaves_f4;
  in.prim_tri_verts = verts; in.prim_tri_index = tri_index; in.prim_visibility = prim_vis;
  in.prim_object = prim_object; in.num_prims = num_prims; in.object_node = object_node;
  in.objects = objects; in.object_stride = SIZEOF_KERNEL_OBJECT; in.object_tfm_offset = KO_TFM;
  in.num_objects = num_objects; in.root = root;
  in.node_unaligned_flag = CY_PATH_RAY_NODE_UNALIGNED; in.primitive_all = CY_PRIMITIVE_ALL;
  in.primitive_triangle = CY_PRIMITIVE_TRIANGLE;
  in.tighten_instances = true;
  Check *ck = new Check(); ck->objects = objects;
  std::string e;
  if (!build_bvh8(in, ck->out, e)) {
    snprintf(err, errlen, "%s", e.c_str()); delete ck; return nullptr;
  }
  return ck;
}

void hostcheck_free(void *p) { delete (Check *)p; }

void hostcheck_info(void *p, uint64_t *nodes, uint64_t *records, uint64_t *tris, uint64_t *insts,
                    uint32_t *depth, float *sah)
{
  Check *ck = (Check *)p;
  *nodes = ck->out.nodes.size(); *records = ck->out.records.size() / 12;
  *tris = ck->out.num_triangles; *insts = ck->out.num_instances; *depth = ck->out.max_depth;
  *sah = ck->out.sah_cost;
}

/* Structural invariants; returns the number of violations (0 = ok).
 * prim_count[prim_addr] receives how many leaf records reference each prim. */
int hostcheck_invariants(void *p, uint32_t *prim_count, size_t num_prims)
{
  Check *ck = (Check *)p;
  int bad = 0;
  std::vector<int> node_refs(ck->out.nodes.size(), 0);
  for (size_t ni = 0; ni < ck->out.nodes.size(); ni++) {
    const BVH8Node &n = ck->out.nodes[ni];
    uint32_t inner_i = 0, expect_off = 0;
    for (int s = 0; s < 8; s++) {
      bool inner = (n.imask >> s) & 1;
      uint8_t meta = n.meta[s];
      if (inner) {
        if (meta != ((1u << 5) | (24 + s))) bad++;
        uint32_t c = n.child_base + inner_i++;
        if (c >= ck->out.nodes.size()) { bad++; continue; }
        node_refs[c]++;
        /* child's own frame must lie inside this slot's quantised box */
        const BVH8Node &cn = ck->out.nodes[c];
        for (int k = 0; k < 3; k++) {
          double sc = std::ldexp(1.0, (int)n.e[k] - 127);
          double lo = (double)n.origin[k] + n.qlo[k][s] * sc;
          if ((double)cn.origin[k] < lo - 1e-9 * std::fabs(lo)) bad++;
        }
      }
      else if (meta) {
        int unary = meta >> 5; int count = unary == 1 ? 1 : unary == 3 ? 2 : unary == 7 ? 3 : -1;
        if (count < 0) { bad++; continue; }
        (void)expect_off;
        uint32_t off = meta & 31; if (off + count > 24) bad++;
        for (int r = 0; r < count; r++) {
          size_t ri = (size_t)n.prim_base + off + r;
          if (12 * ri + 11 >= ck->out.records.size()) { bad++; continue; }
          const float *rec = &ck->out.records[12 * ri];
          int tag = (int)as_uint(rec[3]);
          if (tag >= 0) {
            if ((size_t)tag >= num_prims) { bad++; continue; }
            prim_count[tag]++;
            for (int vtx = 0; vtx < 3; vtx++) for (int k = 0; k < 3; k++) {
              double sc = std::ldexp(1.0, (int)n.e[k] - 127);
              double lo = (double)n.origin[k] + n.qlo[k][s] * sc, hi = (double)n.origin[k] + n.qhi[k][s] * sc;
              double x = rec[4 * vtx + k];
              if (x < lo || x > hi) bad++;
            }
          }
        }
      }
    }
  }
  return bad;
}

void hostcheck_intersect(void *p, const float *rays /* 8 floats each */, size_t n, Hit *hits)
{
  Check *ck = (Check *)p;
  for (size_t i = 0; i < n; i++) {
    const float *r = rays + 8 * i;
    V3 P = {r[0], r[1], r[2]}, D = {r[4], r[5], r[6]};
    float tmax = r[3]; uint32_t vis = as_uint(r[7]);
    Hit h = {tmax, 0, 0, -1, -1}; bool ih = false;
    if (tmax != 0.0f) walk(*ck, ck->out.root, P, D, P, D, -1, vis, tmax, h, ih);
    if (h.prim < 0) h.t = r[3];
    hits[i] = h;
  }
}

}

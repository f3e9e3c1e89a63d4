/* bvh8_build.cpp - host builder of the compressed BVH8 (see bvh8.h).
 *
 * Input: the reference's packed BVH2 arrays exactly as a Device receives them
 * (__bvh_nodes / __bvh_leaf_nodes layout: intern/cycles/bvh/bvh2.cpp:40-116,
 * merged TLAS+BLAS addressing: bvh/bvh.cpp:323-519, traversal semantics:
 * kernel/bvh/bvh_traversal.h:34-227).  The SAH binary tree those arrays encode
 * (built by bvh/bvh_build.cpp) is collapsed 2 -> 8 wide by repeatedly opening
 * the child with the largest surface area, leaves larger than three triangles
 * are split, child slots are ordered per ray octant, and boxes are quantised
 * conservatively.  Plays the role BVH2::pack_nodes plays for the reference.
 */
#include "bvh8_build.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <unordered_map>

namespace b200 {

namespace {

struct Box {
  float lo[3], hi[3];
  void reset()
  {
    for (int k = 0; k < 3; k++) {
      lo[k] = INFINITY;
      hi[k] = -INFINITY;
    }
  }
  void grow(const float *p)
  {
    for (int k = 0; k < 3; k++) {
      lo[k] = std::min(lo[k], p[k]);
      hi[k] = std::max(hi[k], p[k]);
    }
  }
  void grow(const Box &b)
  {
    for (int k = 0; k < 3; k++) {
      lo[k] = std::min(lo[k], b.lo[k]);
      hi[k] = std::max(hi[k], b.hi[k]);
    }
  }
  float half_area() const
  {
    float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    if (!(dx >= 0.0f) || !(dy >= 0.0f) || !(dz >= 0.0f))
      return 0.0f;
    return dx * dy + dy * dz + dz * dx;
  }
};

/* Binary tree decoded from the BVH2 arrays, leaves <= 3 records. */
struct BNode {
  Box box;
  int left, right;     /* -1 for leaves */
  int first, count;    /* leaf: range in `refs` (triangles) */
  int object;          /* instance leaf: object id, else -1 */
  uint32_t visibility; /* instance leaf visibility */
};

struct Builder {
  const BVH2Input &in;
  BVH8Output &out;
  std::vector<BNode> bn;
  std::vector<int> refs; /* prim_addr per leaf record */
  std::unordered_map<int, uint32_t> blas_root8; /* BVH2 encoded root -> BVH8 node */
  std::vector<int> pending_blas;               /* BVH2 roots still to convert */
  std::string error;
  uint32_t max_depth = 0;
  double sah = 0.0;

  Builder(const BVH2Input &i, BVH8Output &o) : in(i), out(o)
  {
  }

  static int as_int(float f)
  {
    int i;
    memcpy(&i, &f, 4);
    return i;
  }
  static float as_float(uint32_t u)
  {
    float f;
    memcpy(&f, &u, 4);
    return f;
  }

  Box tri_box(int prim_addr) const
  {
    Box b;
    b.reset();
    uint32_t vi = in.prim_tri_index[prim_addr];
    for (int k = 0; k < 3; k++)
      b.grow(&in.prim_tri_verts[4 * (size_t)(vi + k)]);
    return b;
  }

  /* Bounds of a whole BVH2 tree (any encoded root). */
  Box bvh2_box(int addr) const
  {
    Box b;
    b.reset();
    if (addr >= 0) {
      const float *n = &in.nodes[4 * (size_t)addr];
      b.lo[0] = std::min(n[4], n[5]), b.hi[0] = std::max(n[6], n[7]);
      b.lo[1] = std::min(n[8], n[9]), b.hi[1] = std::max(n[10], n[11]);
      b.lo[2] = std::min(n[12], n[13]), b.hi[2] = std::max(n[14], n[15]);
    }
    else {
      const float *l = &in.leaf_nodes[4 * (size_t)(-addr - 1)];
      for (int p = as_int(l[0]); p < as_int(l[1]); p++)
        b.grow(tri_box(p));
    }
    return b;
  }

  /* World box of an instance that is the whole top level (no parent entry):
   * the BLAS bounds carried through KernelObject::tfm. */
  Box instance_world_box(int object) const
  {
    Box ob = bvh2_box(in.object_node[object]);
    const float *tfm = (const float *)(in.objects + (size_t)object * in.object_stride +
                                       in.object_tfm_offset);
    Box wb;
    wb.reset();
    for (int c = 0; c < 8; c++) {
      float p[3] = {(c & 1) ? ob.hi[0] : ob.lo[0], (c & 2) ? ob.hi[1] : ob.lo[1],
                    (c & 4) ? ob.hi[2] : ob.lo[2]};
      float q[3];
      for (int r = 0; r < 3; r++)
        q[r] = tfm[4 * r] * p[0] + tfm[4 * r + 1] * p[1] + tfm[4 * r + 2] * p[2] + tfm[4 * r + 3];
      wb.grow(q);
    }
    /* pad: the transform above is float arithmetic, keep it conservative */
    for (int k = 0; k < 3; k++) {
      float pad = 1e-5f * std::max(std::fabs(wb.lo[k]), std::fabs(wb.hi[k])) + 1e-30f;
      wb.lo[k] -= pad;
      wb.hi[k] += pad;
    }
    return wb;
  }

  /* ---- tighter world bounds for instances ----
   * The reference bounds an instance by transforming the mesh's AABB
   * (Object::compute_bounds, render/object.cpp): for a rotated object that is the AABB of
   * a rotated box, up to sqrt(3) wider per axis than the object.  The union of the
   * transformed boxes of the BLAS nodes a few levels down hugs the geometry instead (for
   * the rocks of config 4 the boxes at the mesh AABB's corners are simply not there), and
   * a ray that misses it never pays the instance push / pop - the least efficient part of
   * the traversal kernel.  Both boxes are conservative, so is their intersection. */
  std::unordered_map<int, std::vector<Box>> blas_detail_cache;

  const std::vector<Box> &blas_detail(int blas_addr)
  {
    auto it = blas_detail_cache.find(blas_addr);
    if (it != blas_detail_cache.end())
      return it->second;
    struct Open {
      int addr;
      Box box;
    };
    std::vector<Open> open;
    std::vector<Box> done;
    open.push_back({blas_addr, bvh2_box(blas_addr)});
    /* split the box with the largest surface until there are enough of them */
    /* default: as many boxes as a budget of 8 M box transforms over all instances allows,
     * between 16 and 1024 (measured on config 4, profiles/r02j_instance_bounds_sweep.txt:
     * 64 boxes recover most of it, 1024 the rest at 0.2 s of host time for 10 k instances) */
    size_t want = (size_t)in.instance_detail_boxes;
    if (want == 0) {
      want = ((size_t)8 << 20) / std::max<size_t>(in.num_objects, 1);
      want = std::min<size_t>(std::max<size_t>(want, 16), 1024);
    }
    while (!open.empty() && open.size() + done.size() < want) {
      size_t best = 0;
      for (size_t k = 1; k < open.size(); k++)
        if (open[k].box.half_area() > open[best].box.half_area())
          best = k;
      Open o = open[best];
      open.erase(open.begin() + best);
      if (o.addr < 0 || (size_t)o.addr + 4 > in.num_nodes_f4) {
        done.push_back(o.box); /* a leaf: as tight as it gets here */
        continue;
      }
      const float *n = &in.nodes[4 * (size_t)o.addr];
      Box c0, c1;
      c0.lo[0] = n[4], c1.lo[0] = n[5], c0.hi[0] = n[6], c1.hi[0] = n[7];
      c0.lo[1] = n[8], c1.lo[1] = n[9], c0.hi[1] = n[10], c1.hi[1] = n[11];
      c0.lo[2] = n[12], c1.lo[2] = n[13], c0.hi[2] = n[14], c1.hi[2] = n[15];
      open.push_back({as_int(n[2]), c0});
      open.push_back({as_int(n[3]), c1});
    }
    for (const Open &o : open)
      done.push_back(o.box);
    return blas_detail_cache.emplace(blas_addr, std::move(done)).first->second;
  }

  Box instance_tight_box(int object, const Box &host_box)
  {
    const float *tfm = (const float *)(in.objects + (size_t)object * in.object_stride +
                                       in.object_tfm_offset);
    Box wb;
    wb.reset();
    for (const Box &ob : blas_detail(in.object_node[object])) {
      if (!(ob.lo[0] <= ob.hi[0]))
        continue; /* empty */
      for (int c = 0; c < 8; c++) {
        float p[3] = {(c & 1) ? ob.hi[0] : ob.lo[0], (c & 2) ? ob.hi[1] : ob.lo[1],
                      (c & 4) ? ob.hi[2] : ob.lo[2]};
        float q[3];
        for (int r = 0; r < 3; r++)
          q[r] = tfm[4 * r] * p[0] + tfm[4 * r + 1] * p[1] + tfm[4 * r + 2] * p[2] +
                 tfm[4 * r + 3];
        wb.grow(q);
      }
    }
    /* the transform above is float arithmetic, and the kernel goes the other way through
     * the inverse matrix: keep a margin */
    for (int k = 0; k < 3; k++) {
      float pad = 1e-5f * std::max(std::fabs(wb.lo[k]), std::fabs(wb.hi[k])) +
                  1e-5f * (wb.hi[k] - wb.lo[k]) + 1e-30f;
      wb.lo[k] = std::max(wb.lo[k] - pad, host_box.lo[k]);
      wb.hi[k] = std::min(wb.hi[k] + pad, host_box.hi[k]);
    }
    if (!(wb.lo[0] <= wb.hi[0] && wb.lo[1] <= wb.hi[1] && wb.lo[2] <= wb.hi[2]))
      return host_box; /* not a number somewhere: keep what the host said */
    return wb;
  }

  /* instance leaves of bn[first ..) get their tight box, inner nodes the union of their
   * children again (children are decoded after their parent: one backward sweep) */
  void tighten_instances(int first)
  {
    bool any = false;
    for (int i = first; i < (int)bn.size(); i++)
      if (bn[i].object >= 0 && in.object_node && bn[i].box.lo[0] <= bn[i].box.hi[0]) {
        bn[i].box = instance_tight_box(bn[i].object, bn[i].box);
        any = true;
      }
    if (!any)
      return;
    for (int i = (int)bn.size() - 1; i >= first; i--)
      if (bn[i].left >= 0 && bn[i].right >= 0) {
        Box b = bn[bn[i].left].box;
        b.grow(bn[bn[i].right].box);
        bn[i].box = b;
      }
  }

  /* Split an oversized triangle leaf by the median of the centroids on the
   * longest axis until every leaf holds <= BVH8_MAX_LEAF_RECORDS. */
  int make_tri_leaf(std::vector<int> &prims)
  {
    Box b;
    b.reset();
    for (int p : prims)
      b.grow(tri_box(p));
    int idx = (int)bn.size();
    bn.push_back(BNode());
    bn[idx].box = b;
    bn[idx].object = -1;
    bn[idx].visibility = 0;
    if ((int)prims.size() <= BVH8_MAX_LEAF_RECORDS) {
      bn[idx].left = bn[idx].right = -1;
      bn[idx].first = (int)refs.size();
      bn[idx].count = (int)prims.size();
      refs.insert(refs.end(), prims.begin(), prims.end());
      return idx;
    }
    Box cb;
    cb.reset();
    std::vector<std::pair<float, int>> keyed(prims.size());
    std::vector<Box> boxes(prims.size());
    for (size_t i = 0; i < prims.size(); i++) {
      boxes[i] = tri_box(prims[i]);
      float c[3] = {0.5f * (boxes[i].lo[0] + boxes[i].hi[0]),
                    0.5f * (boxes[i].lo[1] + boxes[i].hi[1]),
                    0.5f * (boxes[i].lo[2] + boxes[i].hi[2])};
      cb.grow(c);
    }
    int axis = 0;
    float ext = cb.hi[0] - cb.lo[0];
    for (int k = 1; k < 3; k++)
      if (cb.hi[k] - cb.lo[k] > ext) {
        ext = cb.hi[k] - cb.lo[k];
        axis = k;
      }
    for (size_t i = 0; i < prims.size(); i++)
      keyed[i] = std::make_pair(0.5f * (boxes[i].lo[axis] + boxes[i].hi[axis]), prims[i]);
    std::stable_sort(keyed.begin(), keyed.end(),
                     [](const std::pair<float, int> &a, const std::pair<float, int> &b) {
                       return a.first < b.first;
                     });
    size_t half = keyed.size() / 2;
    std::vector<int> l, r;
    for (size_t i = 0; i < keyed.size(); i++)
      (i < half ? l : r).push_back(keyed[i].second);
    int li = make_tri_leaf(l);
    int ri = make_tri_leaf(r);
    bn[idx].left = li;
    bn[idx].right = ri;
    bn[idx].first = bn[idx].count = 0;
    return idx;
  }

  /* Decode the BVH2 subtree at encoded address `addr` (>= 0 inner node in float4
   * units, < 0 leaf -addr-1) into bn[]; iterative to survive deep trees. */
  int decode(int root_addr)
  {
    struct Item {
      int addr;
      int parent;
      int side;
      Box box; /* the box the BVH2 parent stores for this child */
    };
    std::vector<Item> stack;
    Box none;
    none.reset();
    stack.push_back({root_addr, -1, 0, none});
    int root_idx = -1;
    while (!stack.empty()) {
      Item it = stack.back();
      stack.pop_back();
      int idx;
      if (it.addr >= 0) {
        if ((size_t)it.addr + 4 > in.num_nodes_f4) {
          error = "BVH2 node address out of range";
          return -1;
        }
        const float *n = &in.nodes[4 * (size_t)it.addr];
        uint32_t vis0 = (uint32_t)as_int(n[0]);
        if (vis0 & in.node_unaligned_flag) {
          error = "unaligned (oriented) BVH2 nodes are outside the hot-path scope (hair only)";
          return -1;
        }
        idx = (int)bn.size();
        bn.push_back(BNode());
        BNode &b = bn[idx];
        b.left = b.right = -1;
        b.first = b.count = 0;
        b.object = -1;
        b.visibility = 0;
        /* bounds = union of the two child boxes (bvh2.cpp:98-113) */
        b.box.lo[0] = std::min(n[4], n[5]);
        b.box.hi[0] = std::max(n[6], n[7]);
        b.box.lo[1] = std::min(n[8], n[9]);
        b.box.hi[1] = std::max(n[10], n[11]);
        b.box.lo[2] = std::min(n[12], n[13]);
        b.box.hi[2] = std::max(n[14], n[15]);
        Box c0, c1;
        c0.lo[0] = n[4], c1.lo[0] = n[5], c0.hi[0] = n[6], c1.hi[0] = n[7];
        c0.lo[1] = n[8], c1.lo[1] = n[9], c0.hi[1] = n[10], c1.hi[1] = n[11];
        c0.lo[2] = n[12], c1.lo[2] = n[13], c0.hi[2] = n[14], c1.hi[2] = n[15];
        stack.push_back({as_int(n[3]), idx, 1, c1});
        stack.push_back({as_int(n[2]), idx, 0, c0});
      }
      else {
        size_t li = (size_t)(-it.addr - 1);
        if (li >= in.num_leaf_nodes_f4) {
          error = "BVH2 leaf address out of range";
          return -1;
        }
        const float *l = &in.leaf_nodes[4 * li];
        int lo = as_int(l[0]), hi = as_int(l[1]);
        if (lo < 0) {
          /* object (instance) leaf: bvh2.cpp:45-48, bvh_traversal.h:188-205 */
          int pa = -lo - 1;
          int object = (int)in.prim_object[pa];
          idx = (int)bn.size();
          bn.push_back(BNode());
          BNode &b = bn[idx];
          b.left = b.right = -1;
          b.first = b.count = 0;
          b.object = object;
          b.visibility = (uint32_t)as_int(l[2]);
          if (it.parent >= 0)
            b.box = it.box;
          else
            b.box = instance_world_box(object);
          int blas = in.object_node[object];
          if (blas_root8.find(blas) == blas_root8.end()) {
            blas_root8[blas] = 0xffffffffu;
            pending_blas.push_back(blas);
          }
        }
        else {
          uint32_t type = (uint32_t)as_int(l[3]);
          if (hi > lo && (type & in.primitive_all) != in.primitive_triangle) {
            error = "only static triangles are inside the hot-path scope (curve / motion leaf found)";
            return -1;
          }
          std::vector<int> prims;
          for (int p = lo; p < hi; p++)
            prims.push_back(p);
          if (prims.empty()) {
            idx = (int)bn.size();
            bn.push_back(BNode());
            BNode &b = bn[idx];
            b.left = b.right = -1;
            b.first = b.count = 0;
            b.object = -1;
            b.visibility = 0;
            b.box.reset();
          }
          else {
            idx = make_tri_leaf(prims);
          }
        }
      }
      if (it.parent < 0) {
        root_idx = idx;
      }
      else {
        if (it.side == 0)
          bn[it.parent].left = idx;
        else
          bn[it.parent].right = idx;
      }
    }
    return root_idx;
  }

  /* ---- SAH-optimal 2 -> 8 collapse (dynamic programme of Ylitie et al. 2017, sec. 3) ----
   * cost[n][i-1] = cheapest way to present binary subtree n to its BVH8 parent in at
   * most i child slots; a slot is a leaf (<= 3 triangles, possibly a merged small
   * subtree) or an inner BVH8 node (which itself distributes 8 slots).  */
  struct DP {
    float cost[7];
    uint8_t split[7]; /* i > 1: slots given to the left subtree, 0 = same as i-1 */
    uint8_t split8;   /* distribution of the 8 slots when n becomes an inner node */
    bool leaf1;       /* the single-slot form is a (merged) leaf */
    int prims;        /* triangles below, big when an instance is below */
  };
  std::vector<DP> dp;

  void run_dp(int root)
  {
    const float c_node = 1.0f, c_prim = 0.3f;
    /* post-order without recursion */
    std::vector<int> order, st;
    st.push_back(root);
    while (!st.empty()) {
      int n = st.back();
      st.pop_back();
      order.push_back(n);
      if (bn[n].left >= 0) {
        st.push_back(bn[n].left);
        st.push_back(bn[n].right);
      }
    }
    if (dp.size() < bn.size())
      dp.resize(bn.size());
    for (size_t oi = order.size(); oi-- > 0;) {
      const int n = order[oi];
      const BNode &b = bn[n];
      DP &d = dp[n];
      const float area = b.box.half_area();
      if (b.left < 0) {
        d.prims = (b.object >= 0) ? 1000000 : b.count;
        const float c = area * c_prim * (b.object >= 0 ? 4.0f : (float)std::max(b.count, 1));
        for (int i = 0; i < 7; i++) {
          d.cost[i] = c;
          d.split[i] = 0;
        }
        d.leaf1 = true;
        d.split8 = 0;
        continue;
      }
      const DP &l = dp[b.left], &r = dp[b.right];
      d.prims = std::min(1000000, l.prims + r.prims);
      /* as an inner BVH8 node: distribute 8 slots over the two subtrees */
      float best8 = INFINITY;
      int k8 = 1;
      for (int k = 1; k <= 7; k++) {
        const float c = l.cost[k - 1] + r.cost[8 - k - 1];
        if (c < best8) {
          best8 = c;
          k8 = k;
        }
      }
      d.split8 = (uint8_t)k8;
      const float c_internal = best8 + area * c_node;
      const float c_leaf = (d.prims <= BVH8_MAX_LEAF_RECORDS) ? area * c_prim * (float)d.prims :
                                                                INFINITY;
      d.leaf1 = c_leaf <= c_internal;
      d.cost[0] = std::min(c_leaf, c_internal);
      d.split[0] = 0;
      for (int i = 2; i <= 7; i++) {
        float best = d.cost[i - 2];
        int bk = 0;
        for (int k = 1; k < i; k++) {
          const float c = l.cost[k - 1] + r.cost[i - k - 1];
          if (c < best) {
            best = c;
            bk = k;
          }
        }
        d.cost[i - 1] = best;
        d.split[i - 1] = (uint8_t)bk;
      }
    }
  }

  struct Slot {
    int bnode;
    bool leaf; /* whole subtree becomes one leaf slot */
  };

  /* children of subtree n when it may use at most i slots */
  void gather(int n, int i, std::vector<Slot> &out)
  {
    struct Item {
      int n, i;
    };
    std::vector<Item> st;
    st.push_back({n, i});
    while (!st.empty()) {
      Item it = st.back();
      st.pop_back();
      const BNode &b = bn[it.n];
      const DP &d = dp[it.n];
      int ii = it.i;
      while (ii > 1 && d.split[ii - 1] == 0)
        ii--;
      if (b.left < 0 || ii == 1) {
        out.push_back({it.n, b.left < 0 ? true : d.leaf1});
        continue;
      }
      const int k = d.split[ii - 1];
      st.push_back({b.right, ii - k});
      st.push_back({b.left, k});
    }
  }

  /* triangles below a (merged) leaf slot */
  void collect_refs(int n, std::vector<int> &out_refs)
  {
    std::vector<int> st;
    st.push_back(n);
    while (!st.empty()) {
      int m = st.back();
      st.pop_back();
      const BNode &b = bn[m];
      if (b.left < 0) {
        for (int t = 0; t < b.count; t++)
          out_refs.push_back(refs[b.first + t]);
      }
      else {
        st.push_back(b.right);
        st.push_back(b.left);
      }
    }
  }

  struct Child {
    int bnode;
    Box box;
  };

  /* Emit the BVH8 node for binary subtree `b` at out.nodes[node_index]; children
   * are appended breadth-first by the caller's queue. */
  struct Work {
    int bnode;
    uint32_t node_index;
    uint32_t depth;
  };

  void collapse(int root_bnode, uint32_t root_index)
  {
    std::vector<Work> queue;
    queue.push_back({root_bnode, root_index, 1});
    size_t head = 0;
    while (head < queue.size()) {
      Work w = queue[head++];
      max_depth = std::max(max_depth, w.depth);
      const BNode &rootb = bn[w.bnode];

      /* children chosen by the dynamic programme */
      std::vector<Slot> slots;
      if (rootb.left < 0) {
        slots.push_back({w.bnode, true}); /* leaf-only root */
      }
      else {
        const int k8 = dp[w.bnode].split8;
        gather(rootb.left, k8, slots);
        gather(rootb.right, 8 - k8, slots);
      }
      /* drop empty leaves */
      slots.erase(std::remove_if(slots.begin(), slots.end(),
                                 [&](const Slot &sl) {
                                   const BNode &b = bn[sl.bnode];
                                   return b.left < 0 && b.object < 0 && b.count == 0;
                                 }),
                  slots.end());
      std::vector<int> ch;
      for (const Slot &sl : slots)
        ch.push_back(sl.bnode);

      Box nb;
      nb.reset();
      for (int c : ch)
        nb.grow(bn[c].box);
      if (ch.empty()) {
        nb.lo[0] = nb.lo[1] = nb.lo[2] = 0.0f;
        nb.hi[0] = nb.hi[1] = nb.hi[2] = 0.0f;
      }

      /* slot assignment: greedy minimum of cost[c][s] = dot(centroid_c - centroid, dir_s),
       * dir_s = (+-1,+-1,+-1) from the slot bits (x = bit 2, y = bit 1, z = bit 0), so
       * that slot ^ (7 - octant) orders children front to back for a ray. */
      int slot_of[8];
      {
        float cen[3] = {0.5f * (nb.lo[0] + nb.hi[0]), 0.5f * (nb.lo[1] + nb.hi[1]),
                        0.5f * (nb.lo[2] + nb.hi[2])};
        float cost[8][8];
        for (size_t i = 0; i < ch.size(); i++) {
          const Box &b = bn[ch[i]].box;
          float d[3] = {0.5f * (b.lo[0] + b.hi[0]) - cen[0], 0.5f * (b.lo[1] + b.hi[1]) - cen[1],
                        0.5f * (b.lo[2] + b.hi[2]) - cen[2]};
          for (int s = 0; s < 8; s++) {
            float sx = (s & 4) ? -1.0f : 1.0f, sy = (s & 2) ? -1.0f : 1.0f,
                  sz = (s & 1) ? -1.0f : 1.0f;
            cost[i][s] = d[0] * sx + d[1] * sy + d[2] * sz;
          }
        }
        bool cused[8] = {false}, sused[8] = {false};
        for (size_t n = 0; n < ch.size(); n++) {
          int bi = -1, bs = -1;
          float bc = INFINITY;
          for (size_t i = 0; i < ch.size(); i++) {
            if (cused[i])
              continue;
            for (int s = 0; s < 8; s++) {
              if (sused[s])
                continue;
              if (cost[i][s] < bc) {
                bc = cost[i][s];
                bi = (int)i;
                bs = s;
              }
            }
          }
          cused[bi] = true;
          sused[bs] = true;
          slot_of[bi] = bs;
        }
      }

      BVH8Node node;
      memset(&node, 0, sizeof(node));
      /* quantisation frame */
      double scale[3];
      for (int k = 0; k < 3; k++) {
        node.origin[k] = nb.lo[k];
        double ext = (double)nb.hi[k] - (double)nb.lo[k];
        int e = -126;
        if (ext > 0.0) {
          e = (int)std::ceil(std::log2(ext / 255.0));
          /* make sure the largest plane fits in 8 bits after rounding up */
          while (std::ceil(ext / std::ldexp(1.0, e)) > 255.0)
            e++;
          e = std::max(e, -126);
          e = std::min(e, 127);
        }
        node.e[k] = (uint8_t)(e + 127);
        scale[k] = std::ldexp(1.0, e);
      }

      /* inner children in ascending slot order */
      int order[8];
      for (size_t i = 0; i < ch.size(); i++)
        order[i] = (int)i;
      for (size_t a = 1; a < ch.size(); a++) { /* insertion sort, <= 8 entries */
        int v = order[a];
        size_t b = a;
        while (b > 0 && slot_of[order[b - 1]] > slot_of[v]) {
          order[b] = order[b - 1];
          b--;
        }
        order[b] = v;
      }

      uint32_t num_inner = 0;
      for (size_t i = 0; i < ch.size(); i++)
        if (!slots[i].leaf)
          num_inner++;
      node.child_base = (uint32_t)out.nodes.size();
      node.prim_base = (uint32_t)(out.records.size() / 12);
      out.nodes.resize(out.nodes.size() + num_inner);

      uint32_t inner_i = 0, rec_off = 0;
      double area = nb.half_area();
      for (size_t oi = 0; oi < ch.size(); oi++) {
        int i = order[oi];
        int s = slot_of[i];
        const BNode &c = bn[ch[i]];
        for (int k = 0; k < 3; k++) {
          double lo = std::floor(((double)c.box.lo[k] - (double)node.origin[k]) / scale[k]);
          double hi = std::ceil(((double)c.box.hi[k] - (double)node.origin[k]) / scale[k]);
          /* guard the float reconstruction origin + q*scale against rounding inwards */
          while (lo > 0.0 && (float)((double)node.origin[k] + lo * scale[k]) > c.box.lo[k])
            lo -= 1.0;
          while (hi < 255.0 && (float)((double)node.origin[k] + hi * scale[k]) < c.box.hi[k])
            hi += 1.0;
          lo = std::min(std::max(lo, 0.0), 255.0);
          hi = std::min(std::max(hi, 0.0), 255.0);
          node.qlo[k][s] = (uint8_t)lo;
          node.qhi[k][s] = (uint8_t)hi;
        }
        if (!slots[i].leaf) {
          node.imask |= (uint8_t)(1u << s);
          node.meta[s] = (uint8_t)((1u << 5) | (24 + s));
          queue.push_back({ch[i], node.child_base + inner_i, w.depth + 1});
          inner_i++;
          if (area > 0.0)
            sah += c.box.half_area() / area;
        }
        else if (c.object >= 0) {
          /* instance record */
          float rec[12];
          memset(rec, 0, sizeof(rec));
          rec[0] = as_float(0u); /* BVH8 root patched after all BLAS are built */
          rec[1] = as_float(c.visibility);
          rec[3] = as_float((uint32_t)~c.object);
          out.instance_patches.push_back(
              std::make_pair(out.records.size() + 0, in.object_node[c.object]));
          out.records.insert(out.records.end(), rec, rec + 12);
          node.meta[s] = (uint8_t)((1u << 5) | rec_off);
          rec_off += 1;
          out.num_instances++;
        }
        else {
          std::vector<int> leaf_refs;
          collect_refs(ch[i], leaf_refs);
          const int count = (int)leaf_refs.size();
          uint32_t unary = (count == 1) ? 1u : (count == 2) ? 3u : 7u;
          node.meta[s] = (uint8_t)((unary << 5) | rec_off);
          for (int t = 0; t < count; t++) {
            int pa = leaf_refs[t];
            uint32_t vi = in.prim_tri_index[pa];
            float rec[12];
            for (int k = 0; k < 3; k++) {
              const float *v = &in.prim_tri_verts[4 * (size_t)(vi + k)];
              rec[4 * k + 0] = v[0];
              rec[4 * k + 1] = v[1];
              rec[4 * k + 2] = v[2];
              rec[4 * k + 3] = 0.0f;
            }
            rec[3] = as_float((uint32_t)pa);
            rec[7] = as_float(in.prim_visibility[pa]);
            out.records.insert(out.records.end(), rec, rec + 12);
            out.num_triangles++;
          }
          rec_off += (uint32_t)count;
          if (area > 0.0)
            sah += (double)count * c.box.half_area() / area;
        }
      }
      out.nodes[w.node_index] = node;
    }
  }

  /* Convert one BVH2 tree (TLAS or one BLAS) into BVH8; returns its root node. */
  uint32_t convert(int root_addr)
  {
    const int first = (int)bn.size();
    int root = decode(root_addr);
    if (root < 0)
      return 0xffffffffu;
    if (in.tighten_instances)
      tighten_instances(first);
    run_dp(root);
    uint32_t root_index = (uint32_t)out.nodes.size();
    out.nodes.resize(out.nodes.size() + 1);
    collapse(root, root_index);
    return root_index;
  }

};

}  // namespace

bool build_bvh8(const BVH2Input &in, BVH8Output &out, std::string &error)
{
  auto t0 = std::chrono::steady_clock::now();
  out = BVH8Output();
  Builder b(in, out);
  uint32_t root = b.convert(in.root);
  if (root == 0xffffffffu) {
    error = b.error;
    return false;
  }
  out.root = root;
  /* bottom-level trees, one per distinct BVH2 root (shared by all its instances) */
  for (size_t i = 0; i < b.pending_blas.size(); i++) {
    int blas = b.pending_blas[i];
    uint32_t r = b.convert(blas);
    if (r == 0xffffffffu) {
      error = b.error;
      return false;
    }
    b.blas_root8[blas] = r;
  }
  for (auto &p : out.instance_patches) {
    uint32_t r = b.blas_root8[p.second];
    memcpy(&out.records[p.first], &r, 4);
  }
  out.object_root8.assign(in.num_objects, -1);
  if (in.object_node) {
    for (size_t o = 0; o < in.num_objects; o++) {
      auto it = b.blas_root8.find(in.object_node[o]);
      if (it != b.blas_root8.end())
        out.object_root8[o] = (int32_t)it->second;
    }
  }
  /* The traversal stack (BVH8_STACK_SIZE = 64 entries, traverse.cuh) holds at most two
   * entries per level of the TLAS and of one BLAS plus two for the instance push; a
   * tree deeper than this bound (a degenerate BVH2 chain) is refused, not truncated. */
  if (4u * b.max_depth + 2u > 64u) {
    error = "BVH8 depth " + std::to_string(b.max_depth) +
            " exceeds what the traversal stack covers (15)";
    return false;
  }
  out.max_depth = b.max_depth;
  out.sah_cost = (float)b.sah;
  out.build_ms =
      std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return true;
}

}  // namespace b200

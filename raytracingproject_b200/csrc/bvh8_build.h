/* bvh8_build.h - host-side BVH2 -> BVH8 conversion (see bvh8_build.cpp). */
#ifndef B200_BVH8_BUILD_H
#define B200_BVH8_BUILD_H

#include <stdint.h>
#include <stddef.h>

#include <string>
#include <utility>
#include <vector>

#include "bvh8.h"

namespace b200 {

/* Host copies of the reference's packed BVH arrays (kernel_textures.h names). */
struct BVH2Input {
  const float *nodes;          /* __bvh_nodes, float4 units */
  size_t num_nodes_f4;
  const float *leaf_nodes;     /* __bvh_leaf_nodes */
  size_t num_leaf_nodes_f4;
  const float *prim_tri_verts; /* __prim_tri_verts */
  const uint32_t *prim_tri_index;
  const uint32_t *prim_visibility;
  const uint32_t *prim_object;
  size_t num_prims;
  const int32_t *object_node;  /* __object_node */
  const uint8_t *objects;      /* __objects (KernelObject records) */
  size_t object_stride, object_tfm_offset;
  size_t num_objects;
  int32_t root;                /* KernelData.bvh.root */
  uint32_t node_unaligned_flag, primitive_all, primitive_triangle;
  /* bound instances by the transformed boxes of their BLAS's upper nodes instead of the
   * host's transformed mesh AABB (see Builder::instance_tight_box) */
  bool tighten_instances = true;
  int instance_detail_boxes = 0; /* BLAS boxes per instance bound, 0 = the default (64) */
};

struct BVH8Output {
  std::vector<BVH8Node> nodes;
  std::vector<float> records; /* 12 floats per leaf record */
  /* (index into records of the blas-root word, BVH2 root it refers to) */
  std::vector<std::pair<size_t, int32_t>> instance_patches;
  uint32_t root = 0;
  /* per object: BVH8 root of its BLAS, -1 for objects that are not instanced */
  std::vector<int32_t> object_root8;
  uint64_t num_triangles = 0, num_instances = 0;
  uint32_t max_depth = 0;
  float sah_cost = 0.0f;
  double build_ms = 0.0;
};

bool build_bvh8(const BVH2Input &in, BVH8Output &out, std::string &error);

}  // namespace b200

#endif

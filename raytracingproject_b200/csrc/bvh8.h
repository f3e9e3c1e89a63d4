/* bvh8.h - the compressed wide-BVH (BVH8) the B200 device traverses.
 *
 * NEW layout (not in the reference, which ships BVH2 / Embree / OptiX only -
 * intern/cycles/bvh/bvh2.cpp:86-116 is the 64 B BVH2 inner node it replaces).
 * An 8-wide node with child boxes quantised to 8 bits per plane relative to the
 * node origin, after Ylitie, Karras, Laine, "Efficient Incoherent Ray Traversal
 * on GPUs Through Compressed Wide BVHs" (HPG 2017).  80 bytes = five 128-bit
 * loads (LDG.E.128), 16 B aligned.
 *
 *   q0: origin.x, origin.y, origin.z, {ex, ey, ez, imask}
 *   q1: child_base, prim_base, meta[0..3], meta[4..7]
 *   q2: qlo_x[0..7], qlo_y[0..7]           (two uint32 each)
 *   q3: qlo_z[0..7], qhi_x[0..7]
 *   q4: qhi_y[0..7], qhi_z[0..7]
 *
 * child box plane = origin + q * 2^(e - 127)   (lo rounded down, hi rounded up:
 * conservative, so no hit the reference BVH2 finds can be culled).
 * meta[slot]: 0 = empty; inner node: 0b001xxxxx with xxxxx = 24 + slot;
 *             leaf: (unary record count 1->001, 2->011, 3->111) << 5 | offset,
 *             offset = first record relative to prim_base (0..23).
 * Inner children are stored contiguously from child_base in ascending slot
 * order; imask has a bit per inner slot.
 *
 * Leaf records are 48 B (three float4), contiguous per node from prim_base:
 *   triangle: a = (v0.xyz, as_float(prim_addr)) b = (v1.xyz, as_float(visibility))
 *             c = (v2.xyz, 0)        - vertices bit-identical to __prim_tri_verts,
 *             prim_addr = index into the reference's packed prim arrays
 *             (what Intersection::prim holds, kernel_types.h:672-686)
 *   instance: a = (as_float(blas_root_node), as_float(visibility), 0, as_float(~object))
 *             - a.w < 0 as an int tells it from a triangle (prim_addr >= 0); stands
 *             for the BVH2 object leaf (bvh2.cpp:45-48); the object's inverse
 *             transform is read from __objects[object].itfm.
 */
#ifndef B200_BVH8_H
#define B200_BVH8_H

#include <stdint.h>

#define BVH8_NODE_BYTES 80
#define BVH8_RECORD_BYTES 48
#define BVH8_MAX_LEAF_RECORDS 3

struct BVH8Node {
  float origin[3];
  uint8_t e[3];
  uint8_t imask;
  uint32_t child_base;
  uint32_t prim_base;
  uint8_t meta[8];
  uint8_t qlo[3][8];
  uint8_t qhi[3][8];
};
#ifdef __cplusplus
static_assert(sizeof(BVH8Node) == BVH8_NODE_BYTES, "BVH8 node must be 80 bytes");
#endif

#endif /* B200_BVH8_H */

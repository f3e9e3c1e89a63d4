/* b200_internal.h - host-side context behind the C ABI (include/b200_cycles.h). */
#ifndef B200_INTERNAL_H
#define B200_INTERNAL_H

#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200_cycles.h"
#include "../../include/cycles_abi.h"
#include "bvh8_build.h"

struct HostArray {
  uint64_t dptr = 0;
  size_t bytes = 0;
  std::vector<uint8_t> host; /* kept only for the arrays the BVH8 build / validation needs */
};

struct PathPool; /* wavefront state, defined in b200_cycles.cu */

/* what the bound SVM program reads beyond the geometry (checked against the scene in
 * check_scope) */
#define SVM_USES_ATTRIBUTES 1u
#define SVM_USES_WINDOW_COORDINATES 2u
#define SVM_USES_EXTENDED_NODES 4u /* anything svm_eval_extended_node dispatches */
#define SVM_USES_TANGENT 8u        /* NODE_GEOM_T: reads generated coordinates when present */
#define SVM_USES_IMAGES 32u        /* an image / environment texture node: needs bound textures */
#define SVM_USES_MULTISCATTER 16u  /* a Multiscatter GGX lobe: kernels with the random walks */

struct b200_ctx {
  int ordinal = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr;     /* stream all work is issued on */
  cudaStream_t own_stream = nullptr; /* the private stream created with the context */
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
  cudaEvent_t ev4 = nullptr, ev5 = nullptr, ev6 = nullptr;
  std::string error;
  std::mutex mutex;

  std::map<uint64_t, size_t> allocs; /* dptr -> bytes */
  size_t mem_used = 0;

  std::map<std::string, HostArray> globals; /* kernel_textures.h name -> binding */
  /* image slots: the reference's TextureInfo records (data = device pointer), uploaded as
   * one array when the scene is prepared */
  std::vector<uint8_t> texture_info;
  void *d_texture_info = nullptr;
  int svm_max_image_slot = -1; /* highest image slot the bound program names */
  std::vector<uint8_t> kernel_data;
  bool scene_dirty = true; /* BVH8 / constant block must be (re)built */
  bool bvh_dirty = true;   /* an array the BVH8 is derived from was (re)bound */
  uint32_t bvh_root8 = 0;  /* root of the BVH8 derived last (BVH2 hosts) */
  /* this context's constant block: the __constant__ DeviceScene is one per GPU, so a
   * context re-uploads its copy when another context of the same GPU used the device
   * last (DeviceUse in b200_cycles.cu) */
  std::vector<uint8_t> constant_block;
  bool have_data = false;
  uint32_t svm_features = 0;     /* SVM_USES_* of the bound __svm_nodes (svm_validate) */
  bool has_subd_patches = false; /* __tri_patch holds a patch index */
  bool force_svm_ext = false; /* a lean batch had to be redone with the full kernels */
  bool has_terminator_offset = false; /* an object with shadow_terminator_offset > 1 */
  bool has_generated_attr = false; /* __attributes_map lists ATTR_STD_GENERATED */

  /* BVH8 on the device */
  void *d_nodes = nullptr;
  void *d_records = nullptr;
  b200_bvh_info bvh_info = {};

  PathPool *pool = nullptr;
  size_t pool_bytes = 0; /* wavefront path pool (device-only memory) */
  int64_t opt_batch_paths = 0;
  int64_t opt_count_traversal = 0;
  int64_t opt_debug_slot = -1;
  float *d_debug = nullptr; /* 16 bounces x 32 floats when "debug_slot" >= 0 */
  int64_t opt_refill_threshold = 0;
  int64_t opt_trace_blocks_per_sm = 0;
  int64_t opt_instance_detail_boxes = 0; /* 0 = builder default */
  int64_t opt_loose_instances = 0; /* A/B: host's instance bounds, no tightening */
  int64_t opt_shade_carveout = 0;  /* A/B: shared-memory carveout of the shade kernels, % */
  int64_t opt_l2_persist_nodes = 0; /* A/B: L2 persisting window over the BVH8 nodes */
  bool l2_window_set = false;
  const void *dev_nodes = nullptr;  /* what the traversal kernels read (prepare_scene) */
  size_t dev_nodes_bytes = 0;
  int64_t opt_sort_tiles = 0;      /* experiment: sort by shader inside 2048-entry tiles */
  int64_t opt_sync_iterations = 0; /* A/B: stop the stream for the counters every bounce */

  /* k_shade_surface<lean / lean + multi-scatter / full / full + render passes>: grid size
   * per SM */
  int shade_blocks_per_sm[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  /* The lean multiscatter and the full shading kernels exist in two block shapes
   * (wavefront.cuh, WIDE: two blocks of 256 threads per SM, one of 512, one of 1024).  Which is
   * faster depends on the shader mix of the scene, so the first batches of a scene
   * alternate between them and the faster one is kept until the SVM programs change.  One
   * probe per kernel kind (0 = lean multiscatter, 1 = full).  choice: -1 probing, 0 / 1 / 2
   * decided. */
  int64_t opt_shade_wide = -1; /* -1 probe, 0 / 1 / 2 forced */
  struct ShadeProbe {
    int choice = -1;
    double ms[3] = {0.0, 0.0, 0.0};
    double paths[3] = {0.0, 0.0, 0.0};
    uint64_t batches = 0;
  } shade_probe[2];

  /* host cancel predicate (task.get_cancel()), polled between wavefront batches */
  b200_cancel_fn cancel_fn = nullptr;
  void *cancel_user = nullptr;
  /* scratch of b200_film_reduce, kept between calls */
  void *reduce_tmp = nullptr;
  size_t reduce_tmp_bytes = 0;

  /* trace_batch work counter + stats */
  unsigned int *d_counters = nullptr; /* small block of device counters */
  unsigned int *h_counters = nullptr; /* pinned mirror */
  b200_stats stats = {};
};

#endif

/* bvh8_host.cpp - the B200 device's BVH as a first-class host layout of the reference:
 * `class BVH8 : public BVH`, standing next to bvh/bvh2.cpp, selected by
 * `BVH_LAYOUT_BVH8` in BVH::create (bvh/bvh.cpp:99-124) when the device's
 * get_bvh_layout_mask() asks for it (BVHParams::best_bvh_layout, bvh/bvh_params.h).
 *
 * What the reference does for a layout (bvh/bvh.cpp:128-178): BVHBuild::run() makes the
 * binary BVHNode tree (SAH, spatial splits), `pack_primitives` fills the prim arrays,
 * `pack_nodes(root)` - the layout's own - flattens the tree into PackedBVH::nodes /
 * leaf_nodes, and GeometryManager::device_update_bvh (render/geometry.cpp:1011-1101)
 * hands those arrays to the device.  BVH8::pack_nodes flattens the binary tree on the
 * host first (the flattening of BVH2, which this class derives from: the 2 -> 8 collapse
 * is a dynamic programme over the WHOLE binary tree, boxes included, so it needs it laid
 * out anyway; per-mesh trees stay in that form, they are the top level's input), then
 * replaces the top level's node arrays by the device layout:
 *     pack.nodes       80-byte BVH8 nodes          (csrc/bvh8.h, five int4 each)
 *     pack.leaf_nodes  48-byte leaf records        (three int4 each)
 *     pack.object_node BVH8 root of each object's BLAS
 *     pack.root_index  BVH8 root
 * through b200_bvh8_pack (include/b200_cycles.h).  The prim arrays the shading code
 * reads (prim_index, prim_object, prim_tri_verts ...) are the reference's, unchanged.
 * No BVH2 array reaches the device, and the device builds nothing at bind time
 * (b200_bvh_info.host_packed = 1).
 *
 * Compiled into libcycles_device_b200.so against the reference headers where they lie;
 * registered with the host application through the hook INTEGRATION.md section 2
 * describes (in a patched tree: `case BVH_LAYOUT_BVH8: return new BVH8(...)`). */
#include "bvh/bvh.h"
#include "bvh/bvh2.h"
#include "bvh/bvh_node.h"
#include "bvh/bvh_params.h"
#include "render/object.h"
#include "util/util_string.h"
#include "util/util_thread.h"
#include "util/util_time.h"

#include <string.h>

#include "../../include/b200_cycles.h"

extern "C" void ref_host_register_bvh_layout(int layout, void *create);

CCL_NAMESPACE_BEGIN

static thread_mutex g_bvh8_mutex;
static string g_bvh8_error;
static b200_bvh_info g_bvh8_info;
static double g_bvh8_pack_seconds = 0.0;

class BVH8 : public BVH2 {
 public:
  BVH8(const BVHParams &params_,
       const vector<Geometry *> &geometry_,
       const vector<Object *> &objects_)
      : BVH2(params_, geometry_, objects_)
  {
  }

 protected:
  virtual void pack_nodes(const BVHNode *root) override
  {
    BVH2::pack_nodes(root);
    if (!params.top_level)
      return;
    const double t0 = time_dt();

    vector<float> tfm(12 * objects.size());
    for (size_t i = 0; i < objects.size(); i++)
      memcpy(&tfm[12 * i], &objects[i]->tfm, 12 * sizeof(float));

    b200_packed_bvh2 in;
    memset(&in, 0, sizeof(in));
    in.nodes = pack.nodes.data();
    in.num_nodes_f4 = pack.nodes.size();
    in.leaf_nodes = pack.leaf_nodes.data();
    in.num_leaf_nodes_f4 = pack.leaf_nodes.size();
    in.prim_tri_verts = pack.prim_tri_verts.data();
    in.prim_tri_index = pack.prim_tri_index.data();
    in.prim_visibility = pack.prim_visibility.data();
    in.prim_object = pack.prim_object.data();
    in.num_prims = pack.prim_tri_index.size();
    in.object_node = pack.object_node.size() ? pack.object_node.data() : NULL;
    in.object_tfm = tfm.size() ? &tfm[0] : NULL;
    in.num_objects = (objects.size() < pack.object_node.size()) ? objects.size() : pack.object_node.size();
    in.root = pack.root_index;

    b200_packed_bvh8 out;
    char err[512] = "";
    const int rc = b200_bvh8_pack(&in, &out, err, sizeof(err));
    thread_scoped_lock lock(g_bvh8_mutex);
    if (rc != B200_OK) {
      /* nothing the device could traverse: leave the arrays empty, the device refuses
       * the scene at bind time and reports this reason */
      g_bvh8_error = err;
      pack.nodes.clear();
      pack.leaf_nodes.clear();
      pack.root_index = -1;
      return;
    }
    g_bvh8_error = "";
    pack.nodes.resize(out.node_bytes / sizeof(int4));
    memcpy(pack.nodes.data(), out.nodes, out.node_bytes);
    pack.leaf_nodes.resize(out.record_bytes / sizeof(int4));
    memcpy(pack.leaf_nodes.data(), out.records, out.record_bytes);
    for (size_t i = 0; i < in.num_objects; i++)
      pack.object_node[i] = out.object_node[i];
    pack.root_index = (int)out.root;
    g_bvh8_info = out.info;
    g_bvh8_pack_seconds = time_dt() - t0;
    b200_bvh8_free(&out);
  }
};

static BVH *bvh8_create(const BVHParams &params,
                        const vector<Geometry *> &geometry,
                        const vector<Object *> &objects)
{
  return new BVH8(params, geometry, objects);
}

namespace {
struct BVH8Registration {
  BVH8Registration()
  {
    ref_host_register_bvh_layout((int)B200_BVH_LAYOUT_BVH8, (void *)&bvh8_create);
  }
} g_bvh8_registration;
}  // namespace

/* for the device shim: why the last top-level pack failed ("" = it did not), and the
 * report of the last successful one */
string bvh8_last_error()
{
  thread_scoped_lock lock(g_bvh8_mutex);
  return g_bvh8_error;
}
void bvh8_last_info(b200_bvh_info *info, double *pack_seconds)
{
  thread_scoped_lock lock(g_bvh8_mutex);
  *info = g_bvh8_info;
  *pack_seconds = g_bvh8_pack_seconds;
}

CCL_NAMESPACE_END

/* Host-side harness over the REFERENCE's own Scene / Device / DeviceTask API
 * (intern/cycles/render/scene.h, device/device.h, device/device_task.h),
 * exported as a small C API for ctypes.  It plays the part Session does in the
 * reference (render/session.cpp:381-486 acquire_tile, :1053-1121 render) without
 * the display / tile-manager machinery, so the very same Scene can be driven
 * either by the reference CPUDevice (the oracle and the CPU baseline) or by an
 * externally created Device* (the B200 device shim, see
 * raytracingproject_b200/csrc/device_b200.cpp).
 *
 * TEST INFRASTRUCTURE ONLY - part of oracle/_ref/libcycles_ref.so; compiled
 * from the reference sources where they lie (see oracle/Makefile). */

/* CPUDevice is defined in a .cpp, not a header; include it here (in place, not
 * copied) so the probe can reach CPUDevice::kernel_globals.  Its three factory
 * functions are renamed on the way in: the host library (libcycles_host.so,
 * ref_host_hooks.cpp) owns the names Device::create calls and forwards to these once
 * this library has registered them (OracleCpuDeviceRegistration below). */
#include "device/device_intern.h"
#define device_cpu_create oracle_device_cpu_create
#define device_cpu_info oracle_device_cpu_info
#define device_cpu_capabilities oracle_device_cpu_capabilities
#include "device/device_cpu.cpp"
#undef device_cpu_create
#undef device_cpu_info
#undef device_cpu_capabilities

#include "app/cycles_xml.h"
#include "bvh/bvh.h"
#include "bvh/bvh_params.h"
#include "render/background.h"
#include "render/buffers.h"
#include "render/camera.h"
#include "render/film.h"
#include "render/integrator.h"
#include "render/light.h"
#include "render/mesh.h"
#include "render/nodes.h"
#include "render/graph.h"
#include "render/image.h"
#include "render/object.h"
#include "render/scene.h"
#include "render/shader.h"
#include "util/util_progress.h"
#include "util/util_task.h"
#include "util/util_time.h"

#include <atomic>
#include <thread>
#include <vector>
#include <execinfo.h>
#include <signal.h>
#include <xmmintrin.h>
#include <unistd.h>
#include <cstdio>
#include <cstring>

#include "ref_probe.h"

using namespace ccl;

extern "C" void ref_host_register_cpu_device(void *create, void *info, void *capabilities);
namespace {
struct OracleCpuDeviceRegistration {
  OracleCpuDeviceRegistration()
  {
    ref_host_register_cpu_device((void *)&ccl::oracle_device_cpu_create,
                                 (void *)&ccl::oracle_device_cpu_info,
                                 (void *)&ccl::oracle_device_cpu_capabilities);
  }
} g_oracle_cpu_device_registration;
}

/* Only widens access to the per-thread KernelGlobals helpers; every virtual is
 * CPUDevice's own, so renders through it ARE the unmodified reference device. */
class ProbeCPUDevice : public CPUDevice {
 public:
  ProbeCPUDevice(DeviceInfo &info, Stats &stats, Profiler &profiler, bool background)
      : CPUDevice(info, stats, profiler, background)
  {
  }
  KernelGlobals kg_init()
  {
    return thread_kernel_globals_init();
  }
  void kg_free(KernelGlobals *kg)
  {
    thread_kernel_globals_free(kg);
  }
};

struct ref_scene {
  Stats stats;
  Profiler profiler;
  DeviceInfo info;
  Device *device;
  bool own_device;
  ProbeCPUDevice *cpu; /* NULL for an external device */
  Scene *scene;
  Progress progress;
  RenderBuffers *buffers;
  vector<Mesh *> meshes;
  std::string error;
  BVH *packed_bvh; /* ref_scene_pack_bvh */
};

static bool g_sched_init = false;

/* REF_BACKTRACE=1: print a symbolised backtrace on SIGSEGV (no gdb in the image). */
static void segv_handler(int sig)
{
  void *frames[64];
  int n = backtrace(frames, 64);
  backtrace_symbols_fd(frames, n, 2);
  signal(sig, SIG_DFL);
  raise(sig);
}

extern "C" {

int ref_version(void)
{
  return 2;
}

/* threads = 0 -> all cores (TaskScheduler::num_threads, util_task.cpp). */
void ref_init(int threads)
{
  if (getenv("REF_BACKTRACE"))
    signal(SIGSEGV, segv_handler);
  if (g_sched_init)
    TaskScheduler::exit();
  TaskScheduler::init(threads);
  g_sched_init = true;
}

int ref_num_threads(void)
{
  return TaskScheduler::num_threads();
}

/* kernel: 0 = generic scalar kernel (parity variant, -ffp-contract=off),
 *         1 = AVX2 kernel (speed variant).  Selected exactly as the reference
 *         does, through DebugFlags().cpu (device_cpu.cpp:89-131). */
ref_scene *ref_scene_new(const char *xml_path, int kernel, void *external_device)
{
  if (!g_sched_init)
    ref_init(0);

  ref_scene *rs = new ref_scene();
  rs->buffers = NULL;
  rs->cpu = NULL;
  rs->packed_bvh = NULL;

  if (external_device) {
    rs->device = (Device *)external_device;
    rs->own_device = false;
  }
  else {
    DebugFlags().cpu.reset();
    if (kernel == 0) {
      DebugFlags().cpu.avx2 = false;
      DebugFlags().cpu.avx = false;
      DebugFlags().cpu.sse41 = false;
      DebugFlags().cpu.sse3 = false;
      DebugFlags().cpu.sse2 = false;
    }
    DebugFlags().cpu.bvh_layout = BVH_LAYOUT_BVH2;
    DebugFlags().cpu.split_kernel = false;
    vector<DeviceInfo> infos;
    oracle_device_cpu_info(infos);
    rs->info = infos[0];
    rs->cpu = new ProbeCPUDevice(rs->info, rs->stats, rs->profiler, true);
    rs->device = rs->cpu;
    rs->own_device = true;
  }

  SceneParams params;
  params.shadingsystem = SHADINGSYSTEM_SVM;
  params.bvh_layout = BVH_LAYOUT_BVH2;
  params.bvh_type = SceneParams::BVH_STATIC;
  params.use_bvh_spatial_split = false;
  params.background = true;
  rs->scene = new Scene(params, rs->device);

  if (xml_path && xml_path[0]) {
    xml_read_file(rs->scene, xml_path);
    /* cycles_standalone.cpp:145 recomputes the view plane after reading the scene; for
     * the panoramic cameras that is what makes the frame cover (u, v) in [0,1]^2 (the
     * perspective / orthographic scenes of this repo keep the view plane the XML reader
     * left, which the committed golden vectors were made with) */
    if (rs->scene->camera->type == CAMERA_PANORAMA) {
      rs->scene->camera->compute_auto_viewplane();
      rs->scene->camera->need_update = true;
    }
  }
  return rs;
}

void ref_scene_free(ref_scene *rs)
{
  if (!rs)
    return;
  delete rs->buffers;
  delete rs->packed_bvh;
  delete rs->scene;
  if (rs->own_device)
    delete rs->device;
  delete rs;
}

const char *ref_scene_error(ref_scene *rs)
{
  return rs->error.c_str();
}

static Shader *find_shader(Scene *scene, const char *name)
{
  foreach (Shader *shader, scene->shaders) {
    if (shader->name == name)
      return shader;
  }
  return NULL;
}

/* Adds a triangle mesh (no object yet).  Returns the mesh handle or -1. */
int ref_scene_add_mesh(ref_scene *rs,
                       const float *P,
                       int num_verts,
                       const int *tris,
                       int num_tris,
                       const char *shader_name,
                       int smooth)
{
  Shader *shader = find_shader(rs->scene, shader_name);
  if (!shader) {
    rs->error = std::string("unknown shader ") + shader_name;
    return -1;
  }
  Mesh *mesh = new Mesh();
  rs->scene->geometry.push_back(mesh);
  mesh->used_shaders.push_back(shader);
  mesh->reserve_mesh(num_verts, num_tris);
  for (int i = 0; i < num_verts; i++)
    mesh->add_vertex(make_float3(P[3 * i], P[3 * i + 1], P[3 * i + 2]));
  for (int i = 0; i < num_tris; i++)
    mesh->add_triangle(tris[3 * i], tris[3 * i + 1], tris[3 * i + 2], 0, smooth != 0);
  /* Attributes the shader asks for, filled the way the standalone XML reader does
   * (app/cycles_xml.cpp:528-533: generated coordinates = vertex positions) plus a
   * planar per-corner UV map (the reader takes UVs from the file; here u, v = x, y of
   * the corner's vertex, so that ATTR_ELEMENT_CORNER float2 data is exercised). */
  if (mesh->need_attribute(rs->scene, ATTR_STD_GENERATED)) {
    Attribute *attr = mesh->attributes.add(ATTR_STD_GENERATED);
    memcpy(attr->data_float3(), mesh->verts.data(), sizeof(float3) * mesh->verts.size());
  }
  if (mesh->need_attribute(rs->scene, ATTR_STD_UV) ||
      mesh->need_attribute(rs->scene, ustring("UVMap"))) {
    Attribute *attr = mesh->attributes.add(ATTR_STD_UV, ustring("UVMap"));
    float2 *uv = attr->data_float2();
    for (int i = 0; i < num_tris; i++)
      for (int c = 0; c < 3; c++) {
        const int v = tris[3 * i + c];
        uv[3 * i + c] = make_float2(P[3 * v], P[3 * v + 1]);
      }
  }
  rs->meshes.push_back(mesh);
  return (int)rs->meshes.size() - 1;
}

/* tfm: 12 floats, row-major 3x4 (ccl::Transform x,y,z rows). */
int ref_scene_add_object(ref_scene *rs, int mesh, const float *tfm)
{
  if (mesh < 0 || mesh >= (int)rs->meshes.size())
    return -1;
  Object *object = new Object();
  object->geometry = rs->meshes[mesh];
  Transform t;
  t.x = make_float4(tfm[0], tfm[1], tfm[2], tfm[3]);
  t.y = make_float4(tfm[4], tfm[5], tfm[6], tfm[7]);
  t.z = make_float4(tfm[8], tfm[9], tfm[10], tfm[11]);
  object->tfm = t;
  rs->scene->objects.push_back(object);
  return (int)rs->scene->objects.size() - 1;
}

/* Object::shadow_terminator_offset (render/object.cpp:105, 544) of one object. */
int ref_scene_set_terminator_offset(ref_scene *rs, int object, float offset)
{
  if (object < 0 || object >= (int)rs->scene->objects.size())
    return -1;
  rs->scene->objects[object]->shadow_terminator_offset = offset;
  return 0;
}

/* UDIM tiles of one Image Texture node.  (The node's tile list is a plain member, not a
 * socket, so the XML reader cannot set it - Blender's exporter does, blender_shader.cpp.)
 * To be called before the first ref_scene_update. */
int ref_scene_set_image_tiles(
    ref_scene *rs, const char *shader_name, const char *node_name, const int *tiles, int num_tiles)
{
  Shader *shader = find_shader(rs->scene, shader_name);
  if (!shader || !shader->graph)
    return 1;
  foreach (ShaderNode *node, shader->graph->nodes) {
    if (node->type == ImageTextureNode::node_type && node->name == node_name) {
      ImageTextureNode *img = (ImageTextureNode *)node;
      img->tiles.clear();
      for (int i = 0; i < num_tiles; i++)
        img->tiles.push_back(tiles[i]);
      shader->tag_update(rs->scene);
      return 0;
    }
  }
  return 1;
}

/* Adds a render pass (PassType) next to the combined one - what BlenderSync::sync_render_passes
 * does from the view layer (blender_sync.cpp) - before the scene is updated.  Returns the
 * number of floats per pixel the film holds now. */
int ref_scene_add_pass(ref_scene *rs, int pass_type)
{
  Scene *scene = rs->scene;
  vector<Pass> passes = scene->passes;
  Pass::add((PassType)pass_type, passes);
  scene->film->tag_passes_update(scene, passes);
  scene->film->tag_update(scene);
  BufferParams bp;
  bp.passes = scene->passes;
  bp.denoising_data_pass = scene->film->denoising_data_pass;
  bp.denoising_clean_pass = scene->film->denoising_clean_pass;
  return bp.get_passes_size();
}

/* Denoising data passes behind the regular ones (film.cpp:604-620; what
 * BlenderSync::sync_view_layer sets from the view layer's denoising settings):
 * normal / albedo / depth with their variances, the two shadowing buffers, the colour
 * with variance, and - with `clean` - the components of `flags` (DenoiseFlag) kept out
 * of the noisy colour in a clean pass of their own.  Returns the floats per pixel. */
int ref_scene_set_denoising(ref_scene *rs, int data, int clean, int flags)
{
  Scene *scene = rs->scene;
  scene->film->denoising_data_pass = data != 0;
  scene->film->denoising_clean_pass = clean != 0;
  scene->film->denoising_flags = flags;
  scene->film->tag_update(scene);
  BufferParams bp;
  bp.passes = scene->passes;
  bp.denoising_data_pass = scene->film->denoising_data_pass;
  bp.denoising_clean_pass = scene->film->denoising_clean_pass;
  return bp.get_passes_size();
}

/* After ref_scene_update: float offsets of the denoising data and clean passes inside a
 * pixel (KernelFilm::pass_denoising_data / _clean, 0 = absent). */
int ref_scene_denoising_offset(ref_scene *rs, int *clean)
{
  if (clean)
    *clean = rs->scene->dscene.data.film.pass_denoising_clean;
  return rs->scene->dscene.data.film.pass_denoising_data;
}

/* Float offset of a pass inside a pixel of the film (RenderBuffers layout: the passes in
 * scene->passes order), -1 when the film does not hold it; `components` as stored. */
int ref_scene_pass_offset(ref_scene *rs, int pass_type, int *components)
{
  int offset = 0;
  foreach (const Pass &pass, rs->scene->passes) {
    if ((int)pass.type == pass_type) {
      if (components)
        *components = pass.components;
      return offset;
    }
    offset += pass.components;
  }
  return -1;
}

/* Equivalent of Session::update_scene (session.cpp:909-950). */
int ref_scene_update(ref_scene *rs, int width, int height)
{
  Scene *scene = rs->scene;
  Camera *cam = scene->camera;
  if (width > 0 && height > 0 && (width != cam->width || height != cam->height)) {
    cam->width = width;
    cam->height = height;
    cam->full_width = width;
    cam->full_height = height;
    cam->compute_auto_viewplane();
    cam->tag_update();
  }
  bool kernel_switch_needed = false;
  scene->update(rs->progress, kernel_switch_needed);
  /* CPUDevice publishes its image table to the kernel globals at the next task_add
   * (device_cpu.cpp:360-366); the probes call kernel functions without a task */
  if (rs->cpu)
    rs->cpu->load_texture_info();
  /* session.cpp:282,702 - size the profiler's per-shader / per-object counters. */
  rs->profiler.reset(scene->shaders.size(), scene->objects.size());
  if (rs->device->have_error()) {
    rs->error = rs->device->error_message();
    return 1;
  }
  if (rs->progress.get_error()) {
    rs->error = rs->progress.get_error_message();
    return 1;
  }
  return 0;
}

int ref_scene_width(ref_scene *rs)
{
  return rs->scene->camera->width;
}
int ref_scene_height(ref_scene *rs)
{
  return rs->scene->camera->height;
}
int ref_scene_pass_stride(ref_scene *rs)
{
  return rs->scene->dscene.data.film.pass_stride;
}

/* The flat device arrays exactly as handed to Device::mem_copy_to, looked up by
 * their kernel_textures.h name among the Scene's DeviceScene vectors
 * (render/scene.h:65-132).  Works for any device (host copies). */
int ref_scene_global(
    ref_scene *rs, const char *name, const void **ptr, uint64_t *count, uint32_t *elem_size)
{
  DeviceScene &d = rs->scene->dscene;
  device_memory *all[] = {
      &d.bvh_nodes, &d.bvh_leaf_nodes, &d.object_node, &d.prim_tri_index, &d.prim_tri_verts,
      &d.prim_type, &d.prim_visibility, &d.prim_index, &d.prim_object, &d.prim_time,
      &d.tri_shader, &d.tri_vnormal, &d.tri_vindex, &d.tri_patch, &d.tri_patch_uv, &d.curves,
      &d.curve_keys, &d.patches, &d.objects, &d.object_motion_pass, &d.object_motion,
      &d.object_flag, &d.object_volume_step, &d.camera_motion, &d.attributes_map,
      &d.attributes_float, &d.attributes_float2, &d.attributes_float3, &d.attributes_uchar4,
      &d.light_distribution, &d.lights, &d.light_background_marginal_cdf,
      &d.light_background_conditional_cdf, &d.particles, &d.svm_nodes, &d.shaders,
      &d.lookup_table, &d.sample_pattern_lut, &d.ies_lights};
  for (size_t i = 0; i < sizeof(all) / sizeof(all[0]); i++) {
    device_memory *m = all[i];
    if (strcmp(m->name, name) == 0) {
      *ptr = m->host_pointer;
      *elem_size = (uint32_t)(m->data_elements * datatype_size(m->data_type));
      *count = (*elem_size) ? m->memory_size() / (*elem_size) : 0;
      if (!m->host_pointer)
        *count = 0;
      return 0;
    }
  }
  return 1;
}

/* Name of the i-th kernel_textures.h entry, NULL past the end. */
const char *ref_global_name(int index)
{
  static const char *names[] = {
#define KERNEL_TEX(type, tname) #tname,
#include "kernel/kernel_textures.h"
      NULL};
  int n = (int)(sizeof(names) / sizeof(names[0])) - 1;
  return (index >= 0 && index < n) ? names[index] : NULL;
}

/* Image textures as the ImageManager handed them to the device (CPUDevice::tex_alloc,
 * device_cpu.cpp:483-504): the TextureInfo record of a slot and the host pixels it points
 * at.  Slots that hold no image report 0 bytes.  Only for the oracle's own CPU device. */
static bool scene_uses_image_slot(Scene *scene, int slot)
{
  /* CPUDevice grows its table 128 slots at a time without clearing them, so the slots in
   * use are taken from the shader graphs' image nodes */
  foreach (Shader *shader, scene->shaders) {
    if (!shader->graph)
      continue;
    foreach (ShaderNode *node, shader->graph->nodes) {
      if (node->special_type != SHADER_SPECIAL_TYPE_IMAGE_SLOT)
        continue;
      ImageHandle &handle = ((ImageSlotTextureNode *)node)->handle;
      for (int t = 0; t < handle.num_tiles(); t++)
        if (handle.svm_slot(t) == slot)
          return true;
    }
  }
  return false;
}
int ref_scene_num_textures(ref_scene *rs)
{
  return rs->cpu ? (int)rs->cpu->texture_info.size() : 0;
}
int ref_scene_texture(
    ref_scene *rs, int slot, void *info_out, uint64_t info_bytes, const void **pixels, uint64_t *bytes)
{
  if (!rs->cpu || slot < 0 || slot >= (int)rs->cpu->texture_info.size() ||
      info_bytes != sizeof(TextureInfo))
    return 1;
  *pixels = NULL;
  *bytes = 0;
  memset(info_out, 0, sizeof(TextureInfo));
  if (!scene_uses_image_slot(rs->scene, slot))
    return 0;
  const TextureInfo &info = rs->cpu->texture_info[slot];
  memcpy(info_out, &info, sizeof(TextureInfo));
  static const uint64_t texel_bytes[IMAGE_DATA_NUM_TYPES] = {16, 4, 8, 4, 1, 2, 8, 2};
  *pixels = (const void *)info.data;
  *bytes = info.data ? (uint64_t)info.width * info.height * (info.depth > 1 ? info.depth : 1u) *
                           texel_bytes[info.data_type % IMAGE_DATA_NUM_TYPES] :
                       0;
  return 0;
}

/* The top-level BVH of the updated scene packed once more in `layout` through the
 * reference's own BVH::create + BVH::build (the few lines of
 * GeometryManager::device_update_bvh, render/geometry.cpp:1019-1034) - without a device,
 * so that a host layout class a device plug-in registered (BVH8, see
 * raytracingproject_b200/csrc/bvh8_host.cpp) can be checked on a machine without a GPU.
 * The arrays stay valid until the next call or ref_scene_free. */
int ref_scene_pack_bvh(ref_scene *rs,
                       int layout,
                       const void **nodes,
                       uint64_t *node_bytes,
                       const void **leaf_nodes,
                       uint64_t *leaf_bytes,
                       const int **object_node,
                       uint64_t *num_objects,
                       int *root)
{
  Scene *scene = rs->scene;
  BVHParams bparams;
  bparams.top_level = true;
  bparams.bvh_layout = (BVHLayout)layout;
  bparams.use_spatial_split = scene->params.use_bvh_spatial_split;
  bparams.use_unaligned_nodes = false;
  bparams.num_motion_triangle_steps = scene->params.num_bvh_time_steps;
  bparams.num_motion_curve_steps = scene->params.num_bvh_time_steps;
  bparams.bvh_type = scene->params.bvh_type;
  bparams.curve_subdivisions = scene->params.curve_subdivisions();
  delete rs->packed_bvh;
  rs->packed_bvh = BVH::create(bparams, scene->geometry, scene->objects);
  if (!rs->packed_bvh) {
    rs->error = "BVH::create does not know this layout";
    return 1;
  }
  rs->packed_bvh->build(rs->progress, &rs->stats);
  PackedBVH &pack = rs->packed_bvh->pack;
  *nodes = pack.nodes.data();
  *node_bytes = pack.nodes.size() * sizeof(int4);
  *leaf_nodes = pack.leaf_nodes.data();
  *leaf_bytes = pack.leaf_nodes.size() * sizeof(int4);
  *object_node = pack.object_node.data();
  *num_objects = pack.object_node.size();
  *root = pack.root_index;
  return 0;
}

int ref_scene_data(ref_scene *rs, const void **ptr, uint64_t *size)
{
  *ptr = &rs->scene->dscene.data;
  *size = sizeof(KernelData);
  return 0;
}

/* Render samples [start_sample, start_sample+num_samples) of the full frame
 * through Device::task_add(RENDER) with tile_size x tile_size tiles handed out
 * from one permanent full-frame RenderBuffers (the `buffers != NULL` branch of
 * Session::acquire_tile, session.cpp:427-446).  `out` receives
 * height*width*pass_stride floats (unnormalised sums, as in the film buffer).
 * accumulate != 0 keeps the current film contents. */
int ref_render(ref_scene *rs,
               int start_sample,
               int num_samples,
               int tile_size,
               int accumulate,
               float *out,
               double *seconds)
{
  Scene *scene = rs->scene;
  Device *device = rs->device;
  const int width = scene->camera->width;
  const int height = scene->camera->height;

  BufferParams bp;
  bp.width = width;
  bp.height = height;
  bp.full_width = width;
  bp.full_height = height;
  bp.passes = scene->passes;
  bp.denoising_data_pass = scene->film->denoising_data_pass;
  bp.denoising_clean_pass = scene->film->denoising_clean_pass;

  if (!rs->buffers || rs->buffers->params.modified(bp)) {
    delete rs->buffers;
    rs->buffers = new RenderBuffers(device);
    rs->buffers->reset(bp);
  }
  else if (!accumulate) {
    rs->buffers->zero();
  }
  RenderBuffers *buffers = rs->buffers;

  if (tile_size <= 0)
    tile_size = max(width, height);
  const int tiles_x = (width + tile_size - 1) / tile_size;
  const int tiles_y = (height + tile_size - 1) / tile_size;
  const int num_tiles = tiles_x * tiles_y;
  std::atomic<int> next_tile(0);

  DeviceTask task(DeviceTask::RENDER);
  task.acquire_tile = [&](Device *, RenderTile &rtile, uint) -> bool {
    int t = next_tile.fetch_add(1);
    if (t >= num_tiles)
      return false;
    int tx = t % tiles_x, ty = t / tiles_x;
    rtile.x = tx * tile_size;
    rtile.y = ty * tile_size;
    rtile.w = min(tile_size, width - rtile.x);
    rtile.h = min(tile_size, height - rtile.y);
    rtile.start_sample = start_sample;
    rtile.num_samples = num_samples;
    rtile.sample = start_sample;
    rtile.resolution = 1;
    rtile.tile_index = t;
    rtile.task = RenderTile::PATH_TRACE;
    buffers->params.get_offset_stride(rtile.offset, rtile.stride);
    rtile.buffer = buffers->buffer.device_pointer;
    rtile.buffers = buffers;
    return true;
  };
  task.release_tile = [](RenderTile &) {};
  task.get_cancel = []() -> bool { return false; };
  task.update_tile_sample = [](RenderTile &) {};
  task.update_progress_sample = [](long, int) {};
  task.need_finish_queue = false;
  task.integrator_branched = false;
  /* Session::render (render/session.cpp:1077-1080) */
  task.adaptive_sampling.use = (scene->integrator->sampling_pattern == SAMPLING_PATTERN_PMJ) &&
                               scene->dscene.data.film.pass_adaptive_aux_buffer;
  task.adaptive_sampling.min_samples = scene->dscene.data.integrator.adaptive_min_samples;
  task.adaptive_sampling.adaptive_step = scene->dscene.data.integrator.adaptive_step;
  task.tile_types = RenderTile::PATH_TRACE;

  /* CPUDevice::render sets FTZ/DAZ (SIMD_SET_FLUSH_TO_ZERO) on whichever thread
   * runs a tile, including this one; do not leak it into the caller. */
  const unsigned int mxcsr = _mm_getcsr();
  double t0 = time_dt();
  device->task_add(task);
  device->task_wait();
  buffers->copy_from_device();
  double t1 = time_dt();
  _mm_setcsr(mxcsr);
  if (seconds)
    *seconds = t1 - t0;

  if (device->have_error()) {
    rs->error = device->error_message();
    return 1;
  }
  if (out) {
    memcpy(out,
           buffers->buffer.data(),
           sizeof(float) * (size_t)width * height * bp.get_passes_size());
  }
  return 0;
}

/* The background, non-progressive shape of Session::acquire_tile / release_tile
 * (render/session.cpp:449-460, 505-521): every tile gets its OWN RenderBuffers,
 * allocated when the device's worker acquires the tile, and release_tile - called on
 * that same worker thread - copies the pixels out (the write_render_tile_cb step) and
 * DELETES the tile's RenderBuffers, i.e. Device::mem_free runs inside the running task.
 * `cancel_after` >= 0 makes task.get_cancel() answer true once that many tiles were
 * released (Session::cancel / progress.set_cancel).  `out` receives the full frame;
 * tiles that were never rendered stay zero.  *tiles_done returns how many tiles were
 * released with all their samples. */
int ref_render_tile_buffers(ref_scene *rs,
                            int start_sample,
                            int num_samples,
                            int tile_size,
                            int cancel_after,
                            float *out,
                            int *tiles_done)
{
  Scene *scene = rs->scene;
  Device *device = rs->device;
  const int width = scene->camera->width;
  const int height = scene->camera->height;

  BufferParams full;
  full.width = width;
  full.height = height;
  full.full_width = width;
  full.full_height = height;
  full.passes = scene->passes;
  full.denoising_data_pass = scene->film->denoising_data_pass;
  full.denoising_clean_pass = scene->film->denoising_clean_pass;
  const int pass_stride = full.get_passes_size();
  memset(out, 0, sizeof(float) * (size_t)width * height * pass_stride);

  if (tile_size <= 0)
    tile_size = max(width, height);
  const int tiles_x = (width + tile_size - 1) / tile_size;
  const int tiles_y = (height + tile_size - 1) / tile_size;
  const int num_tiles = tiles_x * tiles_y;
  std::atomic<int> next_tile(0), released(0), complete(0);
  thread_mutex out_mutex;

  DeviceTask task(DeviceTask::RENDER);
  task.acquire_tile = [&](Device *tile_device, RenderTile &rtile, uint) -> bool {
    int t = next_tile.fetch_add(1);
    if (t >= num_tiles)
      return false;
    int tx = t % tiles_x, ty = t / tiles_x;
    rtile.x = tx * tile_size;
    rtile.y = ty * tile_size;
    rtile.w = min(tile_size, width - rtile.x);
    rtile.h = min(tile_size, height - rtile.y);
    rtile.start_sample = start_sample;
    rtile.num_samples = num_samples;
    rtile.sample = start_sample;
    rtile.resolution = 1;
    rtile.tile_index = t;
    rtile.task = RenderTile::PATH_TRACE;
    BufferParams bp = full;
    bp.full_x = rtile.x;
    bp.full_y = rtile.y;
    bp.width = rtile.w;
    bp.height = rtile.h;
    RenderBuffers *tb = new RenderBuffers(tile_device);
    tb->reset(bp);
    tb->params.get_offset_stride(rtile.offset, rtile.stride);
    rtile.buffer = tb->buffer.device_pointer;
    rtile.buffers = tb;
    return true;
  };
  task.release_tile = [&](RenderTile &rtile) {
    RenderBuffers *tb = rtile.buffers;
    if (rtile.sample == rtile.start_sample + rtile.num_samples) {
      tb->copy_from_device();
      thread_scoped_lock lock(out_mutex);
      const float *src = tb->buffer.data();
      for (int y = 0; y < rtile.h; y++)
        memcpy(out + ((size_t)(rtile.y + y) * width + rtile.x) * pass_stride,
               src + (size_t)y * rtile.w * pass_stride, sizeof(float) * rtile.w * pass_stride);
      complete++;
    }
    delete tb; /* on the device's worker thread, as Session::release_tile does */
    released++;
  };
  task.get_cancel = [&]() -> bool { return cancel_after >= 0 && released >= cancel_after; };
  task.update_tile_sample = [](RenderTile &) {};
  task.update_progress_sample = [](long, int) {};
  task.need_finish_queue = false;
  task.integrator_branched = false;
  /* Session::render (render/session.cpp:1077-1080) */
  task.adaptive_sampling.use = (scene->integrator->sampling_pattern == SAMPLING_PATTERN_PMJ) &&
                               scene->dscene.data.film.pass_adaptive_aux_buffer;
  task.adaptive_sampling.min_samples = scene->dscene.data.integrator.adaptive_min_samples;
  task.adaptive_sampling.adaptive_step = scene->dscene.data.integrator.adaptive_step;
  task.tile_types = RenderTile::PATH_TRACE;

  const unsigned int mxcsr = _mm_getcsr();
  device->task_add(task);
  device->task_wait();
  _mm_setcsr(mxcsr);
  if (tiles_done)
    *tiles_done = complete;
  if (device->have_error()) {
    rs->error = device->error_message();
    return 1;
  }
  return 0;
}

/* DeviceTask::FILM_CONVERT over the film of the last ref_render (what
 * DisplayBuffer::draw_set / Session::tonemap issue, render/buffers.cpp, session.cpp):
 * `out` receives w*h uchar4 (half_float = 0) or w*h half4 (half_float = 1).  Works on
 * whichever Device the scene was created on - the reference CPUDevice or an external
 * one - so the same call produces the oracle bytes and drives the device under test. */
int ref_film_convert(ref_scene *rs, int num_samples, int half_float, void *out)
{
  if (!rs->buffers) {
    rs->error = "film_convert: render first";
    return 1;
  }
  Device *device = rs->device;
  RenderBuffers *buffers = rs->buffers;
  const int width = buffers->params.width, height = buffers->params.height;
  buffers->buffer.copy_to_device();

  device_vector<uchar4> rgba_byte(device, "display_rgba_byte", MEM_READ_WRITE);
  device_vector<half4> rgba_half(device, "display_rgba_half", MEM_READ_WRITE);
  DeviceTask task(DeviceTask::FILM_CONVERT);
  task.x = 0;
  task.y = 0;
  task.w = width;
  task.h = height;
  task.sample = num_samples - 1; /* sample_scale = 1 / (task.sample + 1) */
  buffers->params.get_offset_stride(task.offset, task.stride);
  task.buffer = buffers->buffer.device_pointer;
  if (half_float) {
    rgba_half.alloc(width, height);
    rgba_half.zero_to_device();
    task.rgba_half = rgba_half.device_pointer;
  }
  else {
    rgba_byte.alloc(width, height);
    rgba_byte.zero_to_device();
    task.rgba_byte = rgba_byte.device_pointer;
  }
  const unsigned int mxcsr = _mm_getcsr();
  device->task_add(task);
  device->task_wait();
  _mm_setcsr(mxcsr);
  if (device->have_error()) {
    rs->error = device->error_message();
    return 1;
  }
  if (half_float) {
    rgba_half.copy_from_device(0, width, height);
    memcpy(out, rgba_half.data(), sizeof(half4) * (size_t)width * height);
    rgba_half.free();
  }
  else {
    rgba_byte.copy_from_device(0, width, height);
    memcpy(out, rgba_byte.data(), sizeof(uchar4) * (size_t)width * height);
    rgba_byte.free();
  }
  return 0;
}

/* ---- kernel probes (CPU device only) ---- */

/* The reference renders with FTZ + DAZ on every worker thread
 * (SIMD_SET_FLUSH_TO_ZERO, device_cpu.cpp:905); the probes run on the caller's
 * thread, so they adopt the same mode for the duration of the call. */
struct ScopedFlushToZero {
  unsigned int saved;
  ScopedFlushToZero() : saved(_mm_getcsr())
  {
    _mm_setcsr(saved | 0x8040);
  }
  ~ScopedFlushToZero()
  {
    _mm_setcsr(saved);
  }
};

int ref_intersect(ref_scene *rs, const RefProbeRay *rays, RefProbeHit *hits, uint64_t n)
{
  if (!rs->cpu)
    return 1;
  KernelGlobals kg = rs->cpu->kg_init();
  {
    ScopedFlushToZero ftz;
    ref_probe_intersect(&kg, rays, hits, n);
  }
  rs->cpu->kg_free(&kg);
  return 0;
}

int ref_camera_rays(
    ref_scene *rs, int sample, int x0, int y0, int w, int h, RefProbeRay *rays, uint32_t *rng_hash)
{
  if (!rs->cpu)
    return 1;
  KernelGlobals kg = rs->cpu->kg_init();
  {
    ScopedFlushToZero ftz;
    ref_probe_camera_rays(&kg, sample, x0, y0, w, h, rays, rng_hash);
  }
  rs->cpu->kg_free(&kg);
  return 0;
}

/* Census of the rays the reference traces for samples [start, start+num) of the
 * full frame: counts[0] camera, counts[1] bounce, counts[2] shadow.  Row bands
 * on std::threads (the census is per pixel, order-free). */
int ref_count_rays(ref_scene *rs, int start_sample, int num_samples, unsigned long long *counts)
{
  if (!rs->cpu)
    return 1;
  const int width = rs->scene->camera->width, height = rs->scene->camera->height;
  const int nthreads = std::max(1, TaskScheduler::num_threads());
  std::vector<std::thread> threads;
  std::vector<unsigned long long> partial(3 * (size_t)nthreads, 0ull);
  std::atomic<int> next_row(0);
  for (int t = 0; t < nthreads; t++) {
    threads.emplace_back([&, t]() {
      KernelGlobals kg = rs->cpu->kg_init();
      const unsigned int mxcsr = _mm_getcsr();
      _mm_setcsr(mxcsr | 0x8040); /* FTZ + DAZ as CPUDevice::render */
      for (;;) {
        int y = next_row.fetch_add(4);
        if (y >= height)
          break;
        int rows = std::min(4, height - y);
        for (int s = start_sample; s < start_sample + num_samples; s++)
          ref_probe_count_rays(&kg, s, 0, y, width, rows, &partial[3 * (size_t)t]);
      }
      _mm_setcsr(mxcsr);
      rs->cpu->kg_free(&kg);
    });
  }
  for (auto &th : threads)
    th.join();
  counts[0] = counts[1] = counts[2] = 0;
  for (int t = 0; t < nthreads; t++)
    for (int k = 0; k < 3; k++)
      counts[k] += partial[3 * (size_t)t + k];
  return 0;
}

int ref_path_dump(ref_scene *rs, int sample, int x, int y, float *out)
{
  if (!rs->cpu)
    return 1;
  KernelGlobals kg = rs->cpu->kg_init();
  {
    ScopedFlushToZero ftz;
    ref_probe_path_dump(&kg, sample, x, y, out);
  }
  rs->cpu->kg_free(&kg);
  return 0;
}

/* One SVM node of `nodes` (uint4 words) on one shading point; *next = offset after it. */
int ref_svm_node(ref_scene *rs, const void *nodes, int offset, float *stack,
                 const RefShadingPoint *p, int *next)
{
  if (!rs->cpu)
    return 1;
  KernelGlobals kg = rs->cpu->kg_init();
  {
    ScopedFlushToZero ftz;
    *next = ref_probe_svm_node(&kg, nodes, offset, stack, p);
  }
  rs->cpu->kg_free(&kg);
  return 0;
}

int ref_svm_closure(ref_scene *rs, const void *nodes, int offset, float *stack,
                    const RefShadingPoint *p, const float *closure_weight,
                    unsigned int path_flag, const float *omega_in, float randu, float randv,
                    float *out, int *next)
{
  if (!rs->cpu)
    return 1;
  KernelGlobals kg = rs->cpu->kg_init();
  {
    ScopedFlushToZero ftz;
    *next = ref_probe_svm_closure(&kg, nodes, offset, stack, p, closure_weight, path_flag,
                                  omega_in, randu, randv, out);
  }
  rs->cpu->kg_free(&kg);
  return 0;
}

int ref_shadow_rays(ref_scene *rs, int sample, int x0, int y0, int w, int h, RefProbeRay *rays)
{
  if (!rs->cpu)
    return 1;
  KernelGlobals kg = rs->cpu->kg_init();
  {
    ScopedFlushToZero ftz;
    ref_probe_shadow_rays(&kg, sample, x0, y0, w, h, rays);
  }
  rs->cpu->kg_free(&kg);
  return 0;
}

} /* extern "C" */

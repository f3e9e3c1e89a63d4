/* device_b200.cpp - `class B200Device : public ccl::Device`: the drop-in device
 * type for the reference's intern/cycles/device layer (device/device.h:288-500).
 *
 * Host code stays C++ against the reference's own Device / device_memory /
 * DeviceTask / RenderTile API; everything CUDA goes through the C ABI of
 * include/b200_cycles.h (libb200cycles.so).  In the reference tree this file
 * would live at intern/cycles/device/device_b200.cpp next to device_cuda.cpp and
 * be reached through Device::create (device/device.cpp:367-418) - the ~15 line
 * registration patch is in INTEGRATION.md.  Here it is compiled against the
 * reference headers where they lie and exported through two C functions so
 * the harness (oracle/ref_harness.cpp) or any host can hand the Device* to a
 * reference Scene.
 *
 * Contract followed:
 *   - errors are latched with set_error(), never thrown (device.h:333-348);
 *   - mem_* set device_pointer / device_size and keep Stats current
 *     (device_cpu.cpp:386-460); MEM_GLOBAL uploads are bound by mem.name
 *     (device_cuda_impl.cpp:1088-1096);
 *   - task_add(RENDER) does not block: a DedicatedTaskPool worker runs the
 *     acquire_tile / render / release_tile loop of CUDADevice::thread_run
 *     (device_cuda_impl.cpp:2342-2390); FILM_CONVERT runs on the caller;
 *   - get_bvh_layout_mask() asks the host for the packed BVH2 arrays, from which
 *     the device builds its own compressed BVH8 (as OptiX builds its own
 *     structure in build_optix_bvh, device_optix.cpp:1199).
 */
#include "device/device.h"
#include "device/device_intern.h"
#include "device/device_memory.h"
#include "device/device_task.h"
#include "render/buffers.h"
#include "util/util_foreach.h"
#include "util/util_string.h"
#include "util/util_task.h"
#include "util/util_time.h"

#include "../../include/b200_cycles.h"

CCL_NAMESPACE_BEGIN

/* The unpatched reference enum has no DEVICE_B200; the registration patch adds
 * it after DEVICE_OPTIX.  Standalone builds use the same value. */
static const DeviceType DEVICE_B200_TYPE = (DeviceType)(DEVICE_OPTIX + 1);

class B200Device : public Device {
 public:
  b200_ctx *ctx;
  DedicatedTaskPool task_pool;
  volatile int cancel_flag;
  b200_stats last_stats;

  B200Device(DeviceInfo &info, Stats &stats, Profiler &profiler, bool background_)
      : Device(info, stats, profiler, background_), ctx(NULL), cancel_flag(0)
  {
    memset(&last_stats, 0, sizeof(last_stats));
    char err[512] = {0};
    ctx = b200_create(info.num, err, sizeof(err));
    if (!ctx)
      set_error(string("B200 device: ") + err);
  }

  ~B200Device()
  {
    task_pool.cancel();
    if (ctx)
      b200_destroy(ctx);
  }

  bool check(int rc, const char *what)
  {
    if (rc == B200_OK)
      return true;
    set_error(string_printf("B200 device: %s failed: %s", what, b200_last_error(ctx)));
    return false;
  }

  virtual BVHLayoutMask get_bvh_layout_mask() const
  {
    /* The device consumes the host's packed BVH2 and derives its BVH8 from it. */
    return BVH_LAYOUT_BVH2;
  }

  virtual bool load_kernels(const DeviceRequestedFeatures & /*requested_features*/)
  {
    /* kernels are precompiled sm_100a SASS inside libb200cycles.so */
    return ctx != NULL && !have_error();
  }

  virtual bool show_samples() const
  {
    return false;
  }

  /* ---- memory ---- */

  virtual void mem_alloc(device_memory &mem)
  {
    if (!ctx || mem.device_pointer)
      return;
    if (mem.type == MEM_TEXTURE) {
      set_error("B200 device: image textures are outside the hot-path scope");
      return;
    }
    uint64_t dptr = 0;
    const size_t size = mem.memory_size();
    if (!check(b200_alloc(ctx, size, &dptr), "mem_alloc"))
      return;
    mem.device_pointer = (device_ptr)dptr;
    mem.device_size = size;
    stats.mem_alloc(size);
  }

  virtual void mem_copy_to(device_memory &mem)
  {
    if (!ctx)
      return;
    if (mem.type == MEM_PIXELS) {
      assert(!"mem_copy_to not supported for pixels.");
      return;
    }
    /* (re)allocate on size change, as CUDADevice::global_alloc does */
    if (mem.device_pointer && mem.device_size != mem.memory_size())
      mem_free(mem);
    if (!mem.device_pointer)
      mem_alloc(mem);
    if (!mem.device_pointer)
      return;
    if (mem.host_pointer && mem.memory_size()) {
      if (!check(b200_h2d(ctx, (uint64_t)mem.device_pointer, mem.host_pointer, 0,
                          mem.memory_size()),
                 "mem_copy_to"))
        return;
    }
    if (mem.type == MEM_GLOBAL) {
      check(b200_bind_global(ctx, mem.name, (uint64_t)mem.device_pointer, mem.host_pointer,
                             mem.memory_size()),
            mem.name);
    }
  }

  virtual void mem_copy_from(device_memory &mem, int y, int w, int h, int elem)
  {
    if (!ctx || !mem.device_pointer || !mem.host_pointer)
      return;
    const size_t offset = (size_t)elem * y * w;
    const size_t size = (size_t)elem * w * h;
    check(b200_d2h(ctx, (uint64_t)mem.device_pointer, (char *)mem.host_pointer + offset, offset,
                   size),
          "mem_copy_from");
  }

  virtual void mem_zero(device_memory &mem)
  {
    if (!mem.device_pointer)
      mem_alloc(mem);
    if (!mem.device_pointer)
      return;
    check(b200_zero(ctx, (uint64_t)mem.device_pointer, 0, mem.memory_size()), "mem_zero");
    if (mem.host_pointer)
      memset(mem.host_pointer, 0, mem.memory_size());
  }

  virtual void mem_free(device_memory &mem)
  {
    if (!ctx || !mem.device_pointer)
      return;
    task_pool.wait();
    check(b200_free(ctx, (uint64_t)mem.device_pointer), "mem_free");
    stats.mem_free(mem.device_size);
    mem.device_pointer = 0;
    mem.device_size = 0;
  }

  virtual device_ptr mem_alloc_sub_ptr(device_memory &mem, int offset, int /*size*/)
  {
    return (device_ptr)(((char *)mem.device_pointer) + mem.memory_elements_size(offset));
  }

  virtual void const_copy_to(const char *name, void *host, size_t size)
  {
    if (!ctx)
      return;
    if (strcmp(name, "__data") != 0) {
      set_error(string("B200 device: unknown constant ") + name);
      return;
    }
    check(b200_set_kernel_data(ctx, host, size), "const_copy_to(__data)");
  }

  /* ---- tasks ---- */

  void film_convert(DeviceTask &task)
  {
    const bool half_float = task.rgba_half != 0;
    const float sample_scale = 1.0f / (task.sample + 1);
    check(b200_film_convert(ctx, (uint64_t)task.buffer,
                            (uint64_t)(half_float ? task.rgba_half : task.rgba_byte), half_float,
                            sample_scale, task.x, task.y, task.w, task.h, task.offset, task.stride),
          "film_convert");
  }

  void thread_run(DeviceTask &task)
  {
    if (task.type != DeviceTask::RENDER) {
      set_error("B200 device: only RENDER and FILM_CONVERT tasks are in scope");
      return;
    }
    RenderTile tile;
    while (task.acquire_tile(this, tile, task.tile_types)) {
      if (tile.task == RenderTile::PATH_TRACE) {
        scoped_timer timer(&tile.buffers->render_time);
        b200_work_tile wt;
        wt.x = tile.x;
        wt.y = tile.y;
        wt.w = tile.w;
        wt.h = tile.h;
        wt.start_sample = tile.start_sample;
        wt.num_samples = tile.num_samples;
        wt.offset = tile.offset;
        wt.stride = tile.stride;
        wt.buffer = (uint64_t)tile.buffer;
        cancel_flag = 0;
        const int rc = b200_render(ctx, &wt, &cancel_flag);
        if (rc == B200_OK) {
          tile.sample = tile.start_sample + tile.num_samples;
          b200_get_stats(ctx, &last_stats);
          task.update_progress(&tile, tile.w * tile.h * tile.num_samples);
        }
        else if (rc != B200_ERR_CANCELLED) {
          check(rc, "render");
        }
      }
      else {
        set_error("B200 device: bake / denoise tiles are outside the hot-path scope");
      }
      task.release_tile(tile);
      if (have_error())
        break;
      if (task.get_cancel() || task_pool.canceled()) {
        if (task.need_finish_queue == false)
          break;
      }
    }
  }

  virtual void task_add(DeviceTask &task)
  {
    if (!ctx || have_error())
      return;
    if (task.type == DeviceTask::FILM_CONVERT) {
      /* synchronously on the caller, as CUDADevice does (device_cuda_impl.cpp:2427) */
      film_convert(task);
    }
    else {
      task_pool.push([=] {
        DeviceTask task_copy = task;
        thread_run(task_copy);
      });
    }
  }

  virtual void task_wait()
  {
    task_pool.wait();
  }

  virtual void task_cancel()
  {
    cancel_flag = 1;
    task_pool.cancel();
  }
};

bool device_b200_init()
{
  return b200_device_count() > 0;
}

Device *device_b200_create(DeviceInfo &info, Stats &stats, Profiler &profiler, bool background)
{
  return new B200Device(info, stats, profiler, background);
}

void device_b200_info(vector<DeviceInfo> &devices)
{
  const int n = b200_device_count();
  for (int i = 0; i < n; i++) {
    char name[256] = {0};
    int major = 0, minor = 0, sms = 0;
    uint64_t mem = 0;
    if (b200_device_name(i, name, sizeof(name), &major, &minor, &mem, &sms) != B200_OK)
      continue;
    if (major != 10)
      continue; /* sm_100 only */
    DeviceInfo info;
    info.type = DEVICE_B200_TYPE;
    info.description = string(name);
    info.num = i;
    info.id = string_printf("B200_%s_%d", name, i);
    info.has_half_images = false;
    info.has_volume_decoupled = false;
    info.has_adaptive_stop_per_sample = false;
    info.has_osl = false;
    info.use_split_kernel = false;
    info.has_profiling = false;
    info.has_peer_memory = n > 1;
    info.display_device = false;
    devices.push_back(info);
  }
}

CCL_NAMESPACE_END

/* ---- C entry points for hosts that cannot name ccl:: types (ctypes harness) ---- */

struct b200_host_device {
  ccl::Stats stats;
  ccl::Profiler profiler;
  ccl::DeviceInfo info;
  ccl::B200Device *device;
};

extern "C" {

/* Returns an opaque handle; b200_host_device_ptr() gives the ccl::Device*. */
void *b200_host_device_create(int ordinal, char *err, size_t errlen)
{
  b200_host_device *h = new b200_host_device();
  ccl::vector<ccl::DeviceInfo> infos;
  ccl::device_b200_info(infos);
  bool found = false;
  foreach (ccl::DeviceInfo &info, infos) {
    if (info.num == ordinal) {
      h->info = info;
      found = true;
    }
  }
  if (!found) {
    snprintf(err, errlen, "no sm_100 device with ordinal %d", ordinal);
    delete h;
    return NULL;
  }
  h->device = new ccl::B200Device(h->info, h->stats, h->profiler, true);
  if (h->device->have_error()) {
    snprintf(err, errlen, "%s", h->device->error_message().c_str());
    delete h->device;
    delete h;
    return NULL;
  }
  return h;
}

void *b200_host_device_ptr(void *handle)
{
  return handle ? (void *)static_cast<ccl::Device *>(((b200_host_device *)handle)->device) : NULL;
}

int b200_host_device_stats(void *handle, b200_stats *out)
{
  if (!handle || !out)
    return B200_ERR_INVALID;
  *out = ((b200_host_device *)handle)->device->last_stats;
  return B200_OK;
}

const char *b200_host_device_error(void *handle)
{
  static thread_local std::string msg;
  msg = handle ? ((b200_host_device *)handle)->device->error_message() : "null handle";
  return msg.c_str();
}

void b200_host_device_destroy(void *handle)
{
  if (!handle)
    return;
  b200_host_device *h = (b200_host_device *)handle;
  delete h->device;
  delete h;
}

} /* extern "C" */

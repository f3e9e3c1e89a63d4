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
 *   - get_bvh_layout_mask() asks the host for BVH_LAYOUT_BVH8, the device's own layout,
 *     packed on the host by `BVH8 : BVH` (bvh8_host.cpp) - when the host application
 *     knows that layout.  An unpatched host is asked for the
 *     packed BVH2 arrays instead, from which the device derives the same BVH8 itself (as
 *     OptiX builds its own structure in build_optix_bvh, device_optix.cpp:1199).
 */
#include "device/device.h"
#include "device/device_intern.h"
#include "device/device_memory.h"
#include "device/device_task.h"
#include "render/buffers.h"
#include "util/util_foreach.h"
#include "util/util_map.h"
#include "util/util_string.h"
#include "util/util_task.h"
#include "util/util_thread.h"
#include "util/util_time.h"

#include "../../include/b200_cycles.h"

/* oracle/ref_host_hooks.cpp in this repo's host library; in a patched reference tree the
 * layout is simply part of BVH::create */
extern "C" int ref_host_has_bvh_layout(int layout);

CCL_NAMESPACE_BEGIN
string bvh8_last_error();                                        /* bvh8_host.cpp */
void bvh8_last_info(b200_bvh_info *info, double *pack_seconds);

/* B200_HOST_BVH=bvh2 keeps the device-side derivation even on a host that knows BVH8 (A/B
 * and parity of the two routes) */
static BVHLayoutMask b200_bvh_layout_mask()
{
  const char *force = getenv("B200_HOST_BVH");
  if (force && strcmp(force, "bvh2") == 0)
    return BVH_LAYOUT_BVH2;
  return ref_host_has_bvh_layout((int)B200_BVH_LAYOUT_BVH8) ? (BVHLayoutMask)B200_BVH_LAYOUT_BVH8 :
                                                              (BVHLayoutMask)BVH_LAYOUT_BVH2;
}
CCL_NAMESPACE_END

CCL_NAMESPACE_BEGIN

/* The unpatched reference enum has no DEVICE_B200; the registration patch adds
 * it after DEVICE_OPTIX.  Standalone builds use the same value. */
static const DeviceType DEVICE_B200_TYPE = (DeviceType)(DEVICE_OPTIX + 1);

/* task.get_cancel() as the C ABI's cancel predicate: b200_render asks it between
 * wavefront batches, the way CUDADevice::render polls it every sample step
 * (device_cuda_impl.cpp:1939). */
struct B200CancelProbe {
  DeviceTask *task;
  DedicatedTaskPool *pool;
  static int ask(void *user)
  {
    B200CancelProbe *p = (B200CancelProbe *)user;
    if (p->pool->canceled())
      return 1;
    return (p->task->get_cancel() && !p->task->need_finish_queue) ? 1 : 0;
  }
};

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
    const string bvh_error = bvh8_last_error();
    set_error(string_printf("B200 device: %s failed: %s%s%s", what, b200_last_error(ctx),
                            bvh_error.empty() ? "" : " (host BVH8 pack: ",
                            bvh_error.empty() ? "" : (bvh_error + ")").c_str()));
    return false;
  }

  virtual BVHLayoutMask get_bvh_layout_mask() const
  {
    return b200_bvh_layout_mask();
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
    else if (mem.type == MEM_TEXTURE) {
      /* CUDADevice::tex_alloc (device_cuda_impl.cpp:1105-1304): the ImageManager's pixels
       * are an ordinary allocation; the slot's TextureInfo goes to the device's table with
       * `data` = the device address (the kernels sample with the CPU device's arithmetic,
       * so no CUDA array / texture object is made) */
      device_texture &tex = (device_texture &)mem;
      check(b200_texture_set(ctx, (int)tex.slot, &tex.info, sizeof(TextureInfo),
                             (uint64_t)mem.device_pointer),
            "tex_alloc");
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
    /* No task_pool.wait() here: Session::release_tile deletes a finished tile's
     * RenderBuffers ON the device's worker thread (session.cpp:517-518), so waiting for
     * the pool from mem_free would wait for the caller itself.  b200_free synchronises
     * the context's stream before the memory goes, as CUDADevice::mem_free relies on its
     * context. */
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
    if (task.type == DeviceTask::SHADER && task.shader_eval_type == SHADER_EVAL_BACKGROUND) {
      /* the light manager's evaluation of the world shader for the background importance
       * map (render/light.cpp:38-102; CUDADevice::shader, device_cuda_impl.cpp:2019-2093) */
      check(b200_shader_eval_background(ctx, (uint64_t)task.shader_input,
                                        (uint64_t)task.shader_output, task.shader_x,
                                        task.shader_w),
            "shader(background)");
      return;
    }
    if (task.type != DeviceTask::RENDER) {
      set_error("B200 device: of the SHADER tasks only the background evaluation is in scope "
                "(no baking, no displacement)");
      return;
    }
    B200CancelProbe probe = {&task, &task_pool};
    b200_set_cancel_callback(ctx, B200CancelProbe::ask, &probe);
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
    b200_set_cancel_callback(ctx, NULL, NULL);
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
      /* a new task after a task_cancel(): the flag is cleared here, where tasks are
       * accepted, never by the worker (that could erase a cancel that just arrived) */
      cancel_flag = 0;
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

/* One process, several B200s: the in-process counterpart of the reference's
 * MultiDevice (device/device_multi.cpp), built on the same C ABI.
 *
 * The reference MultiDevice hands different TILES to its sub-devices and stitches
 * the film on the host (mem_copy_from slicing, device_multi.cpp:374-393).  Here every
 * GPU renders the whole tile for its share of the SAMPLES (the Sobol
 * index is the sample number, so the shares jointly equal the single-device sample
 * set) and the per-GPU films are summed on the device over NVLink
 * (b200_film_reduce).  Scene arrays are replicated; device_memory::device_pointer
 * is the pointer on the first GPU and keys a table of the per-GPU pointers. */
class B200MultiDevice : public Device {
 public:
  struct Allocation {
    vector<uint64_t> ptr; /* per GPU */
    size_t size;
  };
  vector<b200_ctx *> ctxs;
  map<device_ptr, Allocation> allocations;
  thread_mutex alloc_mutex;
  DedicatedTaskPool task_pool;
  volatile int cancel_flag;
  b200_stats last_stats;
  bool distinct_gpus; /* one context per GPU: the film sum is an NCCL all-reduce */

  B200MultiDevice(DeviceInfo &info, Stats &stats, Profiler &profiler, bool background_,
                  const vector<int> &ordinals)
      : Device(info, stats, profiler, background_), cancel_flag(0), distinct_gpus(true)
  {
    memset(&last_stats, 0, sizeof(last_stats));
    for (size_t i = 0; i < ordinals.size(); i++)
      for (size_t j = 0; j < i; j++)
        if (ordinals[i] == ordinals[j])
          distinct_gpus = false;
    foreach (int ordinal, ordinals) {
      char err[512] = {0};
      b200_ctx *ctx = b200_create(ordinal, err, sizeof(err));
      if (!ctx) {
        set_error(string("B200 multi device: ") + err);
        break;
      }
      ctxs.push_back(ctx);
    }
  }

  ~B200MultiDevice()
  {
    task_pool.cancel();
    foreach (b200_ctx *ctx, ctxs)
      b200_destroy(ctx);
  }

  bool check(size_t i, int rc, const char *what)
  {
    if (rc == B200_OK)
      return true;
    set_error(string_printf(
        "B200 multi device: %s failed on GPU %d: %s", what, (int)i, b200_last_error(ctxs[i])));
    return false;
  }

  Allocation *find(device_ptr primary)
  {
    thread_scoped_lock lock(alloc_mutex);
    map<device_ptr, Allocation>::iterator it = allocations.find(primary);
    return (it == allocations.end()) ? NULL : &it->second;
  }

  virtual BVHLayoutMask get_bvh_layout_mask() const
  {
    return b200_bvh_layout_mask();
  }
  virtual bool load_kernels(const DeviceRequestedFeatures &)
  {
    return !ctxs.empty() && !have_error();
  }
  virtual bool show_samples() const
  {
    return false;
  }

  virtual void mem_alloc(device_memory &mem)
  {
    if (ctxs.empty() || mem.device_pointer || have_error())
      return;
    Allocation a;
    a.size = mem.memory_size();
    for (size_t i = 0; i < ctxs.size(); i++) {
      uint64_t dptr = 0;
      if (!check(i, b200_alloc(ctxs[i], a.size, &dptr), "mem_alloc")) {
        for (size_t j = 0; j < a.ptr.size(); j++)
          b200_free(ctxs[j], a.ptr[j]);
        return;
      }
      a.ptr.push_back(dptr);
    }
    mem.device_pointer = (device_ptr)a.ptr[0];
    mem.device_size = a.size;
    stats.mem_alloc(a.size * ctxs.size());
    thread_scoped_lock lock(alloc_mutex);
    allocations[mem.device_pointer] = a;
  }

  virtual void mem_copy_to(device_memory &mem)
  {
    if (ctxs.empty() || mem.type == MEM_PIXELS)
      return;
    if (mem.device_pointer && mem.device_size != mem.memory_size())
      mem_free(mem);
    if (!mem.device_pointer)
      mem_alloc(mem);
    Allocation *a = find(mem.device_pointer);
    if (!a)
      return;
    for (size_t i = 0; i < ctxs.size(); i++) {
      if (mem.host_pointer && mem.memory_size()) {
        if (!check(i, b200_h2d(ctxs[i], a->ptr[i], mem.host_pointer, 0, mem.memory_size()),
                   "mem_copy_to"))
          return;
      }
      if (mem.type == MEM_GLOBAL) {
        if (!check(i, b200_bind_global(ctxs[i], mem.name, a->ptr[i], mem.host_pointer,
                                       mem.memory_size()),
                   mem.name))
          return;
      }
      else if (mem.type == MEM_TEXTURE) {
        device_texture &tex = (device_texture &)mem;
        if (!check(i, b200_texture_set(ctxs[i], (int)tex.slot, &tex.info, sizeof(TextureInfo),
                                       a->ptr[i]),
                   "tex_alloc"))
          return;
      }
    }
  }

  virtual void mem_copy_from(device_memory &mem, int y, int w, int h, int elem)
  {
    if (ctxs.empty() || !mem.device_pointer || !mem.host_pointer)
      return;
    /* films are kept summed on the first GPU (thread_run) */
    const size_t offset = (size_t)elem * y * w;
    const size_t size = (size_t)elem * w * h;
    check(0, b200_d2h(ctxs[0], (uint64_t)mem.device_pointer, (char *)mem.host_pointer + offset,
                      offset, size),
          "mem_copy_from");
  }

  virtual void mem_zero(device_memory &mem)
  {
    if (!mem.device_pointer)
      mem_alloc(mem);
    Allocation *a = find(mem.device_pointer);
    if (!a)
      return;
    for (size_t i = 0; i < ctxs.size(); i++)
      check(i, b200_zero(ctxs[i], a->ptr[i], 0, a->size), "mem_zero");
    if (mem.host_pointer)
      memset(mem.host_pointer, 0, mem.memory_size());
  }

  virtual void mem_free(device_memory &mem)
  {
    if (ctxs.empty() || !mem.device_pointer)
      return;
    /* no task_pool.wait(): may run on the worker thread itself (see B200Device) */
    Allocation *a = find(mem.device_pointer);
    if (a) {
      for (size_t i = 0; i < ctxs.size(); i++)
        check(i, b200_free(ctxs[i], a->ptr[i]), "mem_free");
      stats.mem_free(a->size * ctxs.size());
      thread_scoped_lock lock(alloc_mutex);
      allocations.erase(mem.device_pointer);
    }
    mem.device_pointer = 0;
    mem.device_size = 0;
  }

  virtual void const_copy_to(const char *name, void *host, size_t size)
  {
    if (strcmp(name, "__data") != 0) {
      set_error(string("B200 device: unknown constant ") + name);
      return;
    }
    for (size_t i = 0; i < ctxs.size(); i++)
      check(i, b200_set_kernel_data(ctxs[i], host, size), "const_copy_to(__data)");
  }

  void thread_run(DeviceTask &task)
  {
    if (task.type == DeviceTask::SHADER && task.shader_eval_type == SHADER_EVAL_BACKGROUND) {
      /* on the first GPU: its allocation is the one mem_copy_from reads */
      check(0, b200_shader_eval_background(ctxs[0], (uint64_t)task.shader_input,
                                           (uint64_t)task.shader_output, task.shader_x,
                                           task.shader_w),
            "shader(background)");
      return;
    }
    if (task.type != DeviceTask::RENDER) {
      set_error("B200 device: of the SHADER tasks only the background evaluation is in scope "
                "(no baking, no displacement)");
      return;
    }
    const int n = (int)ctxs.size();
    B200CancelProbe probe = {&task, &task_pool};
    for (int i = 0; i < n; i++)
      b200_set_cancel_callback(ctxs[i], B200CancelProbe::ask, &probe);
    RenderTile tile;
    while (task.acquire_tile(this, tile, task.tile_types)) {
      Allocation *film = find(tile.buffer);
      if (tile.task != RenderTile::PATH_TRACE || !film) {
        set_error("B200 multi device: unsupported tile (not a path-trace tile of a known film)");
        task.release_tile(tile);
        break;
      }
      {
        scoped_timer timer(&tile.buffers->render_time);
        /* contiguous sample ranges whose sizes differ by at most one */
        const int base = tile.num_samples / n, rem = tile.num_samples % n;
        vector<int> rcs(n, B200_OK);
        vector<thread *> workers;
        for (int i = 0; i < n; i++) {
          b200_work_tile wt;
          wt.x = tile.x;
          wt.y = tile.y;
          wt.w = tile.w;
          wt.h = tile.h;
          wt.start_sample = tile.start_sample + i * base + min(i, rem);
          wt.num_samples = base + (i < rem ? 1 : 0);
          wt.offset = tile.offset;
          wt.stride = tile.stride;
          wt.buffer = film->ptr[i];
          workers.push_back(new thread([this, i, wt, &rcs] {
            rcs[i] = (wt.num_samples > 0) ? b200_render(ctxs[i], &wt, &cancel_flag) : B200_OK;
          }));
        }
        foreach (thread *w, workers) {
          w->join();
          delete w;
        }
        bool ok = true;
        for (int i = 0; i < n; i++) {
          if (rcs[i] != B200_OK && rcs[i] != B200_ERR_CANCELLED)
            ok = check(i, rcs[i], "render") && ok;
          if (rcs[i] != B200_OK)
            ok = false;
        }
        if (ok && n > 1) {
          /* Sum the per-GPU films: one NCCL all-reduce over NVLink when every context
           * sits on its own GPU; contexts sharing a GPU (single-GPU test boxes) fall back
           * to peer copies + add kernels on the first one.  Then the films of the other
           * GPUs are cleared so that the next tile / sample range starts from zero there
           * (the first one holds the running sum the host reads). */
          if (distinct_gpus)
            ok = check(0, b200_film_allreduce(ctxs.data(), n, film->ptr.data(), film->size / 4),
                       "film_allreduce");
          else
            ok = check(0, b200_film_reduce(ctxs.data(), n, film->ptr.data(), film->size / 4),
                       "film_reduce");
          for (int i = 1; ok && i < n; i++)
            ok = check(i, b200_zero(ctxs[i], film->ptr[i], 0, film->size), "film clear");
        }
        if (ok) {
          tile.sample = tile.start_sample + tile.num_samples;
          memset(&last_stats, 0, sizeof(last_stats));
          for (int i = 0; i < n; i++) {
            b200_stats st;
            b200_get_stats(ctxs[i], &st);
            last_stats.primary_rays += st.primary_rays;
            last_stats.bounce_rays += st.bounce_rays;
            last_stats.shadow_rays += st.shadow_rays;
            last_stats.kernel_launches += st.kernel_launches;
            last_stats.device_ms = max(last_stats.device_ms, st.device_ms);
          }
          task.update_progress(&tile, tile.w * tile.h * tile.num_samples);
        }
      }
      task.release_tile(tile);
      if (have_error())
        break;
      if (task.get_cancel() || task_pool.canceled()) {
        if (task.need_finish_queue == false)
          break;
      }
    }
    for (int i = 0; i < n; i++)
      b200_set_cancel_callback(ctxs[i], NULL, NULL);
  }

  virtual void task_add(DeviceTask &task)
  {
    if (ctxs.empty() || have_error())
      return;
    if (task.type == DeviceTask::FILM_CONVERT) {
      const bool half_float = task.rgba_half != 0;
      check(0, b200_film_convert(ctxs[0], (uint64_t)task.buffer,
                                 (uint64_t)(half_float ? task.rgba_half : task.rgba_byte),
                                 half_float, 1.0f / (task.sample + 1), task.x, task.y, task.w,
                                 task.h, task.offset, task.stride),
            "film_convert");
    }
    else {
      cancel_flag = 0;
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
  /* Device::get_multi_device of a B200-only list (device.cpp:583-655): one device over
   * the sub-devices' GPUs instead of the generic MultiDevice's tile fan-out */
  if (!info.multi_devices.empty()) {
    vector<int> ordinals;
    foreach (const DeviceInfo &sub, info.multi_devices)
      ordinals.push_back(sub.num);
    return new B200MultiDevice(info, stats, profiler, background, ordinals);
  }
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
    /* stable across reboots and re-enumeration: the PCI location, as device_cuda_info()
     * does (device_cuda.cpp:144-152) */
    char pci[32] = {0};
    if (b200_device_pci_id(i, pci, sizeof(pci)) == B200_OK)
      info.id = string_printf("B200_%s_%s", name, pci);
    else
      info.id = string_printf("B200_%s_%d", name, i);
    info.has_half_images = true; /* half images are sampled as stored (svm_image.cuh) */
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

/* In a patched tree Device::create calls the three functions above directly
 * (INTEGRATION.md section 2).  The host library of this repo is built from the read-only
 * reference with the same rows on a copy of device.cpp, reaching them through pointers
 * registered here when this library is loaded (oracle/ref_host_hooks.cpp). */
extern "C" void ref_host_register_b200_device(void *init, void *create, void *info);

namespace {
struct B200DeviceRegistration {
  B200DeviceRegistration()
  {
    ref_host_register_b200_device((void *)&device_b200_init, (void *)&device_b200_create,
                                  (void *)&device_b200_info);
  }
} g_b200_device_registration;
}  // namespace

CCL_NAMESPACE_END

/* ---- C entry points for hosts that cannot name ccl:: types (ctypes harness) ---- */

struct b200_host_device {
  ccl::Stats stats;
  ccl::Profiler profiler;
  ccl::DeviceInfo info;
  ccl::B200Device *device;
  ccl::B200MultiDevice *multi;
  b200_host_device() : device(NULL), multi(NULL)
  {
  }
  ccl::Device *any()
  {
    return device ? static_cast<ccl::Device *>(device) : static_cast<ccl::Device *>(multi);
  }
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

/* One Device over several GPUs (B200MultiDevice): `ordinals` may repeat, which gives
 * several contexts on one GPU (used by the tests on single-GPU boxes). */
void *b200_host_multi_device_create(const int *ordinals, int n, char *err, size_t errlen)
{
  if (!ordinals || n <= 0) {
    snprintf(err, errlen, "no ordinals given");
    return NULL;
  }
  b200_host_device *h = new b200_host_device();
  ccl::vector<ccl::DeviceInfo> infos;
  ccl::device_b200_info(infos);
  if (infos.empty()) {
    snprintf(err, errlen, "no sm_100 device");
    delete h;
    return NULL;
  }
  h->info = infos[0];
  h->info.id = "B200_MULTI";
  ccl::vector<int> ords(ordinals, ordinals + n);
  h->multi = new ccl::B200MultiDevice(h->info, h->stats, h->profiler, true, ords);
  if (h->multi->have_error()) {
    snprintf(err, errlen, "%s", h->multi->error_message().c_str());
    delete h->multi;
    delete h;
    return NULL;
  }
  return h;
}

void *b200_host_device_ptr(void *handle)
{
  return handle ? (void *)((b200_host_device *)handle)->any() : NULL;
}

/* Report of the last top-level BVH8 the host class packed (bvh8_host.cpp) and the seconds
 * BVH8::pack_nodes spent on the 2 -> 8 collapse; `err` receives the reason when the last
 * pack was refused. */
int b200_host_bvh8_report(b200_bvh_info *info, double *pack_seconds, char *err, size_t errlen)
{
  if (!info || !pack_seconds)
    return B200_ERR_INVALID;
  ccl::bvh8_last_info(info, pack_seconds);
  if (err && errlen) {
    const std::string e = ccl::bvh8_last_error();
    strncpy(err, e.c_str(), errlen - 1);
    err[errlen - 1] = 0;
  }
  return B200_OK;
}

/* The BVH the device traverses for the scene bound last (b200_build_bvh on the first
 * context): host_packed = 1 when the host's BVH8 class delivered the device layout. */
int b200_host_device_bvh_info(void *handle, b200_bvh_info *out)
{
  if (!handle || !out)
    return B200_ERR_INVALID;
  b200_host_device *h = (b200_host_device *)handle;
  b200_ctx *ctx = h->device ? h->device->ctx : (h->multi->ctxs.empty() ? NULL : h->multi->ctxs[0]);
  return ctx ? b200_build_bvh(ctx, out) : B200_ERR_INVALID;
}

int b200_host_device_stats(void *handle, b200_stats *out)
{
  if (!handle || !out)
    return B200_ERR_INVALID;
  b200_host_device *h = (b200_host_device *)handle;
  *out = h->device ? h->device->last_stats : h->multi->last_stats;
  return B200_OK;
}

/* For a ccl::Device* that came out of the reference's own registry (Device::create with
 * the "B200" type): the last task's counters if it is one of ours, B200_ERR_INVALID if
 * the registry handed out something else. */
int b200_registered_device_stats(void *device_ptr, b200_stats *out)
{
  ccl::B200Device *dev = dynamic_cast<ccl::B200Device *>((ccl::Device *)device_ptr);
  ccl::B200MultiDevice *multi = dynamic_cast<ccl::B200MultiDevice *>((ccl::Device *)device_ptr);
  if ((!dev && !multi) || !out)
    return B200_ERR_INVALID;
  *out = dev ? dev->last_stats : multi->last_stats;
  return B200_OK;
}

const char *b200_host_device_error(void *handle)
{
  static thread_local std::string msg;
  msg = handle ? ((b200_host_device *)handle)->any()->error_message() : "null handle";
  return msg.c_str();
}

void b200_host_device_destroy(void *handle)
{
  if (!handle)
    return;
  b200_host_device *h = (b200_host_device *)handle;
  delete h->device;
  delete h->multi;
  delete h;
}

} /* extern "C" */

/* Part of oracle/_ref/libcycles_host.so - the reference's HOST code only (Device base
 * class, device_memory, DeviceTask, TaskPool, Scene, BVH builder, SVM compiler; see
 * oracle/Makefile): the library the C++ device shim links against, where a Blender /
 * cycles_standalone build would be the host application.
 *
 * It carries NO kernels and NO CPU device.  Device::create / available_devices
 * (device/device.cpp:382, 529, 559) reference the CPU device's three factory functions;
 * here they forward to hooks that stay empty unless the ORACLE library
 * (libcycles_ref.so: device_cpu.cpp + kernel/kernels/cpu/*.cpp + the harness) was loaded
 * and registered its own - so the split between "reference host code the product may
 * link" and "reference kernels only the checker may run" is structural: `nm -D` of this
 * library and of libcycles_device_b200.so shows no path to kernel_cpu_*. */
#include <cstdio>
#include <string>

#include "device/device.h"
#include "device/device_intern.h"

CCL_NAMESPACE_BEGIN

typedef Device *(*cpu_create_fn)(DeviceInfo &, Stats &, Profiler &, bool);
typedef void (*cpu_info_fn)(vector<DeviceInfo> &);
typedef string (*cpu_capabilities_fn)();

static cpu_create_fn g_cpu_create = NULL;
static cpu_info_fn g_cpu_info = NULL;
static cpu_capabilities_fn g_cpu_capabilities = NULL;

Device *device_cpu_create(DeviceInfo &info, Stats &stats, Profiler &profiler, bool background)
{
  return g_cpu_create ? g_cpu_create(info, stats, profiler, background) : NULL;
}

void device_cpu_info(vector<DeviceInfo> &devices)
{
  if (g_cpu_info)
    g_cpu_info(devices);
}

string device_cpu_capabilities()
{
  return g_cpu_capabilities ? g_cpu_capabilities() : string("");
}

CCL_NAMESPACE_END

extern "C" void ref_host_register_cpu_device(void *create, void *info, void *capabilities)
{
  ccl::g_cpu_create = (ccl::cpu_create_fn)create;
  ccl::g_cpu_info = (ccl::cpu_info_fn)info;
  ccl::g_cpu_capabilities = (ccl::cpu_capabilities_fn)capabilities;
}

/* ---- BVH layouts added by a device plug-in ----
 * BVH::create (bvh/bvh.cpp:99-124) knows BVH2 / Embree / OptiX.  The patched copy of
 * bvh.cpp this library is built with (oracle/Makefile, bvh_layout_hook.sed) asks here
 * first, so that a device library can bring its own `BVH` subclass the way
 * device_optix.cpp brings BVHOptiX - without the subclass living in this tree. */
#include "bvh/bvh.h"
#include "bvh/bvh_params.h"

CCL_NAMESPACE_BEGIN

typedef BVH *(*bvh_create_fn)(const BVHParams &, const vector<Geometry *> &,
                              const vector<Object *> &);
static int g_bvh_hook_layout = 0;
static bvh_create_fn g_bvh_hook_create = NULL;

BVH *bvh_layout_hook_create(const BVHParams &params,
                            const vector<Geometry *> &geometry,
                            const vector<Object *> &objects)
{
  if (g_bvh_hook_create && (int)params.bvh_layout == g_bvh_hook_layout)
    return g_bvh_hook_create(params, geometry, objects);
  return NULL;
}

CCL_NAMESPACE_END

extern "C" void ref_host_register_bvh_layout(int layout, void *create)
{
  ccl::g_bvh_hook_layout = layout;
  ccl::g_bvh_hook_create = (ccl::bvh_create_fn)create;
}

extern "C" int ref_host_has_bvh_layout(int layout)
{
  return ccl::g_bvh_hook_create != NULL && ccl::g_bvh_hook_layout == layout;
}

/* ---- a device type added by a plug-in ----
 * Device::create / type_from_string / string_from_type / available_types /
 * available_devices (device/device.cpp:367-550) know CPU / CUDA / OptiX / OpenCL / network.
 * The patched copy of device.cpp this library is built with (oracle/Makefile,
 * device_registry_hook.sed) carries the "B200" rows of the registration patch; what a
 * patched Blender tree calls directly (device_b200_init / _create / _info, compiled in
 * from device_b200.cpp) is reached here through pointers the device library registers
 * when it is loaded - this library cannot link against it, the dependency runs the other
 * way. */
CCL_NAMESPACE_BEGIN

typedef bool (*b200_init_fn)();
typedef Device *(*b200_create_fn)(DeviceInfo &, Stats &, Profiler &, bool);
typedef void (*b200_info_fn)(vector<DeviceInfo> &);
static b200_init_fn g_b200_init = NULL;
static b200_create_fn g_b200_create = NULL;
static b200_info_fn g_b200_info = NULL;

bool b200_plugin_init()
{
  return g_b200_init && g_b200_init();
}

Device *b200_plugin_create(DeviceInfo &info, Stats &stats, Profiler &profiler, bool background)
{
  return g_b200_create ? g_b200_create(info, stats, profiler, background) : NULL;
}

void b200_plugin_info(vector<DeviceInfo> &devices)
{
  if (g_b200_info)
    g_b200_info(devices);
}

CCL_NAMESPACE_END

extern "C" void ref_host_register_b200_device(void *init, void *create, void *info)
{
  ccl::g_b200_init = (ccl::b200_init_fn)init;
  ccl::g_b200_create = (ccl::b200_create_fn)create;
  ccl::g_b200_info = (ccl::b200_info_fn)info;
}

/* ---- the registry as a host application uses it (cycles_standalone --device NAME,
 * app/cycles_standalone.cpp:370-395; BlenderSync::get_session_params) ---- */
extern "C" int ref_host_device_type_from_string(const char *name)
{
  return (int)ccl::Device::type_from_string(name);
}

extern "C" int ref_host_device_type_name(int type, char *out, int out_size)
{
  const std::string s = ccl::Device::string_from_type((ccl::DeviceType)type).c_str();
  snprintf(out, out_size, "%s", s.c_str());
  return (int)s.size();
}

extern "C" int ref_host_device_type_available(int type)
{
  ccl::vector<ccl::DeviceType> types = ccl::Device::available_types();
  for (size_t i = 0; i < types.size(); i++)
    if ((int)types[i] == type)
      return 1;
  return 0;
}

/* Device::available_devices(mask of `type`): how many, and the id / description of one */
extern "C" int ref_host_available_devices(int type, int index, char *id, int id_size,
                                          char *description, int description_size)
{
  ccl::vector<ccl::DeviceInfo> devices = ccl::Device::available_devices(1u << type);
  if (index >= 0 && index < (int)devices.size()) {
    if (id)
      snprintf(id, id_size, "%s", devices[index].id.c_str());
    if (description)
      snprintf(description, description_size, "%s", devices[index].description.c_str());
  }
  return (int)devices.size();
}

/* Device::create for device `index` of `type` - or, with count > 1, for the multi device
 * Device::get_multi_device makes of `count` of them (device.cpp:583-655; the list wraps
 * around when the box has fewer); NULL when there is none.  The Stats / Profiler the device reports into live as long as the
 * process (a Session owns them in the reference). */
extern "C" void *ref_host_device_create(int type, int index, int count)
{
  static ccl::Stats stats;
  static ccl::Profiler profiler;
  ccl::vector<ccl::DeviceInfo> devices = ccl::Device::available_devices(1u << type);
  if (index < 0 || index >= (int)devices.size())
    return NULL;
  if (count > 1) {
    /* fewer GPUs than asked for: the list wraps around (several contexts on one GPU) */
    ccl::vector<ccl::DeviceInfo> sub;
    for (int k = 0; k < count; k++)
      sub.push_back(devices[(index + k) % devices.size()]);
    ccl::DeviceInfo multi = ccl::Device::get_multi_device(sub, 0, true);
    return ccl::Device::create(multi, stats, profiler, true);
  }
  return ccl::Device::create(devices[index], stats, profiler, true);
}

extern "C" void ref_host_device_free(void *device)
{
  delete (ccl::Device *)device;
}

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

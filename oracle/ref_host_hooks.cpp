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

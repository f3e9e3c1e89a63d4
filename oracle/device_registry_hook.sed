# Applied to a COPY of device/device.cpp at build time (oracle/Makefile; the copy lives
# under oracle/_ref/patched/, git-ignored - no reference source enters the repository).
# It is the registration patch of INTEGRATION.md section 2 in the form this read-only tree
# allows: the DEVICE_B200 enumerator is spelled as the value after DEVICE_OPTIX (device.h
# cannot be edited here), and device_b200_init / _create / _info are reached through the
# plug-in hooks of ref_host_hooks.cpp, which libcycles_device_b200.so fills when loaded.
/^bool Device::need_types_update = true;$/a\
\
/* ---- B200 device plug-in (INTEGRATION.md section 2) ---- */\
#define DEVICE_B200 ((DeviceType)(DEVICE_OPTIX + 1))\
#define DEVICE_MASK_B200 (1 << (DEVICE_OPTIX + 1))\
bool b200_plugin_init();\
Device *b200_plugin_create(DeviceInfo &info, Stats &stats, Profiler &profiler, bool background);\
void b200_plugin_info(vector<DeviceInfo> &devices);\
static vector<DeviceInfo> b200_devices;

/^    return device_multi_create(info, stats, profiler, background);$/i\
    /* a B200-only list is ONE device that splits samples and sums films on the GPUs */\
    {\
      bool all_b200 = true;\
      foreach (const DeviceInfo &sub, info.multi_devices)\
        all_b200 &= (sub.type == DEVICE_B200);\
      if (all_b200) {\
        return b200_plugin_init() ? b200_plugin_create(info, stats, profiler, background) : NULL;\
      }\
    }

/^    default:$/i\
    case DEVICE_B200:\
      if (b200_plugin_init())\
        device = b200_plugin_create(info, stats, profiler, background);\
      else\
        device = NULL;\
      break;

/^  return DEVICE_NONE;$/i\
  else if (strcmp(name, "B200") == 0)\
    return DEVICE_B200;

/^  return "";$/i\
  else if (type == DEVICE_B200)\
    return "B200";

/^  return types;$/i\
  if (b200_plugin_init())\
    types.push_back(DEVICE_B200);

/^  return devices;$/i\
  if (mask & DEVICE_MASK_B200) {\
    if (!(devices_initialized_mask & DEVICE_MASK_B200)) {\
      if (b200_plugin_init()) {\
        b200_plugin_info(b200_devices);\
      }\
      devices_initialized_mask |= DEVICE_MASK_B200;\
    }\
    foreach (DeviceInfo &info, b200_devices) {\
      devices.push_back(info);\
    }\
  }

/^  cpu_devices.free_memory();$/a\
  b200_devices.free_memory();

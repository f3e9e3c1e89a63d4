/* Shared between ref_probe.cpp (kernel side) and ref_harness.cpp (host side).
 * Plain-data batch records; the same layouts are used by include/b200_cycles.h
 * (b200_ray / b200_hit) so a dumped batch feeds both sides unchanged.
 * TEST INFRASTRUCTURE ONLY. */
#ifndef REF_PROBE_H
#define REF_PROBE_H

#include <stddef.h>
#include <stdint.h>

struct RefProbeRay {
  float P[3];
  float t;
  float D[3];
  uint32_t visibility;
};

struct RefProbeHit {
  float t, u, v;
  int32_t prim;
  int32_t object;
  int32_t type;
};

/* One shading point for ref_probe_svm_node (same layout as HostShadingPoint of
 * tests/host_check/svm_tex_host.cpp). */
struct RefShadingPoint {
  float P[3], N[3], I[3], dPdu[3];
  float u, v;
  int32_t object, prim, lamp, shader, backfacing;
};

namespace ccl {
struct KernelGlobals;
int ref_probe_svm_closure(KernelGlobals *kg, const void *nodes, int offset, float *stack,
                          const RefShadingPoint *p, const float *closure_weight,
                          unsigned int path_flag, const float *omega_in, float randu,
                          float randv, float *out);
int ref_probe_svm_node(KernelGlobals *kg, const void *nodes, int offset, float *stack,
                       const RefShadingPoint *p);
void ref_probe_intersect(KernelGlobals *kg, const RefProbeRay *rays, RefProbeHit *hits, size_t n);
void ref_probe_camera_rays(KernelGlobals *kg, int sample, int x0, int y0, int w, int h,
                           RefProbeRay *rays, unsigned int *rng_hash);
void ref_probe_shadow_rays(KernelGlobals *kg, int sample, int x0, int y0, int w, int h,
                           RefProbeRay *rays);
void ref_probe_count_rays(KernelGlobals *kg, int sample, int x0, int y0, int w, int h,
                          unsigned long long counts[3]);
void ref_probe_path_dump(KernelGlobals *kg, int sample, int x, int y, float *out);
void ref_probe_path_trace(KernelGlobals *kg, float *buffer, int sample, int x, int y, int offset,
                          int stride);
}  // namespace ccl

#endif

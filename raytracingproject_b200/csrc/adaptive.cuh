/* adaptive.cuh - adaptive sampling: pixels stop drawing samples once their error estimate
 * is under the threshold.
 *
 * Semantics to match (reference = blender/intern/cycles):
 *   kernel/kernel_adaptive_sampling.h:24-44    kernel_do_adaptive_stopping - the estimate of
 *                                              Dammertz et al.: |all samples - 2 x the even
 *                                              half| against the pixel's brightness
 *   kernel/kernel_adaptive_sampling.h:47-180   kernel_adaptive_post_adjust
 *   kernel/kernel_adaptive_sampling.h:185-245  the 1-pixel dilation in x then y: the
 *                                              neighbours of an unconverged pixel go on too
 *   kernel/kernel_passes.h:392-425             the aux buffer (even samples x 2) and the
 *                                              sample-count pass, negative while in progress
 *   kernel/kernel_path.h:660-666               a converged pixel traces nothing
 *   device/device_cpu.cpp:838-886, 905-945     when the filter runs (every adaptive_step
 *                                              samples past min_samples) and the final
 *                                              rescale to a uniform sample count
 *   device/cuda/device_cuda_impl.cpp:1817-1851 the same as four kernels of a GPU device
 * The wavefront renders `adaptive_step` samples of the whole tile per batch, so the filter
 * runs where the CPU device runs it; init_from_camera leaves converged pixels out of the
 * queue.  Film layout: combined float4 at 0, aux float4 and the sample count where
 * KernelFilm says. */
#ifndef B200_ADAPTIVE_CUH
#define B200_ADAPTIVE_CUH

struct AdaptiveTile {
  float *film;
  int x, y, w, h, offset, stride, pass_stride;
  int aux, sample_count; /* float offsets inside a pixel */
};

CY_DEV float *adaptive_pixel(const AdaptiveTile &t, int x, int y)
{
  return t.film + ((long long)t.offset + x + (long long)y * t.stride) * t.pass_stride;
}

/* kernel_random.h:294-320 */
CY_DEV bool sample_is_even(int pattern, int sample)
{
  if (pattern == CY_SAMPLING_PATTERN_PMJ)
    return (__popc((unsigned int)sample & 0xaaaaaaaau) & 1) != 0;
  return (sample & 1) != 0;
}

__global__ void __launch_bounds__(WF_BLOCK) k_adaptive_stopping(AdaptiveTile t, int sample)
{
  const int n = t.w * t.h;
  const float threshold = kd_float(KD_INT_ADAPTIVE_THRESHOLD);
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    float *buffer = adaptive_pixel(t, t.x + k % t.w, t.y + k / t.w);
    const float4 I = *(const float4 *)buffer;
    const float4 A = *(const float4 *)(buffer + t.aux);
    const float error = (fabsf(I.x - A.x) + fabsf(I.y - A.y) + fabsf(I.z - A.z)) /
                        (sample * 0.0001f + sqrtf(I.x + I.y + I.z));
    if (error < threshold * (float)sample)
      buffer[t.aux + 3] += 1.0f;
  }
}

/* one thread per row / per column, a serial sweep like the reference's: the dilation reads
 * what it wrote one pixel earlier */
__global__ void k_adaptive_filter_x(AdaptiveTile t, unsigned int *any)
{
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= t.h)
    return;
  bool found = false, prev = false;
  for (int x = t.x; x < t.x + t.w; ++x) {
    float *aux = adaptive_pixel(t, x, t.y + row) + t.aux;
    if (aux[3] == 0.0f) {
      found = true;
      if (x > t.x && !prev)
        (adaptive_pixel(t, x - 1, t.y + row) + t.aux)[3] = 0.0f;
      prev = true;
    }
    else {
      if (prev)
        aux[3] = 0.0f;
      prev = false;
    }
  }
  if (found)
    *any = 1u;
}

__global__ void k_adaptive_filter_y(AdaptiveTile t, unsigned int *any)
{
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= t.w)
    return;
  bool found = false, prev = false;
  for (int y = t.y; y < t.y + t.h; ++y) {
    float *aux = adaptive_pixel(t, t.x + col, y) + t.aux;
    if (aux[3] == 0.0f) {
      found = true;
      if (y > t.y && !prev)
        (adaptive_pixel(t, t.x + col, y - 1) + t.aux)[3] = 0.0f;
      prev = true;
    }
    else {
      if (prev)
        aux[3] = 0.0f;
      prev = false;
    }
  }
  if (found)
    *any = 1u;
}

/* adaptive_sampling_post: every pixel as if it had drawn `sample` samples */
__global__ void __launch_bounds__(WF_BLOCK)
    k_adaptive_scale_samples(AdaptiveTile t, int start_sample, int sample)
{
  const int n = t.w * t.h;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    float *buffer = adaptive_pixel(t, t.x + k % t.w, t.y + k / t.w);
    float multiplier;
    if (buffer[t.sample_count] < 0.0f) {
      buffer[t.sample_count] = -buffer[t.sample_count];
      multiplier = (float)sample / fmaxf((float)start_sample + 1.0f, buffer[t.sample_count]);
      if (multiplier == 1.0f)
        continue;
    }
    else {
      multiplier = (float)sample / ((float)sample - 1.0f);
    }
    /* kernel_adaptive_post_adjust for the passes in scope: combined and the aux buffer */
    float4 *c = (float4 *)buffer, *a = (float4 *)(buffer + t.aux);
    float4 cv = *c, av = *a;
    cv.x *= multiplier, cv.y *= multiplier, cv.z *= multiplier, cv.w *= multiplier;
    av.x *= multiplier, av.y *= multiplier, av.z *= multiplier, av.w *= multiplier;
    *c = cv;
    *a = av;
  }
}

#endif /* B200_ADAPTIVE_CUH */

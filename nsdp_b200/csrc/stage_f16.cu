#include "stage_f16.cuh"

#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

#ifndef NSDP_STAGE_F16_DEFAULT
#define NSDP_STAGE_F16_DEFAULT true
#endif

namespace nsdp {
namespace stage16 {

__global__ void absmax_sample_kernel(const float *__restrict__ g, size_t n4, size_t step, unsigned *__restrict__ slot) {
  float m = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i * step < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4 *>(g) + i * step);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(slot, __float_as_uint(m));
}

// process-wide staging format: -1 = not decided yet (first use reads NSDP_STAGE_FMT), 0 = bf16 hi + lo, 1 = fp16
static std::atomic<int> g_fmt{-1};

static int current_fmt() {
  int f = g_fmt.load(std::memory_order_relaxed);
  if (f < 0) {
    const char *e = getenv("NSDP_STAGE_FMT");
    f = e ? (strcmp(e, "fp16") == 0 ? 1 : 0) : (NSDP_STAGE_F16_DEFAULT ? 1 : 0);
    g_fmt.store(f, std::memory_order_relaxed);
  }
  return f;
}

bool enabled() { return current_fmt() == 1; }

int set_format(int fmt) {
  const int prev = current_fmt();
  if (fmt == 0 || fmt == 1) g_fmt.store(fmt, std::memory_order_relaxed);
  return prev;
}

// max|g| over every `step`-th float4 of g (about 2^18 samples), into *slot (zeroed by the caller)
int launch_absmax(const float *g, size_t n, unsigned *slot, cudaStream_t st) {
  const size_t n4 = n / 4;
  if (n4 == 0) return NSDP_OK;
  size_t step = n4 >> 18;
  if (step < 1) step = 1;
  step |= 1;                                    // odd stride: the samples walk through all columns of a row
  const size_t samples = (n4 + step - 1) / step;
  const unsigned blocks = (unsigned)((samples + 255) / 256 < (size_t)(2 * num_sms()) ? (samples + 255) / 256 : 2 * num_sms());
  absmax_sample_kernel<<<blocks, 256, 0, st>>>(g, n4, step, slot);
  return check_launch();
}

}  // namespace stage16
}  // namespace nsdp

extern "C" int nsdp_set_stage_format(int fmt) { return nsdp::stage16::set_format(fmt); }

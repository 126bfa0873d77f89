// Shared helpers for the nsdp_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nsdp_b200.h"

namespace nsdp {

extern thread_local int g_last_cuda_error;

inline int check_launch() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    g_last_cuda_error = (int)e;
    return NSDP_ERR_CUDA;
  }
  return NSDP_OK;
}

inline int cuda_rc(cudaError_t e) {
  if (e != cudaSuccess) {
    g_last_cuda_error = (int)e;
    (void)cudaGetLastError();
    return NSDP_ERR_CUDA;
  }
  return NSDP_OK;
}

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) {
  return (a + b - 1) / b;
}

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

}  // namespace nsdp

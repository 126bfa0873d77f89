// Microbenchmark: how fast can 148 persistent CTAs stream a large buffer HBM -> shared memory
//   mode 0: one thread issues cp.async.bulk (1-D TMA) into an S-stage ring of CHUNK-byte slots (dw_tc's scheme)
//   mode 1: W warps issue cp.async (LDGSTS, 16 B/thread) into the same ring
// usage: stream_bw <mode> <stages> <chunk_bytes> <warps>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../nsdp_b200/csrc/umma.cuh"
using namespace nsdp::umma;

__global__ void __launch_bounds__(288, 1) bulk_kernel(const unsigned char *src, size_t per_cta, int stages, uint32_t chunk, unsigned long long *sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[16], empty[16];
  if (threadIdx.x == 0) { for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); } mbar_fence_init(); }
  __syncthreads();
  const unsigned char *p = src + (size_t)blockIdx.x * per_cta;
  const size_t n = per_cta / chunk;
  if (threadIdx.x == 0) {
    for (size_t it = 0; it < n; ++it) {
      const int s = it % stages; const uint32_t ph = (it / stages) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      mbar_arrive_expect_tx(&full[s], chunk);
      bulk_g2s(smem + (size_t)s * chunk, p + it * chunk, chunk, &full[s]);
    }
  } else if (threadIdx.x == 32) {
    unsigned long long acc = 0;
    for (size_t it = 0; it < n; ++it) {
      const int s = it % stages; const uint32_t ph = (it / stages) & 1;
      mbar_wait(&full[s], ph);
      acc += *reinterpret_cast<volatile uint32_t *>(smem + (size_t)s * chunk);
      mbar_arrive(&empty[s]);
    }
    sink[blockIdx.x] = acc;
  }
}

__global__ void __launch_bounds__(288, 1) ldgsts_kernel(const unsigned char *src, size_t per_cta, int stages, uint32_t chunk, int warps, unsigned long long *sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  const unsigned char *p = src + (size_t)blockIdx.x * per_cta;
  const size_t n = per_cta / chunk;
  const int nthr = warps * 32;
  if ((int)threadIdx.x >= nthr) return;
  unsigned long long acc = 0;
  // classic multistage cp.async pipeline: all `warps` warps copy, commit groups, wait for the oldest
  auto issue = [&](size_t it) {
    const int s = it % stages;
    for (uint32_t off = threadIdx.x * 16; off < chunk; off += nthr * 16) {
      const uint32_t d = smem_u32(smem + (size_t)s * chunk + off);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(p + it * chunk + off) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int i = 0; i < stages - 1; ++i) issue(i);
  for (size_t it = 0; it < n; ++it) {
    if (it + stages - 1 < n) issue(it + stages - 1); else asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(7) : "memory");   // stages - 1 must be <= 7 + 1
    asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
    acc += *reinterpret_cast<volatile uint32_t *>(smem + (size_t)(it % stages) * chunk + (threadIdx.x & 31) * 4);
    asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
  }
  if (threadIdx.x == 0) sink[blockIdx.x] = acc;
}

int main(int argc, char **argv) {
  const int mode = atoi(argv[1]), stages = atoi(argv[2]); const uint32_t chunk = atoi(argv[3]); const int warps = atoi(argv[4]);
  const int ctas = argc > 5 ? atoi(argv[5]) : 148;
  const size_t per_cta = ((size_t)16 << 20) / chunk * chunk;   // 16 MB per CTA -> 2.4 GB total, >> L2
  unsigned char *src; unsigned long long *sink;
  cudaMalloc(&src, per_cta * ctas); cudaMemset(src, 1, per_cta * ctas); cudaMalloc(&sink, 8 * ctas);
  const size_t smem = (size_t)stages * chunk;
  cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(ldgsts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(a);
    if (mode == 0) bulk_kernel<<<ctas, 288, smem>>>(src, per_cta, stages, chunk, sink);
    else ldgsts_kernel<<<ctas, 288, smem>>>(src, per_cta, stages, chunk, warps, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  printf("mode %d stages %d chunk %u warps %d ctas %d: %.3f ms  %.1f GB/s  (%s)\n", mode, stages, chunk, warps, ctas, best,
         per_cta * ctas / best / 1e6, cudaGetErrorString(e));
  return 0;
}

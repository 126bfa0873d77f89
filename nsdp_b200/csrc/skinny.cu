// Weight / bias gradient of a Linear layer with a NARROW input on many rows:
//     dW[n][k] += sum_r dy[r][n] * x[r][k],   db[n] += sum_r dy[r][n]          (K <= 8 input channels)
// The encoder's first layers are such products (enc_sdf: 4 -> 120 on B * 4096 rows, and the q / k / v projections folded
// through it, nsdp_b200/model/encoder/blocks.py _fused_linear_through): cuBLAS runs the [N x R] x [R x K] GEMM at ~0.1 ms
// each (32 768 rows, K = 4: 94 MFLOP), bound by its tile shape, while the job is one coalesced pass over dy (47 MB).
// One thread per output column n and row range, K + 1 accumulators in registers, atomics at the end. Plain fp32.
#include "common.cuh"

namespace nsdp {

constexpr int SK_MAXK = 8;
constexpr int SK_THREADS = 128;

template <int K>
__global__ void __launch_bounds__(SK_THREADS) skinny_dw_kernel(const float *__restrict__ x, const float *__restrict__ dy,
                                                               long long R, int N, float *__restrict__ dW,
                                                               float *__restrict__ db) {
  const int n = blockIdx.x * SK_THREADS + threadIdx.x;
  const long long r0 = R * blockIdx.y / gridDim.y, r1 = R * (blockIdx.y + 1) / gridDim.y;
  float acc[K], sb = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.f;
  if (n < N) {
#pragma unroll 4
    for (long long r = r0; r < r1; ++r) {
      const float d = dy[r * N + n];
      const float *xr = x + r * K;       // the same K values for every thread of the block: served from L1
      sb += d;
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] = fmaf(d, __ldg(xr + k), acc[k]);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) atomicAdd(dW + (size_t)n * K + k, acc[k]);
    if (db) atomicAdd(db + n, sb);
  }
}

}  // namespace nsdp

extern "C" int nsdp_linear_narrow_dw_f32(const float *x, const float *dy, long long R, int K, int N, float *dW, float *db,
                                         void *stream) {
  using namespace nsdp;
  if (!x || !dy || !dW || R < 0 || K < 1 || K > SK_MAXK || N < 1) return NSDP_ERR_INVALID_ARGUMENT;
  if (R == 0) return NSDP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int cb = (N + SK_THREADS - 1) / SK_THREADS;
  long long splits = (4ll * num_sms() + cb - 1) / cb;          // ~4 blocks per SM
  if (splits > (R + 63) / 64) splits = (R + 63) / 64;           // at least 64 rows per block
  if (splits < 1) splits = 1;
  dim3 grid((unsigned)cb, (unsigned)splits);
  switch (K) {
    case 1: skinny_dw_kernel<1><<<grid, SK_THREADS, 0, st>>>(x, dy, R, N, dW, db); break;
    case 2: skinny_dw_kernel<2><<<grid, SK_THREADS, 0, st>>>(x, dy, R, N, dW, db); break;
    case 3: skinny_dw_kernel<3><<<grid, SK_THREADS, 0, st>>>(x, dy, R, N, dW, db); break;
    case 4: skinny_dw_kernel<4><<<grid, SK_THREADS, 0, st>>>(x, dy, R, N, dW, db); break;
    case 5: skinny_dw_kernel<5><<<grid, SK_THREADS, 0, st>>>(x, dy, R, N, dW, db); break;
    case 6: skinny_dw_kernel<6><<<grid, SK_THREADS, 0, st>>>(x, dy, R, N, dW, db); break;
    case 7: skinny_dw_kernel<7><<<grid, SK_THREADS, 0, st>>>(x, dy, R, N, dW, db); break;
    default: skinny_dw_kernel<8><<<grid, SK_THREADS, 0, st>>>(x, dy, R, N, dW, db); break;
  }
  return check_launch();
}

// Hardware self-test of the tcgen05 conventions in umma.cuh: D[128 x N] = A[128 x K] * B[N x K]^T on the tensor
// cores (optionally with the bf16x3 split), operands written by ordinary threads in the canonical no-swizzle
// K-major layout. tests/test_gpu_umma.py compares it with torch; every tensor-core kernel in the library relies on
// exactly these descriptor / layout / TMEM-lane conventions.
#include "common.cuh"
#include "umma.cuh"

namespace nsdp {

__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ D, int N, int K,
                     int split, int mn, int *err) {
  using namespace umma;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_bytes = 128u * K * 2u, b_bytes = (uint32_t)N * K * 2u;
  unsigned char *a_hi = smem, *a_lo = a_hi + a_bytes, *b_hi = a_lo + a_bytes, *b_lo = b_hi + b_bytes;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;

  if (!mn) {
    for (int e = tid; e < 128 * (K / 2); e += 128) {
      const int row = e / (K / 2), k = (e % (K / 2)) * 2;
      uint32_t hi, lo;
      split2(A[row * K + k], A[row * K + k + 1], hi, lo);
      *reinterpret_cast<uint32_t *>(a_hi + canon_off(128, row, k)) = hi;
      *reinterpret_cast<uint32_t *>(a_lo + canon_off(128, row, k)) = lo;
    }
    for (int e = tid; e < N * (K / 2); e += 128) {
      const int row = e / (K / 2), k = (e % (K / 2)) * 2;
      uint32_t hi, lo;
      split2(B[row * K + k], B[row * K + k + 1], hi, lo);
      *reinterpret_cast<uint32_t *>(b_hi + canon_off(N, row, k)) = hi;
      *reinterpret_cast<uint32_t *>(b_lo + canon_off(N, row, k)) = lo;
    }
  } else {
    // operands given as X (K x 128) and Y (K x N): tiles with rows = the reduction index, written exactly like the
    // K-major kernels write an activation tile, consumed as MN-major operands
    for (int e = tid; e < K * 64; e += 128) {
      const int row = e / 64, m = (e % 64) * 2;
      uint32_t hi, lo;
      split2(A[row * 128 + m], A[row * 128 + m + 1], hi, lo);
      *reinterpret_cast<uint32_t *>(a_hi + canon_off(K, row, m)) = hi;
      *reinterpret_cast<uint32_t *>(a_lo + canon_off(K, row, m)) = lo;
    }
    for (int e = tid; e < K * (N / 2); e += 128) {
      const int row = e / (N / 2), n = (e % (N / 2)) * 2;
      uint32_t hi, lo;
      split2(B[row * N + n], B[row * N + n + 1], hi, lo);
      *reinterpret_cast<uint32_t *>(b_hi + canon_off(K, row, n)) = hi;
      *reinterpret_cast<uint32_t *>(b_lo + canon_off(K, row, n)) = lo;
    }
  }
  uint32_t ncols = 32;
  while ((int)ncols < N) ncols *= 2;
  if (warp == 0) tmem_alloc(&tmem_base_s, ncols);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (tid == 0) {
    const uint32_t idesc = mn ? idesc_bf16_mn(128, N) : idesc_bf16(128, N);
    const uint32_t lbo_a = 128 * 16, lbo_b = (uint32_t)N * 16;
    const int passes = split ? 3 : 1;
    bool acc = false;
    for (int p = 0; p < passes; ++p) {
      const unsigned char *pa = (p == 1) ? a_lo : a_hi;  // hi*hi, lo*hi, hi*lo
      const unsigned char *pb = (p == 2) ? b_lo : b_hi;
      for (int ks = 0; ks < K / 16; ++ks) {
        uint64_t ad, bd;
        if (!mn) {
          ad = smem_desc(smem_u32(pa) + ks * 2 * lbo_a, lbo_a, 128);
          bd = smem_desc(smem_u32(pb) + ks * 2 * lbo_b, lbo_b, 128);
        } else {  // k-step = 16 tile rows = two 128-byte core-matrix rows; column groups are K*16 bytes apart
          ad = smem_desc(smem_u32(pa) + ks * 256, 128, (uint32_t)K * 16);
          bd = smem_desc(smem_u32(pb) + ks * 256, 128, (uint32_t)K * 16);
        }
        mma_bf16(tmem_base, ad, bd, idesc, acc);
        acc = true;
      }
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0, err);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) D[row * N + c0 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

}  // namespace nsdp

extern "C" int nsdp_selftest_umma(const float *A, const float *B, float *D, int N, int K, int split, int mn, int *err,
                                  void *stream) {
  using namespace nsdp;
  if (!A || !B || !D || !err || N < 16 || N > 256 || N % 16 || K < 16 || K % 16) return NSDP_ERR_INVALID_ARGUMENT;
  const size_t smem = 2 * (size_t)(128 + N) * K * 2;
  if (smem > 220 * 1024) return NSDP_ERR_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_rc(e);
  umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, N, K, split, mn, err);
  return check_launch();
}

// ---------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): D[256 x N] = A[256 x K] * B[N x K]^T by ONE tcgen05.mma stream issued from the
// leader CTA of a 2-CTA cluster. CTA c holds rows [128 c, 128 c + 128) of A and rows [N/2 c, N/2 c + N/2) of B in its own
// shared memory (same offsets in both CTAs); its TMEM receives rows [128 c, +128) of D, all N columns.
// ---------------------------------------------------------------------------------------------------------------------
namespace nsdp {

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
umma2_selftest_kernel(const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ D, int N, int K,
                      int split, int *err) {
  using namespace umma;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t cta = cluster_ctarank();
  const int NH = N / 2;
  const uint32_t a_bytes = 128u * K * 2u, b_bytes = (uint32_t)NH * K * 2u;
  unsigned char *a_hi = smem, *a_lo = a_hi + a_bytes, *b_hi = a_lo + a_bytes, *b_lo = b_hi + b_bytes;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;

  const float *Ac = A + (size_t)cta * 128 * K;
  const float *Bc = B + (size_t)cta * NH * K;
  for (int e = tid; e < 128 * (K / 2); e += 128) {
    const int row = e / (K / 2), k = (e % (K / 2)) * 2;
    uint32_t hi, lo;
    split2(Ac[row * K + k], Ac[row * K + k + 1], hi, lo);
    *reinterpret_cast<uint32_t *>(a_hi + canon_off(128, row, k)) = hi;
    *reinterpret_cast<uint32_t *>(a_lo + canon_off(128, row, k)) = lo;
  }
  for (int e = tid; e < NH * (K / 2); e += 128) {
    const int row = e / (K / 2), k = (e % (K / 2)) * 2;
    uint32_t hi, lo;
    split2(Bc[row * K + k], Bc[row * K + k + 1], hi, lo);
    *reinterpret_cast<uint32_t *>(b_hi + canon_off(NH, row, k)) = hi;
    *reinterpret_cast<uint32_t *>(b_lo + canon_off(NH, row, k)) = lo;
  }
  uint32_t ncols = 32;
  while ((int)ncols < N) ncols *= 2;
  if (warp == 0) tmem_alloc_pair(&tmem_base_s, ncols);
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  fence_async_smem();
  tc_fence_before();
  cluster_sync_all();          // both CTAs' operands and barriers are in place
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (cta == 0 && tid == 0) {
    const uint32_t idesc = idesc_bf16(256, N);
    const uint32_t lbo_a = 128 * 16, lbo_b = (uint32_t)NH * 16;
    const int passes = split ? 3 : 1;
    uint32_t acc = 0;
    for (int p = 0; p < passes; ++p) {
      const unsigned char *pa = (p == 1) ? a_lo : a_hi;
      const unsigned char *pb = (p == 2) ? b_lo : b_hi;
      for (int ks = 0; ks < K / 16; ++ks) {
        const uint64_t ad = smem_desc(smem_u32(pa) + ks * 2 * lbo_a, lbo_a, 128);
        const uint64_t bd = smem_desc(smem_u32(pb) + ks * 2 * lbo_b, lbo_b, 128);
        mma_bf16_pair(tmem_base, ad, bd, idesc, acc != 0);
        acc = 1;
      }
    }
    mma_commit_pair(&bar);
  }
  mbar_wait(&bar, 0, err);
  tc_fence_after();
  const int row = (int)cta * 128 + warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) D[(size_t)row * N + c0 + i] = v[i];
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair(tmem_base, ncols);
}

}  // namespace nsdp

extern "C" int nsdp_selftest_umma2(const float *A, const float *B, float *D, int N, int K, int split, int *err, void *stream) {
  using namespace nsdp;
  if (!A || !B || !D || !err || N < 32 || N > 256 || N % 16 || K < 16 || K % 16) return NSDP_ERR_INVALID_ARGUMENT;
  const size_t smem = 2 * (size_t)(128 + N / 2) * K * 2;
  if (smem > 220 * 1024) return NSDP_ERR_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(umma2_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_rc(e);
  umma2_selftest_kernel<<<2, 128, smem, (cudaStream_t)stream>>>(A, B, D, N, K, split, err);
  return check_launch();
}

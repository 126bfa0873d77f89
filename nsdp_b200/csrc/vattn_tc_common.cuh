// Pieces shared by the tensor-core vector-attention kernels (forward: vattn_tc.cu, backward: vattn_bwd_tc.cu).
#pragma once
#include <math.h>

#include "umma.cuh"
#include "vattn_common.cuh"

namespace nsdp {
namespace vtc {

using namespace umma;

template <int DP_, int KR_, int NPART_ = 4>
struct TcCfg {
  static constexpr int DP = DP_;                 // padded channel count: K and N of every GEMM
  static constexpr int KR = KR_;                 // rows per centre inside a tile (power of two, 8..32)
  static constexpr int KSTEPS = DP / 16;
  static constexpr int SLAB = DP * 16 * 2;       // one [DP x 16] bf16 K-major weight slab
  static constexpr int STAGE_BYTES = 4 * SLAB;   // GEMM1 stage: W' hi, W' lo, Wd2 hi, Wd2 lo (GEMM2 uses half)
  static constexpr int STAGES = DP > 208 ? 2 : 4;
  static constexpr int A_HALF = 128 * DP * 2;    // bytes of the hi (or lo) A operand
  static constexpr int NPART = NPART_;           // worker warps per TMEM lane quarter (they split the columns)
  static constexpr int WORKER_WARPS = 4 * NPART;
  static constexpr int THREADS = (2 + WORKER_WARPS) * 32;
  static constexpr int CHUNKS = DP / 8;          // 8-column chunks per row
  static constexpr int MAXCH = (CHUNKS + NPART - 1) / NPART;  // chunks per worker thread (upper bound)
  static constexpr uint32_t TMEM_COLS = 512;
  static constexpr uint32_t ACC1_COL = 256;
  static constexpr int CENTRES = 128 / KR;
  // dynamic shared memory carve-up (bytes)
  static constexpr int OFF_A = 0;
  static constexpr int OFF_STAGE = OFF_A + 2 * A_HALF;
  static constexpr int OFF_WD0 = OFF_STAGE + STAGES * STAGE_BYTES;   // float4[DP]
  static constexpr int OFF_PC = OFF_WD0 + DP * 16;                   // float[DP]
  static constexpr int OFF_VC = OFF_PC + DP * 4;                     // float[DP]
  static constexpr int OFF_RED = OFF_VC + DP * 4;                    // float[3][4][DP] cross-warp softmax (KR == 128)
  static constexpr int OFF_BAR = OFF_RED + (KR_ == 128 ? 3 * 4 * DP * 4 : 0);   // mbarriers
  static constexpr int SMEM = OFF_BAR + 256;
  static_assert(DP % 16 == 0 && DP <= 256, "unsupported padded width");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

struct RowInfo {
  int c;       // flattened centre, -1 = inactive
  int n;       // flattened source row, or -(b+1) for the global row
  float rx, ry, rz, flag;
};

template <class C>
__device__ __forceinline__ RowInfo row_info(const nsdp_vattn_args &a, long long tile, int r, int krows) {
  RowInfo ri;
  ri.c = -1; ri.n = 0; ri.rx = ri.ry = ri.rz = 0.f; ri.flag = 0.f;
  const int p = r / C::KR, t = r - p * C::KR;
  const long long ci = tile * C::CENTRES + p;
  if (ci < (long long)a.B * a.M && t < krows) {
    const int b = (int)(ci / a.M);
    ri.c = (int)ci;
    if (t < a.K) {
      const int j = a.idx ? a.idx[ci * a.K + t] : t;
      ri.n = b * a.N + j;
      const float *xc = a.xyz_c + ci * 3;
      const float *xn = a.xyz_n + (size_t)ri.n * 3;
      ri.rx = a.sign * (xc[0] - xn[0]);
      ri.ry = a.sign * (xc[1] - xn[1]);
      ri.rz = a.sign * (xc[2] - xn[2]);
      ri.flag = 1.f;
    } else {
      ri.n = -(b + 1);
    }
  }
  return ri;
}

// sum of v[0..8) over the G lanes of a group; afterwards lane j (j = lane % G < 8) holds the total of v[j] in v[0].
template <int G>
__device__ __forceinline__ float group_transpose_sum(float (&v)[8], int lane) {
  const unsigned full = 0xffffffffu;
#pragma unroll
  for (int off = G / 2; off >= 8; off >>= 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += __shfl_xor_sync(full, v[i], off);
  }
  {
    const bool up = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = up ? v[i] : v[i + 4];
      const float keep = up ? v[i + 4] : v[i];
      v[i] = keep + __shfl_xor_sync(full, send, 4);
    }
  }
  {
    const bool up = lane & 2;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = up ? v[i] : v[i + 2];
      const float keep = up ? v[i + 2] : v[i];
      v[i] = keep + __shfl_xor_sync(full, send, 2);
    }
  }
  {
    const bool up = lane & 1;
    const float send = up ? v[0] : v[1];
    const float keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(full, send, 1);
  }
  return v[0];
}

// max variant of group_transpose_sum: lane j (j = lane % G < 8) ends with the group maximum of v[j] in v[0]
template <int G>
__device__ __forceinline__ float group_transpose_max(float (&v)[8], int lane) {
  const unsigned full = 0xffffffffu;
#pragma unroll
  for (int off = G / 2; off >= 8; off >>= 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], __shfl_xor_sync(full, v[i], off));
  }
  {
    const bool up = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = up ? v[i] : v[i + 4];
      const float keep = up ? v[i + 4] : v[i];
      v[i] = fmaxf(keep, __shfl_xor_sync(full, send, 4));
    }
  }
  {
    const bool up = lane & 2;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = up ? v[i] : v[i + 2];
      const float keep = up ? v[i + 2] : v[i];
      v[i] = fmaxf(keep, __shfl_xor_sync(full, send, 2));
    }
  }
  {
    const bool up = lane & 1;
    const float send = up ? v[0] : v[1];
    const float keep = up ? v[1] : v[0];
    v[0] = fmaxf(keep, __shfl_xor_sync(full, send, 1));
  }
  return v[0];
}

__device__ __forceinline__ int float_order_key(float x) {
  const int i = __float_as_int(x);
  return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float float_from_key(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }


}  // namespace vtc
}  // namespace nsdp

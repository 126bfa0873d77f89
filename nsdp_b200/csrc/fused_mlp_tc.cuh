// Shared configuration and helpers of the fused-MLP tensor-core kernels (fused_mlp_tc.cu forward, fused_mlp_bwd_tc.cu
// backward chain): tile geometry per width, operand-chunk store, TMEM load by width.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace nsdp {
namespace mtc {

using namespace umma;

constexpr int MAX_HIDDEN = 7;

template <int W_>
struct Cfg {
  static constexpr int W = W_;
  static constexpr int KS = W / 16;                    // k-steps (= weight stages) per layer
  // narrow widths are bound by the per-layer handoff latency: small CTAs (few worker warps), many of them per SM
  static constexpr int NWQ = W >= 128 ? 4 : (W >= 64 ? 2 : 1);   // worker warps per TMEM lane quarter
  static constexpr int CPT = (W >= 256 || W <= 32) ? 16 : 8;     // accumulator columns per thread per chunk
  static constexpr int CW = NWQ * CPT;                 // chunk width (columns)
  static constexpr int NCH = W / CW;                   // chunks per layer
  static constexpr int KPC = CW / 16;                  // k-steps per chunk
  static constexpr int WORKERS = 4 * NWQ;              // worker warps
  static constexpr int THREADS = (2 + WORKERS) * 32;   // + weight producer warp + MMA issuing warp
  static constexpr int MIN_CTAS = W >= 256 ? 1 : (W >= 128 ? 2 : (W >= 64 ? 4 : 6));   // co-resident CTAs hide the handoff
  static constexpr int SLAB = W * 16 * 2;              // [W x 16] bf16
  static constexpr int STAGE_BYTES = 2 * SLAB;         // hi + lo
  static constexpr int STAGES = W >= 256 ? 5 : (W >= 64 ? 4 : 8);    // W = 128: two CTAs of 108 KB share an SM; W = 64: four of 53 KB
  static constexpr uint32_t TMEM_COLS = 2 * W < 32 ? 32 : 2 * W;
  static constexpr int A_HALF = 128 * W * 2;
  static constexpr int OFF_X = 0;
  static constexpr int OFF_STAGE = OFF_X + 2 * A_HALF;
  static constexpr int OFF_BIAS = OFF_STAGE + STAGES * STAGE_BYTES;   // float[1 + MAX_HIDDEN][W]
  static constexpr int OFF_WO = OFF_BIAS + (1 + MAX_HIDDEN) * W * 4;  // float4[W]
  static constexpr int OFF_PART = OFF_WO + W * 16;                    // float4[NWQ - 1][128]
  static constexpr int OFF_BAR = OFF_PART + (NWQ - 1) * 128 * 16;
  static constexpr int SMEM = OFF_BAR + 256;
  static_assert(CW % 16 == 0 && NCH * CW == W && NCH <= 4, "chunking");
  static_assert((2 * STAGES + NCH + 1) * 8 + 4 <= 256, "barrier block");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

template <int N>
__device__ __forceinline__ void tmem_ldn(uint32_t taddr, float (&v)[N]);
template <>
__device__ __forceinline__ void tmem_ldn<8>(uint32_t taddr, float (&v)[8]) { tmem_ld8(taddr, v); }
template <>
__device__ __forceinline__ void tmem_ldn<16>(uint32_t taddr, float (&v)[16]) { tmem_ld16(taddr, v); }

// 8 fp32 values of row r, columns [k, k + 8) -> one 16-byte chunk of the hi and of the lo A operand
__device__ __forceinline__ void store_split8(unsigned char *X_hi, unsigned char *X_lo, int r, int k, const float *x) {
  uint4 hi, lo;
  split2(x[0], x[1], hi.x, lo.x);
  split2(x[2], x[3], hi.y, lo.y);
  split2(x[4], x[5], hi.z, lo.z);
  split2(x[6], x[7], hi.w, lo.w);
  const uint32_t off = canon_off(128, r, k);
  *reinterpret_cast<uint4 *>(X_hi + off) = hi;
  *reinterpret_cast<uint4 *>(X_lo + off) = lo;
}


}  // namespace mtc
}  // namespace nsdp

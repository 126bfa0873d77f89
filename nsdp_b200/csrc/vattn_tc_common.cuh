// Pieces shared by the tensor-core vector-attention kernels (forward: vattn_tc.cu, backward: vattn_bwd_tc.cu).
#pragma once
#include <math.h>

#include "umma.cuh"
#include "vattn_common.cuh"

namespace nsdp {
namespace vtc {

using namespace umma;

template <int DP_, int KR_, int NPART_ = 4, bool OH_ = false>
struct TcCfg {
  static constexpr int DP = DP_;                 // padded channel count: K and N of every GEMM
  static constexpr int KR = KR_;                 // rows per centre inside a tile (power of two, 8..32)
  static constexpr int KSTEPS = DP / 16;
  static constexpr int SLAB = DP * 16 * 2;       // one [DP x 16] bf16 K-major weight slab
  static constexpr int STAGE_BYTES = 4 * SLAB;   // GEMM1 stage: W' hi, W' lo, Wd2 hi, Wd2 lo (GEMM2 uses half)
  static constexpr int STAGES = DP > 208 ? 2 : 4;
  // the weight ring is addressed in SLOTS of one matrix k-step (hi slab + lo slab): GEMM1 takes two slots per k-step
  // (W', Wd2), every other GEMM one, so the single-matrix GEMMs prefetch twice as many k-steps ahead
  // OH ("one-hot") variant, decoder cross-attention: the per-row gathers of the K'/V anchor tables are folded into GEMM1
  // as E * [T1 | T2], E = one-hot(neighbour index) [128 x E_COLS] bf16 (exact), T = per-shape tables streamed through the
  // same slot ring; the epilogues then only add per-column constants. Needs N + 1 <= E_COLS (global token = row N).
  static constexpr bool OH = OH_;
  static constexpr int E_COLS = 112;
  static constexpr int E_KSTEPS = E_COLS / 16;
  static constexpr int E_BYTES = OH_ ? 128 * E_COLS * 2 : 0;
  static constexpr int SLOTS = OH_ ? 6 : 2 * STAGES;
  static constexpr int SLOT_BYTES = 2 * SLAB;
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
  static constexpr int OFF_E = OFF_A + 2 * A_HALF;
  static constexpr int OFF_STAGE = OFF_E + E_BYTES;
  static constexpr int OFF_WD0 = OFF_STAGE + SLOTS * SLOT_BYTES;     // float4[DP]
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

// Row description under PER-SHAPE tiling (OH kernels): tile t covers centres [(t % tpb) * CENTRES, +CENTRES) of shape
// b = t / tpb, so a tile never straddles two shapes and the producer can stream that shape's tables.
struct RowInfoPB {
  int c;       // flattened centre b*M + m, -1 = inactive
  int j;       // source index inside the shape (0..N-1), N for the global-token row, -1 = inactive
  float rx, ry, rz, flag;
};

template <class C>
__device__ __forceinline__ RowInfoPB row_info_pb(const nsdp_vattn_args &a, long long tile, int r, int krows, int tpb) {
  RowInfoPB ri;
  ri.c = -1; ri.j = -1; ri.rx = ri.ry = ri.rz = 0.f; ri.flag = 0.f;
  const int b = (int)(tile / tpb);
  const int m = (int)(tile - (long long)b * tpb) * C::CENTRES + r / C::KR;
  const int t = r % C::KR;
  if (b < a.B && m < a.M && t < krows) {
    const long long ci = (long long)b * a.M + m;
    ri.c = (int)ci;
    if (t < a.K) {
      const int j = a.idx ? a.idx[ci * a.K + t] : t;
      ri.j = j;
      const float *xc = a.xyz_c + ci * 3;
      const float *xn = a.xyz_n + ((size_t)b * a.N + j) * 3;
      ri.rx = a.sign * (xc[0] - xn[0]);
      ri.ry = a.sign * (xc[1] - xn[1]);
      ri.rz = a.sign * (xc[2] - xn[2]);
      ri.flag = 1.f;
    } else {
      ri.j = a.N;
    }
  }
  return ri;
}

// In-place 8 x 8 transpose inside every group of 8 consecutive lanes: before, lane l holds row l (v[i] = column i);
// after, lane l holds column l (v[i] = row i). 12 shuffles.
__device__ __forceinline__ void group8_transpose(float (&v)[8], int lane) {
  const unsigned full = 0xffffffffu;
#pragma unroll
  for (int s = 4; s >= 1; s >>= 1) {
    const bool up = lane & s;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i & s) continue;
      const float send = up ? v[i] : v[i | s];
      const float recv = __shfl_xor_sync(full, send, s);
      if (up) v[i] = recv; else v[i | s] = recv;
    }
  }
}

template <class C>
__device__ __forceinline__ RowInfo row_info(const nsdp_vattn_args &a, long long tile, int r, int krows) {
  RowInfo ri;
  ri.c = -1; ri.n = 0; ri.rx = ri.ry = ri.rz = 0.f; ri.flag = 0.f;
  const int p = r / C::KR, t = r - p * C::KR;
  const long long ci = tile * C::CENTRES + p;
  // neighbours sit in rows 0 .. K-1 of the centre's group, the global token ALWAYS in the last row (KR - 1): the
  // weight-gradient reduction picks the global rows out of the staged tiles by row % KR == KR - 1
  (void)krows;
  if (ci < (long long)a.B * a.M && (t < a.K || (a.has_global && t == C::KR - 1))) {
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

// ---- forward -> backward buffer of the OH kernels (nsdp_vattn_args::saved) ---------------------------------------------------------
// Four blocks of `tiles` x SAVED_TILE bytes: [H staged][G staged][a][acc1]. H and G are the operand tiles in the staged
// k-step-major bf16 hi/lo layout of dw_tc.cu (they feed the weight-gradient jobs directly, and G > 0 is the ReLU mask);
// a (GEMM2 result, pre-softmax) and acc1 (GEMM1's second accumulator, s = acc1 + vc) are fp32 in [chunk][row][8] order,
// i.e. byte(r, ch) = (ch * 128 + r) * 32: a warp reads / writes 1 KB contiguous.
template <class C>
constexpr size_t saved_tile_bytes() {
  return (size_t)512 * C::DP;
}
template <class C>
constexpr size_t saved_bytes_total(long long tiles) {
  return (size_t)tiles * 4 * saved_tile_bytes<C>();
}

// ---- per-shape anchor tables of the OH kernels (see vattn_fwd_oh_kernel), packed like the weights ------------------------
template <class C>
constexpr size_t table_bytes_per_shape() {
  return (size_t)C::E_KSTEPS * 2 * C::SLOT_BYTES;   // per k-step: [T1 hi][T1 lo][T2 hi][T2 lo]
}

// Byte offset of element (row n, k inside the k-step) of the hi (lo = 0) or lo (lo = 1) slab inside a two-slab slot.
// pair = 0: [hi slab][lo slab], each canonical over all DP rows. pair = 1 (CTA-pair kernels): the slot is cut by rows into
// the two halves the two CTAs fetch, [CTA 0: hi half slab, lo half slab][CTA 1: ...], each half canonical over DP / 2 rows.
template <class C>
__device__ __forceinline__ uint32_t slot_off(int n, int kin, int lo, int pair) {
  if (!pair) return (uint32_t)(lo * C::SLAB) + canon_off(C::DP, n, kin);
  constexpr int NH = C::DP / 2;
  const int c = n / NH, nn = n - c * NH;
  return (uint32_t)(c * (NH * 64) + lo * (NH * 32) + (kin >> 3) * (NH * 16) + nn * 16 + (kin & 7) * 2);
}

template <class C>
__global__ void pack_tables_kernel(const nsdp_vattn_args a, unsigned char *__restrict__ out, int pair = 0) {
  // one thread per (shape b, table m, column n, anchor pair k)
  const int per = C::DP * (C::E_COLS / 2);
  const long long total = (long long)a.B * 2 * per;
  const int D = a.D, N = a.N;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(e / (2 * per));
    const int rem0 = (int)(e - (long long)b * 2 * per);
    const int m = rem0 / per, rem = rem0 - m * per;
    const int n = rem / (C::E_COLS / 2), k = (rem - n * (C::E_COLS / 2)) * 2;
    float x[2] = {0.f, 0.f};
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = k + u;
      if (n < D) {
        if (j < N) {
          const size_t off = ((size_t)b * N + j) * D + n;
          x[u] = m == 0 ? -a.kp[off] : a.vp[off];
        } else if (j == N) {
          x[u] = m == 0 ? a.gq[(size_t)b * D + n] - a.pc[n] : a.gv[(size_t)b * D + n] - a.vc[n];
        }
      }
    }
    uint32_t hi, lo;
    split2(x[0], x[1], hi, lo);
    const int ks = k >> 4;
    unsigned char *base = out + (size_t)b * table_bytes_per_shape<C>() + (size_t)(ks * 2 + m) * C::SLOT_BYTES;
    *reinterpret_cast<uint32_t *>(base + slot_off<C>(n, k & 15, 0, pair)) = hi;
    *reinterpret_cast<uint32_t *>(base + slot_off<C>(n, k & 15, 1, pair)) = lo;
  }
}

__device__ __forceinline__ int float_order_key(float x) {
  const int i = __float_as_int(x);
  return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float float_from_key(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }


}  // namespace vtc
}  // namespace nsdp

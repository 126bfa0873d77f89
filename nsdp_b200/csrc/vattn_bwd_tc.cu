// Backward of the fused vector attention on tcgen05 tensor cores (sm_100a): the "chain" kernel.
//
// Same contract as the CUDA-core backward (vattn_bwd.cu): inputs + forward result + softmax statistics in, every
// gradient accumulated out; no [pairs, D] activation is ever KEPT. Per tile of 128 pair rows a persistent CTA runs
//
//   H -> [gp | dl] (GEMM1) -> G = relu(gp+P) -> a (GEMM2) -> w, s -> ds = w*dout, da = ds*(s-out)
//     -> dg = da*Wg2 (GEMM3) -> dgp = dg*[g>0] ;  dh = ds*Wd2 (GEMM4a) + dgp*W' (GEMM4b) -> dpre = dh*[h>0]
//
// with the same warp roles as the forward kernel (bulk-copy weight producer, single-thread MMA issuer, 16 worker
// warps that own one TMEM lane = one pair row each). Per-point gradients (d_vp, d_kp, d_qp, d_xyz) are scattered
// with vector fp32 reductions; d_wd0 / d_bd0 are accumulated per column in registers across all tiles of the CTA.
//
// The three d x d weight gradients need 3 x 208 x 208 fp32 of accumulator per CTA, which fits neither TMEM (next to
// the chain's own 416 columns) nor shared memory. The operand tiles H, G, dA, dGP, dS are therefore STAGED as bf16
// hi/lo in the k-step-major layout documented in dw_tc.cu and reduced by dw_tc_kernel:
//     d_wg2t += G^T dA,   d_wpt += H^T dGP,   d_wd2t += H^T dS.
// The op processes the rows in segments so that the staging workspace stays bounded.
#include <stdlib.h>
#include <string.h>

#include "dw_tc.cuh"
#include "stage_f16.cuh"
#include "vattn_tc_common.cuh"

// d_wd0 / d_bd0 of the decoder attention backward by the tensor core (1) or by the fp32 scratch transposition of round 1 (0)
#ifndef NSDP_DWD0_MMA
#define NSDP_DWD0_MMA 1
#endif

namespace nsdp {
namespace vtc {

template <class C>
struct BwdLayout {
  static constexpr int OFF_A = 0;
  static constexpr int OFF_E = OFF_A + 2 * C::A_HALF;
  static constexpr int OFF_STAGE = OFF_E + C::E_BYTES;
  static constexpr int OFF_WD0 = OFF_STAGE + C::SLOTS * C::SLOT_BYTES;    // float4[DP]
  static constexpr int OFF_PC = OFF_WD0 + C::DP * 16;                 // float[DP]
  static constexpr int OFF_VC = OFF_PC + C::DP * 4;                   // float[DP]
  static constexpr int OFF_RELS = OFF_VC + C::DP * 4;                 // float4[128]  rel of every row of the tile
  static constexpr int OFF_RELACC = OFF_RELS + 128 * 16;              // float[NPART][3][128]  d rel, one slot per column part
  static constexpr int OFF_RGV = OFF_RELACC + C::NPART * 3 * 128 * 4;
  static constexpr int OFF_BAR = OFF_RGV;
  static constexpr int SMEM = OFF_BAR + 256;
  // fp32 scratch [col][128] aliasing the A buffer (exactly 2 * A_HALF bytes for D == DP); element (col, row) sits at
  // col*128 + ((row + col) & 127): the rotation keeps both the row-wise writes and the column-wise reads conflict-free
  static constexpr int SCR_LD = 128;
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

// bytes of one staged [128 x DP] operand tile (hi + lo)
template <class C>
constexpr size_t staged_tile_bytes() {
  return (size_t)512 * C::DP;
}
// packed weights: forward regions (GEMM1: 4 slabs / k-step, GEMM2: 2) + three transposed-role regions (2 each)
template <class C>
constexpr size_t bwd_packed_bytes() {
  return (size_t)C::KSTEPS * (4 + 2 + 2 + 2 + 2) * C::SLAB;
}

template <class C>
__global__ void pack_bwd_weights_kernel(const float *__restrict__ wpt, const float *__restrict__ wd2t,
                                        const float *__restrict__ wg2t, int D, unsigned char *__restrict__ out) {
  // matrices: 0 W' fwd, 1 Wd2 fwd, 2 Wg2 fwd, 3 Wg2 bwd (GEMM3), 4 Wd2 bwd (GEMM4a), 5 W' bwd (GEMM4b)
  const int per = C::DP * (C::DP / 2);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < 6 * per; e += gridDim.x * blockDim.x) {
    const int m = e / per;
    const int rem = e - m * per;
    const int n = rem / (C::DP / 2), k = (rem - n * (C::DP / 2)) * 2;
    const float *src = (m == 0 || m == 5) ? wpt : ((m == 1 || m == 4) ? wd2t : wg2t);
    float x0 = 0.f, x1 = 0.f;
    if (m < 3) {  // forward role: B[n][k] = Wt[k][n]
      if (n < D && k < D) x0 = src[(size_t)k * D + n];
      if (n < D && k + 1 < D) x1 = src[(size_t)(k + 1) * D + n];
    } else {      // backward role: B[n][k] = Wt[n][k]
      if (n < D && k < D) x0 = src[(size_t)n * D + k];
      if (n < D && k + 1 < D) x1 = src[(size_t)n * D + k + 1];
    }
    uint32_t hi, lo;
    split2(x0, x1, hi, lo);
    const int ks = k >> 4;
    size_t base;
    if (m < 2)
      base = (size_t)ks * 4 * C::SLAB + (size_t)m * 2 * C::SLAB;
    else
      base = (size_t)C::KSTEPS * (4 + 2 * (m - 2)) * C::SLAB + (size_t)ks * 2 * C::SLAB;
    const uint32_t in_slab = canon_off(C::DP, n, k & 15);
    *reinterpret_cast<uint32_t *>(out + base + in_slab) = hi;
    *reinterpret_cast<uint32_t *>(out + base + C::SLAB + in_slab) = lo;
  }
}

constexpr int dwtc_max_jobs = 24;   // = dwtc::MAX_JOBS (dw_tc.cu)

struct Staging {
  unsigned char *h, *g, *da, *dgp, *ds;  // each: tiles_in_segment * staged_tile_bytes
};

__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <class C>
__device__ __forceinline__ void write_operand(unsigned char *A_hi, unsigned char *A_lo, unsigned char *stage_tile, int r,
                                              int k0, const float (&x)[8]) {
  uint4 hi, lo;
  split2(x[0], x[1], hi.x, lo.x);
  split2(x[2], x[3], hi.y, lo.y);
  split2(x[4], x[5], hi.z, lo.z);
  split2(x[6], x[7], hi.w, lo.w);
  if (A_hi) {
    const uint32_t off = canon_off(128, r, k0);
    *reinterpret_cast<uint4 *>(A_hi + off) = hi;
    *reinterpret_cast<uint4 *>(A_lo + off) = lo;
  }
  if (stage_tile) {  // k-step-major staged layout, see dw_tc.cu
    unsigned char *p = stage_tile + (size_t)(r >> 4) * (2 * C::DP * 32) + (size_t)(k0 >> 3) * 256 + (r & 15) * 16;
    *reinterpret_cast<uint4 *>(p) = hi;
    *reinterpret_cast<uint4 *>(p + C::DP * 32) = lo;
  }
}

// Re-loads this thread's staged chunks (written earlier by THIS thread) and stores them as the A operand: the
// operand tiles ds / dgp are needed again one GEMM later, and keeping them in registers across the wait costs 56
// registers (measured: 1.7 KB of spills per thread and +14 % kernel time).
template <class C>
__device__ __forceinline__ void staged_to_operand(unsigned char *A_hi, unsigned char *A_lo, const unsigned char *stage_tile,
                                                  int r, int ch0, int nch) {
#pragma unroll
  for (int q = 0; q < C::MAXCH; ++q) {
    if (q < nch) {
      const int k0 = (ch0 + q) * 8;
      const unsigned char *p = stage_tile + (size_t)(r >> 4) * (2 * C::DP * 32) + (size_t)(k0 >> 3) * 256 + (r & 15) * 16;
      const uint4 hi = *reinterpret_cast<const uint4 *>(p);
      const uint4 lo = *reinterpret_cast<const uint4 *>(p + C::DP * 32);
      const uint32_t off = canon_off(128, r, k0);
      *reinterpret_cast<uint4 *>(A_hi + off) = hi;
      *reinterpret_cast<uint4 *>(A_lo + off) = lo;
    }
  }
}

template <class C>
__global__ void __launch_bounds__(C::THREADS, 1)
vattn_bwd_tc_kernel(const nsdp_vattn_args a, const float *__restrict__ out, const float *__restrict__ stats,
                    const float *__restrict__ dout, const nsdp_vattn_grads g, const unsigned char *__restrict__ packed,
                    const Staging stg, long long tile_begin, long long tile_end, int *err) {
  using L = BwdLayout<C>;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char *A_hi = smem + L::OFF_A;
  unsigned char *A_lo = A_hi + C::A_HALF;
  float *scratch = reinterpret_cast<float *>(smem + L::OFF_A);  // aliases A once GEMM4b has consumed it
  unsigned char *stage0 = smem + L::OFF_STAGE;
  float4 *wd0s = reinterpret_cast<float4 *>(smem + L::OFF_WD0);
  float *pcs = reinterpret_cast<float *>(smem + L::OFF_PC);
  float *vcs = reinterpret_cast<float *>(smem + L::OFF_VC);
  float4 *rels = reinterpret_cast<float4 *>(smem + L::OFF_RELS);
  float *relacc = reinterpret_cast<float *>(smem + L::OFF_RELACC);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L::OFF_BAR);
  uint64_t *full = bars, *empty = bars + C::SLOTS, *a_ready = bars + 2 * C::SLOTS, *acc_done = a_ready + 1;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_done + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = a.D;
  const int krows = a.K + (a.has_global ? 1 : 0);
  const long long BM = (long long)a.B * a.M;

  for (int kk = tid; kk < C::DP; kk += C::THREADS) {
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    float p = 0.f, v = 0.f;
    if (kk < D) {
      w = make_float4(a.wd0[kk * 3 + 0], a.wd0[kk * 3 + 1], a.wd0[kk * 3 + 2], a.bd0[kk]);
      p = a.pc[kk];
      v = a.vc[kk];
    }
    wd0s[kk] = w;
    pcs[kk] = p;
    vcs[kk] = v;
  }
  if (tid == 0) {
    for (int s = 0; s < C::SLOTS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(a_ready, C::WORKER_WARPS);
    mbar_init(acc_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) {
    // ===================== weight producer =====================
    // the packed image is consumed front to back once per tile (GEMM1 takes 2 slots per k-step, GEMM2..4b one); PL lanes
    // share the copies (lane l serves slots l, l + PL, ...), see vattn_fwd_oh_kernel
    constexpr int PL = C::SLOTS / 2;
    constexpr int PER_TILE = 6 * C::KSTEPS;
    if (lane < PL) {
      const long long first = tile_begin + blockIdx.x;
      const long long my_tiles = first < tile_end ? (tile_end - first + gridDim.x - 1) / gridDim.x : 0;
      const long long total = my_tiles * PER_TILE;
      for (long long it = lane; it < total; it += PL) {
        const int j = (int)(it % PER_TILE);
        const int s = (int)(it % C::SLOTS);
        const uint32_t ph = (uint32_t)(it / C::SLOTS) & 1;
        mbar_wait(&empty[s], ph ^ 1, err);
        mbar_arrive_expect_tx(&full[s], C::SLOT_BYTES);
        bulk_g2s(stage0 + (size_t)s * C::SLOT_BYTES, packed + (size_t)j * C::SLOT_BYTES, C::SLOT_BYTES, &full[s]);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // whole warp runs loops and waits, one elected lane issues (see vattn_bwd_oh_kernel)
    {
      const uint32_t idesc = idesc_bf16(128, C::DP);
      constexpr uint32_t lbo_a = 128 * 16, lbo_b = C::DP * 16;
      constexpr uint64_t A_STEP = (2 * lbo_a) >> 4;
      const uint64_t ah0 = smem_desc(smem_u32(A_hi), lbo_a, 128), al0 = smem_desc(smem_u32(A_lo), lbo_a, 128);
      const uint64_t bh0 = smem_desc(smem_u32(stage0), lbo_b, 128);
      uint32_t slot = 0, slot_phase = 0, ready_phase = 0;
      for (long long tile = tile_begin + blockIdx.x; tile < tile_end; tile += gridDim.x) {
        for (int gi = 0; gi < 5; ++gi) {
          mbar_wait(a_ready, ready_phase, err);
          ready_phase ^= 1;
          tc_fence_after();
          // destination accumulator / accumulate-into flags of the five GEMMs
          const uint32_t dcol = (gi == 3 || gi == 4) ? C::ACC1_COL : 0;
          const bool keep = gi == 4;  // GEMM4b adds to GEMM4a's result
          for (int ks = 0; ks < C::KSTEPS; ++ks) {
            for (int m = 0; m < (gi == 0 ? 2 : 1); ++m) {   // GEMM1: W' -> acc0, then Wd2 -> acc1
              mbar_wait(&full[slot], slot_phase, err);
              tc_fence_after();
              if (elect_one()) {
                const uint64_t ah = ah0 + ks * A_STEP, al = al0 + ks * A_STEP;
                const uint64_t bh = bh0 + (uint64_t)slot * (C::SLOT_BYTES >> 4);
                const uint32_t d = tmem_base + (m ? C::ACC1_COL : dcol);
                const bool acc = keep || ks > 0;
                mma_bf16(d, ah, bh, idesc, acc);
                mma_bf16(d, al, bh, idesc, true);
                mma_bf16(d, ah, bh + (C::SLAB >> 4), idesc, true);
                mma_commit(&empty[slot]);
              }
              if (++slot == C::SLOTS) { slot = 0; slot_phase ^= 1; }
            }
          }
          if (elect_one()) mma_commit(acc_done);
        }
      }
    }
  } else {
    // ===================== workers =====================
    const int ww = warp - 2;
    const int wtid = tid - 64;               // 0 .. 511
    const int quarter = warp & 3;
    const int part = ww >> 2;
    const int r = quarter * 32 + lane;
    const int ch0 = part * C::CHUNKS / C::NPART, ch1 = (part + 1) * C::CHUNKS / C::NPART;
    const int nch = ch1 - ch0;
    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t done_phase = 0;
    // per-column accumulators of d_wd0 / d_bd0 (worker wtid < D owns column wtid for the whole kernel)
    float cw0 = 0.f, cw1 = 0.f, cw2 = 0.f, cb = 0.f;

    auto wait_acc = [&]() {
      mbar_wait(acc_done, done_phase, err);
      done_phase ^= 1;
      tc_fence_after();
    };
    auto publish = [&]() {  // operand written -> MMA may start
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);
    };

    for (long long tile = tile_begin + blockIdx.x; tile < tile_end; tile += gridDim.x) {
      const size_t st_off = (size_t)(tile - tile_begin) * staged_tile_bytes<C>();
      const RowInfo ri = row_info<C>(a, tile, r, krows);
      const bool row_on = ri.c >= 0;
      const bool is_glob = row_on && ri.n < 0;
      if (part == 0) rels[r] = make_float4(ri.rx, ri.ry, ri.rz, ri.flag);
      float R[C::MAXCH][8];
      unsigned long long gmaskbits = 0ull;
      // ---- H operand (+ staged for d_wpt / d_wd2t) ---------------------------------------------------------
#pragma unroll
      for (int q = 0; q < C::MAXCH; ++q) {
        if (q < nch) {
          const int k0 = (ch0 + q) * 8;
          float h[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 w0 = wd0s[k0 + j];
            const float pre = fmaf(w0.x, ri.rx, fmaf(w0.y, ri.ry, fmaf(w0.z, ri.rz, w0.w)));
            h[j] = ri.flag * fmaxf(pre, 0.f);
          }
          write_operand<C>(A_hi, A_lo, stg.h + st_off, r, k0, h);
        }
      }
      publish();
      // ---- gather P while GEMM1 runs ---------------------------------------------------------------------------
      {
        const float *qrow = (row_on && !is_glob && a.qp) ? a.qp + (size_t)ri.c * D : nullptr;
        const float *krow = (row_on && !is_glob && a.kp) ? a.kp + (size_t)ri.n * D : nullptr;
        const float *grow = is_glob ? a.gq + (size_t)(-ri.n - 1) * D : nullptr;
#pragma unroll
        for (int q = 0; q < C::MAXCH; ++q) {
#pragma unroll
          for (int j = 0; j < 8; j += 4) {
            const int col = (ch0 + q) * 8 + j;
            float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q < nch && col < D && row_on) {
              if (grow) {
                p = ldg4(grow + col);
              } else {
                p = *reinterpret_cast<const float4 *>(pcs + col);
                if (qrow) {
                  const float4 t = ldg4(qrow + col);
                  p.x += t.x; p.y += t.y; p.z += t.z; p.w += t.w;
                }
                if (krow) {
                  const float4 t = ldg4(krow + col);
                  p.x -= t.x; p.y -= t.y; p.z -= t.z; p.w -= t.w;
                }
              }
            }
            R[q][j] = p.x; R[q][j + 1] = p.y; R[q][j + 2] = p.z; R[q][j + 3] = p.w;
          }
        }
      }
      // ---- G = relu(gp + P): operand, staging, mask -------------------------------------------------------------
      wait_acc();
#pragma unroll
      for (int q = 0; q < C::MAXCH; ++q) {
        if (q < nch) {
          const int k0 = (ch0 + q) * 8;
          float v[8], gg[8];
          tmem_ld8(trow + k0, v);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            gg[j] = (row_on && k0 + j < D) ? fmaxf(v[j] + R[q][j], 0.f) : 0.f;
            if (gg[j] > 0.f) gmaskbits |= 1ull << (q * 8 + j);
          }
          write_operand<C>(A_hi, A_lo, stg.g + st_off, r, k0, gg);
        }
      }
      publish();
      // ---- gather V while GEMM2 runs ---------------------------------------------------------------------------------
      {
        const float *vrow = (row_on && !is_glob && a.vp) ? a.vp + (size_t)ri.n * D : nullptr;
        const float *gvrow = is_glob ? a.gv + (size_t)(-ri.n - 1) * D : nullptr;
#pragma unroll
        for (int q = 0; q < C::MAXCH; ++q) {
#pragma unroll
          for (int j = 0; j < 8; j += 4) {
            const int col = (ch0 + q) * 8 + j;
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q < nch && col < D && row_on) {
              if (gvrow) {
                t = ldg4(gvrow + col);
              } else {
                t = *reinterpret_cast<const float4 *>(vcs + col);
                if (vrow) {
                  const float4 u = ldg4(vrow + col);
                  t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
                }
              }
            }
            R[q][j] = t.x; R[q][j + 1] = t.y; R[q][j + 2] = t.z; R[q][j + 3] = t.w;
          }
        }
      }
      // ---- w, s -> ds = w*dout, da = ds*(s - out); scatter d_vp / d_gv; operand da ------------------------------------
      wait_acc();
#pragma unroll
      for (int q = 0; q < C::MAXCH; ++q) {
        if (q < nch) {
          const int k0 = (ch0 + q) * 8;
          float av[8], dl[8], ds[8], da[8];
          tmem_ld8(trow + k0, av);
          tmem_ld8(trow + C::ACC1_COL + k0, dl);
#pragma unroll
          for (int j = 0; j < 8; j += 4) {
            const int col = k0 + j;
            float4 mx = make_float4(0.f, 0.f, 0.f, 0.f), iv = mx, go = mx, o = mx;
            const bool on = row_on && col < D;
            if (on) {
              mx = ldg4(stats + (size_t)ri.c * D + col);
              iv = ldg4(stats + ((size_t)BM + ri.c) * D + col);
              go = ldg4(dout + (size_t)ri.c * D + col);
              o = ldg4(out + (size_t)ri.c * D + col);
            }
            const float mxs[4] = {mx.x, mx.y, mx.z, mx.w}, ivs[4] = {iv.x, iv.y, iv.z, iv.w};
            const float gos[4] = {go.x, go.y, go.z, go.w}, os[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float w = on ? __expf(av[j + u] - mxs[u]) * ivs[u] : 0.f;
              const float s = is_glob ? R[q][j + u] : R[q][j + u] + dl[j + u];
              ds[j + u] = w * gos[u];
              da[j + u] = ds[j + u] * (s - os[u]);
            }
            if (on && !is_glob && g.d_vp)
              red_add_v4(g.d_vp + (size_t)ri.n * D + col, ds[j], ds[j + 1], ds[j + 2], ds[j + 3]);
          }
          write_operand<C>(A_hi, A_lo, stg.da + st_off, r, k0, da);
          write_operand<C>(nullptr, nullptr, stg.ds + st_off, r, k0, ds);
        }
      }
      publish();
      // ---- GEMM3 done: A is free -> operand ds (GEMM4a) ------------------------------------------------------------------
      wait_acc();
      staged_to_operand<C>(A_hi, A_lo, stg.ds + st_off, r, ch0, nch);
      publish();
      // ---- dgp = dg * [g > 0] (reads acc0 while GEMM4a fills acc1); scatter d_kp / d_qp / d_gq -------------------------------
#pragma unroll
      for (int q = 0; q < C::MAXCH; ++q) {
        if (q < nch) {
          const int k0 = (ch0 + q) * 8;
          float dg[8];
          tmem_ld8(trow + k0, dg);
#pragma unroll
          for (int j = 0; j < 8; ++j) dg[j] = ((gmaskbits >> (q * 8 + j)) & 1ull) ? dg[j] : 0.f;
#pragma unroll
          for (int j = 0; j < 8; j += 4) {
            const int col = k0 + j;
            if (row_on && col < D && !is_glob) {
              if (g.d_kp) red_add_v4(g.d_kp + (size_t)ri.n * D + col, -dg[j], -dg[j + 1], -dg[j + 2], -dg[j + 3]);
              if (g.d_qp) red_add_v4(g.d_qp + (size_t)ri.c * D + col, dg[j], dg[j + 1], dg[j + 2], dg[j + 3]);
            }
          }
          write_operand<C>(nullptr, nullptr, stg.dgp + st_off, r, k0, dg);
        }
      }
      // ---- GEMM4a done: A is free -> operand dgp (GEMM4b) ----------------------------------------------------------------------
      wait_acc();
      staged_to_operand<C>(A_hi, A_lo, stg.dgp + st_off, r, ch0, nch);
      publish();
      // ---- dpre = dh * [h > 0]; d rel; dpre -> fp32 scratch (aliases A, free once GEMM4b is done) ------------------------------
      wait_acc();
      {
        float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
        for (int q = 0; q < C::MAXCH; ++q) {
          if (q < nch) {
            const int k0 = (ch0 + q) * 8;
            float dh[8];
            tmem_ld8(trow + C::ACC1_COL + k0, dh);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int col = k0 + j;
              const float4 w0 = wd0s[col];
              const float pre = fmaf(w0.x, ri.rx, fmaf(w0.y, ri.ry, fmaf(w0.z, ri.rz, w0.w)));
              const float dp = (ri.flag != 0.f && pre > 0.f) ? dh[j] : 0.f;
              sx = fmaf(dp, w0.x, sx); sy = fmaf(dp, w0.y, sy); sz = fmaf(dp, w0.z, sz);
              if (col < D) scratch[(size_t)col * L::SCR_LD + ((r + col) & 127)] = dp;
            }
          }
        }
        relacc[(part * 3 + 0) * 128 + r] = sx;
        relacc[(part * 3 + 1) * 128 + r] = sy;
        relacc[(part * 3 + 2) * 128 + r] = sz;
      }
      tc_fence_before();
      asm volatile("bar.sync 1, %0;" ::"n"(C::WORKER_WARPS * 32) : "memory");
      // column owners: d_wd0 / d_bd0 partial sums over the 128 rows of the tile
      if (wtid < D) {
        const float *colp = scratch + (size_t)wtid * L::SCR_LD;
#pragma unroll 4
        for (int rr = 0; rr < 128; ++rr) {
          const float dp = colp[(rr + wtid) & 127];
          const float4 rl = rels[rr];
          cw0 = fmaf(dp, rl.x, cw0); cw1 = fmaf(dp, rl.y, cw1); cw2 = fmaf(dp, rl.z, cw2); cb += dp;
        }
      }
      // row owners: d_xyz
      if (part == 0 && ri.flag != 0.f && (g.d_xyz_c || g.d_xyz_n)) {
        float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
        for (int pp = 0; pp < C::NPART; ++pp) {
          sx += relacc[(pp * 3 + 0) * 128 + r];
          sy += relacc[(pp * 3 + 1) * 128 + r];
          sz += relacc[(pp * 3 + 2) * 128 + r];
        }
        if (g.d_xyz_c) {
          float *dst = g.d_xyz_c + (size_t)ri.c * 3;
          atomicAdd(dst, a.sign * sx); atomicAdd(dst + 1, a.sign * sy); atomicAdd(dst + 2, a.sign * sz);
        }
        if (g.d_xyz_n) {
          float *dst = g.d_xyz_n + (size_t)ri.n * 3;
          atomicAdd(dst, -a.sign * sx); atomicAdd(dst + 1, -a.sign * sy); atomicAdd(dst + 2, -a.sign * sz);
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(C::WORKER_WARPS * 32) : "memory");
    }
    if (wtid < D) {
      if (g.d_wd0) {
        atomicAdd(g.d_wd0 + wtid * 3 + 0, cw0); atomicAdd(g.d_wd0 + wtid * 3 + 1, cw1); atomicAdd(g.d_wd0 + wtid * 3 + 2, cw2);
      }
      if (g.d_bd0) atomicAdd(g.d_bd0 + wtid, cb);
    }
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// tmp (zero-filled scatter target) -> dst += tmp ; colsum[c] += sign * sum_rows tmp[row][c]
__global__ void finalize_scatter_kernel(const float *__restrict__ tmp, float *__restrict__ dst, float *__restrict__ colsum,
                                        float sign, long long rows, int D) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  const long long r0 = rows * blockIdx.y / gridDim.y, r1 = rows * (blockIdx.y + 1) / gridDim.y;
  float s = 0.f;
#pragma unroll 8
  for (long long r = r0; r < r1; ++r) {   // ~32 rows per block, loads batched by the unroll (the pass is pure streaming)
    const float v = tmp[r * D + c];
    s += v;
    if (dst) dst[r * D + c] += v;
  }
  if (colsum) atomicAdd(colsum + c, sign * s);
}

// staging workspace = 5 tensors x 4144 tiles x 104 KB = 2.2 GB per segment (NSDP_VATTN_SEG overrides the segment length)
static long long segment_tiles() {
  static const long long v = [] {
    const char *e = getenv("NSDP_VATTN_SEG");
    const long long t = e ? atoll(e) : 4144;   // 28 x 148 SMs: whole waves (sweep 296 ... 12 432: longer is slightly better)
    return t < 1 || t > 16576 ? 4144ll : t;
  }();
  return v;
}
#define kSegmentTiles segment_tiles()

template <class C>
static size_t bwd_workspace_bytes(const nsdp_vattn_args &a) {
  const long long tiles = ceil_div((long long)a.B * a.M, (long long)C::CENTRES);
  const long long seg = tiles < kSegmentTiles ? tiles : kSegmentTiles;
  return bwd_packed_bytes<C>() + 256 + 2 * sizeof(float) * (size_t)a.B * a.N * a.D + 5 * (size_t)seg * staged_tile_bytes<C>();
}

template <class C>
static int launch_bwd(const nsdp_vattn_args &a, const float *out, const float *stats, const float *dout,
                      const nsdp_vattn_grads &g, void *workspace, size_t ws_bytes, cudaStream_t st) {
  if (!workspace || ws_bytes < bwd_workspace_bytes<C>(a)) return NSDP_ERR_WORKSPACE;
  if (!g.d_wd2t || !g.d_wpt || !g.d_wg2t) return NSDP_ERR_INVALID_ARGUMENT;
  unsigned char *packed = (unsigned char *)workspace;
  int *err = (int *)(packed + bwd_packed_bytes<C>());
  // d_vp / d_kp are scattered into zeroed temporaries so that d_vc = sum d_vp and d_pc = -sum d_kp (the bias
  // gradients are column sums over the non-global rows, which is exactly what lands in those tables) can be derived
  // without touching whatever running sum the caller keeps in its own buffers
  const size_t tbl = (size_t)a.B * a.N * a.D;
  float *tmp_vp = (float *)(packed + bwd_packed_bytes<C>() + 256);
  float *tmp_kp = tmp_vp + tbl;
  unsigned char *stage_base = (unsigned char *)(tmp_kp + tbl);
  const long long tiles = ceil_div((long long)a.B * a.M, (long long)C::CENTRES);
  const long long seg = tiles < kSegmentTiles ? tiles : kSegmentTiles;
  const size_t per = (size_t)seg * staged_tile_bytes<C>();
  Staging stg{stage_base, stage_base + per, stage_base + 2 * per, stage_base + 3 * per, stage_base + 4 * per};
  cudaError_t e = cudaMemsetAsync(err, 0, 256 + 2 * tbl * sizeof(float), st);
  if (e != cudaSuccess) return cuda_rc(e);
  nsdp_vattn_grads gk = g;   // what the kernel scatters into
  gk.d_vp = tmp_vp;
  gk.d_kp = tmp_kp;
  pack_bwd_weights_kernel<C><<<96, 256, 0, st>>>(a.wpt, a.wd2t, a.wg2t, a.D, packed);
  int rc = check_launch();
  if (rc != NSDP_OK) return rc;
  auto kern = vattn_bwd_tc_kernel<C>;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdLayout<C>::SMEM);
  if (e != cudaSuccess) return cuda_rc(e);
  for (long long t0 = 0; t0 < tiles; t0 += seg) {
    const long long t1 = t0 + seg < tiles ? t0 + seg : tiles;
    const long long n = t1 - t0;
    const int grid = (int)(n < num_sms() ? n : num_sms());
    kern<<<grid, C::THREADS, BwdLayout<C>::SMEM, st>>>(a, out, stats, dout, gk, packed, stg, t0, t1, err);
    rc = check_launch();
    if (rc != NSDP_OK) return rc;
    dwtc::Job jobs[3] = {
        {stg.g, stg.da, g.d_wg2t, C::DP, C::DP, a.D, a.D, a.D, nullptr, nullptr, 0, 0, 0},
        // the global-token rows (row % KR == KR-1) of dGP / dS sum up to d_gq / d_gv per shape
        {stg.h, stg.dgp, g.d_wpt, C::DP, C::DP, a.D, a.D, a.D, nullptr, a.has_global ? g.d_gq : nullptr, C::KR, a.M,
         t0 * C::CENTRES},
        {stg.h, stg.ds, g.d_wd2t, C::DP, C::DP, a.D, a.D, a.D, nullptr, a.has_global ? g.d_gv : nullptr, C::KR, a.M,
         t0 * C::CENTRES},
    };
    rc = dw_tc_launch(jobs, 3, n, err, st);
    if (rc != NSDP_OK) return rc;
  }
  const long long rows = (long long)a.B * a.N;
  dim3 fgrid((unsigned)ceil_div(a.D, 128), (unsigned)(rows < 1024 ? ceil_div(rows, 8ll) : (rows / 32 < 2048 ? rows / 32 : 2048)));
  finalize_scatter_kernel<<<fgrid, 128, 0, st>>>(tmp_vp, a.vp ? g.d_vp : nullptr, g.d_vc, 1.f, rows, a.D);
  finalize_scatter_kernel<<<fgrid, 128, 0, st>>>(tmp_kp, a.kp ? g.d_kp : nullptr, g.d_pc, -1.f, rows, a.D);
  return check_launch();
}


// =====================================================================================================================
// OH variant of the chain kernel (decoder cross-attention; forward counterpart: vattn_fwd_oh_kernel in vattn_tc.cu).
//   * GEMM1 = [H | E] * [W' ; T1_b] -> acc0 and [H | E] * [Wd2 ; T2_b] -> acc1: no per-row gathers of the anchor tables,
//     the epilogues add per-column constants only (G = relu(acc0 + pc), s = acc1 + vc).
//   * the scatters d_kp / d_vp / d_gq / d_gv become products with the same one-hot operand, dT1_b = E^T dGP and
//     dT2_b = E^T dS, reduced by dw_tc_kernel from the staged tiles like the weight gradients (no atomics).
//   * dS and dGP are needed a second time as A operands one GEMM later: they are PARKED in TMEM (already split into bf16
//     hi/lo words, in the columns of the accumulator they were derived from) instead of being re-read from global memory.
//   * the per-centre inputs of the softmax backward (max, 1/sum, d_out, out) are loaded one half-chunk ahead.
//   * staging precision: `lo` = 0 stages only the bf16 hi half of every operand tile (reductions over millions of rows).
// =====================================================================================================================
struct OhStaging {
  unsigned char *h, *g, *da, *dgp, *ds, *e;
  int lo;   // 1: [hi slab][lo slab] per k-step, 0: hi slab only
  int f16;  // 1 (with lo == 0): the single slab holds fp16 values, gradient tiles scaled by 2^k (stage_f16.cuh)
  const unsigned *gmax;   // float bits of max|d_out| (sampled) the scale is derived from
};

template <class C>
__global__ void __launch_bounds__(C::THREADS, 1)
vattn_bwd_oh_kernel(const nsdp_vattn_args a, const float *__restrict__ out, const float *__restrict__ stats,
                    const float *__restrict__ dout, const nsdp_vattn_grads g, const unsigned char *__restrict__ packed,
                    const unsigned char *__restrict__ tables, const OhStaging stg, int tpb, long long tile_begin,
                    long long tile_end, int *err, unsigned long long *trace) {
  static_assert(C::OH && C::KR == 8, "one-hot kernel: 8 rows per centre");
  using L = BwdLayout<C>;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char *A_hi = smem + L::OFF_A;
  unsigned char *A_lo = A_hi + C::A_HALF;
  unsigned char *E = smem + L::OFF_E;
  unsigned char *stage0 = smem + L::OFF_STAGE;
  float4 *wd0s = reinterpret_cast<float4 *>(smem + L::OFF_WD0);
  float *pcs = reinterpret_cast<float *>(smem + L::OFF_PC);
  float *vcs = reinterpret_cast<float *>(smem + L::OFF_VC);
  float *relacc = reinterpret_cast<float *>(smem + L::OFF_RELACC);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L::OFF_BAR);
  // kready[ks]: the 16 operand columns of k-step ks are in place (8 warps: 2 chunks x 4 lane quarters); e_ready: the
  // one-hot operand (GEMM1) / the [rel | 1] operand (d_wd0 product); acc_free: every worker is done reading the
  // accumulator the next GEMM overwrites. The GEMMs trail the workers' epilogues k-step by k-step (vattn_fwd_oh_kernel).
  uint64_t *full = bars, *empty = bars + C::SLOTS, *acc_done = bars + 2 * C::SLOTS, *kready = acc_done + 1;
  uint64_t *e_ready = kready + C::KSTEPS, *acc_free = e_ready + 1;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_free + 1);
  static_assert((2 * C::SLOTS + C::KSTEPS + 3) * 8 + 4 <= 256, "mbarrier area");

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = a.D;
  const int krows = a.K + 1;
  const long long BM = (long long)a.B * a.M;

  for (int kk = tid; kk < C::DP; kk += C::THREADS) {
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    float p = 0.f, v = 0.f;
    if (kk < D) {
      w = make_float4(a.wd0[kk * 3 + 0], a.wd0[kk * 3 + 1], a.wd0[kk * 3 + 2], a.bd0[kk]);
      p = a.pc[kk];
      v = a.vc[kk];
    }
    wd0s[kk] = w;
    pcs[kk] = p;
    vcs[kk] = v;
  }
  if (tid == 0) {
    for (int s = 0; s < C::SLOTS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int ks = 0; ks < C::KSTEPS; ++ks) mbar_init(&kready[ks], 8);
    mbar_init(e_ready, C::WORKER_WARPS);
    mbar_init(acc_free, C::WORKER_WARPS);
    mbar_init(acc_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr int W1 = 2 * C::KSTEPS;     // slots of GEMM1's weight part (W', Wd2 per k-step)
  constexpr int T1 = 2 * C::E_KSTEPS;   // slots of GEMM1's table part (T1_b, T2_b per k-step)

  if (warp == 0) {
    // ===================== producer: weights + this shape's tables, one slot at a time =====================
    // PL lanes share the work (lane l serves slots l, l + PL, ...), see vattn_fwd_oh_kernel
    constexpr int PL = C::SLOTS / 2;
    constexpr int PER_TILE = T1 + 6 * C::KSTEPS;
    if (lane < PL) {
      const long long first = tile_begin + blockIdx.x;
      const long long my_tiles = first < tile_end ? (tile_end - first + gridDim.x - 1) / gridDim.x : 0;
      const long long total = my_tiles * PER_TILE;
      for (long long it = lane; it < total; it += PL) {
        const long long tile = first + (it / PER_TILE) * gridDim.x;
        const int j = (int)(it % PER_TILE);
        const unsigned char *tb = tables + (size_t)(tile / tpb) * table_bytes_per_shape<C>();
        const unsigned char *src = j < W1 ? packed + (size_t)j * C::SLOT_BYTES
                                          : (j < W1 + T1 ? tb + (size_t)(j - W1) * C::SLOT_BYTES
                                                         : packed + (size_t)(j - T1) * C::SLOT_BYTES);
        const int s = (int)(it % C::SLOTS);
        const uint32_t ph = (uint32_t)(it / C::SLOTS) & 1;
        mbar_wait(&empty[s], ph ^ 1, err);
        mbar_arrive_expect_tx(&full[s], C::SLOT_BYTES);
        bulk_g2s(stage0 + (size_t)s * C::SLOT_BYTES, src, C::SLOT_BYTES, &full[s]);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the loops and the waits; one elected lane issues. Per slot the issuing thread must stay under
    // the ~310 cycles the tensor pipe needs for three MMAs, so the descriptors are advanced by adds (the start-address
    // field counts 16-byte units) and the ring position is a running counter.
    {
      const uint32_t idesc = idesc_bf16(128, C::DP);
      constexpr uint32_t lbo_a = 128 * 16, lbo_b = C::DP * 16;
      constexpr uint64_t A_STEP = (2 * lbo_a) >> 4;            // one k-step further inside an A operand
      const uint64_t ah0 = smem_desc(smem_u32(A_hi), lbo_a, 128), al0 = smem_desc(smem_u32(A_lo), lbo_a, 128);
      const uint64_t eh0 = smem_desc(smem_u32(E), lbo_a, 128);
      const uint64_t bh0 = smem_desc(smem_u32(stage0), lbo_b, 128);
      uint32_t slot = 0, slot_phase = 0, kphase = 0, ephase = 0, fphase = 0;
      bool first_tile = true;
      auto wait_kstep = [&](int ks) {
        mbar_wait_poll(&kready[ks], kphase, err);
        tc_fence_after();
      };
      auto wait_bar = [&](uint64_t *bar, uint32_t &phase) {
        mbar_wait_poll(bar, phase, err);
        phase ^= 1;
        tc_fence_after();
      };
      // waits for the next slot of the ring; returns the B descriptor of its hi slab (lo slab = + SLAB / 16)
      auto take_slot = [&](uint64_t &bh, uint64_t *&release) {
        mbar_wait_poll(&full[slot], slot_phase, err);
        tc_fence_after();
        bh = bh0 + (uint64_t)slot * (C::SLOT_BYTES >> 4);
        release = &empty[slot];
        if (++slot == C::SLOTS) { slot = 0; slot_phase ^= 1; }
      };
      for (long long tile = tile_begin + blockIdx.x; tile < tile_end; tile += gridDim.x) {
        // ---- GEMM1: [H | E] -> acc0 (W', T1), acc1 (Wd2, T2) ----
        TR(100);
        wait_bar(acc_free, fphase);
        TR(101);
        for (int ks = 0; ks < C::KSTEPS; ++ks) {
          const uint64_t ah = ah0 + ks * A_STEP, al = al0 + ks * A_STEP;
          wait_kstep(ks);
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            uint64_t bh, *rel;
            take_slot(bh, rel);
            if (elect_one()) {
              const uint32_t d = tmem_base + (m ? C::ACC1_COL : 0);
              mma_bf16(d, ah, bh, idesc, ks > 0);
              mma_bf16(d, al, bh, idesc, true);
              mma_bf16(d, ah, bh + (C::SLAB >> 4), idesc, true);
              mma_commit(rel);
            }
          }
        }
        kphase ^= 1;
        wait_bar(e_ready, ephase);
        for (int ks = 0; ks < C::E_KSTEPS; ++ks) {
          const uint64_t eh = eh0 + ks * A_STEP;
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            uint64_t bh, *rel;
            take_slot(bh, rel);
            if (elect_one()) {
              const uint32_t d = tmem_base + (m ? C::ACC1_COL : 0);
              mma_bf16(d, eh, bh, idesc, true);
              mma_bf16(d, eh, bh + (C::SLAB >> 4), idesc, true);
              mma_commit(rel);
            }
          }
        }
        TR(102);
        if (elect_one()) mma_commit(acc_done);
        // ---- GEMM2 (-> acc0), GEMM3 (-> acc0), GEMM4a (-> acc1), GEMM4b (acc1 +=) ----
        for (int gi = 1; gi < 5; ++gi) {
          TR(100 + 10 * gi);
          wait_bar(acc_free, fphase);
          TR(101 + 10 * gi);
          const uint32_t d = tmem_base + (gi >= 3 ? C::ACC1_COL : 0);
          for (int ks = 0; ks < C::KSTEPS; ++ks) {
            uint64_t bh, *rel;
            wait_kstep(ks);
            take_slot(bh, rel);
            if (elect_one()) {
              const uint64_t ah = ah0 + ks * A_STEP, al = al0 + ks * A_STEP;
              mma_bf16(d, ah, bh, idesc, gi == 4 || ks > 0);
              mma_bf16(d, al, bh, idesc, true);
              mma_bf16(d, ah, bh + (C::SLAB >> 4), idesc, true);
              mma_commit(rel);
            }
          }
          TR(102 + 10 * gi);
          if (elect_one()) mma_commit(acc_done);
          kphase ^= 1;
        }
        // ---- d_wd0 / d_bd0 += dpre^T [rel | 1]: the A buffer holds dpre (hi / lo), the first 8 KB of the E buffer the
        //      [128 x 16] operand (rx, ry, rz, 1, 0 ...) of the rows. Transposed product (both operands MN-major, as in
        //      dw_tc.cu): M = channel, K = the tile's 128 rows, N = 16. Two M-tiles accumulate over ALL tiles of this CTA
        //      in the spare TMEM columns next to acc0 / acc1.
        for (int ks = 0; ks < C::KSTEPS; ++ks) wait_kstep(ks);     // K runs over the tile's rows here: all of dpre
        kphase ^= 1;
        wait_bar(e_ready, ephase);
        if (elect_one()) {
          const uint32_t idesc_t = idesc_bf16_mn(128, 16);
          const uint64_t dh0 = smem_desc(smem_u32(A_hi), 128, 2048), dl0 = smem_desc(smem_u32(A_lo), 128, 2048);
          const uint64_t rh0 = smem_desc(smem_u32(E), 128, 2048), rl0 = smem_desc(smem_u32(E + 4096), 128, 2048);
          for (int j = 0; j < 2; ++j) {
            const uint32_t d = tmem_base + (j ? C::ACC1_COL + C::DP : C::DP);
            for (int ks = 0; ks < 8; ++ks) {     // 16 rows per k-step = two core matrices = 256 bytes along K
              const uint64_t xh = dh0 + (uint64_t)j * (32768 >> 4) + (uint64_t)ks * (256 >> 4);
              const uint64_t xl = dl0 + (uint64_t)j * (32768 >> 4) + (uint64_t)ks * (256 >> 4);
              const uint64_t yh = rh0 + (uint64_t)ks * (256 >> 4), yl = rl0 + (uint64_t)ks * (256 >> 4);
              mma_bf16(d, xh, yh, idesc_t, !(first_tile && ks == 0));
              mma_bf16(d, xl, yh, idesc_t, true);
              mma_bf16(d, xh, yl, idesc_t, true);
            }
          }
          mma_commit(acc_done);
        }
        first_tile = false;
      }
    }
  } else {
    // ===================== workers =====================
    const int ww = warp - 2;
    const int quarter = warp & 3;
    const int part = ww >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t done_phase = 0;
#ifdef NSDP_TRACE
    const bool tr_on = (tid == 64);
#define TW(id) do { if (tr_on) TR(id); } while (0)
#else
#define TW(id) do { } while (0)
#endif
    constexpr int NQ = (C::CHUNKS + C::NPART - 1) / C::NPART;           // chunk rounds per thread (chunk = part + q*NPART)
    constexpr int NE = (C::E_COLS / 8 + C::NPART - 1) / C::NPART;
    static_assert(NQ * 8 <= 64, "ReLU mask of G is kept in one 64-bit register");
    const int slabs = stg.lo ? 2 : 1;
    const uint32_t a_base = (uint32_t)((r >> 3) * 128 + (r & 7) * 16);                         // + chunk * 2048
    const size_t st_row = (size_t)(r >> 4) * (size_t)(slabs * C::DP * 32) + (size_t)(r & 15) * 16;   // + chunk * 256
    const size_t st_tile = (size_t)slabs * 256 * C::DP;
    const size_t ste_row = (size_t)(r >> 4) * (size_t)(C::E_COLS * 32) + (size_t)(r & 15) * 16;

    auto wait_acc = [&]() {
      mbar_wait(acc_done, done_phase, err);
      done_phase ^= 1;
      tc_fence_after();
    };
    // chunk ch of the A operand (8 columns of this warp's 32 rows) is written: one arrival on its k-step's barrier
    auto chunk_done = [&](int ch) {
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&kready[ch >> 1]);
    };
    // this warp no longer reads the accumulator the next GEMM overwrites (one arrival per GEMM 1 .. 4b)
    auto release_acc = [&]() {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_free);
    };
    // x[8] -> bf16 hi/lo words; optional A operand store, optional staged store
    // gradient tiles are staged times gsc (a power of two, 1 unless fp16 staging is on)
    const float gsc = stg.f16 ? stage16::scale_from_max(*stg.gmax) : 1.f;
    auto emit = [&](const float (&x)[8], int ch, bool to_a, unsigned char *stage_tile, uint4 &hi, uint4 &lo, float sc = 1.f) {
      split2(x[0], x[1], hi.x, lo.x);
      split2(x[2], x[3], hi.y, lo.y);
      split2(x[4], x[5], hi.z, lo.z);
      split2(x[6], x[7], hi.w, lo.w);
      if (to_a) {
        *reinterpret_cast<uint4 *>(A_hi + a_base + ch * 2048) = hi;
        *reinterpret_cast<uint4 *>(A_lo + a_base + ch * 2048) = lo;
      }
      if (stage_tile) {
        unsigned char *p = stage_tile + st_row + (size_t)ch * 256;
        if (stg.f16) {
          *reinterpret_cast<uint4 *>(p) = stage16::pack8(x, sc);
        } else {
          *reinterpret_cast<uint4 *>(p) = hi;
          if (stg.lo) *reinterpret_cast<uint4 *>(p + C::DP * 32) = lo;
        }
      }
    };

    RowInfoPB ri = row_info_pb<C>(a, tile_begin + blockIdx.x, r, krows, tpb);
    bool any_tile = false;
    for (long long tile = tile_begin + blockIdx.x; tile < tile_end; tile += gridDim.x) {
      TW(200);
      if (any_tile) wait_acc();     // the previous tile's d_wd0 product has read dpre (A) and the rel operand (E)
      any_tile = true;
      release_acc();                // GEMM1 overwrites acc0 / acc1: the previous tile's readers are all past their loads
      const size_t tl = (size_t)(tile - tile_begin);
      const bool row_on = ri.c >= 0;
      const int b = (int)(tile / tpb);
      unsigned long long gmaskbits = 0ull;
      // ---- operands E (smem + staged) and H (A + staged); GEMM1 starts on the first finished k-step ---------------------
#pragma unroll
      for (int q = 0; q < NE; ++q) {
        const int ch = part + q * C::NPART;
        if (ch < C::E_COLS / 8) {
          uint4 e = make_uint4(0u, 0u, 0u, 0u);
          if ((ri.j >> 3) == ch) {
            const uint32_t one = (ri.j & 1) ? 0x3F800000u : 0x00003F80u;
            const int w = (ri.j & 7) >> 1;
            e.x = w == 0 ? one : 0u; e.y = w == 1 ? one : 0u; e.z = w == 2 ? one : 0u; e.w = w == 3 ? one : 0u;
          }
          *reinterpret_cast<uint4 *>(E + a_base + ch * 2048) = e;
          if (stg.f16) {   // the staged copy feeds an fp16 product: 1.0 = 0x3C00 there (bf16: 0x3F80)
            e.x = e.x ? (e.x > 0xffffu ? 0x3C000000u : 0x00003C00u) : 0u;
            e.y = e.y ? (e.y > 0xffffu ? 0x3C000000u : 0x00003C00u) : 0u;
            e.z = e.z ? (e.z > 0xffffu ? 0x3C000000u : 0x00003C00u) : 0u;
            e.w = e.w ? (e.w > 0xffffu ? 0x3C000000u : 0x00003C00u) : 0u;
          }
          *reinterpret_cast<uint4 *>(stg.e + tl * (size_t)(256 * C::E_COLS) + ste_row + (size_t)ch * 256) = e;
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(e_ready);
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int ch = part + q * C::NPART;
        if (ch < C::CHUNKS) {
          float h[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 w0 = wd0s[ch * 8 + j];
            const float pre = fmaf(w0.x, ri.rx, fmaf(w0.y, ri.ry, fmaf(w0.z, ri.rz, w0.w)));
            h[j] = ri.flag * fmaxf(pre, 0.f);
          }
          uint4 hi, lo;
          emit(h, ch, true, stg.h + tl * st_tile, hi, lo);
          chunk_done(ch);
        }
      }
      TW(201);
      // ---- G = relu(acc0 + pc): operand, staged, mask. acc0 moves to registers first: GEMM2 overwrites it and starts on
      //      the first finished k-step ---------------------------------------------------------------------------------------
      wait_acc();
      TW(202);
      {
        uint32_t gp[NQ][8];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int ch = part + q * C::NPART;
          if (ch < C::CHUNKS) tmem_ld8_nowait(trow + ch * 8, gp[q]);
        }
        tmem_ld_wait();
        release_acc();
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int ch = part + q * C::NPART;
          if (ch < C::CHUNKS) {
            float gg[8];
            pin8(gp[q]);     // keeps this round's arithmetic behind the previous round's hand-over
            const float4 p0 = *reinterpret_cast<const float4 *>(pcs + ch * 8), p1 = *reinterpret_cast<const float4 *>(pcs + ch * 8 + 4);
            const float pv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              gg[j] = fmaxf(__uint_as_float(gp[q][j]) + pv[j], 0.f);
              if (gg[j] > 0.f) gmaskbits |= 1ull << (q * 8 + j);
            }
            uint4 hi, lo;
            emit(gg, ch, true, stg.g + tl * st_tile, hi, lo);
            chunk_done(ch);
          }
        }
      }
      TW(203);
      // ---- while GEMM2 runs: the next tile's row description -------------------------------------------------------------
      const RowInfoPB nxt = row_info_pb<C>(a, tile + gridDim.x, r, krows, tpb);
      // ---- w = exp(a - max) / sum, s = acc1 + vc; ds = w * dout, da = ds * (s - out) ---------------------------------------
      //      Two passes, so that GEMM3 (which overwrites acc0 = a) can trail the second one: (1) w of every chunk, kept as
      //      fp32 in the chunk's own 32 bytes of the (free) A buffer, then acc0 is released; (2) da -> A operand (over w) +
      //      staged, ds -> staged + parked (split words) in acc1's columns.
      {
        const size_t crow = (size_t)(row_on ? ri.c : 0) * D;
        const float *st_mx = stats + crow;
        const float *st_iv = stats + (size_t)BM * D + crow;
        const float *p_go = dout + crow;
        const float *p_o = out + crow;
        float4 cmx[2], civ[2];      // inputs of the current chunk (loaded one chunk ahead)
        auto load_stats = [&](int q, float4 (&mx)[2], float4 (&iv)[2]) {
          const int col = (part + q * C::NPART) * 8;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            mx[h] = iv[h] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row_on && col + h * 4 < D) { mx[h] = ldg4(st_mx + col + h * 4); iv[h] = ldg4(st_iv + col + h * 4); }
          }
        };
        load_stats(0, cmx, civ);
        wait_acc();
        TW(204);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int ch = part + q * C::NPART;
          float4 nmx[2], niv[2];
          load_stats(q + 1, nmx, niv);      // beyond the last chunk: col >= D, nothing is loaded
          if (ch < C::CHUNKS) {
            float av[8];
            tmem_ld8(trow + ch * 8, av);
            const float mxs[8] = {cmx[0].x, cmx[0].y, cmx[0].z, cmx[0].w, cmx[1].x, cmx[1].y, cmx[1].z, cmx[1].w};
            const float ivs[8] = {civ[0].x, civ[0].y, civ[0].z, civ[0].w, civ[1].x, civ[1].y, civ[1].z, civ[1].w};
            float w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w[j] = (row_on && ch * 8 + (j & 4) < D) ? __expf(av[j] - mxs[j]) * ivs[j] : 0.f;
            *reinterpret_cast<float4 *>(A_hi + a_base + ch * 2048) = make_float4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<float4 *>(A_lo + a_base + ch * 2048) = make_float4(w[4], w[5], w[6], w[7]);
          }
          cmx[0] = nmx[0]; cmx[1] = nmx[1]; civ[0] = niv[0]; civ[1] = niv[1];
        }
        release_acc();
        float4 cgo[2], co[2];
        auto load_grads = [&](int q, float4 (&go)[2], float4 (&o)[2]) {
          const int col = (part + q * C::NPART) * 8;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            go[h] = o[h] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row_on && col + h * 4 < D) { go[h] = ldg4(p_go + col + h * 4); o[h] = ldg4(p_o + col + h * 4); }
          }
        };
        load_grads(0, cgo, co);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int ch = part + q * C::NPART;
          float4 ngo[2], no[2];
          load_grads(q + 1, ngo, no);
          if (ch < C::CHUNKS) {
            float sv[8], ds[8], da[8];
            tmem_ld8(trow + C::ACC1_COL + ch * 8, sv);
            const float4 w0 = *reinterpret_cast<const float4 *>(A_hi + a_base + ch * 2048);
            const float4 w1 = *reinterpret_cast<const float4 *>(A_lo + a_base + ch * 2048);
            const float4 v0 = *reinterpret_cast<const float4 *>(vcs + ch * 8), v1 = *reinterpret_cast<const float4 *>(vcs + ch * 8 + 4);
            const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
            const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            const float gos[8] = {cgo[0].x, cgo[0].y, cgo[0].z, cgo[0].w, cgo[1].x, cgo[1].y, cgo[1].z, cgo[1].w};
            const float os[8] = {co[0].x, co[0].y, co[0].z, co[0].w, co[1].x, co[1].y, co[1].z, co[1].w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              ds[j] = w[j] * gos[j];
              da[j] = ds[j] * (sv[j] + vv[j] - os[j]);
            }
            uint4 hi, lo;
            emit(da, ch, true, stg.da + tl * st_tile, hi, lo, gsc);
            chunk_done(ch);
            emit(ds, ch, false, stg.ds + tl * st_tile, hi, lo, gsc);
            const uint32_t pk[8] = {hi.x, hi.y, hi.z, hi.w, lo.x, lo.y, lo.z, lo.w};
            tmem_st8(trow + C::ACC1_COL + ch * 8, pk);
          }
          cgo[0] = ngo[0]; cgo[1] = ngo[1]; co[0] = no[0]; co[1] = no[1];
        }
      }
      tmem_st_wait();
      TW(205);
      // ---- GEMM3 done: A is free -> operand ds from its parked words; they leave acc1 first (GEMM4a overwrites it) -----------
      wait_acc();
      TW(206);
      {
        uint32_t pk[NQ][8];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int ch = part + q * C::NPART;
          if (ch < C::CHUNKS) tmem_ld8_nowait(trow + C::ACC1_COL + ch * 8, pk[q]);
        }
        tmem_ld_wait();
        release_acc();
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int ch = part + q * C::NPART;
          if (ch < C::CHUNKS) {
            *reinterpret_cast<uint4 *>(A_hi + a_base + ch * 2048) = make_uint4(pk[q][0], pk[q][1], pk[q][2], pk[q][3]);
            *reinterpret_cast<uint4 *>(A_lo + a_base + ch * 2048) = make_uint4(pk[q][4], pk[q][5], pk[q][6], pk[q][7]);
            chunk_done(ch);
          }
        }
      }
      TW(207);
      // ---- dgp = dg * [g > 0] (reads acc0 while GEMM4a fills acc1): staged + parked in acc0's columns ---------------------------
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int ch = part + q * C::NPART;
        if (ch < C::CHUNKS) {
          float dg[8];
          tmem_ld8(trow + ch * 8, dg);
#pragma unroll
          for (int j = 0; j < 8; ++j) dg[j] = ((gmaskbits >> (q * 8 + j)) & 1ull) ? dg[j] : 0.f;
          uint4 hi, lo;
          emit(dg, ch, false, stg.dgp + tl * st_tile, hi, lo, gsc);
          const uint32_t pk[8] = {hi.x, hi.y, hi.z, hi.w, lo.x, lo.y, lo.z, lo.w};
          tmem_st8(trow + ch * 8, pk);
        }
      }
      tmem_st_wait();
      TW(208);
      // ---- GEMM4a done: A is free -> operand dgp (GEMM4b accumulates into acc1, which nobody reads now) -----------------------
      wait_acc();
      TW(209);
      release_acc();
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int ch = part + q * C::NPART;
        if (ch < C::CHUNKS) {
          uint32_t pk[8];
          tmem_ld8u(trow + ch * 8, pk);
          *reinterpret_cast<uint4 *>(A_hi + a_base + ch * 2048) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4 *>(A_lo + a_base + ch * 2048) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          chunk_done(ch);
        }
      }
      TW(210);
      // ---- dpre = dh * [h > 0]; d rel ----------------------------------------------------------------------------------------
      wait_acc();
      TW(211);
      {
        // dpre becomes one more operand (A buffer, GEMM4b is done with it) and the tensor core forms d_wd0 / d_bd0 from it
        // (transposed product with the [rel | 1] operand in the E buffer)
        if (part == 0) {
          // row r of the [128 x 16] operand (rx, ry, rz, 1, 0, ...), MN-major: n-chunk 0 at (r/8)*128 + (r%8)*16, chunk 1 (all
          // zero) 2048 bytes further; hi image at E, lo image at E + 4096
          uint4 hi, lo;
          split2(ri.rx, ri.ry, hi.x, lo.x);
          split2(ri.rz, 1.f, hi.y, lo.y);
          hi.z = hi.w = lo.z = lo.w = 0u;
          *reinterpret_cast<uint4 *>(E + a_base) = hi;
          *reinterpret_cast<uint4 *>(E + a_base + 2048) = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4 *>(E + 4096 + a_base) = lo;
          *reinterpret_cast<uint4 *>(E + 4096 + a_base + 2048) = make_uint4(0u, 0u, 0u, 0u);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(e_ready);
        float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int ch = part + q * C::NPART;
          if (ch < C::CHUNKS) {
            float dh[8], dp[8];
            tmem_ld8(trow + C::ACC1_COL + ch * 8, dh);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 w0 = wd0s[ch * 8 + j];
              const float pre = fmaf(w0.x, ri.rx, fmaf(w0.y, ri.ry, fmaf(w0.z, ri.rz, w0.w)));
              dp[j] = (ri.flag != 0.f && pre > 0.f) ? dh[j] : 0.f;
              sx = fmaf(dp[j], w0.x, sx); sy = fmaf(dp[j], w0.y, sy); sz = fmaf(dp[j], w0.z, sz);
            }
            uint4 hi, lo;
            emit(dp, ch, true, nullptr, hi, lo);
            chunk_done(ch);
          }
        }
        relacc[(part * 3 + 0) * 128 + r] = sx;
        relacc[(part * 3 + 1) * 128 + r] = sy;
        relacc[(part * 3 + 2) * 128 + r] = sz;
      }
      tc_fence_before();            // this tile's reads of acc1 are ordered before the barrier: the next GEMM1 may overwrite it
      asm volatile("bar.sync 1, %0;" ::"n"(C::WORKER_WARPS * 32) : "memory");     // relacc of every column part is in place
      TW(212);
      if (part == 0 && ri.flag != 0.f && (g.d_xyz_c || g.d_xyz_n)) {   // row owners: d_xyz
        float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
        for (int pp = 0; pp < C::NPART; ++pp) {
          sx += relacc[(pp * 3 + 0) * 128 + r];
          sy += relacc[(pp * 3 + 1) * 128 + r];
          sz += relacc[(pp * 3 + 2) * 128 + r];
        }
        if (g.d_xyz_c) {
          float *dst = g.d_xyz_c + (size_t)ri.c * 3;
          atomicAdd(dst, a.sign * sx); atomicAdd(dst + 1, a.sign * sy); atomicAdd(dst + 2, a.sign * sz);
        }
        if (g.d_xyz_n) {
          float *dst = g.d_xyz_n + ((size_t)b * a.N + ri.j) * 3;
          atomicAdd(dst, -a.sign * sx); atomicAdd(dst + 1, -a.sign * sy); atomicAdd(dst + 2, -a.sign * sz);
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(C::WORKER_WARPS * 32) : "memory");
      TW(213);
      ri = nxt;
    }
    if (any_tile) {
      wait_acc();                      // the last tile's d_wd0 product
      if (part == 0) {                 // one warp per TMEM lane quarter: lane = channel m (M-tile 0) / 128 + m (M-tile 1)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float v[8];
          tmem_ld8(trow + (j ? C::ACC1_COL + C::DP : C::DP), v);
          const int m = j * 128 + r;
          if (m < D) {
            if (g.d_wd0) {
              atomicAdd(g.d_wd0 + m * 3 + 0, v[0]); atomicAdd(g.d_wd0 + m * 3 + 1, v[1]); atomicAdd(g.d_wd0 + m * 3 + 2, v[2]);
            }
            if (g.d_bd0) atomicAdd(g.d_bd0 + m, v[3]);
          }
        }
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}


// =====================================================================================================================
// "Saved" variant of the OH chain kernel: the forward call kept H, G (staged operand tiles) and the fp32 blocks a / acc1
// of every tile in nsdp_vattn_args::saved, so the backward starts at the softmax backward: no H / E operands, no GEMM1,
// no GEMM2, no G epilogue. Per tile: (a, acc1, mask <- saved) -> dS, dA -> GEMM3 -> dGP -> GEMM4a/b -> dpre. The
// weight-gradient jobs read H and G straight from the saved buffer.
// =====================================================================================================================
template <class C>
__global__ void __launch_bounds__(C::THREADS, 1)
vattn_bwd_sv_kernel(const nsdp_vattn_args a, const float *__restrict__ out, const float *__restrict__ stats,
                    const float *__restrict__ dout, const nsdp_vattn_grads g, const unsigned char *__restrict__ packed,
                    const unsigned char *__restrict__ saved, long long tiles_total, const OhStaging stg, int tpb,
                    long long tile_begin, long long tile_end, int *err) {
  static_assert(C::OH && C::KR == 8, "one-hot kernel: 8 rows per centre");
  using L = BwdLayout<C>;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char *A_hi = smem + L::OFF_A;
  unsigned char *A_lo = A_hi + C::A_HALF;
  float *scratch = reinterpret_cast<float *>(smem + L::OFF_A);  // aliases A once GEMM4b has consumed it
  unsigned char *stage0 = smem + L::OFF_STAGE;
  float4 *wd0s = reinterpret_cast<float4 *>(smem + L::OFF_WD0);
  float *vcs = reinterpret_cast<float *>(smem + L::OFF_VC);
  float4 *rels = reinterpret_cast<float4 *>(smem + L::OFF_RELS);
  float *relacc = reinterpret_cast<float *>(smem + L::OFF_RELACC);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L::OFF_BAR);
  uint64_t *full = bars, *empty = bars + C::SLOTS, *a_ready = bars + 2 * C::SLOTS, *acc_done = a_ready + 1;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_done + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = a.D;
  const int krows = a.K + 1;
  const long long BM = (long long)a.B * a.M;
  constexpr size_t TB = saved_tile_bytes<C>();
  const unsigned char *sv_g = saved + (size_t)tiles_total * TB;        // staged G tiles (hi slab = ReLU mask source)
  const unsigned char *sv_a = saved + (size_t)2 * tiles_total * TB;    // fp32 a blocks
  const unsigned char *sv_s = saved + (size_t)3 * tiles_total * TB;    // fp32 acc1 blocks

  for (int kk = tid; kk < C::DP; kk += C::THREADS) {
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    float v = 0.f;
    if (kk < D) {
      w = make_float4(a.wd0[kk * 3 + 0], a.wd0[kk * 3 + 1], a.wd0[kk * 3 + 2], a.bd0[kk]);
      v = a.vc[kk];
    }
    wd0s[kk] = w;
    vcs[kk] = v;
  }
  if (tid == 0) {
    for (int s = 0; s < C::SLOTS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(a_ready, C::WORKER_WARPS);
    mbar_init(acc_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr int PER_TILE = 3 * C::KSTEPS;            // GEMM3, GEMM4a, GEMM4b
  constexpr int FIRST_SLOT = 3 * C::KSTEPS;          // their weights follow GEMM1 (2 slots / k-step) and GEMM2 in the image

  if (warp == 0) {
    // ===================== producer: backward-role weights; L2 prefetch of the next tile's saved blocks =====================
    constexpr int PL = C::SLOTS / 2;
    const long long first = tile_begin + blockIdx.x;
    const long long my_tiles = first < tile_end ? (tile_end - first + gridDim.x - 1) / gridDim.x : 0;
    if (lane < PL) {
      const long long total = my_tiles * PER_TILE;
      for (long long it = lane; it < total; it += PL) {
        const int j = (int)(it % PER_TILE);
        const int s = (int)(it % C::SLOTS);
        const uint32_t ph = (uint32_t)(it / C::SLOTS) & 1;
        mbar_wait(&empty[s], ph ^ 1, err);
        mbar_arrive_expect_tx(&full[s], C::SLOT_BYTES);
        bulk_g2s(stage0 + (size_t)s * C::SLOT_BYTES, packed + (size_t)(FIRST_SLOT + j) * C::SLOT_BYTES, C::SLOT_BYTES, &full[s]);
        if (j == 0) {   // one tile ahead: pull the saved a / acc1 / G blocks into L2 while this tile computes
          const long long nt = first + (it / PER_TILE + 1) * gridDim.x;
          if (nt < tile_end) {
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(sv_a + (size_t)nt * TB), "r"((uint32_t)TB) : "memory");
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(sv_s + (size_t)nt * TB), "r"((uint32_t)TB) : "memory");
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(sv_g + (size_t)nt * TB), "r"((uint32_t)TB) : "memory");
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (see vattn_bwd_oh_kernel) =====================
    {
      const uint32_t idesc = idesc_bf16(128, C::DP);
      constexpr uint32_t lbo_a = 128 * 16, lbo_b = C::DP * 16;
      constexpr uint64_t A_STEP = (2 * lbo_a) >> 4;
      const uint64_t ah0 = smem_desc(smem_u32(A_hi), lbo_a, 128), al0 = smem_desc(smem_u32(A_lo), lbo_a, 128);
      const uint64_t bh0 = smem_desc(smem_u32(stage0), lbo_b, 128);
      uint32_t slot = 0, slot_phase = 0, ready_phase = 0;
      for (long long tile = tile_begin + blockIdx.x; tile < tile_end; tile += gridDim.x) {
        for (int gi = 2; gi < 5; ++gi) {   // GEMM3 (-> acc0), GEMM4a (-> acc1), GEMM4b (acc1 +=)
          mbar_wait(a_ready, ready_phase, err);
          ready_phase ^= 1;
          tc_fence_after();
          const uint32_t d = tmem_base + (gi >= 3 ? C::ACC1_COL : 0);
          for (int ks = 0; ks < C::KSTEPS; ++ks) {
            mbar_wait(&full[slot], slot_phase, err);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t ah = ah0 + ks * A_STEP, al = al0 + ks * A_STEP;
              const uint64_t bh = bh0 + (uint64_t)slot * (C::SLOT_BYTES >> 4);
              mma_bf16(d, ah, bh, idesc, gi == 4 || ks > 0);
              mma_bf16(d, al, bh, idesc, true);
              mma_bf16(d, ah, bh + (C::SLAB >> 4), idesc, true);
              mma_commit(&empty[slot]);
            }
            if (++slot == C::SLOTS) { slot = 0; slot_phase ^= 1; }
          }
          if (elect_one()) mma_commit(acc_done);
        }
      }
    }
  } else {
    // ===================== workers =====================
    const int ww = warp - 2;
    const int wtid = tid - 64;
    const int quarter = warp & 3;
    const int part = ww >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t done_phase = 0;
    constexpr int NQ = (C::CHUNKS + C::NPART - 1) / C::NPART;
    constexpr int NE = (C::E_COLS / 8 + C::NPART - 1) / C::NPART;
    static_assert(NQ * 8 <= 64, "ReLU mask of G is kept in one 64-bit register");
    const int slabs = stg.lo ? 2 : 1;
    const uint32_t a_base = (uint32_t)((r >> 3) * 128 + (r & 7) * 16);
    const size_t st_row = (size_t)(r >> 4) * (size_t)(slabs * C::DP * 32) + (size_t)(r & 15) * 16;
    const size_t st_tile = (size_t)slabs * 256 * C::DP;
    const size_t ste_row = (size_t)(r >> 4) * (size_t)(C::E_COLS * 32) + (size_t)(r & 15) * 16;
    const size_t svg_row = (size_t)(r >> 4) * (size_t)(2 * C::DP * 32) + (size_t)(r & 15) * 16;   // saved G: always hi + lo
    const size_t svf_row = (size_t)r * 32;
    float cw0 = 0.f, cw1 = 0.f, cw2 = 0.f, cb = 0.f;

    auto wait_acc = [&]() {
      mbar_wait(acc_done, done_phase, err);
      done_phase ^= 1;
      tc_fence_after();
    };
    auto publish = [&]() {
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);
    };
    auto emit = [&](const float (&x)[8], int ch, bool to_a, unsigned char *stage_tile, uint4 &hi, uint4 &lo) {
      split2(x[0], x[1], hi.x, lo.x);
      split2(x[2], x[3], hi.y, lo.y);
      split2(x[4], x[5], hi.z, lo.z);
      split2(x[6], x[7], hi.w, lo.w);
      if (to_a) {
        *reinterpret_cast<uint4 *>(A_hi + a_base + ch * 2048) = hi;
        *reinterpret_cast<uint4 *>(A_lo + a_base + ch * 2048) = lo;
      }
      if (stage_tile) {
        unsigned char *p = stage_tile + st_row + (size_t)ch * 256;
        *reinterpret_cast<uint4 *>(p) = hi;
        if (stg.lo) *reinterpret_cast<uint4 *>(p + C::DP * 32) = lo;
      }
    };

    RowInfoPB ri = row_info_pb<C>(a, tile_begin + blockIdx.x, r, krows, tpb);
    for (long long tile = tile_begin + blockIdx.x; tile < tile_end; tile += gridDim.x) {
      const size_t tl = (size_t)(tile - tile_begin);
      const bool row_on = ri.c >= 0;
      const int b = (int)(tile / tpb);
      if (part == 0) rels[r] = make_float4(ri.rx, ri.ry, ri.rz, ri.flag);
      unsigned long long gmaskbits = 0ull;
      // ---- one-hot tile for the table-gradient jobs (staged only: no GEMM reads it here) ----------------------------------
#pragma unroll
      for (int q = 0; q < NE; ++q) {
        const int ch = part + q * C::NPART;
        if (ch < C::E_COLS / 8) {
          uint4 e = make_uint4(0u, 0u, 0u, 0u);
          if ((ri.j >> 3) == ch) {
            const uint32_t one = (ri.j & 1) ? 0x3F800000u : 0x00003F80u;
            const int w = (ri.j & 7) >> 1;
            e.x = w == 0 ? one : 0u; e.y = w == 1 ? one : 0u; e.z = w == 2 ? one : 0u; e.w = w == 3 ? one : 0u;
          }
          *reinterpret_cast<uint4 *>(stg.e + tl * (size_t)(256 * C::E_COLS) + ste_row + (size_t)ch * 256) = e;
        }
      }
      // ---- w = exp(a - max) / sum, s = acc1 + vc; ds = w * dout, da = ds * (s - out) ---------------------------------------
      //      a, acc1 and the ReLU mask come from the forward's saved blocks (L2-prefetched one tile ahead), the per-centre
      //      softmax inputs as in vattn_bwd_oh_kernel; everything is loaded one half-chunk ahead
      {
        const float *st_mx = stats + (size_t)(row_on ? ri.c : 0) * D;
        const float *st_iv = stats + ((size_t)BM + (row_on ? ri.c : 0)) * D;
        const float *p_go = dout + (size_t)(row_on ? ri.c : 0) * D;
        const float *p_o = out + (size_t)(row_on ? ri.c : 0) * D;
        const unsigned char *pa = sv_a + (size_t)tile * TB + svf_row;
        const unsigned char *ps = sv_s + (size_t)tile * TB + svf_row;
        const unsigned char *pg = sv_g + (size_t)tile * TB + svg_row;
        struct Half { float4 mx, iv, go, o, av, sv; };
        auto load_half = [&](int hc, Half &h) {
          const int ch = part + (hc >> 1) * C::NPART;
          const int col = ch * 8 + (hc & 1) * 4;
          h.mx = h.iv = h.go = h.o = h.av = h.sv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row_on && col < D) {
            h.mx = ldg4(st_mx + col); h.iv = ldg4(st_iv + col); h.go = ldg4(p_go + col); h.o = ldg4(p_o + col);
            h.av = __ldg(reinterpret_cast<const float4 *>(pa + (size_t)ch * 4096 + (hc & 1) * 16));
            h.sv = __ldg(reinterpret_cast<const float4 *>(ps + (size_t)ch * 4096 + (hc & 1) * 16));
          }
        };
        Half cur, nxt2;
        load_half(0, cur);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int ch = part + q * C::NPART;
          float ds[8], da[8];
          uint4 gh = make_uint4(0u, 0u, 0u, 0u);
          if (ch < C::CHUNKS && row_on) gh = __ldg(reinterpret_cast<const uint4 *>(pg + (size_t)ch * 256));
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            load_half(q * 2 + half + 1, nxt2);
            if (ch < C::CHUNKS) {
              const float4 v0 = *reinterpret_cast<const float4 *>(vcs + ch * 8 + half * 4);
              const float avs[4] = {cur.av.x, cur.av.y, cur.av.z, cur.av.w}, svs[4] = {cur.sv.x, cur.sv.y, cur.sv.z, cur.sv.w};
              const float mxs[4] = {cur.mx.x, cur.mx.y, cur.mx.z, cur.mx.w}, ivs[4] = {cur.iv.x, cur.iv.y, cur.iv.z, cur.iv.w};
              const float gos[4] = {cur.go.x, cur.go.y, cur.go.z, cur.go.w}, os[4] = {cur.o.x, cur.o.y, cur.o.z, cur.o.w};
              const float vv[4] = {v0.x, v0.y, v0.z, v0.w};
              const bool on = row_on && ch * 8 + half * 4 < D;
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int j = half * 4 + u;
                const float w = on ? __expf(avs[u] - mxs[u]) * ivs[u] : 0.f;
                ds[j] = w * gos[u];
                da[j] = ds[j] * (svs[u] + vv[u] - os[u]);
              }
            }
            cur = nxt2;
          }
          if (ch < C::CHUNKS) {
            const uint32_t gw[4] = {gh.x, gh.y, gh.z, gh.w};
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if ((gw[j >> 1] >> ((j & 1) * 16)) & 0x7fffu) gmaskbits |= 1ull << (q * 8 + j);
            uint4 hi, lo;
            emit(da, ch, true, stg.da + tl * st_tile, hi, lo);
            emit(ds, ch, false, stg.ds + tl * st_tile, hi, lo);
            const uint32_t pk[8] = {hi.x, hi.y, hi.z, hi.w, lo.x, lo.y, lo.z, lo.w};
            tmem_st8(trow + C::ACC1_COL + ch * 8, pk);
          }
        }
      }
      tmem_st_wait();
      publish();
      // ---- while GEMM3 runs: the next tile's row description -------------------------------------------------------------
      const RowInfoPB nxt = row_info_pb<C>(a, tile + gridDim.x, r, krows, tpb);
      // ---- GEMM3 done: A is free -> operand ds from its parked words (GEMM4a) ---------------------------------------------
      wait_acc();
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int ch = part + q * C::NPART;
        if (ch < C::CHUNKS) {
          uint32_t pk[8];
          tmem_ld8u(trow + C::ACC1_COL + ch * 8, pk);
          *reinterpret_cast<uint4 *>(A_hi + a_base + ch * 2048) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4 *>(A_lo + a_base + ch * 2048) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
      publish();
      // ---- dgp = dg * [g > 0] (reads acc0 while GEMM4a fills acc1): staged + parked in acc0's columns ---------------------------
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int ch = part + q * C::NPART;
        if (ch < C::CHUNKS) {
          float dg[8];
          tmem_ld8(trow + ch * 8, dg);
#pragma unroll
          for (int j = 0; j < 8; ++j) dg[j] = ((gmaskbits >> (q * 8 + j)) & 1ull) ? dg[j] : 0.f;
          uint4 hi, lo;
          emit(dg, ch, false, stg.dgp + tl * st_tile, hi, lo);
          const uint32_t pk[8] = {hi.x, hi.y, hi.z, hi.w, lo.x, lo.y, lo.z, lo.w};
          tmem_st8(trow + ch * 8, pk);
        }
      }
      tmem_st_wait();
      // ---- GEMM4a done: A is free -> operand dgp (GEMM4b) ----------------------------------------------------------------------
      wait_acc();
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int ch = part + q * C::NPART;
        if (ch < C::CHUNKS) {
          uint32_t pk[8];
          tmem_ld8u(trow + ch * 8, pk);
          *reinterpret_cast<uint4 *>(A_hi + a_base + ch * 2048) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4 *>(A_lo + a_base + ch * 2048) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
      publish();
      // ---- dpre = dh * [h > 0]; d rel; dpre -> fp32 scratch (aliases A, free once GEMM4b is done) ------------------------------
      wait_acc();
      {
        float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int ch = part + q * C::NPART;
          if (ch < C::CHUNKS) {
            float dh[8];
            tmem_ld8(trow + C::ACC1_COL + ch * 8, dh);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int col = ch * 8 + j;
              const float4 w0 = wd0s[col];
              const float pre = fmaf(w0.x, ri.rx, fmaf(w0.y, ri.ry, fmaf(w0.z, ri.rz, w0.w)));
              const float dp = (ri.flag != 0.f && pre > 0.f) ? dh[j] : 0.f;
              sx = fmaf(dp, w0.x, sx); sy = fmaf(dp, w0.y, sy); sz = fmaf(dp, w0.z, sz);
              if (col < D) scratch[(size_t)col * L::SCR_LD + ((r + col) & 127)] = dp;
            }
          }
        }
        relacc[(part * 3 + 0) * 128 + r] = sx;
        relacc[(part * 3 + 1) * 128 + r] = sy;
        relacc[(part * 3 + 2) * 128 + r] = sz;
      }
      tc_fence_before();
      asm volatile("bar.sync 1, %0;" ::"n"(C::WORKER_WARPS * 32) : "memory");
      if (wtid < D) {
        const float *colp = scratch + (size_t)wtid * L::SCR_LD;
#pragma unroll 4
        for (int rr = 0; rr < 128; ++rr) {
          const float dp = colp[(rr + wtid) & 127];
          const float4 rl = rels[rr];
          cw0 = fmaf(dp, rl.x, cw0); cw1 = fmaf(dp, rl.y, cw1); cw2 = fmaf(dp, rl.z, cw2); cb += dp;
        }
      }
      if (part == 0 && ri.flag != 0.f && (g.d_xyz_c || g.d_xyz_n)) {
        float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
        for (int pp = 0; pp < C::NPART; ++pp) {
          sx += relacc[(pp * 3 + 0) * 128 + r];
          sy += relacc[(pp * 3 + 1) * 128 + r];
          sz += relacc[(pp * 3 + 2) * 128 + r];
        }
        if (g.d_xyz_c) {
          float *dst = g.d_xyz_c + (size_t)ri.c * 3;
          atomicAdd(dst, a.sign * sx); atomicAdd(dst + 1, a.sign * sy); atomicAdd(dst + 2, a.sign * sz);
        }
        if (g.d_xyz_n) {
          float *dst = g.d_xyz_n + ((size_t)b * a.N + ri.j) * 3;
          atomicAdd(dst, -a.sign * sx); atomicAdd(dst + 1, -a.sign * sy); atomicAdd(dst + 2, -a.sign * sz);
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(C::WORKER_WARPS * 32) : "memory");
      ri = nxt;
    }
    if (wtid < D) {
      if (g.d_wd0) {
        atomicAdd(g.d_wd0 + wtid * 3 + 0, cw0); atomicAdd(g.d_wd0 + wtid * 3 + 1, cw1); atomicAdd(g.d_wd0 + wtid * 3 + 2, cw2);
      }
      if (g.d_bd0) atomicAdd(g.d_bd0 + wtid, cb);
    }
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// dT1 / dT2 [B][E_COLS][D] (rows <= N valid) -> d_kp, d_gq, d_pc / d_vp, d_gv, d_vc (see the table definitions in
// vattn_fwd_oh_kernel): one thread per (shape, column)
template <class C>
__global__ void finalize_tables_kernel(const float *__restrict__ dt1, const float *__restrict__ dt2, const nsdp_vattn_grads g,
                                       int B, int N, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int b = i / D, c = i - b * D;
  const float *t1 = dt1 + (size_t)b * C::E_COLS * D + c, *t2 = dt2 + (size_t)b * C::E_COLS * D + c;
  float s1 = 0.f, s2 = 0.f;
  for (int j = 0; j < N; ++j) {
    const float x1 = t1[(size_t)j * D], x2 = t2[(size_t)j * D];
    s1 += x1; s2 += x2;
    if (g.d_kp) g.d_kp[((size_t)b * N + j) * D + c] -= x1;
    if (g.d_vp) g.d_vp[((size_t)b * N + j) * D + c] += x2;
  }
  if (g.d_gq) g.d_gq[(size_t)b * D + c] += t1[(size_t)N * D];
  if (g.d_gv) g.d_gv[(size_t)b * D + c] += t2[(size_t)N * D];
  if (g.d_pc) atomicAdd(g.d_pc + c, s1);
  if (g.d_vc) atomicAdd(g.d_vc + c, s2);
}

// Staging precision of the operand tiles that feed the weight / table gradient reductions. Default: bf16 hi + lo
// (fp32-grade). NSDP_STAGE_LO=0 stages only the hi half (plain bf16 operands): half the staged traffic, but a measured
// ~2e-3 relative error on weight gradients whose per-row terms cancel (tools/big_grad_check.py), so it is opt-in.
static bool stage_lo_for(long long /*pair_rows*/) {
  static const int forced = [] { const char *e = getenv("NSDP_STAGE_LO"); return e ? atoi(e) : 1; }();
  return forced != 0 && !stage16::enabled();
}

static bool no_saved() {
  static const bool v = [] { const char *e = getenv("NSDP_NO_SAVED"); return e && atoi(e) != 0; }();
  return v;
}

template <class C>
static size_t bwd_oh_workspace_bytes(const nsdp_vattn_args &a) {
  const int tpb = (a.M + C::CENTRES - 1) / C::CENTRES;
  const long long tiles = (long long)a.B * tpb;
  const long long seg = tiles < kSegmentTiles ? tiles : kSegmentTiles;
  const size_t slabs = stage_lo_for(tiles * 128) ? 2 : 1;
  return bwd_packed_bytes<C>() + (size_t)a.B * table_bytes_per_shape<C>() + 256 +
         2 * sizeof(float) * (size_t)a.B * C::E_COLS * a.D + (size_t)seg * (5 * slabs * 256 * C::DP + 256 * C::E_COLS);
}

template <class C>
static int launch_bwd_oh(const nsdp_vattn_args &a, const float *out, const float *stats, const float *dout,
                         const nsdp_vattn_grads &g, void *workspace, size_t ws_bytes, cudaStream_t st) {
  if (!workspace || ws_bytes < bwd_oh_workspace_bytes<C>(a)) return NSDP_ERR_WORKSPACE;
  if (!g.d_wd2t || !g.d_wpt || !g.d_wg2t) return NSDP_ERR_INVALID_ARGUMENT;
  const int tpb = (a.M + C::CENTRES - 1) / C::CENTRES;
  const long long tiles = (long long)a.B * tpb;
  const long long seg = tiles < kSegmentTiles ? tiles : kSegmentTiles;
  const int lo = stage_lo_for(tiles * 128) ? 1 : 0;
  const size_t slabs = lo ? 2 : 1;
  unsigned char *packed = (unsigned char *)workspace;
  unsigned char *tables = packed + bwd_packed_bytes<C>();
  int *err = (int *)(tables + (size_t)a.B * table_bytes_per_shape<C>());
  float *dt1 = (float *)((unsigned char *)err + 256);
  const size_t tbl = (size_t)a.B * C::E_COLS * a.D;
  float *dt2 = dt1 + tbl;
  unsigned char *sbase = (unsigned char *)(dt2 + tbl);
  const size_t per = (size_t)seg * slabs * 256 * C::DP;
  const int f16 = (!lo && stage16::enabled()) ? 1 : 0;
  unsigned *gmax = (unsigned *)err + 16;     // inside the zeroed 256-byte header
  OhStaging stg{sbase, sbase + per, sbase + 2 * per, sbase + 3 * per, sbase + 4 * per, sbase + 5 * per, lo, f16, gmax};
  cudaError_t e = cudaMemsetAsync(err, 0, 256 + 2 * tbl * sizeof(float), st);
  if (e != cudaSuccess) return cuda_rc(e);
  if (f16) {
    const int r0 = stage16::launch_absmax(dout, (size_t)a.B * a.M * a.D, gmax, st);
    if (r0 != NSDP_OK) return r0;
  }
  pack_bwd_weights_kernel<C><<<96, 256, 0, st>>>(a.wpt, a.wd2t, a.wg2t, a.D, packed);
  int rc = check_launch();
  if (rc != NSDP_OK) return rc;
  pack_tables_kernel<C><<<128, 256, 0, st>>>(a, tables);
  rc = check_launch();
  if (rc != NSDP_OK) return rc;
  const bool use_saved = a.saved && a.saved_bytes >= saved_bytes_total<C>(tiles) && !no_saved() && !f16;
  const unsigned char *saved = (const unsigned char *)a.saved;
  auto kern = vattn_bwd_oh_kernel<C>;
  auto kern_sv = vattn_bwd_sv_kernel<C>;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdLayout<C>::SMEM);
  if (e != cudaSuccess) return cuda_rc(e);
  e = cudaFuncSetAttribute(kern_sv, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdLayout<C>::SMEM);
  if (e != cudaSuccess) return cuda_rc(e);
  constexpr size_t TB = saved_tile_bytes<C>();
  for (long long t0 = 0; t0 < tiles; t0 += seg) {
    const long long t1 = t0 + seg < tiles ? t0 + seg : tiles;
    const long long n = t1 - t0;
    const int grid = (int)(n < num_sms() ? n : num_sms());
    if (use_saved) {
      kern_sv<<<grid, C::THREADS, BwdLayout<C>::SMEM, st>>>(a, out, stats, dout, g, packed, saved, tiles, stg, tpb, t0, t1, err);
    } else {
      unsigned long long *trace = nullptr;
#ifdef NSDP_TRACE
      if (const char *tp = getenv("NSDP_TRACE_BWD_PTR")) trace = (unsigned long long *)strtoull(tp, nullptr, 0);
#endif
      kern<<<grid, C::THREADS, BwdLayout<C>::SMEM, st>>>(a, out, stats, dout, g, packed, tables, stg, tpb, t0, t1, err, trace);
    }
    rc = check_launch();
    if (rc != NSDP_OK) return rc;
    // weight gradients over the whole segment + table gradients per shape (tile ranges relative to the segment)
    dwtc::Job jobs[dwtc_max_jobs];
    int nj = 0;
    auto flush = [&]() -> int {
      const int r2 = nj ? dw_tc_launch(jobs, nj, n, err, st) : NSDP_OK;
      nj = 0;
      return r2;
    };
    auto wjob = [&](const unsigned char *x, int xlo, const unsigned char *y, float *o) {
      dwtc::Job j{x, y, o, C::DP, C::DP, a.D, a.D, a.D, nullptr, nullptr, 0, 0, 0};
      j.x_lo = xlo; j.y_lo = lo;
      j.f16 = f16; j.gmax = f16 ? gmax : nullptr;
      return j;
    };
    // H and G: from this segment's staging, or (saved variant) straight from the forward's buffer, always hi + lo there
    const unsigned char *xh = use_saved ? saved + (size_t)t0 * TB : stg.h;
    const unsigned char *xg = use_saved ? saved + (size_t)(tiles + t0) * TB : stg.g;
    const int xlo = use_saved ? 1 : lo;
    jobs[nj++] = wjob(xg, xlo, stg.da, g.d_wg2t);
    jobs[nj++] = wjob(xh, xlo, stg.dgp, g.d_wpt);
    jobs[nj++] = wjob(xh, xlo, stg.ds, g.d_wd2t);
    // chunk-aligned launch first: weight and table jobs walk the same tile chunks, so dGP / dS / H are read from HBM once
    {
      const long long b0 = t0 / tpb;
      long long bounds[40];
      int ns = 0;
      bounds[0] = 0;
      for (long long b = b0; b * tpb < t1 && ns < 38; ++b) bounds[++ns] = ((b + 1) * tpb < t1 ? (b + 1) * tpb : t1) - t0;
      if (bounds[ns] == n) {
        dwtc::Job cj[5] = {jobs[0], jobs[1], jobs[2], {}, {}};
        for (int m = 0; m < 2; ++m) {
          dwtc::Job j{stg.e, m == 0 ? stg.dgp : stg.ds, (m == 0 ? dt1 : dt2) + (size_t)b0 * C::E_COLS * a.D, C::E_COLS, C::DP,
                      a.N + 1, a.D, a.D, nullptr, nullptr, 0, 0, 0};
          j.x_lo = 0; j.y_lo = lo; j.out_shape_stride = (long long)C::E_COLS * a.D;
          j.f16 = f16; j.gmax = f16 ? gmax : nullptr;
          cj[3 + m] = j;
        }
        rc = dw_tc_launch_chunked(cj, 5, n, bounds, ns, err, st);
        if (rc == NSDP_OK) continue;
        if (rc != NSDP_ERR_UNSUPPORTED) return rc;
      }
    }
    for (long long b = t0 / tpb; b * tpb < t1; ++b) {
      const long long r0 = (b * tpb > t0 ? b * tpb : t0) - t0, r1 = ((b + 1) * tpb < t1 ? (b + 1) * tpb : t1) - t0;
      for (int m = 0; m < 2; ++m) {
        if (nj == dwtc_max_jobs) {
          rc = flush();
          if (rc != NSDP_OK) return rc;
        }
        dwtc::Job j{stg.e, m == 0 ? stg.dgp : stg.ds, (m == 0 ? dt1 : dt2) + (size_t)b * C::E_COLS * a.D, C::E_COLS, C::DP,
                    a.N + 1, a.D, a.D, nullptr, nullptr, 0, 0, 0};
        j.x_lo = 0; j.y_lo = lo; j.t0 = r0; j.t1 = r1;
        j.f16 = f16; j.gmax = f16 ? gmax : nullptr;
        jobs[nj++] = j;
      }
    }
    rc = flush();
    if (rc != NSDP_OK) return rc;
  }
  finalize_tables_kernel<C><<<(unsigned)ceil_div(a.B * a.D, 128), 128, 0, st>>>(dt1, dt2, g, a.B, a.N, a.D);
  return check_launch();
}

static bool bwd_oh_ok(const nsdp_vattn_args &a) {
  static const bool off = [] { const char *e = getenv("NSDP_NO_ONEHOT"); return e && atoi(e) != 0; }();
  return !off && a.has_global && !a.qp && a.kp && a.vp && a.gq && a.gv && a.idx && a.N + 1 <= 112 && a.K + 1 <= 8 && a.M >= 16;
}

static int pick_bwd(const nsdp_vattn_args &a) {
  const int krows = a.K + (a.has_global ? 1 : 0);
  if (a.D % 4 != 0 || a.D > 256 || !a.kp || !a.vp) return 0;
  if (krows <= 8 && a.D > 128 && a.D <= 208 && (!a.has_global || a.M >= 16)) return 1;
  if (krows <= 16 && !a.has_global) return a.D <= 128 ? 2 : 3;
  if (krows <= 128 && !a.has_global && a.D > 128) return 4;
  return 0;
}

}  // namespace vtc

#ifndef NSDP_BWD_NPART
#define NSDP_BWD_NPART 4
#endif
constexpr int BWD_NPART = NSDP_BWD_NPART;  // worker warps per TMEM lane quarter in the backward chain kernel

size_t vattn_bwd_tc_workspace_bytes(const nsdp_vattn_args *a) {
  switch (vtc::pick_bwd(*a)) {
    case 1:
      if (vtc::bwd_oh_ok(*a)) return vtc::bwd_oh_workspace_bytes<vtc::TcCfg<208, 8, 4, true>>(*a);
      return vtc::bwd_workspace_bytes<vtc::TcCfg<208, 8, BWD_NPART>>(*a);
    case 2: return vtc::bwd_workspace_bytes<vtc::TcCfg<128, 16, BWD_NPART>>(*a);
    case 3: return vtc::bwd_workspace_bytes<vtc::TcCfg<256, 16, BWD_NPART>>(*a);
    case 4: return vtc::bwd_workspace_bytes<vtc::TcCfg<256, 128, BWD_NPART>>(*a);
    default: return 0;
  }
}

int vattn_bwd_tc_dispatch(const nsdp_vattn_args *a, const float *out, const float *stats, const float *dout,
                          const nsdp_vattn_grads *g, void *workspace, size_t ws_bytes, cudaStream_t st, bool *handled) {
  *handled = true;
  switch (vtc::pick_bwd(*a)) {
    case 1:
      if (vtc::bwd_oh_ok(*a))
        return vtc::launch_bwd_oh<vtc::TcCfg<208, 8, 4, true>>(*a, out, stats, dout, *g, workspace, ws_bytes, st);
      return vtc::launch_bwd<vtc::TcCfg<208, 8, BWD_NPART>>(*a, out, stats, dout, *g, workspace, ws_bytes, st);
    case 2: return vtc::launch_bwd<vtc::TcCfg<128, 16, BWD_NPART>>(*a, out, stats, dout, *g, workspace, ws_bytes, st);
    case 3: return vtc::launch_bwd<vtc::TcCfg<256, 16, BWD_NPART>>(*a, out, stats, dout, *g, workspace, ws_bytes, st);
    case 4: return vtc::launch_bwd<vtc::TcCfg<256, 128, BWD_NPART>>(*a, out, stats, dout, *g, workspace, ws_bytes, st);
    default: *handled = false; return NSDP_OK;
  }
}

}  // namespace nsdp

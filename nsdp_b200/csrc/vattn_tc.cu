// Fused vector attention forward on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same math and same C-ABI arguments as the CUDA-core kernel in vattn.cu (see nsdp_vattn_args); this is the
// fast path for the shapes that dominate the model (decoder cross-attention: D = 200, 7 neighbours + global
// token; encoder blocks with 16 neighbours).
//
// A persistent CTA (one per SM) walks over tiles of 128 pair rows = 128/KR centres x KR rows. Per tile, with
// every [128 x D] activation living only in shared memory / TMEM:
//
//   workers   H = relu(Wd0*rel + b)  (K = 3, CUDA cores)  -> bf16 hi/lo A operand in smem
//   tensor    [gp | dl] = H * [W' ; Wd2]^T                 -> TMEM acc0, acc1   (GEMM1, N = 2*DP)
//   workers   G = relu(gp + P)                             -> A operand (overwrites H)
//   tensor    a = G * Wg2^T                                -> TMEM acc0         (GEMM2)
//   workers   w = softmax over the KR rows of a centre, out = sum w * (V + dl)  -> global
//
// Warp roles: warp 0 streams pre-packed weight slabs from L2 with 1-D bulk async copies into a 4-stage
// mbarrier ring; warp 1 (one elected lane) issues tcgen05.mma and commits completions to mbarriers; warps 2..9
// are the workers (each owns one TMEM lane = one pair row, two warps per lane quarter split the columns).
//
// Precision: fp32 operands are split into bf16 hi + lo and every product is hi*hi + lo*hi + hi*lo accumulated in
// fp32 (umma.cuh); the result matches the fp32 CUDA-core kernel to ~1e-6 relative.
#include <math.h>
#include <stdlib.h>

#include "vattn_tc_common.cuh"

namespace nsdp {
namespace vtc {

using namespace umma;

// Packed weight image (global memory), produced by pack_weights_kernel:
//   GEMM1 region: for ks in [0, KSTEPS): [W' hi slab][W' lo slab][Wd2 hi slab][Wd2 lo slab]
//   GEMM2 region: for ks in [0, KSTEPS): [Wg2 hi slab][Wg2 lo slab]
// slab(ks) = B[:, 16ks : 16ks+16] of the (N = DP) x (K = DP) operand B[n][k] = Wt[k][n], canonical no-swizzle layout.
template <class C>
constexpr size_t packed_bytes() {
  return (size_t)C::KSTEPS * 6 * C::SLAB;
}

template <class C>
__global__ void pack_weights_kernel(const float *__restrict__ wpt, const float *__restrict__ wd2t,
                                    const float *__restrict__ wg2t, int D, unsigned char *__restrict__ out, int pair = 0) {
  // one thread per (matrix m, n, k-pair)
  const int total = 3 * C::DP * (C::DP / 2);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int m = e / (C::DP * (C::DP / 2));
    const int rem = e - m * (C::DP * (C::DP / 2));
    const int n = rem / (C::DP / 2), k = (rem - n * (C::DP / 2)) * 2;
    const float *src = m == 0 ? wpt : (m == 1 ? wd2t : wg2t);
    float x0 = 0.f, x1 = 0.f;
    if (n < D && k < D) x0 = src[(size_t)k * D + n];
    if (n < D && k + 1 < D) x1 = src[(size_t)(k + 1) * D + n];
    uint32_t hi, lo;
    split2(x0, x1, hi, lo);
    const int ks = k >> 4, kin = k & 15;
    size_t base;
    if (m < 2)
      base = (size_t)ks * 4 * C::SLAB + (size_t)m * 2 * C::SLAB;
    else
      base = (size_t)C::KSTEPS * 4 * C::SLAB + (size_t)ks * 2 * C::SLAB;
    *reinterpret_cast<uint32_t *>(out + base + slot_off<C>(n, kin, 0, pair)) = hi;
    *reinterpret_cast<uint32_t *>(out + base + slot_off<C>(n, kin, 1, pair)) = lo;
  }
}

template <class C>
__global__ void __launch_bounds__(C::THREADS, 1)
vattn_fwd_tc_kernel(const nsdp_vattn_args a, const unsigned char *__restrict__ packed, float *__restrict__ out,
                    float *__restrict__ stats, long long tiles, int *err) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char *A_hi = smem + C::OFF_A;
  unsigned char *A_lo = A_hi + C::A_HALF;
  unsigned char *stage0 = smem + C::OFF_STAGE;
  float4 *wd0s = reinterpret_cast<float4 *>(smem + C::OFF_WD0);
  float *pcs = reinterpret_cast<float *>(smem + C::OFF_PC);
  float *vcs = reinterpret_cast<float *>(smem + C::OFF_VC);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::OFF_BAR);
  uint64_t *full = bars;                    // [SLOTS]
  uint64_t *empty = bars + C::SLOTS;        // [SLOTS]
  uint64_t *a_ready = bars + 2 * C::SLOTS;  // workers -> MMA (count = worker warps)
  uint64_t *acc_done = a_ready + 1;         // MMA -> workers (tcgen05.commit)
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_done + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = a.D;
  const int krows = a.K + (a.has_global ? 1 : 0);

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
    // the packed image is consumed front to back once per tile: 2 slots per k-step of GEMM1, then 1 slot per k-step of
    // GEMM2; PL lanes share the copies (lane l serves slots l, l + PL, ...), see vattn_fwd_oh_kernel
    constexpr int PL = C::SLOTS / 2;
    constexpr int PER_TILE = 3 * C::KSTEPS;
    if (lane < PL) {
      const long long my_tiles = (long long)blockIdx.x < tiles ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
      const long long total = my_tiles * PER_TILE;
      for (long long it = lane; it < total; it += PL) {
        const int j = (int)(it % PER_TILE);
        const int s = (int)(it % C::SLOTS);
        const uint32_t ph = (uint32_t)(it / C::SLOTS) & 1;
        mbar_wait(&empty[s], ph ^ 1, err);  // first round passes immediately (fresh barrier, parity trick)
        mbar_arrive_expect_tx(&full[s], C::SLOT_BYTES);
        bulk_g2s(stage0 + (size_t)s * C::SLOT_BYTES, packed + (size_t)j * C::SLOT_BYTES, C::SLOT_BYTES, &full[s]);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // whole warp runs loops and waits, one elected lane issues (see vattn_fwd_oh_kernel)
    {
      const uint32_t idesc = idesc_bf16(128, C::DP);
      constexpr uint32_t lbo_a = 128 * 16, lbo_b = C::DP * 16;
      constexpr uint64_t A_STEP = (2 * lbo_a) >> 4;
      const uint64_t ah0 = smem_desc(smem_u32(A_hi), lbo_a, 128), al0 = smem_desc(smem_u32(A_lo), lbo_a, 128);
      const uint64_t bh0 = smem_desc(smem_u32(stage0), lbo_b, 128);
      uint32_t slot = 0, slot_phase = 0, ready_phase = 0;
      for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        for (int g = 0; g < 2; ++g) {
          mbar_wait(a_ready, ready_phase, err);
          ready_phase ^= 1;
          tc_fence_after();
          for (int ks = 0; ks < C::KSTEPS; ++ks) {
            // GEMM1: slot 0 = W' -> acc0, slot 1 = Wd2 -> acc1 ; GEMM2: one slot, Wg2 -> acc0
            for (int m = 0; m < (g == 0 ? 2 : 1); ++m) {
              mbar_wait(&full[slot], slot_phase, err);
              tc_fence_after();
              if (elect_one()) {
                const uint64_t ah = ah0 + ks * A_STEP, al = al0 + ks * A_STEP;
                const uint64_t bh = bh0 + (uint64_t)slot * (C::SLOT_BYTES >> 4);
                const uint32_t d = tmem_base + (m ? C::ACC1_COL : 0);
                mma_bf16(d, ah, bh, idesc, ks > 0);
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
    const int quarter = warp & 3;            // TMEM lane quarter this warp may touch
    const int part = ww >> 2;                // which slice of the columns
    const int r = quarter * 32 + lane;       // pair row inside the tile == TMEM lane
    const int ch0 = part * C::CHUNKS / C::NPART, ch1 = (part + 1) * C::CHUNKS / C::NPART;
    const int nch = ch1 - ch0;               // 8-column chunks owned by this thread (<= MAXCH)
    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t done_phase = 0;
    const long long BM = (long long)a.B * a.M;

    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const RowInfo ri = row_info<C>(a, tile, r, krows);
      const bool row_on = ri.c >= 0;
      const bool is_glob = row_on && ri.n < 0;
      // ---- H operand ---------------------------------------------------------------------------------
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
          uint4 hi, lo;
          split2(h[0], h[1], hi.x, lo.x);
          split2(h[2], h[3], hi.y, lo.y);
          split2(h[4], h[5], hi.z, lo.z);
          split2(h[6], h[7], hi.w, lo.w);
          const uint32_t off = canon_off(128, r, k0);
          *reinterpret_cast<uint4 *>(A_hi + off) = hi;
          *reinterpret_cast<uint4 *>(A_lo + off) = lo;
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);

      // ---- gather P = pc + qp - kp (or gq for the global row) for this thread's columns while GEMM1 runs ---------
      float P[C::MAXCH][8];
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
            P[q][j] = p.x; P[q][j + 1] = p.y; P[q][j + 2] = p.z; P[q][j + 3] = p.w;
          }
        }
      }

      // ---- epilogue 1: G = relu(gp + P) -> A operand ------------------------------------------------------------
      mbar_wait(acc_done, done_phase, err);
      done_phase ^= 1;
      tc_fence_after();
#pragma unroll
      for (int q = 0; q < C::MAXCH; ++q) {
        if (q < nch) {
          const int k0 = (ch0 + q) * 8;
          float v[8];
          tmem_ld8(trow + k0, v);
          float g[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] = (row_on && k0 + j < D) ? fmaxf(v[j] + P[q][j], 0.f) : 0.f;
          uint4 hi, lo;
          split2(g[0], g[1], hi.x, lo.x);
          split2(g[2], g[3], hi.y, lo.y);
          split2(g[4], g[5], hi.z, lo.z);
          split2(g[6], g[7], hi.w, lo.w);
          const uint32_t off = canon_off(128, r, k0);
          *reinterpret_cast<uint4 *>(A_hi + off) = hi;
          *reinterpret_cast<uint4 *>(A_lo + off) = lo;
        }
      }
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);

      // ---- gather V = vc + vp (or gv) while GEMM2 runs (re-uses the P registers) ----------------------------------
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
            P[q][j] = t.x; P[q][j + 1] = t.y; P[q][j + 2] = t.z; P[q][j + 3] = t.w;
          }
        }
      }

      // ---- epilogue 2: softmax over the KR rows of a centre, out = sum w * (V + dl) ---------------------------------
      mbar_wait(acc_done, done_phase, err);
      done_phase ^= 1;
      tc_fence_after();
      const long long ci = tile * C::CENTRES + r / C::KR;   // centre of this lane's group
      if constexpr (C::KR < 128) {
        const int gl = lane & (C::KR - 1);
#pragma unroll
        for (int q = 0; q < C::MAXCH; ++q) {
          if (q < nch) {
            const int k0 = (ch0 + q) * 8;
            float av[8], dl[8];
            tmem_ld8(trow + k0, av);
            tmem_ld8(trow + C::ACC1_COL + k0, dl);
            float e[8], es[8], mxv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float x = row_on ? av[j] : -INFINITY;
              float mx = x;
#pragma unroll
              for (int off = 1; off < C::KR; off <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
              mxv[j] = mx;
              const float ex = row_on ? __expf(x - mx) : 0.f;
              const float sv = is_glob ? P[q][j] : P[q][j] + dl[j];
              e[j] = ex;
              es[j] = ex * sv;
            }
            const float se = group_transpose_sum<C::KR>(e, lane);
            const float ses = group_transpose_sum<C::KR>(es, lane);
            if (gl < 8) {
              const int col = k0 + gl;
              if (ci < BM && col < D) {
                const float inv = 1.f / se;
                out[ci * D + col] = ses * inv;
                if (stats) {
                  float m = mxv[0];
#pragma unroll
                  for (int j = 1; j < 8; ++j) m = (gl == j) ? mxv[j] : m;
                  stats[ci * D + col] = m;
                  stats[(BM + ci) * D + col] = inv;
                }
              }
            }
          }
        }
      } else {
        // one centre per tile: the softmax spans all four lane quarters -> warp reduction, then an exchange
        // through shared memory (two named barriers per tile)
        float *red = reinterpret_cast<float *>(smem + C::OFF_RED);   // [3][4][DP]: max, sum e, sum e*s
#pragma unroll
        for (int q = 0; q < C::MAXCH; ++q) {
          if (q < nch) {
            const int k0 = (ch0 + q) * 8;
            float av[8];
            tmem_ld8(trow + k0, av);
#pragma unroll
            for (int j = 0; j < 8; ++j) av[j] = row_on ? av[j] : -INFINITY;
            const float m = group_transpose_max<32>(av, lane);
            if (lane < 8) red[(0 * 4 + quarter) * C::DP + k0 + lane] = m;
          }
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");
#pragma unroll
        for (int q = 0; q < C::MAXCH; ++q) {
          if (q < nch) {
            const int k0 = (ch0 + q) * 8;
            float av[8], dl[8], e[8], es[8];
            tmem_ld8(trow + k0, av);
            tmem_ld8(trow + C::ACC1_COL + k0, dl);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int col = k0 + j;
              const float mx = fmaxf(fmaxf(red[0 * C::DP + col], red[1 * C::DP + col]),
                                     fmaxf(red[2 * C::DP + col], red[3 * C::DP + col]));
              const float ex = row_on ? __expf(av[j] - mx) : 0.f;
              e[j] = ex;
              es[j] = ex * (P[q][j] + dl[j]);
            }
            const float se = group_transpose_sum<32>(e, lane);
            const float ses = group_transpose_sum<32>(es, lane);
            if (lane < 8) {
              red[(1 * 4 + quarter) * C::DP + k0 + lane] = se;
              red[(2 * 4 + quarter) * C::DP + k0 + lane] = ses;
            }
          }
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (quarter == 0 && ci < BM) {
#pragma unroll
          for (int q = 0; q < C::MAXCH; ++q) {
            if (q < nch && lane < 8) {
              const int col = (ch0 + q) * 8 + lane;
              if (col < D) {
                float mx = -INFINITY, se = 0.f, ses = 0.f;
#pragma unroll
                for (int w4 = 0; w4 < 4; ++w4) {
                  mx = fmaxf(mx, red[(0 * 4 + w4) * C::DP + col]);
                  se += red[(1 * 4 + w4) * C::DP + col];
                  ses += red[(2 * 4 + w4) * C::DP + col];
                }
                const float inv = 1.f / se;
                out[ci * D + col] = ses * inv;
                if (stats) {
                  stats[ci * D + col] = mx;
                  stats[(BM + ci) * D + col] = inv;
                }
              }
            }
          }
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");   // `red` is rewritten by the next tile
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}


// =====================================================================================================================
// OH variant (decoder cross-attention, C::OH): no per-row gathers. GEMM1 is extended by E_KSTEPS k-steps over the
// one-hot operand E (row r has a single 1 in column j_r = neighbour index, or N for the global-token row):
//     acc0 = H * W'^T + E * T1_b,   T1_b[j] = -kp[b][j] (j < N),  T1_b[N] = gq[b] - pc      ->  G = relu(acc0 + pc)
//     acc1 = H * Wd2^T + E * T2_b,  T2_b[j] =  vp[b][j] (j < N),  T2_b[N] = gv[b] - vc      ->  s = acc1 + vc
// (H = 0 on the global row, so that row sees exactly gq / gv.) Tiles never straddle shapes (row_info_pb), the tables of
// shape b are pre-packed like the weights and stream through the same slot ring. E is exact in bf16: 2 MMAs per k-step.
// =====================================================================================================================
template <class C>
__global__ void __launch_bounds__(C::THREADS, 1)
vattn_fwd_oh_kernel(const nsdp_vattn_args a, const unsigned char *__restrict__ packed,
                    const unsigned char *__restrict__ tables, float *__restrict__ out, float *__restrict__ stats,
                    unsigned char *__restrict__ saved, int tpb, long long tiles, int *err,
                    unsigned long long *trace) {
  // Schedule of one tile t (the tensor pipe trails the workers column chunk by column chunk, and the s-half of the NEXT
  // tile's GEMM1 runs underneath this tile's softmax epilogue):
  //   workers                                              tensor
  //   gp = acc0 -> registers, release acc0                 .
  //   G = relu(gp + pc) -> A, k-step by k-step             GEMM2 (acc0 = G Wg2^T), starts on the first finished k-step
  //   s = acc1 -> registers, release acc1                  .
  //   E(t+1), H(t+1) -> E buffer / A, k-step by k-step     GEMM1b(t+1) (acc1 = H Wd2^T + E T2)
  //   softmax over the centre's rows of a = acc0; out      GEMM1b(t+1) continues
  //   release acc0                                         GEMM1a(t+1) (acc0 = H W'^T + E T1)
  // G overwrites H in place, so epilogue 1 must wait for BOTH halves of GEMM1: the s-half goes first.
  static_assert(C::OH && C::KR == 8, "one-hot kernel: 8 rows per centre");
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char *A_hi = smem + C::OFF_A;
  unsigned char *A_lo = A_hi + C::A_HALF;
  unsigned char *E = smem + C::OFF_E;
  unsigned char *stage0 = smem + C::OFF_STAGE;
  float4 *wd0s = reinterpret_cast<float4 *>(smem + C::OFF_WD0);
  float *pcs = reinterpret_cast<float *>(smem + C::OFF_PC);
  float *vcs = reinterpret_cast<float *>(smem + C::OFF_VC);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::OFF_BAR);
  // kready[ks]: the 16 operand columns of k-step ks are in place (8 warps: 2 chunks x 4 lane quarters); e_ready: the
  // one-hot operand; free0 / free1: every worker is done reading acc0 / acc1, the next GEMM into it may start
  uint64_t *full = bars, *empty = bars + C::SLOTS, *acc_done = bars + 2 * C::SLOTS, *kready = acc_done + 1;
  uint64_t *e_ready = kready + C::KSTEPS, *free0 = e_ready + 1, *free1 = free0 + 1;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(free1 + 1);
  static_assert((2 * C::SLOTS + C::KSTEPS + 4) * 8 + 4 <= 256, "mbarrier area");

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = a.D;
  const int krows = a.K + 1;

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
    mbar_init(free0, C::WORKER_WARPS);
    mbar_init(free1, C::WORKER_WARPS);
    mbar_init(acc_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr int KS = C::KSTEPS, ES = C::E_KSTEPS;
  constexpr int PER_TILE = 3 * KS + 2 * ES;   // slots: Wd2, T2, W', T1, Wg2

  if (warp == 0) {
    // ===================== producer: weights + this shape's tables, one slot at a time =====================
    // PL lanes share the work (lane l serves slots l, l + PL, ...): a single thread sustains only about one bulk copy
    // per ~500 cycles (serial wait / expect_tx / issue chain), less than the tensor pipe consumes.
    // Packed images: GEMM1 region slot 2 ks + m (m = 0: W', 1: Wd2), tables slot 2 ks + m (0: T1, 1: T2), then Wg2.
    constexpr int PL = C::SLOTS / 2;
    if (lane < PL) {
      const long long my_tiles = (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
      const long long total = my_tiles * PER_TILE;
      for (long long it = lane; it < total; it += PL) {
        const long long tile = blockIdx.x + (it / PER_TILE) * gridDim.x;
        const int j = (int)(it % PER_TILE);
        const unsigned char *tb = tables + (size_t)(tile / tpb) * table_bytes_per_shape<C>();
        const unsigned char *src;
        if (j < KS) src = packed + (size_t)(2 * j + 1) * C::SLOT_BYTES;                                // Wd2
        else if (j < KS + ES) src = tb + (size_t)(2 * (j - KS) + 1) * C::SLOT_BYTES;                   // T2
        else if (j < 2 * KS + ES) src = packed + (size_t)(2 * (j - KS - ES)) * C::SLOT_BYTES;          // W'
        else if (j < 2 * KS + 2 * ES) src = tb + (size_t)(2 * (j - 2 * KS - ES)) * C::SLOT_BYTES;      // T1
        else src = packed + (size_t)(2 * KS + (j - 2 * KS - 2 * ES)) * C::SLOT_BYTES;                  // Wg2
        const int s = (int)(it % C::SLOTS);
        const uint32_t ph = (uint32_t)(it / C::SLOTS) & 1;
        mbar_wait(&empty[s], ph ^ 1, err);
        mbar_arrive_expect_tx(&full[s], C::SLOT_BYTES);
        bulk_g2s(stage0 + (size_t)s * C::SLOT_BYTES, src, C::SLOT_BYTES, &full[s]);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the loops and the waits; one elected lane issues (elect_one: straight UTCHMMA issue instead of a
    // per-lane loop). The issuing thread must stay under the ~310 cycles the tensor pipe needs for three MMAs per slot:
    // descriptors advance by adds (the start-address field counts 16-byte units), the ring position is a running counter.
    {
      const uint32_t idesc = idesc_bf16(128, C::DP);
      constexpr uint32_t lbo_a = 128 * 16, lbo_b = C::DP * 16;
      constexpr uint64_t A_STEP = (2 * lbo_a) >> 4;
      const uint64_t ah0 = smem_desc(smem_u32(A_hi), lbo_a, 128), al0 = smem_desc(smem_u32(A_lo), lbo_a, 128);
      const uint64_t eh0 = smem_desc(smem_u32(E), lbo_a, 128);
      const uint64_t bh0 = smem_desc(smem_u32(stage0), lbo_b, 128);
      uint32_t slot = 0, slot_phase = 0, kphase = 0, ephase = 0, f0phase = 0, f1phase = 0;
      auto wait_kstep = [&](int ks) {
        mbar_wait_poll(&kready[ks], kphase, err);
        tc_fence_after();
      };
      auto wait_bar = [&](uint64_t *bar, uint32_t &phase) {
        mbar_wait_poll(bar, phase, err);
        phase ^= 1;
        tc_fence_after();
      };
      auto take_slot = [&](uint64_t &bh, uint64_t *&release) {
        mbar_wait_poll(&full[slot], slot_phase, err);
        tc_fence_after();
        bh = bh0 + (uint64_t)slot * (C::SLOT_BYTES >> 4);
        release = &empty[slot];
        if (++slot == C::SLOTS) { slot = 0; slot_phase ^= 1; }
      };
      // d (+)= [A operand] * slab^T over all k-steps, three bf16 terms; `follow`: the operand is still being written
      auto gemm_a = [&](uint32_t d, bool follow) {
        for (int ks = 0; ks < KS; ++ks) {
          uint64_t bh, *rel;
          if (follow) wait_kstep(ks);
          take_slot(bh, rel);
          if (elect_one()) {
            const uint64_t ah = ah0 + ks * A_STEP, al = al0 + ks * A_STEP;
            mma_bf16(d, ah, bh, idesc, ks > 0);
            mma_bf16(d, al, bh, idesc, true);
            mma_bf16(d, ah, bh + (C::SLAB >> 4), idesc, true);
            mma_commit(rel);
          }
        }
      };
      // d += E * table^T (E is exact in bf16: two terms)
      auto gemm_e = [&](uint32_t d) {
        for (int ks = 0; ks < ES; ++ks) {
          uint64_t bh, *rel;
          take_slot(bh, rel);
          if (elect_one()) {
            const uint64_t eh = eh0 + ks * A_STEP;
            mma_bf16(d, eh, bh, idesc, true);
            mma_bf16(d, eh, bh + (C::SLAB >> 4), idesc, true);
            mma_commit(rel);
          }
        }
      };
      for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        // ---- GEMM1b: acc1 = H Wd2^T + E T2 (trails the workers writing H) ----
        TR(100);
        wait_bar(free1, f1phase);
        TR(101);
        gemm_a(tmem_base + C::ACC1_COL, true);
        kphase ^= 1;
        wait_bar(e_ready, ephase);
        gemm_e(tmem_base + C::ACC1_COL);
        // ---- GEMM1a: acc0 = H W'^T + E T1 ----
        TR(102);
        wait_bar(free0, f0phase);
        TR(103);
        gemm_a(tmem_base, false);
        gemm_e(tmem_base);
        if (elect_one()) mma_commit(acc_done);
        // ---- GEMM2: acc0 = G Wg2^T (trails the workers writing G) ----
        TR(110);
        wait_bar(free0, f0phase);
        TR(111);
        gemm_a(tmem_base, true);
        if (elect_one()) mma_commit(acc_done);
        TR(113);
        kphase ^= 1;
      }
    }
  } else {
    // ===================== workers: one TMEM lane = one pair row; NPART warps per lane quarter interleave the chunks =====
    const int ww = warp - 2;
    const int quarter = warp & 3;
    const int part = ww >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t done_phase = 0;
    constexpr int NQ = (C::CHUNKS + C::NPART - 1) / C::NPART;           // chunk rounds per thread
    constexpr int NE = (C::E_COLS / 8 + C::NPART - 1) / C::NPART;       // one-hot chunk rounds per thread
    const int g8 = lane & 7;
    // forward -> backward buffer (training): staged H / G tiles and the fp32 a / acc1 blocks of every tile
    constexpr size_t TB = saved_tile_bytes<C>();
    const size_t sv_st = (size_t)(r >> 4) * (size_t)(2 * C::DP * 32) + (size_t)(r & 15) * 16;   // + chunk * 256 (staged layout)
    const size_t sv_f32 = (size_t)r * 32;                                                       // + chunk * 4096 (fp32 blocks)

    // chunk ch of the A operand (8 columns of this warp's 32 rows) is written: one arrival on its k-step's barrier
    auto chunk_done = [&](int ch) {
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&kready[ch >> 1]);
    };
    auto release = [&](uint64_t *bar) {     // this warp's TMEM reads of an accumulator are complete
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
#ifdef NSDP_TRACE
    const bool tr_on = (tid == 64);
#define TWF(id) do { if (tr_on) TR(id); } while (0)
#else
#define TWF(id) do { } while (0)
#endif
    // operands E (one-hot) and H (bf16 hi/lo) of one tile; GEMM1b starts on the first finished k-step
    auto gen_operands = [&](const RowInfoPB &ri, long long tile) {
#pragma unroll
      for (int q = 0; q < NE; ++q) {
        const int ch = part + q * C::NPART;
        if (ch < C::E_COLS / 8) {
          uint4 e = make_uint4(0u, 0u, 0u, 0u);
          if ((ri.j >> 3) == ch) {   // ri.j = -1 on inactive rows: never matches
            const uint32_t one = (ri.j & 1) ? 0x3F800000u : 0x00003F80u;   // bf16 1.0 in the high / low half
            const int w = (ri.j & 7) >> 1;
            e.x = w == 0 ? one : 0u; e.y = w == 1 ? one : 0u; e.z = w == 2 ? one : 0u; e.w = w == 3 ? one : 0u;
          }
          *reinterpret_cast<uint4 *>(E + canon_off(128, r, ch * 8)) = e;
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(e_ready);
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int ch = part + q * C::NPART;
        if (ch < C::CHUNKS) {
          const int k0 = ch * 8;
          float h[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 w0 = wd0s[k0 + j];
            const float pre = fmaf(w0.x, ri.rx, fmaf(w0.y, ri.ry, fmaf(w0.z, ri.rz, w0.w)));
            h[j] = ri.flag * fmaxf(pre, 0.f);
          }
          uint4 hi, lo;
          split2(h[0], h[1], hi.x, lo.x);
          split2(h[2], h[3], hi.y, lo.y);
          split2(h[4], h[5], hi.z, lo.z);
          split2(h[6], h[7], hi.w, lo.w);
          const uint32_t off = canon_off(128, r, k0);
          *reinterpret_cast<uint4 *>(A_hi + off) = hi;
          *reinterpret_cast<uint4 *>(A_lo + off) = lo;
          if (saved) {
            unsigned char *p = saved + (size_t)tile * TB + sv_st + (size_t)ch * 256;
            *reinterpret_cast<uint4 *>(p) = hi;
            *reinterpret_cast<uint4 *>(p + C::DP * 32) = lo;
          }
          chunk_done(ch);
        }
      }
    };

    release(free0);                // nothing to protect before the first tile's GEMMs
    release(free1);
    RowInfoPB ri = row_info_pb<C>(a, blockIdx.x, r, krows, tpb);
    if ((long long)blockIdx.x < tiles) gen_operands(ri, blockIdx.x);
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const bool row_on = ri.c >= 0;
      TWF(200);
      // ---- epilogue 1: G = relu(acc0 + pc) -> A operand ------------------------------------------------------------
      //      acc0 moves to registers first, so GEMM2 (which overwrites it) can start while G is still being produced
      mbar_wait(acc_done, done_phase, err);
      done_phase ^= 1;
      tc_fence_after();
      TWF(202);
      {
        uint32_t gp[NQ][8];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int ch = part + q * C::NPART;
          if (ch < C::CHUNKS) tmem_ld8_nowait(trow + ch * 8, gp[q]);
        }
        tmem_ld_wait();
        release(free0);
        TWF(203);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int ch = part + q * C::NPART;
          if (ch < C::CHUNKS) {
            const int k0 = ch * 8;
            float g[8];
            pin8(gp[q]);     // keeps this round's arithmetic behind the previous round's hand-over
            const float4 p0 = *reinterpret_cast<const float4 *>(pcs + k0), p1 = *reinterpret_cast<const float4 *>(pcs + k0 + 4);
            const float pv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
            // no masking needed: padded columns have zero weights and pc = 0 (G = 0), inactive rows only feed their own
            // (ignored) rows of GEMM2
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] = fmaxf(__uint_as_float(gp[q][j]) + pv[j], 0.f);
            uint4 hi, lo;
            split2(g[0], g[1], hi.x, lo.x);
            split2(g[2], g[3], hi.y, lo.y);
            split2(g[4], g[5], hi.z, lo.z);
            split2(g[6], g[7], hi.w, lo.w);
            const uint32_t off = canon_off(128, r, k0);
            *reinterpret_cast<uint4 *>(A_hi + off) = hi;
            *reinterpret_cast<uint4 *>(A_lo + off) = lo;
            if (saved) {
              unsigned char *p = saved + (size_t)(tiles + tile) * TB + sv_st + (size_t)ch * 256;
              *reinterpret_cast<uint4 *>(p) = hi;
              *reinterpret_cast<uint4 *>(p + C::DP * 32) = lo;
            }
            chunk_done(ch);
          }
        }
      }
      TWF(204);
      // ---- while GEMM2 runs: the next tile's row description (index + coordinates: two dependent L2 reads) --------
      const long long ci_grp = __shfl_sync(0xffffffffu, ri.c, lane & ~7);   // row 0 of the group: the centre (or -1)
      const bool has_next = tile + gridDim.x < tiles;
      const RowInfoPB nxt = row_info_pb<C>(a, tile + gridDim.x, r, krows, tpb);

      // ---- GEMM2 done: s = acc1 -> registers (acc1 is free for the next tile's GEMM1b), then the next tile's operands
      TWF(205);
      mbar_wait(acc_done, done_phase, err);
      done_phase ^= 1;
      tc_fence_after();
      TWF(206);
      uint32_t sreg[NQ][8];
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int ch = part + q * C::NPART;
        if (ch < C::CHUNKS) tmem_ld8_nowait(trow + C::ACC1_COL + ch * 8, sreg[q]);
      }
      tmem_ld_wait();
      release(free1);
      if (has_next) gen_operands(nxt, tile + gridDim.x);
      TWF(208);

      // ---- epilogue 2: softmax over the 8 rows of a centre; out = sum w * (acc1 + vc) -------------------------------------
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int ch = part + q * C::NPART;
        if (ch < C::CHUNKS) {
          const int k0 = ch * 8;
          float av[8], sv[8];
          tmem_ld8(trow + k0, av);
#pragma unroll
          for (int j = 0; j < 8; ++j) sv[j] = __uint_as_float(sreg[q][j]);
          if (saved) {
            float4 *pa = reinterpret_cast<float4 *>(saved + (size_t)(2 * tiles + tile) * TB + sv_f32 + (size_t)ch * 4096);
            float4 *ps = reinterpret_cast<float4 *>(saved + (size_t)(3 * tiles + tile) * TB + sv_f32 + (size_t)ch * 4096);
            pa[0] = make_float4(av[0], av[1], av[2], av[3]); pa[1] = make_float4(av[4], av[5], av[6], av[7]);
            ps[0] = make_float4(sv[0], sv[1], sv[2], sv[3]); ps[1] = make_float4(sv[4], sv[5], sv[6], sv[7]);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) av[j] = row_on ? av[j] : -INFINITY;
          // transpose inside the 8-lane group: afterwards this lane holds column k0 + g8 of all 8 rows of the centre
          group8_transpose(av, lane);
          group8_transpose(sv, lane);
          const int col = k0 + g8;
          const float vcc = vcs[col];
          float mx = av[0];
#pragma unroll
          for (int i = 1; i < 8; ++i) mx = fmaxf(mx, av[i]);
          float se = 0.f, ses = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float e = __expf(av[i] - mx);     // exp(-inf) = 0 on inactive rows (row 0 of a live centre is on)
            se += e;
            ses = fmaf(e, sv[i] + vcc, ses);
          }
          if (ci_grp >= 0 && col < D) {
            const float inv = 1.f / se;
            out[ci_grp * D + col] = ses * inv;
            if (stats) {
              stats[ci_grp * D + col] = mx;
              stats[((long long)a.B * a.M + ci_grp) * D + col] = inv;
            }
          }
        }
      }
      TWF(207);
      release(free0);              // acc0 is consumed: the next tile's GEMM1a may overwrite it
      ri = nxt;
    }
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// =====================================================================================================================
// CTA-PAIR version of the one-hot forward kernel (tcgen05 cta_group::2; conventions: umma.cuh, nsdp_selftest_umma2).
// Two CTAs on the two SMs of a TPC walk over PAIRS of tiles of the same shape: CTA c owns tile 2 j + c (its own A / E
// operands, its own 128 TMEM lanes, its own epilogues), the leader (rank 0) issues every MMA as ONE M = 256 instruction.
// Each CTA streams only ITS HALF of every weight / table slab (rows [104 c, 104 c + 104) of the [208 x 16] slab): half the
// L2 -> SM weight traffic, half the shared-memory fill and B-operand reads per SM -- the single-CTA kernel is bound by
// exactly these (44.5 KB of shared-memory traffic per k-step against 312 MMA cycles).
// Schedule per tile: the one of vattn_fwd_oh_kernel (hand-over per k-step, GEMM1b of the next tile under the softmax
// epilogue). Barriers live at the same offsets in both CTAs; the workers of BOTH CTAs arrive on the LEADER's kready /
// e_ready / free barriers (cluster-space address), the MMA completions are multicast to both CTAs' acc_done / empty[],
// the peer's warp 1 relays "my half of slot s has landed" to the leader's pfull[s].
// =====================================================================================================================
template <class C>
struct PairCfg {
  static constexpr int NH = C::DP / 2;                 // B rows per CTA
  static constexpr int HRUN = NH * 16;                 // one 8-wide k chunk of a half slab (contiguous in the packed image)
  static constexpr int HSLAB = 2 * HRUN;               // [NH x 16] bf16
  static constexpr int HSLOT = 2 * HSLAB;              // hi + lo
  static constexpr int SLOTS = 12;
  static constexpr int OFF_A = 0;
  static constexpr int OFF_E = OFF_A + 2 * C::A_HALF;
  static constexpr int OFF_STAGE = OFF_E + C::E_BYTES;
  static constexpr int OFF_WD0 = OFF_STAGE + SLOTS * HSLOT;
  static constexpr int OFF_PC = OFF_WD0 + C::DP * 16;
  static constexpr int OFF_VC = OFF_PC + C::DP * 4;
  static constexpr int OFF_BAR = OFF_VC + C::DP * 4;
  static constexpr int SMEM = OFF_BAR + 512;
  static_assert(C::DP % 32 == 0 || (C::DP / 2) % 8 == 0, "half slabs are whole core matrices");
  static_assert((3 * SLOTS + C::KSTEPS + 5) * 8 + 4 <= 512, "mbarrier area");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

template <class C>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(C::THREADS, 1)
vattn_fwd_oh2_kernel(const nsdp_vattn_args a, const unsigned char *__restrict__ packed,
                     const unsigned char *__restrict__ tables, float *__restrict__ out, float *__restrict__ stats,
                     unsigned char *__restrict__ saved, int tpb, long long tiles, int *err,
                     unsigned long long *trace) {
  static_assert(C::OH && C::KR == 8, "one-hot kernel: 8 rows per centre");
  using P = PairCfg<C>;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char *A_hi = smem + P::OFF_A;
  unsigned char *A_lo = A_hi + C::A_HALF;
  unsigned char *E = smem + P::OFF_E;
  unsigned char *stage0 = smem + P::OFF_STAGE;
  float4 *wd0s = reinterpret_cast<float4 *>(smem + P::OFF_WD0);
  float *pcs = reinterpret_cast<float *>(smem + P::OFF_PC);
  float *vcs = reinterpret_cast<float *>(smem + P::OFF_VC);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + P::OFF_BAR);
  uint64_t *full = bars, *empty = full + P::SLOTS, *pfull = empty + P::SLOTS, *acc_done = pfull + P::SLOTS;
  uint64_t *kready = acc_done + 1, *e_ready = kready + C::KSTEPS, *free0 = e_ready + 1, *free1 = free0 + 1;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(free1 + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t cta = cluster_ctarank();
  const int D = a.D;
  const int krows = a.K + 1;
  const int tpp = (tpb + 1) / 2;                               // tile pairs per shape
  const long long pairs = (long long)a.B * tpp;
  const long long pair0 = blockIdx.x >> 1, pstride = gridDim.x >> 1;

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
    for (int s = 0; s < P::SLOTS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
      mbar_init(&pfull[s], 1);
    }
    for (int ks = 0; ks < C::KSTEPS; ++ks) mbar_init(&kready[ks], 16);     // 8 warps of each CTA
    mbar_init(e_ready, 2 * C::WORKER_WARPS);
    mbar_init(free0, 2 * C::WORKER_WARPS);
    mbar_init(free1, 2 * C::WORKER_WARPS);
    mbar_init(acc_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr int KS = C::KSTEPS, ES = C::E_KSTEPS;
  constexpr int PER_TILE = 3 * KS + 2 * ES;   // slots: Wd2, T2, W', T1, Wg2

  if (warp == 0) {
    // ===================== producer: this CTA's half of every slot =====================
    constexpr int PL = 6;
    if (lane < PL) {
      const long long mine = pair0 < pairs ? (pairs - pair0 + pstride - 1) / pstride : 0;
      const long long total = mine * PER_TILE;
      for (long long it = lane; it < total; it += PL) {
        const long long pr = pair0 + (it / PER_TILE) * pstride;
        const int j = (int)(it % PER_TILE);
        const unsigned char *tb = tables + (size_t)(pr / tpp) * table_bytes_per_shape<C>();
        const unsigned char *src;
        if (j < KS) src = packed + (size_t)(2 * j + 1) * C::SLOT_BYTES;                                // Wd2
        else if (j < KS + ES) src = tb + (size_t)(2 * (j - KS) + 1) * C::SLOT_BYTES;                   // T2
        else if (j < 2 * KS + ES) src = packed + (size_t)(2 * (j - KS - ES)) * C::SLOT_BYTES;          // W'
        else if (j < 2 * KS + 2 * ES) src = tb + (size_t)(2 * (j - 2 * KS - ES)) * C::SLOT_BYTES;      // T1
        else src = packed + (size_t)(2 * KS + (j - 2 * KS - 2 * ES)) * C::SLOT_BYTES;                  // Wg2
        const int s = (int)(it % P::SLOTS);
        const uint32_t ph = (uint32_t)(it / P::SLOTS) & 1;
        mbar_wait(&empty[s], ph ^ 1, err);
        mbar_arrive_expect_tx(&full[s], P::HSLOT);
        unsigned char *dst = stage0 + (size_t)s * P::HSLOT;
        bulk_g2s(dst, src + (size_t)cta * P::HSLOT, P::HSLOT, &full[s]);     // the slot image is cut per CTA (slot_off)
      }
    }
  } else if (warp == 1) {
    if (cta != 0) {
      // ===================== peer: relay "my half of slot s has landed" to the leader =====================
      const long long mine = pair0 < pairs ? (pairs - pair0 + pstride - 1) / pstride : 0;
      const long long total = mine * PER_TILE;
      uint32_t slot = 0, slot_phase = 0;
      for (long long it = 0; it < total; ++it) {
        mbar_spin(&full[slot], slot_phase, err);
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&pfull[slot]), 0));
        __syncwarp();
        if (++slot == (uint32_t)P::SLOTS) { slot = 0; slot_phase ^= 1; }
      }
    } else {
      // ===================== leader: MMA issuer for the pair =====================
      const uint32_t idesc = idesc_bf16(256, C::DP);
      constexpr uint32_t lbo_a = 128 * 16, lbo_b = P::HRUN;
      constexpr uint64_t A_STEP = (2 * lbo_a) >> 4;
      const uint64_t ah0 = smem_desc(smem_u32(A_hi), lbo_a, 128), al0 = smem_desc(smem_u32(A_lo), lbo_a, 128);
      const uint64_t eh0 = smem_desc(smem_u32(E), lbo_a, 128);
      const uint64_t bh0 = smem_desc(smem_u32(stage0), lbo_b, 128);
      uint32_t slot = 0, slot_phase = 0, kphase = 0, ephase = 0, f0phase = 0, f1phase = 0;
      auto wait_kstep = [&](int ks) {
        mbar_wait(&kready[ks], kphase, err);
        tc_fence_after();
      };
      auto wait_bar = [&](uint64_t *bar, uint32_t &phase) {
        mbar_wait(bar, phase, err);
        phase ^= 1;
        tc_fence_after();
      };
      auto take_slot = [&](uint64_t &bh, uint64_t *&release) {
#ifndef NSDP_PAIR_NOWAIT     // (experiment: garbage results, pure MMA issue / execution rate)
        mbar_spin(&full[slot], slot_phase, err);
        mbar_spin(&pfull[slot], slot_phase, err);
#endif
        tc_fence_after();
        bh = bh0 + (uint64_t)slot * (P::HSLOT >> 4);
        release = &empty[slot];
        if (++slot == (uint32_t)P::SLOTS) { slot = 0; slot_phase ^= 1; }
      };
      auto gemm_a = [&](uint32_t d, bool follow) {
        for (int ks = 0; ks < KS; ++ks) {
          uint64_t bh, *rel;
          if (follow) wait_kstep(ks);
          take_slot(bh, rel);
          if (elect_one()) {
            const uint64_t ah = ah0 + ks * A_STEP, al = al0 + ks * A_STEP;
            mma_bf16_pair(d, ah, bh, idesc, ks > 0);
            mma_bf16_pair(d, al, bh, idesc, true);
            mma_bf16_pair(d, ah, bh + (P::HSLAB >> 4), idesc, true);
            mma_commit_pair(rel);
          }
        }
      };
      auto gemm_e = [&](uint32_t d) {
        for (int ks = 0; ks < ES; ++ks) {
          uint64_t bh, *rel;
          take_slot(bh, rel);
          if (elect_one()) {
            const uint64_t eh = eh0 + ks * A_STEP;
            mma_bf16_pair(d, eh, bh, idesc, true);
            mma_bf16_pair(d, eh, bh + (P::HSLAB >> 4), idesc, true);
            mma_commit_pair(rel);
          }
        }
      };
      for (long long pr = pair0; pr < pairs; pr += pstride) {
        TR(100);
        wait_bar(free1, f1phase);
        TR(101);
        gemm_a(tmem_base + C::ACC1_COL, true);          // GEMM1b: acc1 = H Wd2^T + E T2
        kphase ^= 1;
        wait_bar(e_ready, ephase);
        gemm_e(tmem_base + C::ACC1_COL);
        TR(102);
        wait_bar(free0, f0phase);
        TR(103);
        gemm_a(tmem_base, false);                       // GEMM1a: acc0 = H W'^T + E T1
        gemm_e(tmem_base);
        if (elect_one()) mma_commit_pair(acc_done);
        TR(110);
        wait_bar(free0, f0phase);
        TR(111);
        gemm_a(tmem_base, true);                        // GEMM2: acc0 = G Wg2^T
        if (elect_one()) mma_commit_pair(acc_done);
        TR(113);
        kphase ^= 1;
      }
    }
  } else {
    // ===================== workers (both CTAs): one TMEM lane = one pair row of this CTA's tile =====================
    const int ww = warp - 2;
    const int quarter = warp & 3;
    const int part = ww >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t done_phase = 0;
    constexpr int NQ = (C::CHUNKS + C::NPART - 1) / C::NPART;
    constexpr int NE = (C::E_COLS / 8 + C::NPART - 1) / C::NPART;
    const int g8 = lane & 7;
    constexpr size_t TB = saved_tile_bytes<C>();
    const size_t sv_st = (size_t)(r >> 4) * (size_t)(2 * C::DP * 32) + (size_t)(r & 15) * 16;
    const size_t sv_f32 = (size_t)r * 32;
    // the leader's barriers, cluster-space addresses (valid from both CTAs)
    const uint32_t kready0 = mapa_u32(smem_u32(kready), 0), e_ready_l = mapa_u32(smem_u32(e_ready), 0);
    const uint32_t free0_l = mapa_u32(smem_u32(free0), 0), free1_l = mapa_u32(smem_u32(free1), 0);

    auto chunk_done = [&](int ch) {
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(kready0 + (uint32_t)(ch >> 1) * 8u);
    };
    auto release = [&](uint32_t bar_l) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(bar_l);
    };
    // tile of this CTA inside pair pr: shape b = pr / tpp, tile 2 (pr % tpp) + cta of the shape (may not exist: all rows off)
    auto tile_of = [&](long long pr, long long &tile) -> bool {
      const long long b = pr / tpp;
      const int tin = 2 * (int)(pr - b * tpp) + (int)cta;
      tile = b * tpb + tin;
      return pr < pairs && tin < tpb;
    };
    auto row_info = [&](long long pr) {
      long long tile;
      RowInfoPB ri;
      if (tile_of(pr, tile)) {
        ri = row_info_pb<C>(a, tile, r, krows, tpb);
      } else {
        ri.c = -1; ri.j = -1; ri.rx = ri.ry = ri.rz = 0.f; ri.flag = 0.f;
      }
      return ri;
    };
    auto gen_operands = [&](const RowInfoPB &ri, long long pr) {
      long long tile;
      const bool sv = tile_of(pr, tile) && saved;
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
          *reinterpret_cast<uint4 *>(E + canon_off(128, r, ch * 8)) = e;
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(e_ready_l);
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int ch = part + q * C::NPART;
        if (ch < C::CHUNKS) {
          const int k0 = ch * 8;
          float h[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 w0 = wd0s[k0 + j];
            const float pre = fmaf(w0.x, ri.rx, fmaf(w0.y, ri.ry, fmaf(w0.z, ri.rz, w0.w)));
            h[j] = ri.flag * fmaxf(pre, 0.f);
          }
          uint4 hi, lo;
          split2(h[0], h[1], hi.x, lo.x);
          split2(h[2], h[3], hi.y, lo.y);
          split2(h[4], h[5], hi.z, lo.z);
          split2(h[6], h[7], hi.w, lo.w);
          const uint32_t off = canon_off(128, r, k0);
          *reinterpret_cast<uint4 *>(A_hi + off) = hi;
          *reinterpret_cast<uint4 *>(A_lo + off) = lo;
          if (sv) {
            unsigned char *p = saved + (size_t)tile * TB + sv_st + (size_t)ch * 256;
            *reinterpret_cast<uint4 *>(p) = hi;
            *reinterpret_cast<uint4 *>(p + C::DP * 32) = lo;
          }
          chunk_done(ch);
        }
      }
    };

    release(free0_l);
    release(free1_l);
    RowInfoPB ri = row_info(pair0);
    if (pair0 < pairs) gen_operands(ri, pair0);
    for (long long pr = pair0; pr < pairs; pr += pstride) {
      long long tile;
      const bool sv = tile_of(pr, tile) && saved;
      const bool row_on = ri.c >= 0;
      if (tid == 64) TR(200);
      // ---- epilogue 1: G = relu(acc0 + pc) -> A operand (acc0 to registers first: GEMM2 overwrites it) ----------------
      mbar_wait(acc_done, done_phase, err);
      done_phase ^= 1;
      tc_fence_after();
      if (tid == 64) TR(202);
      {
        uint32_t gp[NQ][8];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int ch = part + q * C::NPART;
          if (ch < C::CHUNKS) tmem_ld8_nowait(trow + ch * 8, gp[q]);
        }
        tmem_ld_wait();
        release(free0_l);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int ch = part + q * C::NPART;
          if (ch < C::CHUNKS) {
            const int k0 = ch * 8;
            float g[8];
            pin8(gp[q]);
            const float4 p0 = *reinterpret_cast<const float4 *>(pcs + k0), p1 = *reinterpret_cast<const float4 *>(pcs + k0 + 4);
            const float pv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] = fmaxf(__uint_as_float(gp[q][j]) + pv[j], 0.f);
            uint4 hi, lo;
            split2(g[0], g[1], hi.x, lo.x);
            split2(g[2], g[3], hi.y, lo.y);
            split2(g[4], g[5], hi.z, lo.z);
            split2(g[6], g[7], hi.w, lo.w);
            const uint32_t off = canon_off(128, r, k0);
            *reinterpret_cast<uint4 *>(A_hi + off) = hi;
            *reinterpret_cast<uint4 *>(A_lo + off) = lo;
            if (sv) {
              unsigned char *p = saved + (size_t)(tiles + tile) * TB + sv_st + (size_t)ch * 256;
              *reinterpret_cast<uint4 *>(p) = hi;
              *reinterpret_cast<uint4 *>(p + C::DP * 32) = lo;
            }
            chunk_done(ch);
          }
        }
      }
      if (tid == 64) TR(204);
      // ---- while GEMM2 runs: the next pair's row description ----------------------------------------------------------
      const long long ci_grp = __shfl_sync(0xffffffffu, ri.c, lane & ~7);
      const bool has_next = pr + pstride < pairs;
      const RowInfoPB nxt = row_info(pr + pstride);

      // ---- GEMM2 done: s = acc1 -> registers, then the next tile's operands (GEMM1b of the next pair trails them) ------
      mbar_wait(acc_done, done_phase, err);
      done_phase ^= 1;
      tc_fence_after();
      if (tid == 64) TR(206);
      uint32_t sreg[NQ][8];
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int ch = part + q * C::NPART;
        if (ch < C::CHUNKS) tmem_ld8_nowait(trow + C::ACC1_COL + ch * 8, sreg[q]);
      }
      tmem_ld_wait();
      release(free1_l);
      if (has_next) gen_operands(nxt, pr + pstride);
      if (tid == 64) TR(208);

      // ---- epilogue 2: softmax over the 8 rows of a centre; out = sum w * (acc1 + vc) ---------------------------------
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int ch = part + q * C::NPART;
        if (ch < C::CHUNKS) {
          const int k0 = ch * 8;
          float av[8], sv8[8];
          tmem_ld8(trow + k0, av);
#pragma unroll
          for (int j = 0; j < 8; ++j) sv8[j] = __uint_as_float(sreg[q][j]);
          if (sv) {
            float4 *pa = reinterpret_cast<float4 *>(saved + (size_t)(2 * tiles + tile) * TB + sv_f32 + (size_t)ch * 4096);
            float4 *ps = reinterpret_cast<float4 *>(saved + (size_t)(3 * tiles + tile) * TB + sv_f32 + (size_t)ch * 4096);
            pa[0] = make_float4(av[0], av[1], av[2], av[3]); pa[1] = make_float4(av[4], av[5], av[6], av[7]);
            ps[0] = make_float4(sv8[0], sv8[1], sv8[2], sv8[3]); ps[1] = make_float4(sv8[4], sv8[5], sv8[6], sv8[7]);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) av[j] = row_on ? av[j] : -INFINITY;
          group8_transpose(av, lane);
          group8_transpose(sv8, lane);
          const int col = k0 + g8;
          const float vcc = vcs[col];
          float mx = av[0];
#pragma unroll
          for (int i = 1; i < 8; ++i) mx = fmaxf(mx, av[i]);
          float se = 0.f, ses = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float e = __expf(av[i] - mx);
            se += e;
            ses = fmaf(e, sv8[i] + vcc, ses);
          }
          if (ci_grp >= 0 && col < D) {
            const float inv = 1.f / se;
            out[ci_grp * D + col] = ses * inv;
            if (stats) {
              stats[ci_grp * D + col] = mx;
              stats[((long long)a.B * a.M + ci_grp) * D + col] = inv;
            }
          }
        }
      }
      if (tid == 64) TR(207);
      release(free0_l);
      ri = nxt;
    }
  }
  tc_fence_before();
  cluster_sync_all();            // nobody leaves while the peer may still touch this CTA's barriers / shared memory
  if (warp == 1) tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
}

// NSDP_FWD_PAIR=1 selects the CTA-pair kernel. It is bit-compatible with the single-CTA kernel (tests/test_gpu_vattn.py runs the
// decoder cases through both) but NOT the default: measured on B200 (8 x 50 000 queries) the pair kernel needs 33.0 k cycles
// per tile pair against 29.5 k per tile of the single-CTA kernel on each of the two SMs, i.e. 3.15 ms against 2.75 ms. With the
// k-step hand-over the tiles are bound by the workers' epilogues (softmax, operand generation: ~18 k of the 29.5 k cycles),
// not by the weight stream the pair halves; and the pair's slot ring has a longer round trip (commit -> both producers ->
// L2 -> relay to the leader: ~5.5 k cycles for 12 half slots, 460 cycles per k-step against the 357 the MMAs need), which
// the shared memory left over (7 KB) cannot cover with more slots. Without slot waits (garbage results, timing only) the
// pair tile takes 28.0 k cycles: the ceiling of the variant is ~5 %.
static bool fwd_pair_on() {
  static const bool v = [] { const char *e = getenv("NSDP_FWD_PAIR"); return e && atoi(e) != 0; }();
  return v;
}

template <class C>
static int launch_oh(const nsdp_vattn_args &a, float *out, float *stats, void *workspace, size_t ws_bytes, cudaStream_t st) {
  const size_t tb = (size_t)a.B * table_bytes_per_shape<C>();
  const size_t need = packed_bytes<C>() + tb + 16;
  if (!workspace || ws_bytes < need) return NSDP_ERR_WORKSPACE;
  unsigned char *packed = (unsigned char *)workspace;
  unsigned char *tables = packed + packed_bytes<C>();
  int *err = (int *)(tables + tb);
  cudaError_t e = cudaMemsetAsync(err, 0, sizeof(int), st);
  if (e != cudaSuccess) return cuda_rc(e);
  const int pair = (fwd_pair_on() && C::NPART == 4) ? 1 : 0;
  pack_weights_kernel<C><<<64, 256, 0, st>>>(a.wpt, a.wd2t, a.wg2t, a.D, packed, pair);
  int rc = check_launch();
  if (rc != NSDP_OK) return rc;
  pack_tables_kernel<C><<<128, 256, 0, st>>>(a, tables, pair);
  rc = check_launch();
  if (rc != NSDP_OK) return rc;
  const int tpb = (a.M + C::CENTRES - 1) / C::CENTRES;
  const long long tiles = (long long)a.B * tpb;
  auto kern = vattn_fwd_oh_kernel<C>;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
  if (e != cudaSuccess) return cuda_rc(e);
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  unsigned char *saved = (stats && a.saved && a.saved_bytes >= saved_bytes_total<C>(tiles)) ? (unsigned char *)a.saved : nullptr;
  unsigned long long *trace = nullptr;
#ifdef NSDP_TRACE
  if (const char *tp = getenv("NSDP_TRACE_FWD_PTR")) trace = (unsigned long long *)strtoull(tp, nullptr, 0);
#endif
  if (pair) {
    auto kern2 = vattn_fwd_oh2_kernel<C>;
    static bool ready = false;
    if (!ready) {
      e = cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg<C>::SMEM);
      if (e != cudaSuccess) return cuda_rc(e);
      ready = true;
    }
    const long long pairs = (long long)a.B * ((tpb + 1) / 2);
    const long long slots = num_sms() / 2;
    const int grid2 = 2 * (int)(pairs < slots ? pairs : slots);
    kern2<<<grid2, C::THREADS, PairCfg<C>::SMEM, st>>>(a, packed, tables, out, stats, saved, tpb, tiles, err, trace);
    return check_launch();
  }
  kern<<<grid, C::THREADS, C::SMEM, st>>>(a, packed, tables, out, stats, saved, tpb, tiles, err, trace);
  return check_launch();
}

// the one-hot kernel serves the decoder cross-attention: global token, per-shape query (no qp), anchor tables that
// fit the one-hot width
static bool oh_ok(const nsdp_vattn_args &a) {
  return a.has_global && !a.qp && a.kp && a.vp && a.gq && a.gv && a.idx && a.N + 1 <= 112 && a.K + 1 <= 8;
}

template <class C>
static int launch(const nsdp_vattn_args &a, float *out, float *stats, void *workspace, size_t ws_bytes, cudaStream_t st) {
  const size_t need = packed_bytes<C>() + 16;
  if (!workspace || ws_bytes < need) return NSDP_ERR_WORKSPACE;
  unsigned char *packed = (unsigned char *)workspace;
  int *err = (int *)(packed + packed_bytes<C>());
  cudaError_t e = cudaMemsetAsync(err, 0, sizeof(int), st);
  if (e != cudaSuccess) return cuda_rc(e);
  pack_weights_kernel<C><<<64, 256, 0, st>>>(a.wpt, a.wd2t, a.wg2t, a.D, packed);
  int rc = check_launch();
  if (rc != NSDP_OK) return rc;
  const long long tiles = ceil_div((long long)a.B * a.M, (long long)C::CENTRES);
  auto kern = vattn_fwd_tc_kernel<C>;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
  if (e != cudaSuccess) return cuda_rc(e);
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  kern<<<grid, C::THREADS, C::SMEM, st>>>(a, packed, out, stats, tiles, err);
  return check_launch();
}

// decoder shape: worker warps per lane quarter (NSDP_FWD_NPART_DEC = 4 | 7)
static int dec_npart_fwd() {
  static const int v = [] {
    const char *e = getenv("NSDP_FWD_NPART_DEC");
    return e && atoi(e) == 7 ? 7 : 4;   // measured at HEAD of round 1: 3.06 ms (4) vs 3.12 ms (7)
  }();
  return v;
}

static bool no_onehot() {
  static const bool v = [] { const char *e = getenv("NSDP_NO_ONEHOT"); return e && atoi(e) != 0; }();
  return v;
}

// Which (DP, KR) instantiation serves these arguments; 0 = not supported by the tensor-core path.
//   1: <208, 8>   decoder: D in (128, 208], 7 neighbours + global token
//   2: <128, 16>  encoder d_reduced: D <= 128, up to 16 rows per centre (k = 10 is padded to 16)
//   3: <256, 16>  encoder d_transformer: D <= 256, up to 16 rows per centre
//   4: <256, 128> full attention over up to 128 source points (group_all, one centre per tile)
static int pick(const nsdp_vattn_args &a) {
  const int krows = a.K + (a.has_global ? 1 : 0);
  if (a.D % 4 != 0 || a.D > 256) return 0;
  if (krows <= 8 && a.D > 128 && a.D <= 208) return 1;
  if (krows <= 16 && !a.has_global) return a.D <= 128 ? 2 : 3;
  if (krows <= 128 && !a.has_global && a.D > 128) return 4;
  return 0;
}

}  // namespace vtc
}  // namespace nsdp

extern "C" size_t nsdp_vattn_fwd_workspace_bytes(const nsdp_vattn_args *args) {
  using namespace nsdp;
  if (!args) return 0;
  switch (vtc::pick(*args)) {
    case 1:
      return vtc::packed_bytes<vtc::TcCfg<208, 8>>() + 16 +
             (vtc::oh_ok(*args) ? (size_t)args->B * vtc::table_bytes_per_shape<vtc::TcCfg<208, 8, 4, true>>() : 0);
    case 2: return vtc::packed_bytes<vtc::TcCfg<128, 16>>() + 16;
    case 3: return vtc::packed_bytes<vtc::TcCfg<256, 16>>() + 16;
    case 4: return vtc::packed_bytes<vtc::TcCfg<256, 128>>() + 16;
    default: return 0;
  }
}

extern "C" size_t nsdp_vattn_saved_bytes(const nsdp_vattn_args *args) {
  using namespace nsdp;
  if (!args || args->impl == 1 || vtc::pick(*args) != 1 || !vtc::oh_ok(*args) || vtc::no_onehot()) return 0;
  using Cfg = vtc::TcCfg<208, 8, 4, true>;
  const long long tiles = (long long)args->B * ((args->M + Cfg::CENTRES - 1) / Cfg::CENTRES);
  return vtc::saved_bytes_total<Cfg>(tiles);
}

namespace nsdp {
int vattn_fwd_tc_dispatch(const nsdp_vattn_args *args, float *out, float *stats, void *workspace, size_t ws_bytes,
                          cudaStream_t st, bool *handled) {
  *handled = true;
  switch (vtc::pick(*args)) {
    case 1:
      if (vtc::oh_ok(*args) && !vtc::no_onehot()) {
        if (vtc::dec_npart_fwd() == 7)
          return vtc::launch_oh<vtc::TcCfg<208, 8, 7, true>>(*args, out, stats, workspace, ws_bytes, st);
        return vtc::launch_oh<vtc::TcCfg<208, 8, 4, true>>(*args, out, stats, workspace, ws_bytes, st);
      }
      return vtc::launch<vtc::TcCfg<208, 8>>(*args, out, stats, workspace, ws_bytes, st);
    case 2: return vtc::launch<vtc::TcCfg<128, 16>>(*args, out, stats, workspace, ws_bytes, st);
    case 3: return vtc::launch<vtc::TcCfg<256, 16>>(*args, out, stats, workspace, ws_bytes, st);
    case 4: return vtc::launch<vtc::TcCfg<256, 128>>(*args, out, stats, workspace, ws_bytes, st);
    default: *handled = false; return NSDP_OK;
  }
}
}  // namespace nsdp

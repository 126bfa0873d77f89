// Backward of the fused neural-field MLP on tcgen05 tensor cores (sm_100a): chain kernel + staged weight gradients.
//
// Per tile of 128 query rows a persistent CTA recomputes the forward exactly like fused_mlp_tc.cu (same chunk-pipelined
// layer handoff, two ping-pong accumulators in TMEM), keeps the ReLU masks of h_0 .. h_L as bit masks, and then walks the
// layers backwards through the SAME pipeline — every step is again a [128 x W] x [W x W] product:
//     dz_{L-1} = (d_out W_out) * [h_L > 0]                      (fp32 FMAs, N = O <= 4)
//     for l = L-1 .. 0:   dh_l = dz_l W_l ;  dz_{l-1} = dh_l * [h_l > 0]          (dz_{-1} = dz_in)
//     d_x = dz_in W_in                                           (fp32 FMAs, K = Cin <= 4)
// so a tile is 2L GEMMs with 2L chunked operand handoffs. Weight / bias gradients: the operand tiles x, h_0..h_L, dz_in,
// dz_0..dz_{L-1} and d_out are staged (bf16 hi/lo, layout of dw_tc.cu) segment by segment and reduced by dw_tc_kernel
// (d_W^T = X^T Y, d_b = column sums of Y).
#include <stdlib.h>

#include "dw_tc.cuh"
#include "fused_mlp_tc.cuh"

namespace nsdp {
namespace mbtc {

using namespace umma;
using mtc::Cfg;
using mtc::MAX_HIDDEN;
using mtc::store_split8;
using mtc::tmem_ldn;

// staged tensors of one segment (tile stride = 512 * width bytes)
struct Staging {
  unsigned char *x;                     // width 16 (columns >= Cin are zero)
  unsigned char *h[MAX_HIDDEN + 1];     // h_0 .. h_L, width W
  unsigned char *dz[MAX_HIDDEN + 1];    // dz[0] = dz_in, dz[1 + l] = dz_l, width W
  unsigned char *dout;                  // width 16 (columns >= O are zero)
};

template <int W>
__device__ __forceinline__ void stage8(unsigned char *tile, int r, int k0, const float *x) {
  uint4 hi, lo;
  split2(x[0], x[1], hi.x, lo.x);
  split2(x[2], x[3], hi.y, lo.y);
  split2(x[4], x[5], hi.z, lo.z);
  split2(x[6], x[7], hi.w, lo.w);
  unsigned char *p = tile + (size_t)(r >> 4) * (2 * W * 32) + (size_t)(k0 >> 3) * 256 + (r & 15) * 16;
  *reinterpret_cast<uint4 *>(p) = hi;
  *reinterpret_cast<uint4 *>(p + W * 32) = lo;
}

// A operand chunk + staged tile chunk from the same 8 values
template <int W>
__device__ __forceinline__ void store_both8(unsigned char *X_hi, unsigned char *X_lo, unsigned char *tile, int r, int k0,
                                            const float *x, bool to_operand) {
  uint4 hi, lo;
  split2(x[0], x[1], hi.x, lo.x);
  split2(x[2], x[3], hi.y, lo.y);
  split2(x[4], x[5], hi.z, lo.z);
  split2(x[6], x[7], hi.w, lo.w);
  if (to_operand) {
    const uint32_t off = canon_off(128, r, k0);
    *reinterpret_cast<uint4 *>(X_hi + off) = hi;
    *reinterpret_cast<uint4 *>(X_lo + off) = lo;
  }
  unsigned char *p = tile + (size_t)(r >> 4) * (2 * W * 32) + (size_t)(k0 >> 3) * 256 + (r & 15) * 16;
  *reinterpret_cast<uint4 *>(p) = hi;
  *reinterpret_cast<uint4 *>(p + W * 32) = lo;
}

template <class C>
constexpr size_t packed_bytes(int n_hidden) {
  return (size_t)2 * n_hidden * C::KS * C::STAGE_BYTES;
}

// grid.x = 2L matrices in consumption order: forward layers 0 .. L-1 (B[n][k] = W_l[n][k] = w_h_t[l][k][n]), then the data-
// gradient products of layers L-1 .. 0 (B[n][k] = W_l[k][n] = w_h_t[l][n][k])
template <class C>
__global__ void pack_mlp_bwd_weights_kernel(const nsdp_mlp_args a, unsigned char *__restrict__ out) {
  constexpr int W = C::W;
  const int L = a.n_hidden;
  const int m = blockIdx.x;
  const bool bwd = m >= L;
  const int l = bwd ? 2 * L - 1 - m : m;
  const float *wt = a.w_h_t + (size_t)l * W * W;
  unsigned char *o = out + (size_t)m * C::KS * C::STAGE_BYTES;
  for (int e = threadIdx.x; e < W * (W / 2); e += blockDim.x) {
    int n, k;
    float x0, x1;
    if (!bwd) {
      n = e % W; k = (e / W) * 2;
      x0 = __ldg(wt + (size_t)k * W + n); x1 = __ldg(wt + (size_t)(k + 1) * W + n);
    } else {
      k = (e % (W / 2)) * 2; n = e / (W / 2);
      x0 = __ldg(wt + (size_t)n * W + k); x1 = __ldg(wt + (size_t)n * W + k + 1);
    }
    uint32_t hi, lo;
    split2(x0, x1, hi, lo);
    const size_t base = (size_t)(k >> 4) * C::STAGE_BYTES + canon_off(W, n, k & 15);
    *reinterpret_cast<uint32_t *>(o + base) = hi;
    *reinterpret_cast<uint32_t *>(o + base + C::SLAB) = lo;
  }
}

template <class C>
__global__ void __launch_bounds__(C::THREADS, C::MIN_CTAS)
fused_mlp_bwd_tc_kernel(const nsdp_mlp_args a, const float *__restrict__ dout, float *__restrict__ d_x,
                        const unsigned char *__restrict__ packed, const Staging stg, long long t0, long long t1, int *err) {
  constexpr int W = C::W, STAGES = C::STAGES, NCH = C::NCH, CPT = C::CPT, CW = C::CW;
  constexpr int MW = (W / C::NWQ + 31) / 32;   // mask words per thread and layer
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char *X_hi = smem + C::OFF_X, *X_lo = X_hi + C::A_HALF;
  unsigned char *stage0 = smem + C::OFF_STAGE;
  float *bias = reinterpret_cast<float *>(smem + C::OFF_BIAS);
  float4 *wos = reinterpret_cast<float4 *>(smem + C::OFF_WO);
  float4 *part = reinterpret_cast<float4 *>(smem + C::OFF_PART);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::OFF_BAR);
  uint64_t *full = bars, *empty = bars + STAGES, *a_ready = bars + 2 * STAGES, *acc_done = a_ready + NCH;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_done + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = a.n_hidden, Cin = a.Cin, O = a.O;
  const long long tiles = t1 - t0;   // tiles of this segment; tile index below is segment-local

  for (int i = tid; i < (1 + L) * W; i += C::THREADS) bias[i] = i < W ? a.b_in[i] : a.b_h[i - W];
  for (int c = tid; c < W; c += C::THREADS) {
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    w.x = a.w_out_t[(size_t)c * O + 0];
    if (O > 1) w.y = a.w_out_t[(size_t)c * O + 1];
    if (O > 2) w.z = a.w_out_t[(size_t)c * O + 2];
    if (O > 3) w.w = a.w_out_t[(size_t)c * O + 3];
    wos[c] = w;
  }
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int c = 0; c < NCH; ++c) mbar_init(&a_ready[c], C::WORKERS);
    mbar_init(acc_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int per_tile = 2 * L * C::KS;

  if (warp == 0) {
    constexpr int PL = 2;
    if (lane < PL) {
      const long long my_tiles = (long long)blockIdx.x < tiles ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
      const long long total = my_tiles * per_tile;
      for (long long it = lane; it < total; it += PL) {
        const int st = (int)(it % per_tile);
        const int s = (int)(it % STAGES);
        const uint32_t ph = (uint32_t)(it / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1, err);
        mbar_arrive_expect_tx(&full[s], C::STAGE_BYTES);
        bulk_g2s(stage0 + (size_t)s * C::STAGE_BYTES, packed + (size_t)st * C::STAGE_BYTES, C::STAGE_BYTES, &full[s]);
      }
    }
  } else if (warp == 1) {
    // MMA issuer: 2L identical GEMM shapes per tile (L forward, L data-gradient), operand chunks as they land
    const uint32_t idesc = idesc_bf16(128, W);
    constexpr uint32_t lbo_a = 128 * 16, lbo_b = W * 16;
    constexpr uint64_t A_STEP = (2 * lbo_a) >> 4;
    const uint64_t xhi = smem_desc(smem_u32(X_hi), lbo_a, 128), xlo = smem_desc(smem_u32(X_lo), lbo_a, 128);
    const uint64_t bh0 = smem_desc(smem_u32(stage0), lbo_b, 128);
    uint32_t slot = 0, slot_phase = 0, ready_phase = 0, g = 0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      for (int l = 0; l < 2 * L; ++l) {
        const uint32_t col = tmem_base + (g & 1u) * W;
        for (int c = 0; c < NCH; ++c) {
          mbar_wait(&a_ready[c], ready_phase, err);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < C::KPC; ++kk) {
            const int ks = c * C::KPC + kk;
            mbar_wait(&full[slot], slot_phase, err);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t ah = xhi + ks * A_STEP, al = xlo + ks * A_STEP;
              const uint64_t bh = bh0 + (uint64_t)slot * (C::STAGE_BYTES >> 4);
              mma_bf16(col, ah, bh, idesc, ks != 0);
              mma_bf16(col, al, bh, idesc, true);
              mma_bf16(col, ah, bh + (C::SLAB >> 4), idesc, true);
              mma_commit(&empty[slot]);
            }
            __syncwarp();
            if (++slot == STAGES) { slot = 0; slot_phase ^= 1; }
          }
        }
        if (elect_one()) mma_commit(acc_done);
        __syncwarp();
        ready_phase ^= 1;
        ++g;
      }
    }
  } else {
    const int quarter = warp & 3;
    const int p = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t done_phase = 0, g = 0;
    uint32_t masks[(MAX_HIDDEN + 1) * MW];   // ReLU masks of h_0 .. h_L: bit (c * CPT + u) = this thread's column u of chunk c

    auto chunk_done = [&](int c) {
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_ready[c]);
    };
    auto wait_acc = [&]() -> uint32_t {
      mbar_wait(acc_done, done_phase, err);
      done_phase ^= 1;
      tc_fence_after();
      const uint32_t acc = trow + (g & 1u) * W;
      ++g;
      return acc;
    };

    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const long long grow = (t0 + tile) * 128 + r;
      const bool on = grow < a.R;
      const size_t toff = (size_t)tile * 512;   // x width = byte offset of this tile inside a staged tensor
      float xin[4] = {0.f, 0.f, 0.f, 0.f}, dov[4] = {0.f, 0.f, 0.f, 0.f};
      if (on) {
#pragma unroll
        for (int ci = 0; ci < 4; ++ci)
          if (ci < Cin) xin[ci] = __ldg(a.x + (size_t)grow * Cin + ci);
#pragma unroll
        for (int o = 0; o < 4; ++o)
          if (o < O) dov[o] = __ldg(dout + (size_t)grow * O + o);
      }
      if (p == 0) {   // the two narrow operands of the first / last layer's weight gradients (zero rows beyond R)
        const float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const float x8[8] = {xin[0], xin[1], xin[2], xin[3], 0.f, 0.f, 0.f, 0.f};
        const float d8[8] = {dov[0], dov[1], dov[2], dov[3], 0.f, 0.f, 0.f, 0.f};
        stage8<16>(stg.x + toff * 16, r, 0, x8);
        stage8<16>(stg.x + toff * 16, r, 8, z8);
        stage8<16>(stg.dout + toff * 16, r, 0, d8);
        stage8<16>(stg.dout + toff * 16, r, 8, z8);
      }
      // ---- layer 0: h_0 = relu(x W_in + b_in) -> first A operand, staged, mask ----------------------------------------
      {
        uint32_t mk[MW];
#pragma unroll
        for (int w = 0; w < MW; ++w) mk[w] = 0u;
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
          const int col0 = c * CW + p * CPT;
          uint32_t bits = 0u;
#pragma unroll
          for (int j = 0; j < CPT; j += 8) {
            float v[8];
            {
              const float4 b0 = *reinterpret_cast<const float4 *>(bias + col0 + j);
              const float4 b1 = *reinterpret_cast<const float4 *>(bias + col0 + j + 4);
              v[0] = b0.x; v[1] = b0.y; v[2] = b0.z; v[3] = b0.w; v[4] = b1.x; v[5] = b1.y; v[6] = b1.z; v[7] = b1.w;
            }
#pragma unroll
            for (int ci = 0; ci < 4; ++ci) {
              if (ci < Cin) {
                const float4 w0 = ldg4(a.w_in_t + (size_t)ci * W + col0 + j);
                const float4 w1 = ldg4(a.w_in_t + (size_t)ci * W + col0 + j + 4);
                const float xv = xin[ci];
                v[0] = fmaf(xv, w0.x, v[0]); v[1] = fmaf(xv, w0.y, v[1]); v[2] = fmaf(xv, w0.z, v[2]); v[3] = fmaf(xv, w0.w, v[3]);
                v[4] = fmaf(xv, w1.x, v[4]); v[5] = fmaf(xv, w1.y, v[5]); v[6] = fmaf(xv, w1.z, v[6]); v[7] = fmaf(xv, w1.w, v[7]);
              }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              bits |= (v[u] > 0.f ? 1u : 0u) << (j + u);
              v[u] = fmaxf(v[u], 0.f);
            }
            store_both8<W>(X_hi, X_lo, stg.h[0] + toff * W, r, col0 + j, v, true);
          }
          chunk_done(c);
          const int bit0 = c * CPT;
#pragma unroll
          for (int w = 0; w < MW; ++w)
            if ((bit0 >> 5) == w) mk[w] |= bits << (bit0 & 31);
        }
#pragma unroll
        for (int w = 0; w < MW; ++w) masks[w] = mk[w];
      }
      // ---- forward layers: h_{l+1} = relu(h_l W_l^T + b_l); the last one turns straight into dz_{L-1} ----------------
      for (int l = 0; l < L; ++l) {
        const uint32_t acc = wait_acc();
        const float *bl = bias + (1 + l) * W;
        const bool last = l + 1 == L;
        unsigned char *htile = stg.h[l + 1] + toff * W;
        unsigned char *dztile = stg.dz[L] + toff * W;   // dz_{L-1}
        uint32_t mk[MW];
#pragma unroll
        for (int w = 0; w < MW; ++w) mk[w] = 0u;
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
          const int col0 = c * CW + p * CPT;
          float v[CPT];
          tmem_ldn<CPT>(acc + col0, v);
          uint32_t bits = 0u;
#pragma unroll
          for (int j = 0; j < CPT; j += 8) {
            const float4 b0 = *reinterpret_cast<const float4 *>(bl + col0 + j);
            const float4 b1 = *reinterpret_cast<const float4 *>(bl + col0 + j + 4);
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            float x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float z = v[j + u] + bv[u];
              bits |= (z > 0.f ? 1u : 0u) << (j + u);
              x[u] = fmaxf(z, 0.f);
            }
            store_both8<W>(X_hi, X_lo, htile, r, col0 + j, x, !last);
            if (last) {
              float dz[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const float4 w = wos[col0 + j + u];
                const float s = fmaf(dov[0], w.x, fmaf(dov[1], w.y, fmaf(dov[2], w.z, dov[3] * w.w)));
                dz[u] = x[u] > 0.f ? s : 0.f;
              }
              store_both8<W>(X_hi, X_lo, dztile, r, col0 + j, dz, true);
            }
          }
          chunk_done(c);
          const int bit0 = c * CPT;
#pragma unroll
          for (int w = 0; w < MW; ++w)
            if ((bit0 >> 5) == w) mk[w] |= bits << (bit0 & 31);
        }
#pragma unroll
        for (int w = 0; w < MW; ++w) masks[(l + 1) * MW + w] = mk[w];
      }
      // ---- backward layers: dh_l = dz_l W_l ; dz_{l-1} = dh_l * [h_l > 0] -----------------------------------------------
      for (int l = L - 1; l >= 0; --l) {
        const uint32_t acc = wait_acc();
        uint32_t mk[MW];
#pragma unroll
        for (int w = 0; w < MW; ++w) mk[w] = masks[l * MW + w];
        unsigned char *dztile = stg.dz[l] + toff * W;   // dz_{l-1} (l == 0: dz_in)
        float dx[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
          const int col0 = c * CW + p * CPT;
          float v[CPT];
          tmem_ldn<CPT>(acc + col0, v);
          const int bit0 = c * CPT;
          uint32_t bits = 0u;
#pragma unroll
          for (int w = 0; w < MW; ++w)
            if ((bit0 >> 5) == w) bits = mk[w] >> (bit0 & 31);
#pragma unroll
          for (int j = 0; j < CPT; j += 8) {
            float dz[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) dz[u] = ((bits >> (j + u)) & 1u) ? v[j + u] : 0.f;
            store_both8<W>(X_hi, X_lo, dztile, r, col0 + j, dz, l > 0);
            if (l == 0 && d_x) {
#pragma unroll
              for (int ci = 0; ci < 4; ++ci) {
                if (ci < Cin) {
                  const float4 w0 = ldg4(a.w_in_t + (size_t)ci * W + col0 + j);
                  const float4 w1 = ldg4(a.w_in_t + (size_t)ci * W + col0 + j + 4);
                  dx[ci] = fmaf(dz[0], w0.x, fmaf(dz[1], w0.y, fmaf(dz[2], w0.z, fmaf(dz[3], w0.w, dx[ci]))));
                  dx[ci] = fmaf(dz[4], w1.x, fmaf(dz[5], w1.y, fmaf(dz[6], w1.z, fmaf(dz[7], w1.w, dx[ci]))));
                }
              }
            }
          }
          if (l > 0) chunk_done(c);
        }
        if (l == 0) {
          tc_fence_before();
          if (p > 0) part[(p - 1) * 128 + r] = make_float4(dx[0], dx[1], dx[2], dx[3]);
          asm volatile("bar.sync 1, %0;" ::"n"(C::WORKERS * 32) : "memory");
          if (p == 0 && on && d_x) {
#pragma unroll
            for (int q = 0; q < C::NWQ - 1; ++q) {
              const float4 t = part[q * 128 + r];
              dx[0] += t.x; dx[1] += t.y; dx[2] += t.z; dx[3] += t.w;
            }
            for (int ci = 0; ci < Cin; ++ci) d_x[grow * Cin + ci] = dx[ci];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

static long long segment_tiles() {
  static const long long v = [] {
    const char *e = getenv("NSDP_MLP_SEG");
    const long long t = e ? atoll(e) : 592;   // 4 x 148 SMs: whole waves; 1.1 GB of staging per segment at W = 256, L = 6
    return t < 1 || t > 4736 ? 592ll : t;
  }();
  return v;
}

template <class C>
static size_t staged_bytes_per_tile(int L) {
  return (size_t)512 * ((size_t)2 * (L + 1) * C::W + 32);
}

template <class C>
static size_t workspace_bytes(const nsdp_mlp_args &a) {
  const long long tiles = ceil_div((long long)a.R, 128ll);
  const long long seg = tiles < segment_tiles() ? tiles : segment_tiles();
  return packed_bytes<C>(a.n_hidden) + 256 + (size_t)seg * staged_bytes_per_tile<C>(a.n_hidden);
}

template <class C>
static int launch(const nsdp_mlp_args &a, const float *dout, const nsdp_mlp_grads &g, void *workspace, size_t ws_bytes,
                  cudaStream_t st) {
  constexpr int W = C::W;
  const int L = a.n_hidden;
  if (!workspace || ws_bytes < workspace_bytes<C>(a)) return NSDP_ERR_WORKSPACE;
  unsigned char *packed = (unsigned char *)workspace;
  int *err = (int *)(packed + packed_bytes<C>(L));
  unsigned char *sbase = packed + packed_bytes<C>(L) + 256;
  const long long tiles = ceil_div((long long)a.R, 128ll);
  const long long seg = tiles < segment_tiles() ? tiles : segment_tiles();
  Staging stg;
  {
    unsigned char *p = sbase;
    auto take = [&](int width) { unsigned char *q = p; p += (size_t)seg * 512 * width; return q; };
    stg.x = take(16);
    for (int i = 0; i <= MAX_HIDDEN; ++i) stg.h[i] = i <= L ? take(W) : nullptr;
    for (int i = 0; i <= MAX_HIDDEN; ++i) stg.dz[i] = i <= L ? take(W) : nullptr;
    stg.dout = take(16);
  }
  cudaError_t e = cudaMemsetAsync(err, 0, sizeof(int), st);
  if (e != cudaSuccess) return cuda_rc(e);
  pack_mlp_bwd_weights_kernel<C><<<2 * L, 256, 0, st>>>(a, packed);
  int rc = check_launch();
  if (rc != NSDP_OK) return rc;
  auto kern = fused_mlp_bwd_tc_kernel<C>;
  static int per_sm = 0;
  if (per_sm == 0) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return cuda_rc(e);
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return cuda_rc(e);
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess) return cuda_rc(e);
    const int regs_per_warp = ((fa.numRegs * 32 + 255) / 256) * 256;
    int occ = 65536 / (regs_per_warp * (C::THREADS / 32));
    const int by_smem = (228 * 1024) / (C::SMEM + 1024), by_threads = 2048 / C::THREADS;
    const int by_tmem = 512 / (int)C::TMEM_COLS;
    occ = occ < by_smem ? occ : by_smem;
    occ = occ < by_threads ? occ : by_threads;
    occ = occ < by_tmem ? occ : by_tmem;
    per_sm = occ < 1 ? 1 : occ;
  }
  for (long long t0 = 0; t0 < tiles; t0 += seg) {
    const long long t1 = t0 + seg < tiles ? t0 + seg : tiles;
    const long long n = t1 - t0;
    const long long slots = (long long)num_sms() * per_sm;
    const int grid = (int)(n < slots ? n : slots);
    kern<<<grid, C::THREADS, C::SMEM, st>>>(a, dout, g.d_x, packed, stg, t0, t1, err);
    rc = check_launch();
    if (rc != NSDP_OK) return rc;
    dwtc::Job jobs[MAX_HIDDEN + 2];
    int nj = 0;
    // first layer: d_w_in_t[ci][n] = x^T dz_in ; d_b_in = colsum(dz_in)
    jobs[nj++] = {stg.x, stg.dz[0], g.d_w_in_t, 16, W, a.Cin, W, W, g.d_b_in, nullptr, 0, 0, 0};
    // hidden layers: d_w_h_t[l][k][n] = h_l^T dz_l ; d_b_h[l] = colsum(dz_l)
    for (int l = 0; l < L; ++l)
      jobs[nj++] = {stg.h[l], stg.dz[1 + l], g.d_w_h_t + (size_t)l * W * W, W, W, W, W, W, g.d_b_h + (size_t)l * W, nullptr, 0, 0, 0};
    // last layer: d_w_out_t[k][o] = h_L^T d_out ; d_b_out = colsum(d_out)
    jobs[nj++] = {stg.h[L], stg.dout, g.d_w_out_t, W, 16, W, a.O, a.O, g.d_b_out, nullptr, 0, 0, 0};
    const long long bounds[2] = {0, n};
    rc = dw_tc_launch_chunked(jobs, nj, n, bounds, 1, err, st);
    if (rc == NSDP_ERR_UNSUPPORTED) rc = dw_tc_launch(jobs, nj, n, err, st);
    if (rc != NSDP_OK) return rc;
  }
  return NSDP_OK;
}

static int pick(const nsdp_mlp_args &a) {
  if (a.n_hidden < 1 || a.n_hidden > MAX_HIDDEN) return 0;
  switch (a.W) {
    case 16: case 32: case 64: case 128: case 256: return a.W;
    default: return 0;
  }
}

}  // namespace mbtc

size_t mlp_bwd_tc_workspace_bytes(const nsdp_mlp_args *a) {
  switch (mbtc::pick(*a)) {
    case 16: return mbtc::workspace_bytes<mtc::Cfg<16>>(*a);
    case 32: return mbtc::workspace_bytes<mtc::Cfg<32>>(*a);
    case 64: return mbtc::workspace_bytes<mtc::Cfg<64>>(*a);
    case 128: return mbtc::workspace_bytes<mtc::Cfg<128>>(*a);
    case 256: return mbtc::workspace_bytes<mtc::Cfg<256>>(*a);
    default: return 0;
  }
}

int mlp_bwd_tc_dispatch(const nsdp_mlp_args *a, const float *dout, const nsdp_mlp_grads *g, void *workspace, size_t ws_bytes,
                        cudaStream_t st) {
  switch (mbtc::pick(*a)) {
    case 16: return mbtc::launch<mtc::Cfg<16>>(*a, dout, *g, workspace, ws_bytes, st);
    case 32: return mbtc::launch<mtc::Cfg<32>>(*a, dout, *g, workspace, ws_bytes, st);
    case 64: return mbtc::launch<mtc::Cfg<64>>(*a, dout, *g, workspace, ws_bytes, st);
    case 128: return mbtc::launch<mtc::Cfg<128>>(*a, dout, *g, workspace, ws_bytes, st);
    case 256: return mbtc::launch<mtc::Cfg<256>>(*a, dout, *g, workspace, ws_bytes, st);
    default: return NSDP_ERR_UNSUPPORTED;
  }
}

}  // namespace nsdp

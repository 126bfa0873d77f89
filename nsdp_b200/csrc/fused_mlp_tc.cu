// Fused neural-field MLP forward on tcgen05 tensor cores, sm_100a (nsdp_mlp_args; BASELINE.json configs[3]).
//
// Same math as fused_mlp.cu. A persistent CTA walks over tiles of 128 query rows (one row per TMEM lane). The
// activations never leave the SM: layer 0 (K = Cin <= 4) is plain fp32 FMAs in the worker threads, every hidden layer
// is a [128 x W] x [W x W] product in bf16x3 split precision (hi*hi + lo*hi + hi*lo, fp32 accumulate in TMEM), the last
// layer (N = O <= 4) is again fp32 FMAs on the worker side. Packed weights (bf16 hi/lo, canonical K-major slabs, one
// "stage" per 16-wide k-step) stream from L2 through a bulk-copy ring in consumption order.
//
// What is different from the ResNet-FC tail kernel: the GEMM -> epilogue -> GEMM chain of a tile is PIPELINED inside the
// tile. Two accumulators ping-pong in TMEM and the A operand is handed over in column CHUNKS: the workers turn
// accumulator columns [c*CW, (c+1)*CW) into relu(h + b) operand columns and signal chunk c; the issuing warp starts the
// k-steps of the NEXT layer that only need those columns while the workers are still converting the following chunks.
// So the tensor pipe idles only for the first chunk of every layer instead of for the whole epilogue.
#include <cstdlib>

#include "fused_mlp_tc.cuh"

namespace nsdp {
namespace mtc {

using namespace umma;

template <class C>
constexpr size_t packed_bytes(int n_hidden) {
  return (size_t)n_hidden * C::KS * C::STAGE_BYTES;
}

// grid.x = hidden layer; stage (l * KS + ks) = [hi slab][lo slab] of B[n][k] = w_h_t[l][k][n], k in [16 ks, 16 ks + 16)
template <class C>
__global__ void pack_mlp_weights_kernel(const nsdp_mlp_args a, unsigned char *__restrict__ out) {
  constexpr int W = C::W;
  const float *wt = a.w_h_t + (size_t)blockIdx.x * W * W;
  unsigned char *o = out + (size_t)blockIdx.x * C::KS * C::STAGE_BYTES;
  for (int e = threadIdx.x; e < W * (W / 2); e += blockDim.x) {
    const int n = e % W, k = (e / W) * 2;
    uint32_t hi, lo;
    split2(__ldg(wt + (size_t)k * W + n), __ldg(wt + (size_t)(k + 1) * W + n), hi, lo);
    const size_t base = (size_t)(k >> 4) * C::STAGE_BYTES + canon_off(W, n, k & 15);
    *reinterpret_cast<uint32_t *>(o + base) = hi;
    *reinterpret_cast<uint32_t *>(o + base + C::SLAB) = lo;
  }
}

template <class C>
__global__ void __launch_bounds__(C::THREADS, C::MIN_CTAS)
fused_mlp_tc_kernel(const nsdp_mlp_args a, const unsigned char *__restrict__ packed, float *__restrict__ out,
                    long long tiles, int *err) {
  constexpr int W = C::W, STAGES = C::STAGES, NCH = C::NCH, CPT = C::CPT, CW = C::CW;
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
  const int per_tile = L * C::KS;

  if (warp == 0) {
    // weight producer: two lanes share the bulk copies (one thread sustains about one copy per ~500 cycles)
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
    // MMA issuer: the whole warp runs the loops and the waits, one elected lane issues
    const uint32_t idesc = idesc_bf16(128, W);
    constexpr uint32_t lbo_a = 128 * 16, lbo_b = W * 16;
    constexpr uint64_t A_STEP = (2 * lbo_a) >> 4;
    const uint64_t xhi = smem_desc(smem_u32(X_hi), lbo_a, 128), xlo = smem_desc(smem_u32(X_lo), lbo_a, 128);
    const uint64_t bh0 = smem_desc(smem_u32(stage0), lbo_b, 128);
    uint32_t slot = 0, slot_phase = 0, ready_phase = 0, g = 0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      for (int l = 0; l < L; ++l) {
        const uint32_t col = tmem_base + (g & 1u) * W;
        for (int c = 0; c < NCH; ++c) {
          mbar_wait(&a_ready[c], ready_phase, err);   // operand columns [c*CW, (c+1)*CW) of this layer are in place
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
    const int quarter = warp & 3;          // TMEM lane quarter this warp may access
    const int p = (warp - 2) >> 2;         // which CPT-wide share of every chunk
    const int r = quarter * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t done_phase = 0, g = 0;

    auto chunk_done = [&](int c) {
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_ready[c]);
    };

    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const long long grow = tile * 128 + r;
      const bool on = grow < a.R;
      float xin[4] = {0.f, 0.f, 0.f, 0.f};
      if (on) {
#pragma unroll
        for (int ci = 0; ci < 4; ++ci)
          if (ci < Cin) xin[ci] = __ldg(a.x + (size_t)grow * Cin + ci);
      }
      // ---- layer 0: h = relu(x W_in + b_in), fp32 FMAs -> first A operand -----------------------------------------
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        const int col0 = c * CW + p * CPT;
#pragma unroll
        for (int j = 0; j < CPT; j += 8) {
          // 128-bit loads: the (warp-uniform) weight reads of this layer are LSU-issue bound, not latency bound
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
          for (int u = 0; u < 8; ++u) v[u] = fmaxf(v[u], 0.f);
          store_split8(X_hi, X_lo, r, col0 + j, v);
        }
        chunk_done(c);
      }
      // ---- hidden layers ---------------------------------------------------------------------------------------------
      for (int l = 0; l < L; ++l) {
        mbar_wait(acc_done, done_phase, err);
        done_phase ^= 1;
        tc_fence_after();
        const uint32_t acc = trow + (g & 1u) * W;
        ++g;
        const float *bl = bias + (1 + l) * W;
        if (l + 1 < L) {
#pragma unroll 1
          for (int c = 0; c < NCH; ++c) {
            const int col0 = c * CW + p * CPT;
            float v[CPT];
            tmem_ldn<CPT>(acc + col0, v);
#pragma unroll
            for (int j = 0; j < CPT; j += 8) {
              const float4 b0 = *reinterpret_cast<const float4 *>(bl + col0 + j);
              const float4 b1 = *reinterpret_cast<const float4 *>(bl + col0 + j + 4);
              const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              float x[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) x[u] = fmaxf(v[j + u] + bv[u], 0.f);
              store_split8(X_hi, X_lo, r, col0 + j, x);
            }
            chunk_done(c);
          }
        } else {
          // ---- out = relu(h) W_out + b_out ---------------------------------------------------------------------------
          float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
#pragma unroll 1
          for (int c = 0; c < NCH; ++c) {
            const int col0 = c * CW + p * CPT;
            float v[CPT];
            tmem_ldn<CPT>(acc + col0, v);
#pragma unroll
            for (int u = 0; u < CPT; ++u) {
              const float x = fmaxf(v[u] + bl[col0 + u], 0.f);
              const float4 w = wos[col0 + u];
              o0 = fmaf(x, w.x, o0); o1 = fmaf(x, w.y, o1); o2 = fmaf(x, w.z, o2); o3 = fmaf(x, w.w, o3);
            }
          }
          tc_fence_before();
          if (p > 0) part[(p - 1) * 128 + r] = make_float4(o0, o1, o2, o3);
          asm volatile("bar.sync 1, %0;" ::"n"(C::WORKERS * 32) : "memory");
          if (p == 0 && on) {
            float res[4] = {o0, o1, o2, o3};
#pragma unroll
            for (int q = 0; q < C::NWQ - 1; ++q) {
              const float4 t = part[q * 128 + r];
              res[0] += t.x; res[1] += t.y; res[2] += t.z; res[3] += t.w;
            }
            for (int o = 0; o < O; ++o) out[grow * O + o] = res[o] + __ldg(a.b_out + o);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// =====================================================================================================================
// Narrow widths (W <= 64): a [128 x W] x [W x W] product is a handful of 8..32-cycle MMAs, so a tile-at-a-time CTA spends
// its time in the worker -> issuer -> worker round trip of every layer (~2.5 k cycles), not in math. Here one CTA carries
// RT row tiles through the layers in LOCK STEP: per layer and 16-column chunk the workers convert the accumulator columns
// of all RT tiles, hand the chunk over ONCE, and the issuing thread runs that k-step for all RT tiles against the same
// weight slab (one commit). The round trip and the weight stream are paid once per RT x 128 rows; TMEM holds the RT
// accumulator pairs (RT x 2 x W = 256 columns). Worker warps split the ROW TILES (group g owns tiles g, g + G, ...), each
// thread converts whole 16-column chunks of its rows.
// =====================================================================================================================
template <int W_, int RT_, int G_, int CTAS_>
struct NCfg {
  static constexpr int W = W_;
  static constexpr int KS = W / 16;                    // k-steps = chunks per layer
  static constexpr int RT = RT_;                       // row tiles in flight per CTA
  static constexpr int G = G_;                         // worker warp groups (each: 4 warps = the 4 TMEM lane quarters)
  static constexpr int TPG = RT / G;                   // row tiles per group
  static constexpr int CTAS = CTAS_;                   // co-resident CTAs per SM this geometry is sized for
  // one k-step per layer (W = 16): the next layer's MMAs start only after every accumulator column has been read, so a
  // single accumulator per tile is enough; wider layers hand chunks over early and need the ping-pong pair
  static constexpr int ACCS = KS == 1 ? 1 : 2;
  static constexpr int WORKERS = 4 * G;
  static constexpr int THREADS = (2 + WORKERS) * 32;
  static constexpr int SLAB = W * 16 * 2;
  static constexpr int STAGE_BYTES = 2 * SLAB;
  static constexpr int STAGES = W >= 64 ? 4 : 8;
  static constexpr uint32_t TMEM_NEED = RT * ACCS * W;
  static constexpr uint32_t TMEM_COLS = TMEM_NEED <= 32 ? 32 : (TMEM_NEED <= 64 ? 64 : (TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512)));
  static constexpr int A_HALF = 128 * W * 2;           // hi (or lo) operand of ONE row tile
  static constexpr int OFF_X = 0;                      // [RT][hi | lo]
  static constexpr int OFF_STAGE = OFF_X + RT * 2 * A_HALF;
  static constexpr int OFF_BIAS = OFF_STAGE + STAGES * STAGE_BYTES;   // float[1 + MAX_HIDDEN][W]
  static constexpr int OFF_WO = OFF_BIAS + (1 + MAX_HIDDEN) * W * 4;  // float4[W]
  static constexpr int OFF_BAR = OFF_WO + W * 16;
  static constexpr int SMEM = OFF_BAR + 256;
  static_assert(RT % G == 0 && TMEM_NEED <= 512, "row-tile batching");
  static_assert(CTAS * TMEM_COLS <= 512 && CTAS * (SMEM + 1024) <= 228 * 1024 && CTAS * THREADS <= 2048, "occupancy");
  static_assert((2 * STAGES + KS + 1) * 8 + 4 <= 256, "barrier block");
};

template <class C>
__global__ void __launch_bounds__(C::THREADS, C::CTAS)
fused_mlp_tc_narrow_kernel(const nsdp_mlp_args a, const unsigned char *__restrict__ packed, float *__restrict__ out,
                           long long groups, int *err) {
  constexpr int W = C::W, STAGES = C::STAGES, KS = C::KS, RT = C::RT;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char *X = smem + C::OFF_X;
  unsigned char *stage0 = smem + C::OFF_STAGE;
  float *bias = reinterpret_cast<float *>(smem + C::OFF_BIAS);
  float4 *wos = reinterpret_cast<float4 *>(smem + C::OFF_WO);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::OFF_BAR);
  uint64_t *full = bars, *empty = bars + STAGES, *a_ready = bars + 2 * STAGES, *acc_done = a_ready + KS;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_done + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = a.n_hidden, Cin = a.Cin, O = a.O;

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
    for (int c = 0; c < KS; ++c) mbar_init(&a_ready[c], C::WORKERS);
    mbar_init(acc_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int per_group = L * KS;

  if (warp == 0) {
    if (lane == 0) {
      const long long mine = (long long)blockIdx.x < groups ? (groups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
      const long long total = mine * per_group;
      for (long long it = 0; it < total; ++it) {
        const int st = (int)(it % per_group);
        const int s = (int)(it % STAGES);
        const uint32_t ph = (uint32_t)(it / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1, err);
        mbar_arrive_expect_tx(&full[s], C::STAGE_BYTES);
        bulk_g2s(stage0 + (size_t)s * C::STAGE_BYTES, packed + (size_t)st * C::STAGE_BYTES, C::STAGE_BYTES, &full[s]);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = idesc_bf16(128, W);
    constexpr uint32_t lbo_a = 128 * 16, lbo_b = W * 16;
    constexpr uint64_t A_STEP = (2 * lbo_a) >> 4;
    const uint64_t x0 = smem_desc(smem_u32(X), lbo_a, 128);
    const uint64_t bh0 = smem_desc(smem_u32(stage0), lbo_b, 128);
    uint32_t slot = 0, slot_phase = 0, ready_phase = 0, g = 0;
    for (long long grp = blockIdx.x; grp < groups; grp += gridDim.x) {
      for (int l = 0; l < L; ++l) {
        const uint32_t col = tmem_base + (C::ACCS == 2 ? (g & 1u) * W : 0u);
        for (int ks = 0; ks < KS; ++ks) {
          mbar_wait(&a_ready[ks], ready_phase, err);   // operand columns [16 ks, 16 ks + 16) of all RT tiles are in place
          tc_fence_after();
          mbar_wait(&full[slot], slot_phase, err);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t bh = bh0 + (uint64_t)slot * (C::STAGE_BYTES >> 4);
#pragma unroll
            for (int t = 0; t < RT; ++t) {
              const uint64_t ah = x0 + (uint64_t)t * ((2 * C::A_HALF) >> 4) + ks * A_STEP, al = ah + (C::A_HALF >> 4);
              const uint32_t d = col + (uint32_t)t * C::ACCS * W;
              mma_bf16(d, ah, bh, idesc, ks != 0);
              mma_bf16(d, al, bh, idesc, true);
              mma_bf16(d, ah, bh + (C::SLAB >> 4), idesc, true);
            }
            mma_commit(&empty[slot]);
          }
          __syncwarp();
          if (++slot == STAGES) { slot = 0; slot_phase ^= 1; }
        }
        if (elect_one()) mma_commit(acc_done);
        __syncwarp();
        ready_phase ^= 1;
        ++g;
      }
    }
  } else {
    const int quarter = warp & 3;          // TMEM lane quarter this warp may access
    const int grp_w = (warp - 2) >> 2;     // worker group: owns row tiles grp_w, grp_w + G, ...
    const int r = quarter * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t done_phase = 0, g = 0;

    auto chunk_done = [&](int c) {
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_ready[c]);
    };

    for (long long grp = blockIdx.x; grp < groups; grp += gridDim.x) {
      float xin[C::TPG][4];
      long long grow[C::TPG];
#pragma unroll
      for (int i = 0; i < C::TPG; ++i) {
        const int t = grp_w + i * C::G;
        grow[i] = (grp * RT + t) * 128 + r;
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) xin[i][ci] = (grow[i] < a.R && ci < Cin) ? __ldg(a.x + (size_t)grow[i] * Cin + ci) : 0.f;
      }
      // ---- layer 0: h = relu(x W_in + b_in), fp32 FMAs -> first A operand -----------------------------------------
#pragma unroll 1
      for (int c = 0; c < KS; ++c) {
        const int col0 = c * 16;
#pragma unroll
        for (int j = 0; j < 16; j += 8) {
          // 128-bit warp-uniform weight loads (L1 hits after the first tile)
          float wv[4][8];
#pragma unroll
          for (int ci = 0; ci < 4; ++ci) {
            const float4 w0 = ci < Cin ? ldg4(a.w_in_t + (size_t)ci * W + col0 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 w1 = ci < Cin ? ldg4(a.w_in_t + (size_t)ci * W + col0 + j + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            wv[ci][0] = w0.x; wv[ci][1] = w0.y; wv[ci][2] = w0.z; wv[ci][3] = w0.w;
            wv[ci][4] = w1.x; wv[ci][5] = w1.y; wv[ci][6] = w1.z; wv[ci][7] = w1.w;
          }
#pragma unroll
          for (int i = 0; i < C::TPG; ++i) {
            const int t = grp_w + i * C::G;
            unsigned char *X_hi = X + (size_t)t * 2 * C::A_HALF, *X_lo = X_hi + C::A_HALF;
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              float acc = bias[col0 + j + u];
#pragma unroll
              for (int ci = 0; ci < 4; ++ci) acc = fmaf(xin[i][ci], wv[ci][u], acc);
              v[u] = fmaxf(acc, 0.f);
            }
            store_split8(X_hi, X_lo, r, col0 + j, v);
          }
        }
        chunk_done(c);
      }
      // ---- hidden layers ---------------------------------------------------------------------------------------------
      for (int l = 0; l < L; ++l) {
        mbar_wait(acc_done, done_phase, err);
        done_phase ^= 1;
        tc_fence_after();
        const uint32_t acc = trow + (C::ACCS == 2 ? (g & 1u) * W : 0u);
        ++g;
        const float *bl = bias + (1 + l) * W;
        if (l + 1 < L) {
#pragma unroll 1
          for (int c = 0; c < KS; ++c) {
            const int col0 = c * 16;
            float bv[16];
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 b4 = *reinterpret_cast<const float4 *>(bl + col0 + j);
              bv[j] = b4.x; bv[j + 1] = b4.y; bv[j + 2] = b4.z; bv[j + 3] = b4.w;
            }
#pragma unroll
            for (int i = 0; i < C::TPG; ++i) {
              const int t = grp_w + i * C::G;
              unsigned char *X_hi = X + (size_t)t * 2 * C::A_HALF, *X_lo = X_hi + C::A_HALF;
              float v[16];
              tmem_ld16(acc + (uint32_t)t * C::ACCS * W + col0, v);
#pragma unroll
              for (int j = 0; j < 16; j += 8) {
                float x[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) x[u] = fmaxf(v[j + u] + bv[j + u], 0.f);
                store_split8(X_hi, X_lo, r, col0 + j, x);
              }
            }
            chunk_done(c);
          }
        } else {
          // ---- out = relu(h) W_out + b_out ---------------------------------------------------------------------------
#pragma unroll
          for (int i = 0; i < C::TPG; ++i) {
            const int t = grp_w + i * C::G;
            float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
#pragma unroll 1
            for (int c = 0; c < KS; ++c) {
              const int col0 = c * 16;
              float v[16];
              tmem_ld16(acc + (uint32_t)t * C::ACCS * W + col0, v);
#pragma unroll
              for (int u = 0; u < 16; ++u) {
                const float x = fmaxf(v[u] + bl[col0 + u], 0.f);
                const float4 w = wos[col0 + u];
                o0 = fmaf(x, w.x, o0); o1 = fmaf(x, w.y, o1); o2 = fmaf(x, w.z, o2); o3 = fmaf(x, w.w, o3);
              }
            }
            if (grow[i] < a.R) {
              const float res[4] = {o0, o1, o2, o3};
              for (int o = 0; o < O; ++o) out[grow[i] * O + o] = res[o] + __ldg(a.b_out + o);
            }
          }
          tc_fence_before();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

template <class C>
static int launch_narrow(const nsdp_mlp_args &a, float *out, void *workspace, size_t ws_bytes, cudaStream_t st) {
  using P = Cfg<C::W>;     // the packed weight image is the one of the wide kernel (same slabs, same order)
  const size_t pb = packed_bytes<P>(a.n_hidden);
  if (!workspace || ws_bytes < pb + 16) return NSDP_ERR_WORKSPACE;
  unsigned char *packed = (unsigned char *)workspace;
  int *err = (int *)(packed + pb);
  cudaError_t e = cudaMemsetAsync(err, 0, sizeof(int), st);
  if (e != cudaSuccess) return cuda_rc(e);
  if (!a.reuse_packed) {
    pack_mlp_weights_kernel<P><<<a.n_hidden, 256, 0, st>>>(a, packed);
    int rc = check_launch();
    if (rc != NSDP_OK) return rc;
  }
  auto kern = fused_mlp_tc_narrow_kernel<C>;
  static bool ready = false;
  if (!ready) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return cuda_rc(e);
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return cuda_rc(e);
    ready = true;
  }
  const long long groups = ceil_div((long long)a.R, (long long)(128 * C::RT));
  const long long slots = (long long)num_sms() * C::CTAS;
  const int grid = (int)(groups < slots ? groups : slots);
  kern<<<grid, C::THREADS, C::SMEM, st>>>(a, packed, out, groups, err);
  return check_launch();
}

// Geometry of the row-tile batched kernel for W <= 64. Default: the fastest of a sweep on B200 (1M rows, 8 layers; tile-at-a-
// time kernel -> batched): W = 16: 69.8 -> 60.8 us (8 tiles, 2 warp groups, 3 CTAs / SM), W = 32: 120.7 -> 106.8 us (2 tiles, 1
// group, 4 CTAs), W = 64: 255.2 -> 219.6 us (2 tiles, 1 group, 2 CTAs). NSDP_MLP_NARROW=<n> picks variant n of the switch
// below for tuning runs, 0 the tile-at-a-time kernel.
static int narrow_variant(int best) {
  static const int v = [] { const char *e = getenv("NSDP_MLP_NARROW"); return e ? atoi(e) : -1; }();
  return v < 0 ? best : v;
}

template <class C>
static int launch(const nsdp_mlp_args &a, float *out, void *workspace, size_t ws_bytes, cudaStream_t st) {
  const size_t pb = packed_bytes<C>(a.n_hidden);
  if (!workspace || ws_bytes < pb + 16) return NSDP_ERR_WORKSPACE;
  unsigned char *packed = (unsigned char *)workspace;
  int *err = (int *)(packed + pb);
  cudaError_t e = cudaMemsetAsync(err, 0, sizeof(int), st);
  if (e != cudaSuccess) return cuda_rc(e);
  if (!a.reuse_packed) {
    pack_mlp_weights_kernel<C><<<a.n_hidden, 256, 0, st>>>(a, packed);
    int rc = check_launch();
    if (rc != NSDP_OK) return rc;
  }
  auto kern = fused_mlp_tc_kernel<C>;
  static int per_sm = 0;   // resident CTAs per SM: shared memory, threads, registers and TMEM columns
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
    static const int forced_occ = [] { const char *env = getenv("NSDP_MLP_CTAS_PER_SM"); return env ? atoi(env) : 0; }();
    if (forced_occ > 0) occ = forced_occ;   // tuning experiments
    per_sm = occ < 1 ? 1 : occ;
  }
  const long long tiles = ceil_div((long long)a.R, 128ll);
  const long long slots = (long long)num_sms() * per_sm;
  const int grid = (int)(tiles < slots ? tiles : slots);
  kern<<<grid, C::THREADS, C::SMEM, st>>>(a, packed, out, tiles, err);
  return check_launch();
}

static int pick(const nsdp_mlp_args &a) {
  if (a.n_hidden < 1 || a.n_hidden > MAX_HIDDEN) return 0;
  switch (a.W) {
    case 16: case 32: case 64: case 128: case 256: return a.W;
    default: return 0;
  }
}

}  // namespace mtc

size_t mlp_tc_workspace_bytes(const nsdp_mlp_args *a) {
  switch (mtc::pick(*a)) {
    case 16: return mtc::packed_bytes<mtc::Cfg<16>>(a->n_hidden) + 16;
    case 32: return mtc::packed_bytes<mtc::Cfg<32>>(a->n_hidden) + 16;
    case 64: return mtc::packed_bytes<mtc::Cfg<64>>(a->n_hidden) + 16;
    case 128: return mtc::packed_bytes<mtc::Cfg<128>>(a->n_hidden) + 16;
    case 256: return mtc::packed_bytes<mtc::Cfg<256>>(a->n_hidden) + 16;
    default: return 0;
  }
}

int mlp_tc_dispatch(const nsdp_mlp_args *a, float *out, void *workspace, size_t ws_bytes, cudaStream_t st, bool *handled) {
  *handled = true;
  switch (mtc::pick(*a)) {
    case 16:
      switch (mtc::narrow_variant(2)) {
        case 1: return mtc::launch_narrow<mtc::NCfg<16, 8, 4, 2>>(*a, out, workspace, ws_bytes, st);
        case 2: return mtc::launch_narrow<mtc::NCfg<16, 8, 2, 3>>(*a, out, workspace, ws_bytes, st);
        case 3: return mtc::launch_narrow<mtc::NCfg<16, 4, 1, 5>>(*a, out, workspace, ws_bytes, st);
        case 4: return mtc::launch_narrow<mtc::NCfg<16, 4, 2, 3>>(*a, out, workspace, ws_bytes, st);
        case 5: return mtc::launch_narrow<mtc::NCfg<16, 2, 1, 6>>(*a, out, workspace, ws_bytes, st);
        default: return mtc::launch<mtc::Cfg<16>>(*a, out, workspace, ws_bytes, st);
      }
    case 32:
      switch (mtc::narrow_variant(3)) {
        case 1: return mtc::launch_narrow<mtc::NCfg<32, 4, 4, 2>>(*a, out, workspace, ws_bytes, st);
        case 2: return mtc::launch_narrow<mtc::NCfg<32, 4, 2, 2>>(*a, out, workspace, ws_bytes, st);
        case 3: return mtc::launch_narrow<mtc::NCfg<32, 2, 1, 4>>(*a, out, workspace, ws_bytes, st);
        case 4: return mtc::launch_narrow<mtc::NCfg<32, 2, 2, 4>>(*a, out, workspace, ws_bytes, st);
        default: return mtc::launch<mtc::Cfg<32>>(*a, out, workspace, ws_bytes, st);
      }
    case 64:
      switch (mtc::narrow_variant(2)) {
        case 1: return mtc::launch_narrow<mtc::NCfg<64, 2, 2, 2>>(*a, out, workspace, ws_bytes, st);
        case 2: return mtc::launch_narrow<mtc::NCfg<64, 2, 1, 2>>(*a, out, workspace, ws_bytes, st);
        case 3: return mtc::launch_narrow<mtc::NCfg<64, 1, 1, 4>>(*a, out, workspace, ws_bytes, st);
        default: return mtc::launch<mtc::Cfg<64>>(*a, out, workspace, ws_bytes, st);
      }
    case 128: return mtc::launch<mtc::Cfg<128>>(*a, out, workspace, ws_bytes, st);
    case 256: return mtc::launch<mtc::Cfg<256>>(*a, out, workspace, ws_bytes, st);
    default: *handled = false; return NSDP_OK;
  }
}

}  // namespace nsdp

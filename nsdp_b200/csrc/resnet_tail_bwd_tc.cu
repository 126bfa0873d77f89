// Backward of the fused decoder tail on tcgen05 tensor cores (sm_100a): chain kernel + staged weight gradients.
//
// Per tile of 128 query rows a persistent CTA recomputes the forward (net accumulating in TMEM exactly like
// resnet_tail_tc.cu), keeps the ReLU masks of relu(net_i) / relu(h_i) as bit masks in registers, and then walks the
// blocks backwards:
//     dnet = (dout*Wo^T)*[x_n > 0]
//     for i = n-1 .. 0:   dy = dnet*W1_i ; dhh = dy*[y_i > 0] ; dx = dhh*W0_i ; dn = dnet + dx*[x_i > 0]
//                         dlat += dn*Wc_{i+1} ; dnet = dn
//     dlat += dnet*Wc_0                                   (d_lat accumulates in TMEM over the six products)
// Weight / bias gradients: the operand tiles lat, x_i, y_i, dhh_i, dn_i and dout are staged (bf16 hi/lo, layout of
// dw_tc.cu) and reduced by dw_tc_kernel (d_w = X^T Y, d_b = column sums of Y).
#include <stdlib.h>

#include "common.cuh"
#include "dw_tc.cuh"
#include "stage_f16.cuh"
#include "umma.cuh"

namespace nsdp {
namespace tbtc {

using namespace umma;

constexpr int H = 128;
constexpr int KS_H = H / 16;
constexpr int SLAB_H = H * 16 * 2;        // [128 x 16] bf16 slab
constexpr int SLOT_BYTES = 2 * SLAB_H;   // 8 KB: one 128-row matrix k-step (hi + lo slab) or ONE slab of a wider matrix
constexpr int NPART = 4;
constexpr int WORKER_WARPS = 4 * NPART;
constexpr int THREADS = (2 + WORKER_WARPS) * 32;
constexpr int MAXB = 5;                   // blocks supported by the register mask budget
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t ACC_H = 128, ACC_DLAT = 256;
constexpr int XPT = H / NPART;            // hidden columns per worker thread (32)

template <int CP_>
struct Cfg {
  static constexpr int CP = CP_;
  static constexpr int KS_C = CP / 16;
  static constexpr int SLAB_C = CP * 16 * 2;                       // [CP x 16] slab (d_lat products: N = CP)
  // The weight ring is addressed in 8 KB slots. A k-step of an H-row matrix (hi + lo slab) takes one slot; a k-step of a
  // CP-row matrix (d_lat products) takes one slot if both slabs fit, else two (hi slab, lo slab). With 3 stages of
  // 13 KB the ring could not cover the L2 latency: the MMAs ran at 170 cycles instead of 64 (trace, round 1).
  static constexpr bool C_SPLIT = 2 * SLAB_C > SLOT_BYTES;
  static constexpr int C_SLOTS = C_SPLIT ? 2 : 1;                  // slots per k-step of a CP-row matrix
  static constexpr int STAGES = CP_ > 128 ? 6 : 8;                 // ring depth in slots
  static constexpr int STAGE_BYTES = SLOT_BYTES;
  static constexpr int A_LAT_HALF = 128 * CP * 2;
  static constexpr int A_X_HALF = 128 * H * 2;
  static constexpr int LCHUNKS = CP / 8;
  static constexpr int LMAXCH = (LCHUNKS + NPART - 1) / NPART;
  static constexpr int OFF_LAT = 0;
  static constexpr int OFF_X = OFF_LAT + 2 * A_LAT_HALF;
  static constexpr int OFF_STAGE = OFF_X + 2 * A_X_HALF;
  static constexpr int OFF_BSUM = OFF_STAGE + STAGES * SLOT_BYTES;    // float[MAXB + 1][H]
  static constexpr int OFF_B0 = OFF_BSUM + (MAXB + 1) * H * 4;         // float[MAXB][H]
  static constexpr int OFF_WO = OFF_B0 + MAXB * H * 4;                 // float4[H]
  static constexpr int OFF_DO = OFF_WO + H * 16;                       // float4[128]  dout rows of the tile
  static constexpr int OFF_BAR = OFF_DO + 128 * 16;
  static constexpr int SMEM = OFF_BAR + 256;
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

// ---- packed weights -----------------------------------------------------------------------------------------
// forward region (stage = 2 * SLAB_H): init, c0, {w0_i, c_{i+1}, w1_i}
// backward region: for i = n-1..0: w1b_i (2*SLAB_H), w0b_i (2*SLAB_H), wcb_{i+1} (2*SLAB_C); then wcb_0
template <class C>
__host__ __device__ constexpr size_t fwd_region_bytes(int nb) {
  return (size_t)((1 + nb) * C::KS_C + 2 * nb * KS_H) * 2 * SLAB_H;
}
template <class C>
__host__ __device__ constexpr size_t bwd_region_bytes(int nb) {
  return (size_t)nb * 2 * KS_H * 2 * SLAB_H + (size_t)(nb + 1) * KS_H * 2 * C::SLAB_C;
}
template <class C>
constexpr size_t packed_bytes(int nb) {
  return fwd_region_bytes<C>(nb) + bwd_region_bytes<C>(nb);
}

// B[n][k] = src[n * sn + k * sk] for n < nv, k < kv (zero-padded to nrows x kpad), as consecutive k-step stages
// [hi slab][lo slab] of an (nrows x 16) slab each.
__device__ __forceinline__ void pack_generic(const float *__restrict__ src, size_t sn, size_t sk, int nv, int kv, int nrows,
                                             int kpad, unsigned char *__restrict__ out, int tid, int nthreads) {
  const int slab = nrows * 32;
  const int total = nrows * (kpad / 2);
  for (int e = tid; e < total; e += nthreads) {
    const int n = e / (kpad / 2), k = (e - n * (kpad / 2)) * 2;
    const float x0 = (n < nv && k < kv) ? src[n * sn + k * sk] : 0.f;
    const float x1 = (n < nv && k + 1 < kv) ? src[n * sn + (k + 1) * sk] : 0.f;
    uint32_t hi, lo;
    split2(x0, x1, hi, lo);
    const size_t base = (size_t)(k >> 4) * 2 * slab + canon_off(nrows, n, k & 15);
    *reinterpret_cast<uint32_t *>(out + base) = hi;
    *reinterpret_cast<uint32_t *>(out + base + slab) = lo;
  }
}

template <class C>
__global__ void pack_tail_bwd_weights_kernel(const nsdp_tail_args a, unsigned char *__restrict__ out) {
  const int nb = a.n_blocks;
  const size_t wld = (size_t)(1 + nb) * H;
  const int m = blockIdx.x;
  int idx = 0;
  size_t off = 0;
  const int tid = threadIdx.x, nt = blockDim.x;
  // ---- forward role: B[n = out channel][k = in] = Wt[k][n]
  for (int j = 0; j < 2 && j <= nb; ++j) {
    if (idx++ == m) { pack_generic(a.wc_t + (size_t)j * H, 1, wld, H, a.C, H, C::CP, out + off, tid, nt); return; }
    off += (size_t)C::KS_C * 2 * SLAB_H;
    if (nb == 0) break;
  }
  for (int i = 0; i < nb; ++i) {
    if (idx++ == m) { pack_generic(a.w0_t + (size_t)i * H * H, 1, H, H, H, H, H, out + off, tid, nt); return; }
    off += (size_t)KS_H * 2 * SLAB_H;
    if (i + 1 < nb) {
      if (idx++ == m) { pack_generic(a.wc_t + (size_t)(i + 2) * H, 1, wld, H, a.C, H, C::CP, out + off, tid, nt); return; }
      off += (size_t)C::KS_C * 2 * SLAB_H;
    }
    if (idx++ == m) { pack_generic(a.w1_t + (size_t)i * H * H, 1, H, H, H, H, H, out + off, tid, nt); return; }
    off += (size_t)KS_H * 2 * SLAB_H;
  }
  // ---- backward role: B[n = in][k = out channel] = Wt[n][k]
  for (int i = nb - 1; i >= 0; --i) {
    if (idx++ == m) { pack_generic(a.w1_t + (size_t)i * H * H, H, 1, H, H, H, H, out + off, tid, nt); return; }
    off += (size_t)KS_H * 2 * SLAB_H;
    if (idx++ == m) { pack_generic(a.w0_t + (size_t)i * H * H, H, 1, H, H, H, H, out + off, tid, nt); return; }
    off += (size_t)KS_H * 2 * SLAB_H;
    if (idx++ == m) { pack_generic(a.wc_t + (size_t)(i + 1) * H, wld, 1, a.C, H, C::CP, H, out + off, tid, nt); return; }
    off += (size_t)KS_H * 2 * C::SLAB_C;
  }
  if (idx++ == m) pack_generic(a.wc_t, wld, 1, a.C, H, C::CP, H, out + off, tid, nt);
}

template <class C>
__host__ __device__ constexpr int num_pack_matrices(int nb) {
  return (nb == 0 ? 1 : 2 + 2 * nb + (nb - 1)) + 3 * nb + 1;
}

// staged tensors of one segment (tile stride = 512 * width bytes)
struct Staging {
  unsigned char *lat;              // width CP
  unsigned char *x[MAXB + 1];      // relu(net_i), i = 0..n   (x[n] feeds fc_out)
  unsigned char *y[MAXB];          // relu(h_i)
  unsigned char *dhh[MAXB];
  unsigned char *dn[MAXB + 1];     // dn[i] = d n_i (i < n), dn[n] = d net_n
  unsigned char *dout;             // width 16
  int lo;                          // 1: [hi slab][lo slab] per k-step (fp32-grade), 0: hi slab only (plain bf16 operand tiles)
  int f16;                         // 1 (lo == 0): the slab holds fp16 values, gradient tiles scaled by 2^k (stage_f16.cuh)
  const unsigned *gmax;            // float bits of max|d_out| the scale is derived from
};

template <int W>
__device__ __forceinline__ void stage_write(unsigned char *tile, int r, int k0, const uint4 &hi, const uint4 &lo, int has_lo) {
  unsigned char *p = tile + (size_t)(r >> 4) * ((has_lo ? 2 : 1) * W * 32) + (size_t)(k0 >> 3) * 256 + (r & 15) * 16;
  *reinterpret_cast<uint4 *>(p) = hi;
  if (has_lo) *reinterpret_cast<uint4 *>(p + W * 32) = lo;
}

template <int W>
__device__ __forceinline__ void stage_write16(unsigned char *tile, int r, int k0, const float (&x)[8], float sc) {
  unsigned char *p = tile + (size_t)(r >> 4) * (W * 32) + (size_t)(k0 >> 3) * 256 + (r & 15) * 16;
  *reinterpret_cast<uint4 *>(p) = stage16::pack8(x, sc);
}

__device__ __forceinline__ void split8(const float (&x)[8], uint4 &hi, uint4 &lo) {
  split2(x[0], x[1], hi.x, lo.x);
  split2(x[2], x[3], hi.y, lo.y);
  split2(x[4], x[5], hi.z, lo.z);
  split2(x[6], x[7], hi.w, lo.w);
}

template <class C>
__global__ void __launch_bounds__(THREADS, 1)
resnet_tail_bwd_tc_kernel(const nsdp_tail_args a, const float *__restrict__ dout, float *__restrict__ d_lat,
                          const unsigned char *__restrict__ packed, const Staging stg, long long tile_begin,
                          long long tile_end, int *err, unsigned long long *trace) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char *L_hi = smem + C::OFF_LAT, *L_lo = L_hi + C::A_LAT_HALF;
  unsigned char *X_hi = smem + C::OFF_X, *X_lo = X_hi + C::A_X_HALF;
  unsigned char *stage0 = smem + C::OFF_STAGE;
  float *bsum = reinterpret_cast<float *>(smem + C::OFF_BSUM);
  float *b0s = reinterpret_cast<float *>(smem + C::OFF_B0);
  float4 *wos = reinterpret_cast<float4 *>(smem + C::OFF_WO);
  float4 *dos = reinterpret_cast<float4 *>(smem + C::OFF_DO);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::OFF_BAR);
  constexpr int STAGES = C::STAGES;
  uint64_t *full = bars, *empty = bars + STAGES, *a_ready = bars + 2 * STAGES, *acc_done = a_ready + 1;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_done + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nb = a.n_blocks, Cin = a.C, O = a.O;

  for (int c = tid; c < H; c += THREADS) {
    float s = a.bc[c];
    for (int i = 0; i <= nb; ++i) {
      if (i < nb) s += a.bc[(size_t)(i + 1) * H + c];
      bsum[i * H + c] = s;
      if (i < nb) {
        b0s[i * H + c] = a.b0[(size_t)i * H + c];
        s += a.b1[(size_t)i * H + c];
      }
    }
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    w.x = a.wo_t[(size_t)c * O + 0];
    if (O > 1) w.y = a.wo_t[(size_t)c * O + 1];
    if (O > 2) w.z = a.wo_t[(size_t)c * O + 2];
    if (O > 3) w.w = a.wo_t[(size_t)c * O + 3];
    wos[c] = w;
  }
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(a_ready, WORKER_WARPS);
    mbar_init(acc_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t ST_H = 2 * SLAB_H, ST_C = 2 * C::SLAB_C;

  if (warp == 0) {
    // ===================== weight producer: streams the packed image once per tile, in order =====================
    // PL lanes share the bulk copies (lane l serves stages l, l + PL, ...): one thread sustains only about one copy per
    // ~500 cycles, less than the tensor pipe drains
    constexpr int PL = 2;
    if (lane < PL) {
      const size_t fwd_bytes = fwd_region_bytes<C>(nb);
      const int nfwd = (1 + nb) * C::KS_C + 2 * nb * KS_H;
      constexpr int CS = C::C_SLOTS;
      constexpr int PER_BLK = 2 * KS_H + CS * KS_H;                       // w1b_i, w0b_i (2 KS_H slots), wcb_{i+1} (CS x KS_H slots)
      constexpr size_t BLK_BYTES = (size_t)2 * KS_H * ST_H + (size_t)KS_H * ST_C;
      constexpr uint32_t C_BYTES = ST_C / CS;                              // one slot of a CP-row k-step
      const int per_tile = nfwd + nb * PER_BLK + CS * KS_H;               // ... + wcb_0
      const long long first = tile_begin + blockIdx.x;
      const long long my_tiles = first < tile_end ? (tile_end - first + gridDim.x - 1) / gridDim.x : 0;
      const long long total = my_tiles * per_tile;
      for (long long it = lane; it < total; it += PL) {
        const int st = (int)(it % per_tile);
        const unsigned char *src;
        uint32_t bytes;
        if (st < nfwd) {
          src = packed + (size_t)st * ST_H; bytes = ST_H;
        } else {
          const int u = st - nfwd, blk = u / PER_BLK, w = u - blk * PER_BLK;
          if (blk < nb) {
            src = packed + fwd_bytes + (size_t)blk * BLK_BYTES +
                  (w < 2 * KS_H ? (size_t)w * ST_H : (size_t)2 * KS_H * ST_H + (size_t)(w - 2 * KS_H) * C_BYTES);
            bytes = w < 2 * KS_H ? ST_H : C_BYTES;
          } else {
            src = packed + fwd_bytes + (size_t)nb * BLK_BYTES + (size_t)(u - nb * PER_BLK) * C_BYTES; bytes = C_BYTES;
          }
        }
        const int s = (int)(it % STAGES);
        const uint32_t ph = (uint32_t)(it / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1, err);
        mbar_arrive_expect_tx(&full[s], bytes);
        bulk_g2s(stage0 + (size_t)s * C::STAGE_BYTES, src, bytes, &full[s]);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // whole warp runs loops and waits, one elected lane issues (elect_one: straight UTCHMMA issue); descriptors advance
    // by adds, the ring position is a running counter (the issuing thread must keep ahead of the tensor pipe)
    {
      const uint32_t idesc_h = idesc_bf16(128, H), idesc_c = idesc_bf16(128, C::CP);
      constexpr uint32_t lbo_a = 128 * 16;
      constexpr uint64_t A_STEP = (2 * lbo_a) >> 4;
      const uint64_t lhi = smem_desc(smem_u32(L_hi), lbo_a, 128), llo = smem_desc(smem_u32(L_lo), lbo_a, 128);
      const uint64_t xhi = smem_desc(smem_u32(X_hi), lbo_a, 128), xlo = smem_desc(smem_u32(X_lo), lbo_a, 128);
      const uint32_t stage_addr = smem_u32(stage0);
      uint32_t slot = 0, slot_phase = 0, ready_phase = 0;
      // A (hi/lo descriptors, `ksteps` k-steps) x the next weight k-steps of N = nrows -> TMEM column `col`
      auto gemm = [&](uint64_t a_hi, uint64_t a_lo, int ksteps, int nrows, uint32_t idesc, uint32_t col, bool fresh) {
        const uint32_t lbo_b = nrows * 16, slab = nrows * 32;
        const uint64_t bh0 = smem_desc(stage_addr, lbo_b, 128);
        const bool split = C::C_SPLIT && nrows > H;   // hi slab and lo slab arrive in two slots
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t ah = a_hi + ks * A_STEP, al = a_lo + ks * A_STEP;
          mbar_wait(&full[slot], slot_phase, err);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t bh = bh0 + (uint64_t)slot * (SLOT_BYTES >> 4);
            mma_bf16(tmem_base + col, ah, bh, idesc, !(fresh && ks == 0));
            mma_bf16(tmem_base + col, al, bh, idesc, true);
            if (!split) mma_bf16(tmem_base + col, ah, bh + (slab >> 4), idesc, true);
            mma_commit(&empty[slot]);
          }
          if (++slot == STAGES) { slot = 0; slot_phase ^= 1; }
          if (split) {
            mbar_wait(&full[slot], slot_phase, err);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t bl = bh0 + (uint64_t)slot * (SLOT_BYTES >> 4);
              mma_bf16(tmem_base + col, ah, bl, idesc, true);
              mma_commit(&empty[slot]);
            }
            if (++slot == STAGES) { slot = 0; slot_phase ^= 1; }
          }
        }
      };
      auto commit_acc = [&]() {
        TR(102);
        if (elect_one()) mma_commit(acc_done);
      };
      auto wait_ready = [&]() {
        TR(100);
        mbar_wait(a_ready, ready_phase, err);
        ready_phase ^= 1;
        tc_fence_after();
        TR(101);
      };
      for (long long tile = tile_begin + blockIdx.x; tile < tile_end; tile += gridDim.x) {
        // ---- forward recompute ----
        wait_ready();
        gemm(lhi, llo, C::KS_C, H, idesc_h, 0, true);
        if (nb > 0) gemm(lhi, llo, C::KS_C, H, idesc_h, 0, false);
        commit_acc();
        for (int i = 0; i < nb; ++i) {
          wait_ready();
          gemm(xhi, xlo, KS_H, H, idesc_h, ACC_H, true);
          commit_acc();
          if (i + 1 < nb) gemm(lhi, llo, C::KS_C, H, idesc_h, 0, false);
          wait_ready();
          gemm(xhi, xlo, KS_H, H, idesc_h, 0, false);
          commit_acc();
        }
        // ---- backward ----
        for (int i = nb - 1; i >= 0; --i) {
          wait_ready();                                            // dnet (= d net_{i+1}) operand in X
          if (i < nb - 1) gemm(xhi, xlo, KS_H, C::CP, idesc_c, ACC_DLAT, i == nb - 2);   // dlat += dn_{i+1} * Wc_{i+2}
          gemm(xhi, xlo, KS_H, H, idesc_h, ACC_H, true);           // dy = dnet * W1_i
          commit_acc();
          wait_ready();                                            // dhh operand
          gemm(xhi, xlo, KS_H, H, idesc_h, ACC_H, true);           // dx = dhh * W0_i
          commit_acc();
        }
        wait_ready();                                              // dn_0 operand
        if (nb > 0) gemm(xhi, xlo, KS_H, C::CP, idesc_c, ACC_DLAT, nb == 1);             // dlat += dn_0 * Wc_1
        gemm(xhi, xlo, KS_H, C::CP, idesc_c, ACC_DLAT, nb == 0);                          // dlat += dn_0 * Wc_0
        commit_acc();
      }
    }
  } else {
    // ===================== workers =====================
    const int ww = warp - 2;
    const int quarter = warp & 3;
    const int part = ww >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int xb = part * XPT;
    const int lch0 = part * C::LCHUNKS / NPART, lch1 = (part + 1) * C::LCHUNKS / NPART;
    uint32_t done_phase = 0;

#ifdef NSDP_TRACE
    const bool tr_on = (tid == 64);
#define TW(id) do { if (tr_on) TR(id); } while (0)
#else
#define TW(id) do { } while (0)
#endif
    auto wait_acc = [&]() {
      TW(200);
      mbar_wait(acc_done, done_phase, err);
      done_phase ^= 1;
      tc_fence_after();
      TW(201);
    };
    auto publish = [&]() {
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);
      TW(202);
    };
    // v[0..32) of this thread's hidden columns -> X operand (+ optional staged copy)
    const float gsc = stg.f16 ? stage16::scale_from_max(*stg.gmax) : 1.f;   // gradient tiles are staged times gsc
    auto put_x = [&](const float (&v)[XPT], unsigned char *stage_tile, float sc = 1.f) {
#pragma unroll
      for (int j = 0; j < XPT; j += 8) {
        const float x[8] = {v[j], v[j + 1], v[j + 2], v[j + 3], v[j + 4], v[j + 5], v[j + 6], v[j + 7]};
        uint4 hi, lo;
        split8(x, hi, lo);
        const uint32_t off = canon_off(128, r, xb + j);
        *reinterpret_cast<uint4 *>(X_hi + off) = hi;
        *reinterpret_cast<uint4 *>(X_lo + off) = lo;
        if (stage_tile) {
          if (stg.f16) stage_write16<H>(stage_tile, r, xb + j, x, sc);
          else stage_write<H>(stage_tile, r, xb + j, hi, lo, stg.lo);
        }
      }
    };
    auto load_acc = [&](uint32_t col, float (&v)[XPT]) {
      float t[16];
      tmem_ld16(trow + col + xb, t);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = t[j];
      tmem_ld16(trow + col + xb + 16, t);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[16 + j] = t[j];
    };

    for (long long tile = tile_begin + blockIdx.x; tile < tile_end; tile += gridDim.x) {
      const long long grow = tile * 128 + r;
      const bool on = grow < a.R;
      const size_t tstride = stg.lo ? 512 : 256;   // staged bytes per tile and unit of width
      const size_t toff_h = (size_t)(tile - tile_begin) * tstride * H;
      const size_t toff_c = (size_t)(tile - tile_begin) * tstride * C::CP;
      uint32_t mx[MAXB + 1], my[MAXB];
      // ---- lat tile -> operand + staging; dout row -> smem + staging ------------------------------------------------------
      {
        const float *lrow = a.lat + (size_t)grow * Cin;
        for (int ch = lch0; ch < lch1; ++ch) {
          const int k0 = ch * 8;
          float4 u0 = make_float4(0.f, 0.f, 0.f, 0.f), u1 = u0;
          if (on && k0 < Cin) u0 = __ldg(reinterpret_cast<const float4 *>(lrow + k0));
          if (on && k0 + 4 < Cin) u1 = __ldg(reinterpret_cast<const float4 *>(lrow + k0 + 4));
          const float x[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
          uint4 hi, lo;
          split8(x, hi, lo);
          const uint32_t off = canon_off(128, r, k0);
          *reinterpret_cast<uint4 *>(L_hi + off) = hi;
          *reinterpret_cast<uint4 *>(L_lo + off) = lo;
          if (stg.f16) stage_write16<C::CP>(stg.lat + toff_c, r, k0, x, 1.f);
          else stage_write<C::CP>(stg.lat + toff_c, r, k0, hi, lo, stg.lo);
        }
        if (part == 0) {
          float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
          if (on) {
            d.x = dout[grow * O + 0];
            if (O > 1) d.y = dout[grow * O + 1];
            if (O > 2) d.z = dout[grow * O + 2];
            if (O > 3) d.w = dout[grow * O + 3];
          }
          dos[r] = d;
          const float x0[8] = {d.x, d.y, d.z, d.w, 0.f, 0.f, 0.f, 0.f}, x1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          uint4 hi, lo;
          unsigned char *dt = stg.dout + (size_t)(tile - tile_begin) * tstride * 16;
          if (stg.f16) {
            stage_write16<16>(dt, r, 0, x0, gsc);
            stage_write16<16>(dt, r, 8, x1, gsc);
          } else {
            split8(x0, hi, lo);
            stage_write<16>(dt, r, 0, hi, lo, stg.lo);
            split8(x1, hi, lo);
            stage_write<16>(dt, r, 8, hi, lo, stg.lo);
          }
        }
      }
      publish();
      // ---- forward recompute: x_i, y_i operands + staging + masks ---------------------------------------------------------------
      float v[XPT];
      for (int i = 0; i < nb; ++i) {
        wait_acc();
        load_acc(0, v);
        uint32_t m = 0;
#pragma unroll
        for (int j = 0; j < XPT; ++j) {
          v[j] = fmaxf(v[j] + bsum[i * H + xb + j], 0.f);
          if (v[j] > 0.f) m |= 1u << j;
        }
        mx[i] = m;
        put_x(v, stg.x[i] + toff_h);
        publish();
        wait_acc();
        load_acc(ACC_H, v);
        m = 0;
#pragma unroll
        for (int j = 0; j < XPT; ++j) {
          v[j] = fmaxf(v[j] + b0s[i * H + xb + j], 0.f);
          if (v[j] > 0.f) m |= 1u << j;
        }
        my[i] = m;
        put_x(v, stg.y[i] + toff_h);
        publish();
      }
      // ---- x_n = relu(net_n) (staged for d_wo) ; dnet = (dout * Wo^T) * [x_n > 0] --------------------------------------------------
      wait_acc();
      load_acc(0, v);
      float dnet[XPT];
      {
        const float4 d = dos[r];
        uint32_t m = 0;
#pragma unroll
        for (int j = 0; j < XPT; ++j) {
          v[j] = fmaxf(v[j] + bsum[nb * H + xb + j], 0.f);
          if (v[j] > 0.f) m |= 1u << j;
          const float4 w = wos[xb + j];
          const float s = fmaf(d.x, w.x, fmaf(d.y, w.y, fmaf(d.z, w.z, d.w * w.w)));
          dnet[j] = v[j] > 0.f ? s : 0.f;
        }
        mx[nb] = m;
        // x_n goes to staging only (it is no MMA operand)
#pragma unroll
        for (int j = 0; j < XPT; j += 8) {
          const float x[8] = {v[j], v[j + 1], v[j + 2], v[j + 3], v[j + 4], v[j + 5], v[j + 6], v[j + 7]};
          if (stg.f16) {
            stage_write16<H>(stg.x[nb] + toff_h, r, xb + j, x, 1.f);
          } else {
            uint4 hi, lo;
            split8(x, hi, lo);
            stage_write<H>(stg.x[nb] + toff_h, r, xb + j, hi, lo, stg.lo);
          }
        }
      }
      put_x(dnet, stg.dn[nb] + toff_h, gsc);
      publish();
      // ---- backward through the blocks ---------------------------------------------------------------------------------------------
      for (int i = nb - 1; i >= 0; --i) {
        wait_acc();                       // dy
        load_acc(ACC_H, v);
#pragma unroll
        for (int j = 0; j < XPT; ++j) v[j] = ((my[i] >> j) & 1u) ? v[j] : 0.f;
        put_x(v, stg.dhh[i] + toff_h, gsc);
        publish();
        wait_acc();                       // dx
        load_acc(ACC_H, v);
#pragma unroll
        for (int j = 0; j < XPT; ++j) dnet[j] += ((mx[i] >> j) & 1u) ? v[j] : 0.f;
        put_x(dnet, stg.dn[i] + toff_h, gsc);
        publish();
      }
      // ---- d_lat ----------------------------------------------------------------------------------------------------------------------
      wait_acc();
      {
        // tcgen05.ld is warp-collective (.sync.aligned): every lane issues it, only the stores are predicated
        float *drow = d_lat + (size_t)grow * Cin;
        for (int ch = lch0; ch < lch1; ++ch) {
          const int k0 = ch * 8;
          float t[8];
          tmem_ld8(trow + ACC_DLAT + k0, t);
          if (on && k0 < Cin) *reinterpret_cast<float4 *>(drow + k0) = make_float4(t[0], t[1], t[2], t[3]);
          if (on && k0 + 4 < Cin) *reinterpret_cast<float4 *>(drow + k0 + 4) = make_float4(t[4], t[5], t[6], t[7]);
        }
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

static long long tail_segment_tiles() {
  static const long long v = [] {
    const char *e = getenv("NSDP_TAIL_SEG");
    const long long t = e ? atoll(e) : 1184;   // 8 x 148 SMs: whole waves; 1.85 GB of staging per segment
    return t < 1 || t > 4736 ? 1184ll : t;
  }();
  return v;
}
#define kSegmentTiles tail_segment_tiles()
//   // 4 x 148 SMs: whole waves. staging: (CP + 22*128 + 16) * 512 B per tile ~ 1.5 MB -> 0.8 GB per segment

template <class C>
static size_t staged_bytes_per_tile(int nb) {
  return (size_t)512 * (C::CP + (size_t)(nb + 1 + nb + nb + nb + 1) * H + 16);
}

template <class C>
static size_t workspace_bytes(const nsdp_tail_args &a) {
  const long long tiles = ceil_div((long long)a.R, 128ll);
  const long long seg = tiles < kSegmentTiles ? tiles : kSegmentTiles;
  return packed_bytes<C>(a.n_blocks) + 256 + (size_t)seg * staged_bytes_per_tile<C>(a.n_blocks);
}

template <class C>
static int launch(const nsdp_tail_args &a, const float *dout, const nsdp_tail_grads &g, void *workspace, size_t ws_bytes,
                  cudaStream_t st) {
  const int nb = a.n_blocks;
  if (!workspace || ws_bytes < workspace_bytes<C>(a)) return NSDP_ERR_WORKSPACE;
  unsigned char *packed = (unsigned char *)workspace;
  int *err = (int *)(packed + packed_bytes<C>(nb));
  unsigned char *sbase = packed + packed_bytes<C>(nb) + 256;
  const long long tiles = ceil_div((long long)a.R, 128ll);
  const long long seg = tiles < kSegmentTiles ? tiles : kSegmentTiles;
  // staging precision of the operand tiles of the weight-gradient reductions: fp32-grade (hi + lo) unless
  // NSDP_STAGE_LO=0 opts into plain bf16 tiles (see vattn_bwd_tc.cu: stage_lo_for)
  static const int forced_lo = [] { const char *e = getenv("NSDP_STAGE_LO"); return e ? atoi(e) : 1; }();
  const int f16 = stage16::enabled() ? 1 : 0;
  const int lo = forced_lo != 0 && !f16;
  unsigned *gmax = (unsigned *)err + 16;   // inside the 256-byte header
  Staging stg;
  stg.lo = lo;
  stg.f16 = f16;
  stg.gmax = gmax;
  {
    unsigned char *p = sbase;
    auto take = [&](int width) { unsigned char *q = p; p += (size_t)seg * (lo ? 512 : 256) * width; return q; };
    stg.lat = take(C::CP);
    for (int i = 0; i <= nb; ++i) stg.x[i] = take(H);
    for (int i = 0; i < nb; ++i) stg.y[i] = take(H);
    for (int i = 0; i < nb; ++i) stg.dhh[i] = take(H);
    for (int i = 0; i <= nb; ++i) stg.dn[i] = take(H);
    stg.dout = take(16);
  }
  cudaError_t e = cudaMemsetAsync(err, 0, 256, st);
  if (e != cudaSuccess) return cuda_rc(e);
  if (f16) {
    const int r0 = stage16::launch_absmax(dout, (size_t)a.R * a.O, gmax, st);
    if (r0 != NSDP_OK) return r0;
  }
  pack_tail_bwd_weights_kernel<C><<<num_pack_matrices<C>(nb), 256, 0, st>>>(a, packed);
  int rc = check_launch();
  if (rc != NSDP_OK) return rc;
  auto kern = resnet_tail_bwd_tc_kernel<C>;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
  if (e != cudaSuccess) return cuda_rc(e);
  const int wld = (1 + nb) * H;
  for (long long t0 = 0; t0 < tiles; t0 += seg) {
    const long long t1 = t0 + seg < tiles ? t0 + seg : tiles;
    const long long n = t1 - t0;
    const int grid = (int)(n < num_sms() ? n : num_sms());
    unsigned long long *trace = nullptr;
#ifdef NSDP_TRACE
    if (const char *tp = getenv("NSDP_TRACE_PTR")) trace = (unsigned long long *)strtoull(tp, nullptr, 0);
#endif
    kern<<<grid, THREADS, C::SMEM, st>>>(a, dout, g.d_lat, packed, stg, t0, t1, err, trace);
    rc = check_launch();
    if (rc != NSDP_OK) return rc;
    dwtc::Job jobs[24];
    int nj = 0;
    // fc_out: d_wo_t[k][o] = x_n^T dout ; d_bo = colsum(dout)
    jobs[nj++] = {stg.x[nb], stg.dout, g.d_wo_t, H, 16, H, a.O, a.O, g.d_bo, nullptr, 0, 0, 0};
    for (int i = 0; i < nb; ++i) {
      // fc_1[i]: d_w1_t[k][c] = y_i^T d net_{i+1} ; d_b1 = colsum
      jobs[nj++] = {stg.y[i], stg.dn[i + 1], g.d_w1_t + (size_t)i * H * H, H, H, H, H, H, g.d_b1 + (size_t)i * H, nullptr, 0, 0, 0};
      // fc_0[i]: d_w0_t[k][c] = x_i^T dhh_i ; d_b0 = colsum
      jobs[nj++] = {stg.x[i], stg.dhh[i], g.d_w0_t + (size_t)i * H * H, H, H, H, H, H, g.d_b0 + (size_t)i * H, nullptr, 0, 0, 0};
      // fc_c[i]: d_wc_t[kc][(i+1)H + c] = lat^T dn_i ; d_bc slice = colsum
      jobs[nj++] = {stg.lat, stg.dn[i], g.d_wc_t + (size_t)(i + 1) * H, C::CP, H, a.C, H, wld, g.d_bc + (size_t)(i + 1) * H, nullptr, 0, 0, 0};
    }
    // init_enc: d pre_0 = d net_0 = dn_0 as well
    jobs[nj++] = {stg.lat, stg.dn[0], g.d_wc_t, C::CP, H, a.C, H, wld, g.d_bc, nullptr, 0, 0, 0};
    for (int j = 0; j < nj; ++j) {
      jobs[j].x_lo = jobs[j].y_lo = lo;
      jobs[j].f16 = f16;
      jobs[j].gmax = f16 ? gmax : nullptr;
    }
    // chunk-aligned launch: every job walks the same tile chunks at the same time, so `lat` (six readers), dn_i (two) and
    // x_i / y_i come from HBM once and from L2 for the other readers
    const long long bounds[2] = {0, n};
    rc = dw_tc_launch_chunked(jobs, nj, n, bounds, 1, err, st);
    if (rc == NSDP_ERR_UNSUPPORTED) rc = dw_tc_launch(jobs, nj, n, err, st);
    if (rc != NSDP_OK) return rc;
  }
  return NSDP_OK;
}

static int pick(const nsdp_tail_args &a) {
  if (a.H != H || a.O > 4 || a.C % 4 != 0 || a.n_blocks > MAXB || a.n_blocks < 1) return 0;
  if (a.C <= 128) return 128;
  if (a.C <= 208) return 208;
  return 0;
}

}  // namespace tbtc

size_t tail_bwd_tc_workspace_bytes(const nsdp_tail_args *a) {
  switch (tbtc::pick(*a)) {
    case 128: return tbtc::workspace_bytes<tbtc::Cfg<128>>(*a);
    case 208: return tbtc::workspace_bytes<tbtc::Cfg<208>>(*a);
    default: return 0;
  }
}

int tail_bwd_tc_dispatch(const nsdp_tail_args *a, const float *dout, const nsdp_tail_grads *g, void *workspace,
                         size_t ws_bytes, cudaStream_t st, bool *handled) {
  *handled = true;
  switch (tbtc::pick(*a)) {
    case 128: return tbtc::launch<tbtc::Cfg<128>>(*a, dout, *g, workspace, ws_bytes, st);
    case 208: return tbtc::launch<tbtc::Cfg<208>>(*a, dout, *g, workspace, ws_bytes, st);
    default: *handled = false; return NSDP_OK;
  }
}

}  // namespace nsdp

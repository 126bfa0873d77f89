// Fused decoder tail (ResNet-FC) forward on tcgen05 tensor cores, sm_100a.
//
// Same math and arguments as resnet_tail.cu (nsdp_tail_args). A persistent CTA per SM walks over tiles of 128 query
// rows. The `lat` tile is converted once to a bf16 hi/lo A operand in shared memory and feeds init_enc and all fc_c
// layers; `net` never leaves TMEM (fp32 accumulator acc0: every fc_c / fc_1 product is accumulated IN PLACE by the
// tensor core), the hidden activation goes through a second accumulator acc1; only relu(net) / relu(h) pass through
// registers to become the next A operand. Weight slabs (pre-packed bf16 hi/lo, K-major canonical layout) stream
// from L2 through a 5-stage bulk-copy ring in exactly the order the MMA warp consumes them:
//     init, fc_c[0], { fc_0[i], fc_c[i+1], fc_1[i] } for each block i.
// fc_c[i+1] only needs `lat`, so it is issued right after fc_0[i] and runs on the tensor core while the workers are
// busy turning acc1 into relu(h).
// The relu(net) / relu(h) operand is handed over in four 32-column chunks (x_ready[c]): the k-steps of the next GEMM
// that only need chunk c are issued while the workers are still converting chunks c+1.. (the scheme of
// fused_mlp_tc.cu; the source accumulator and the destination accumulator of such a pair are always different).
#include "common.cuh"
#include "umma.cuh"

namespace nsdp {
namespace ttc {

using namespace umma;

constexpr int H = 128;
constexpr int KS_H = H / 16;
constexpr int SLAB = H * 16 * 2;          // [128 x 16] bf16 slab
constexpr int STAGE_BYTES = 2 * SLAB;     // hi + lo
constexpr int STAGES = 5;
constexpr int WORKER_WARPS = 8;
constexpr int THREADS = (2 + WORKER_WARPS) * 32;
constexpr int MAX_BLOCKS = 8;
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t ACC1_COL = 128;
constexpr int XCH = 4;                    // chunks of the X operand handoff
constexpr int XCW = H / XCH;              // 32 columns = 2 k-steps per chunk
constexpr int XKPC = XCW / 16;

template <int CP_>
struct Cfg {
  static constexpr int CP = CP_;
  static constexpr int KS_C = CP / 16;
  static constexpr int A_LAT_HALF = 128 * CP * 2;
  static constexpr int A_X_HALF = 128 * H * 2;
  static constexpr int OFF_LAT = 0;
  static constexpr int OFF_X = OFF_LAT + 2 * A_LAT_HALF;
  static constexpr int OFF_STAGE = OFF_X + 2 * A_X_HALF;
  static constexpr int OFF_BSUM = OFF_STAGE + STAGES * STAGE_BYTES;       // float[(1+MAX_BLOCKS)][H]
  static constexpr int OFF_B0 = OFF_BSUM + (1 + MAX_BLOCKS) * H * 4;      // float[MAX_BLOCKS][H]
  static constexpr int OFF_WO = OFF_B0 + MAX_BLOCKS * H * 4;              // float[H][4]
  static constexpr int OFF_PART = OFF_WO + H * 16;                        // float[128][4]
  static constexpr int OFF_BAR = OFF_PART + 128 * 16;
  static constexpr int SMEM = OFF_BAR + 256;
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

// number of weight stages per tile and bytes of the packed image
template <class C>
__host__ __device__ constexpr int stages_per_tile(int nb) {
  return (1 + nb) * C::KS_C + 2 * nb * KS_H;
}
template <class C>
constexpr size_t packed_bytes(int nb) {
  return (size_t)stages_per_tile<C>(nb) * STAGE_BYTES;
}

// Packs one [H x K] operand (B[n][k] = Wt[k*ldw + n], Wt K-major / transposed like the fp32 kernel's arguments) into
// consecutive k-step stages [hi slab][lo slab] starting at `out`.
template <class C>
__device__ __forceinline__ void pack_matrix(const float *__restrict__ wt, int ldw, int kdim, int kpad,
                                            unsigned char *__restrict__ out, int tid, int nthreads) {
  const int total = H * (kpad / 2);
  for (int e = tid; e < total; e += nthreads) {
    const int n = e / (kpad / 2), k = (e - n * (kpad / 2)) * 2;
    const float x0 = k < kdim ? wt[(size_t)k * ldw + n] : 0.f;
    const float x1 = k + 1 < kdim ? wt[(size_t)(k + 1) * ldw + n] : 0.f;
    uint32_t hi, lo;
    split2(x0, x1, hi, lo);
    const size_t base = (size_t)(k >> 4) * STAGE_BYTES + canon_off(H, n, k & 15);
    *reinterpret_cast<uint32_t *>(out + base) = hi;
    *reinterpret_cast<uint32_t *>(out + base + SLAB) = lo;
  }
}

// grid.x = number of matrices in consumption order: init, c0, then per block (w0_i, c_{i+1} if any, w1_i)
template <class C>
__global__ void pack_tail_weights_kernel(const nsdp_tail_args a, unsigned char *__restrict__ out) {
  const int nb = a.n_blocks;
  const int wld = (1 + nb) * H;
  int m = blockIdx.x;
  size_t off = 0;
  // walk the consumption order until matrix m
  int idx = 0;
  auto stage_off = [&](int ksteps) { size_t o = off; off += (size_t)ksteps * STAGE_BYTES; return o; };
  for (int j = 0; j < 2 && j <= nb; ++j) {  // init and fc_c[0]
    const size_t o = stage_off(C::KS_C);
    if (idx++ == m) { pack_matrix<C>(a.wc_t + (size_t)j * H, wld, a.C, C::CP, out + o, threadIdx.x, blockDim.x); return; }
    if (nb == 0) break;
  }
  for (int i = 0; i < nb; ++i) {
    size_t o = stage_off(KS_H);
    if (idx++ == m) { pack_matrix<C>(a.w0_t + (size_t)i * H * H, H, H, H, out + o, threadIdx.x, blockDim.x); return; }
    if (i + 1 < nb) {
      o = stage_off(C::KS_C);
      if (idx++ == m) { pack_matrix<C>(a.wc_t + (size_t)(i + 2) * H, wld, a.C, C::CP, out + o, threadIdx.x, blockDim.x); return; }
    }
    o = stage_off(KS_H);
    if (idx++ == m) { pack_matrix<C>(a.w1_t + (size_t)i * H * H, H, H, H, out + o, threadIdx.x, blockDim.x); return; }
  }
}

template <class C>
__global__ void __launch_bounds__(THREADS, 1)
resnet_tail_tc_kernel(const nsdp_tail_args a, const unsigned char *__restrict__ packed, float *__restrict__ out,
                      long long tiles, int *err) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char *L_hi = smem + C::OFF_LAT, *L_lo = L_hi + C::A_LAT_HALF;
  unsigned char *X_hi = smem + C::OFF_X, *X_lo = X_hi + C::A_X_HALF;
  unsigned char *stage0 = smem + C::OFF_STAGE;
  float *bsum = reinterpret_cast<float *>(smem + C::OFF_BSUM);
  float *b0s = reinterpret_cast<float *>(smem + C::OFF_B0);
  float4 *wos = reinterpret_cast<float4 *>(smem + C::OFF_WO);
  float4 *part = reinterpret_cast<float4 *>(smem + C::OFF_PART);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::OFF_BAR);
  uint64_t *full = bars, *empty = bars + STAGES, *a_ready = bars + 2 * STAGES, *acc_done = a_ready + 1;
  uint64_t *x_ready = acc_done + 1;   // [XCH]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(x_ready + XCH);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nb = a.n_blocks, Cin = a.C, O = a.O;

  // per-channel bias prefix sums: bsum[i] = b_init + sum_{j<=i} bc[j] ... see file header of resnet_tail.cu
  for (int c = tid; c < H; c += THREADS) {
    float s = a.bc[c];
    for (int i = 0; i <= nb; ++i) {
      if (i < nb) s += a.bc[(size_t)(i + 1) * H + c];
      bsum[i * H + c] = s;              // i < nb: input of block i ; i == nb: final net (no further fc_c)
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
    for (int c = 0; c < XCH; ++c) mbar_init(&x_ready[c], WORKER_WARPS);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int per_tile = stages_per_tile<C>(nb);

  if (warp == 0) {
    // weight producer: PL lanes share the bulk copies (lane l serves stages l, l + PL, ...): one thread sustains only
    // about one copy per ~500 cycles, the tensor pipe drains an 8 KB stage in ~200
    constexpr int PL = 2;
    if (lane < PL) {
      const long long my_tiles = (long long)blockIdx.x < tiles ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
      const long long total = my_tiles * per_tile;
      for (long long it = lane; it < total; it += PL) {
        const int st = (int)(it % per_tile);
        const int s = (int)(it % STAGES);
        const uint32_t ph = (uint32_t)(it / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1, err);
        mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
        bulk_g2s(stage0 + (size_t)s * STAGE_BYTES, packed + (size_t)st * STAGE_BYTES, STAGE_BYTES, &full[s]);
      }
    }
  } else if (warp == 1) {
    // MMA issuer: whole warp runs loops and waits, one elected lane issues (elect_one: straight UTCHMMA issue);
    // descriptors advance by adds, the ring position is a running counter
    {
      const uint32_t idesc = idesc_bf16(128, H);
      constexpr uint32_t lbo_a = 128 * 16, lbo_b = H * 16;
      constexpr uint64_t A_STEP = (2 * lbo_a) >> 4;
      const uint64_t lhi = smem_desc(smem_u32(L_hi), lbo_a, 128), llo = smem_desc(smem_u32(L_lo), lbo_a, 128);
      const uint64_t xhi = smem_desc(smem_u32(X_hi), lbo_a, 128), xlo = smem_desc(smem_u32(X_lo), lbo_a, 128);
      const uint64_t bh0 = smem_desc(smem_u32(stage0), lbo_b, 128);
      uint32_t slot = 0, slot_phase = 0, ready_phase = 0, x_phase = 0;
      // one GEMM: A (hi/lo descriptors, k-steps [ks0, ks1)) x the next weight stages -> tmem column `col`
      auto gemm_part = [&](uint64_t a_hi, uint64_t a_lo, int ks0, int ks1, uint32_t col, bool fresh) {
        for (int ks = ks0; ks < ks1; ++ks) {
          mbar_wait(&full[slot], slot_phase, err);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t ah = a_hi + ks * A_STEP, al = a_lo + ks * A_STEP;
            const uint64_t bh = bh0 + (uint64_t)slot * (STAGE_BYTES >> 4);
            mma_bf16(tmem_base + col, ah, bh, idesc, !(fresh && ks == 0));
            mma_bf16(tmem_base + col, al, bh, idesc, true);
            mma_bf16(tmem_base + col, ah, bh + (SLAB >> 4), idesc, true);
            mma_commit(&empty[slot]);
          }
          if (++slot == STAGES) { slot = 0; slot_phase ^= 1; }
        }
      };
      auto gemm = [&](uint64_t a_hi, uint64_t a_lo, int ksteps, uint32_t col, bool fresh) {
        gemm_part(a_hi, a_lo, 0, ksteps, col, fresh);
      };
      // X-operand GEMM, issued chunk by chunk as the workers hand the operand columns over
      auto gemm_x = [&](uint32_t col, bool fresh) {
        for (int c = 0; c < XCH; ++c) {
          mbar_wait(&x_ready[c], x_phase, err);
          tc_fence_after();
          gemm_part(xhi, xlo, c * XKPC, (c + 1) * XKPC, col, fresh);
        }
        x_phase ^= 1;
      };
      auto commit_acc = [&]() {
        if (elect_one()) mma_commit(acc_done);
      };
      auto wait_ready = [&]() {
        mbar_wait(a_ready, ready_phase, err);
        ready_phase ^= 1;
        tc_fence_after();
      };
      for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        wait_ready();                                   // lat operand written
        gemm(lhi, llo, C::KS_C, 0, true);               // init_enc
        if (nb > 0) gemm(lhi, llo, C::KS_C, 0, false);  // fc_c[0]
        commit_acc();
        for (int i = 0; i < nb; ++i) {
          gemm_x(ACC1_COL, true);                       // fc_0[i] -> acc1, chunks of x = relu(net + bsum_i) as they land
          commit_acc();
          if (i + 1 < nb) gemm(lhi, llo, C::KS_C, 0, false);  // fc_c[i+1] -> acc0 while the workers build y
          gemm_x(0, false);                             // fc_1[i] -> acc0, chunks of y = relu(h + b0_i)
          commit_acc();
        }
      }
    }
  } else {
    const int ww = warp - 2;
    const int quarter = warp & 3;
    const int half = ww >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t done_phase = 0;
    constexpr int LPT = C::CP / 2;   // lat columns per thread
    constexpr int XPT = H / 2;       // hidden columns per thread
    const int xb = half * XPT;

    // acc (TMEM column base `col`) + bias -> relu -> bf16 hi/lo X operand, chunk by chunk: this warp converts columns
    // [c*XCW + half*16, +16) of chunk c and signals x_ready[c]
    auto to_x = [&](uint32_t col, const float *bias) {
#pragma unroll 1
      for (int c = 0; c < XCH; ++c) {
        const int k0 = c * XCW + half * (XCW / 2);
        float v[16];
        tmem_ld16(trow + col + k0, v);
#pragma unroll
        for (int j = 0; j < 16; j += 8) {
          float x[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) x[u] = fmaxf(v[j + u] + bias[k0 + j + u], 0.f);
          uint4 hi, lo;
          split2(x[0], x[1], hi.x, lo.x);
          split2(x[2], x[3], hi.y, lo.y);
          split2(x[4], x[5], hi.z, lo.z);
          split2(x[6], x[7], hi.w, lo.w);
          const uint32_t off = canon_off(128, r, k0 + j);
          *reinterpret_cast<uint4 *>(X_hi + off) = hi;
          *reinterpret_cast<uint4 *>(X_lo + off) = lo;
        }
        tc_fence_before();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&x_ready[c]);
      }
    };
    auto wait_acc = [&]() {
      mbar_wait(acc_done, done_phase, err);
      done_phase ^= 1;
      tc_fence_after();
    };

    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const long long grow = tile * 128 + r;
      const bool on = grow < a.R;
      const float *lrow = a.lat + (size_t)grow * Cin;
      // ---- lat tile -> bf16 hi/lo A operand ----------------------------------------------------------------
#pragma unroll 2
      for (int k0 = half * LPT; k0 < half * LPT + LPT; k0 += 8) {
        float4 u0 = make_float4(0.f, 0.f, 0.f, 0.f), u1 = u0;
        if (on && k0 < Cin) u0 = __ldg(reinterpret_cast<const float4 *>(lrow + k0));
        if (on && k0 + 4 < Cin) u1 = __ldg(reinterpret_cast<const float4 *>(lrow + k0 + 4));
        uint4 hi, lo;
        split2(u0.x, u0.y, hi.x, lo.x);
        split2(u0.z, u0.w, hi.y, lo.y);
        split2(u1.x, u1.y, hi.z, lo.z);
        split2(u1.z, u1.w, hi.w, lo.w);
        const uint32_t off = canon_off(128, r, k0);
        *reinterpret_cast<uint4 *>(L_hi + off) = hi;
        *reinterpret_cast<uint4 *>(L_lo + off) = lo;
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);

      for (int i = 0; i < nb; ++i) {
        wait_acc();
        to_x(0, bsum + i * H);            // x = relu(net + bias prefix)
        wait_acc();
        to_x(ACC1_COL, b0s + i * H);      // y = relu(h + b0)
      }
      // ---- out = relu(net) * Wo + bo -----------------------------------------------------------------------------
      wait_acc();
      float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
      const float *bl = bsum + nb * H;
#pragma unroll 1
      for (int k0 = xb; k0 < xb + XPT; k0 += 16) {
        float v[16];
        tmem_ld16(trow + k0, v);
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const float x = fmaxf(v[u] + bl[k0 + u], 0.f);
          const float4 w = wos[k0 + u];
          o0 = fmaf(x, w.x, o0); o1 = fmaf(x, w.y, o1); o2 = fmaf(x, w.z, o2); o3 = fmaf(x, w.w, o3);
        }
      }
      tc_fence_before();
      if (half == 1) part[r] = make_float4(o0, o1, o2, o3);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (half == 0 && on) {
        const float4 p = part[r];
        const float res[4] = {o0 + p.x, o1 + p.y, o2 + p.z, o3 + p.w};
        for (int o = 0; o < O; ++o) out[grow * O + o] = res[o] + a.bo[o];
      }
    }
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <class C>
static int launch(const nsdp_tail_args &a, float *out, void *workspace, size_t ws_bytes, cudaStream_t st) {
  const size_t pb = packed_bytes<C>(a.n_blocks);
  if (!workspace || ws_bytes < pb + 16) return NSDP_ERR_WORKSPACE;
  unsigned char *packed = (unsigned char *)workspace;
  int *err = (int *)(packed + pb);
  cudaError_t e = cudaMemsetAsync(err, 0, sizeof(int), st);
  if (e != cudaSuccess) return cuda_rc(e);
  const int nmat = a.n_blocks == 0 ? 1 : (2 + 2 * a.n_blocks + (a.n_blocks - 1));
  pack_tail_weights_kernel<C><<<nmat, 256, 0, st>>>(a, packed);
  int rc = check_launch();
  if (rc != NSDP_OK) return rc;
  const long long tiles = ceil_div((long long)a.R, 128ll);
  auto kern = resnet_tail_tc_kernel<C>;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
  if (e != cudaSuccess) return cuda_rc(e);
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  kern<<<grid, THREADS, C::SMEM, st>>>(a, packed, out, tiles, err);
  return check_launch();
}

static int pick(const nsdp_tail_args &a) {
  if (a.H != H || a.O > 4 || a.C % 4 != 0 || a.n_blocks > MAX_BLOCKS) return 0;
  if (a.C <= 128) return 128;
  if (a.C <= 208) return 208;
  return 0;
}

}  // namespace ttc

size_t tail_tc_workspace_bytes(const nsdp_tail_args *a) {
  switch (ttc::pick(*a)) {
    case 128: return ttc::packed_bytes<ttc::Cfg<128>>(a->n_blocks) + 16;
    case 208: return ttc::packed_bytes<ttc::Cfg<208>>(a->n_blocks) + 16;
    default: return 0;
  }
}

int tail_tc_dispatch(const nsdp_tail_args *a, float *out, void *workspace, size_t ws_bytes, cudaStream_t st, bool *handled) {
  *handled = true;
  switch (ttc::pick(*a)) {
    case 128: return ttc::launch<ttc::Cfg<128>>(*a, out, workspace, ws_bytes, st);
    case 208: return ttc::launch<ttc::Cfg<208>>(*a, out, workspace, ws_bytes, st);
    default: *handled = false; return NSDP_OK;
  }
}

}  // namespace nsdp

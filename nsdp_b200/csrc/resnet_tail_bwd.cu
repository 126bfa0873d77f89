// Backward of the fused decoder tail (fp32 CUDA-core version).
//
// Reference behaviour being replaced: autograd over 17 Linear layers keeps every [B*Q, 128] activation in
// HBM (model/decoder/crosstransformer_decoder.py:63-69). Here a persistent CTA walks over tiles of 64 query
// rows; per tile it recomputes the forward, parking relu(net_i) / relu(h_i) in a small per-CTA scratch slot
// (11 x 64 x 128 fp32 = 352 KB per CTA, 52 MB for 148 CTAs: L2-resident, never a per-row HBM tensor), then
// runs the chain backwards: data gradients with the un-transposed weights, weight gradients reduced over
// the tile on chip and added to the global accumulators with one atomic per element per tile.
#include "common.cuh"

namespace nsdp {
namespace tailb {

constexpr int H = 128;
constexpr int R = 64;
constexpr int TX = 32, CN = 4, TY = 16, RM = 4;  // [64][128] tiles: 512 threads
constexpr int THREADS = TX * TY;
constexpr int LDX = H + 4;
constexpr int KM = 8;
// d_lat tile [64][C<=256]: 16 row groups x 4 rows, 32 column groups x 8 columns
constexpr int LTX = 32, LCN = 8, LTY = 16, LRM = 4;

__host__ __device__ inline int lat_ld(int C) { return C + 4; }
inline size_t smem_bytes(int C) { return sizeof(float) * ((size_t)R * lat_ld(C) + 3 * (size_t)R * LDX + 64 * 4 + 2 * H); }
constexpr size_t kSlotFloats = (size_t)11 * R * H;

// acc[i][c] += sum_k A[(r0+i)*lda + k] * W[k*ldw + c0 + c]
template <int RM_, int CN_>
__device__ __forceinline__ void gemm_nn(float (&acc)[RM_][CN_], const float *__restrict__ A, int lda, int r0,
                                        const float *__restrict__ W, int ldw, int kdim, int c0, int ncols) {
  for (int kk = 0; kk < kdim; kk += 4) {
    float4 av[RM_];
#pragma unroll
    for (int i = 0; i < RM_; ++i) av[i] = *reinterpret_cast<const float4 *>(A + (size_t)(r0 + i) * lda + kk);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float w[CN_];
#pragma unroll
      for (int c = 0; c < CN_; c += 4) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 + c < ncols) t = ldg4(W + (size_t)(kk + u) * ldw + c0 + c);
        w[c] = t.x; w[c + 1] = t.y; w[c + 2] = t.z; w[c + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < RM_; ++i) {
        const float x = u == 0 ? av[i].x : (u == 1 ? av[i].y : (u == 2 ? av[i].z : av[i].w));
#pragma unroll
        for (int c = 0; c < CN_; ++c) acc[i][c] = fmaf(x, w[c], acc[i][c]);
      }
    }
  }
}

// out[k*ldo + c] += sum_r A[r*lda + k] * Bm[r*LDX + c], k < kdim, c < H
__device__ __forceinline__ void gemm_tn_atomic(const float *__restrict__ A, int lda, int kdim,
                                               const float *__restrict__ Bm, float *__restrict__ out, int ldo, int tx,
                                               int ty) {
  const int c0 = tx * CN;
  for (int kbase = 0; kbase < kdim; kbase += TY * KM) {
    const int k0 = kbase + ty * KM;
    if (k0 >= kdim) continue;
    float acc[KM][CN];
#pragma unroll
    for (int j = 0; j < KM; ++j)
#pragma unroll
      for (int c = 0; c < CN; ++c) acc[j][c] = 0.f;
#pragma unroll 2
    for (int r = 0; r < R; ++r) {
      const float4 a0 = *reinterpret_cast<const float4 *>(A + (size_t)r * lda + k0);
      const float4 a1 = *reinterpret_cast<const float4 *>(A + (size_t)r * lda + k0 + 4);
      const float4 b = *reinterpret_cast<const float4 *>(Bm + (size_t)r * LDX + c0);
      const float av[KM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int j = 0; j < KM; ++j) {
        acc[j][0] = fmaf(av[j], b.x, acc[j][0]);
        acc[j][1] = fmaf(av[j], b.y, acc[j][1]);
        acc[j][2] = fmaf(av[j], b.z, acc[j][2]);
        acc[j][3] = fmaf(av[j], b.w, acc[j][3]);
      }
    }
#pragma unroll
    for (int j = 0; j < KM; ++j)
      if (k0 + j < kdim) {
#pragma unroll
        for (int c = 0; c < CN; ++c) atomicAdd(out + (size_t)(k0 + j) * ldo + c0 + c, acc[j][c]);
      }
  }
}

__device__ __forceinline__ void store_tile(const float (&v)[RM][CN], float *__restrict__ dst, int ld, int r0, int c0,
                                           bool relu) {
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    float4 t = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
    if (relu) {
      t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f);
    }
    *reinterpret_cast<float4 *>(dst + (size_t)(r0 + i) * ld + c0) = t;
  }
}

__device__ __forceinline__ void add_bias(float (&acc)[RM][CN], const float *__restrict__ bias, int c0) {
  const float4 b = ldg4(bias + c0);
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    acc[i][0] += b.x; acc[i][1] += b.y; acc[i][2] += b.z; acc[i][3] += b.w;
  }
}

// column sums of a [R][H] shared tile -> atomically added to dst[0..H)
__device__ __forceinline__ void colsum_atomic(const float *__restrict__ buf, float *__restrict__ dst, int tid) {
  if (tid < H) {
    float s = 0.f;
    for (int r = 0; r < R; ++r) s += buf[(size_t)r * LDX + tid];
    atomicAdd(dst + tid, s);
  }
}

struct Weights {  // un-transposed (out, in) matrices for the data-gradient GEMMs
  const float *wc;  // ((1+n)*H, C)
  const float *w0;  // (n, H, H)
  const float *w1;  // (n, H, H)
};

__global__ void __launch_bounds__(THREADS, 1)
resnet_tail_bwd_kernel(const nsdp_tail_args a, const Weights wu, const float *__restrict__ dout, const nsdp_tail_grads g,
                       float *__restrict__ scratch, long long tiles) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C = a.C, O = a.O, nb = a.n_blocks;
  const int ldl = lat_ld(C);
  float *latS = reinterpret_cast<float *>(smem_raw);  // [R][ldl]
  float *bufA = latS + (size_t)R * ldl;               // [R][LDX]
  float *bufB = bufA + (size_t)R * LDX;
  float *bufC = bufB + (size_t)R * LDX;
  float *doS = bufC + (size_t)R * LDX;                // [R][4]
  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int c0 = tx * CN, r0 = ty * RM;
  const int ltx = tid % LTX, lty = tid / LTX;
  const int lc0 = ltx * LCN, lr0 = lty * LRM;
  float *slot = scratch + (size_t)blockIdx.x * kSlotFloats;  // [11][R][H]
  const int wld = (1 + nb) * H;

  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long row0 = tile * R;
    const int nrows = (int)min((long long)R, (long long)a.R - row0);
    __syncthreads();  // previous tile fully consumed
    const int c4 = C / 4;
    for (int t = tid; t < R * c4; t += THREADS) {
      const int r = t / c4, q = t - r * c4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nrows) v = ldg4(a.lat + (size_t)(row0 + r) * C + q * 4);
      *reinterpret_cast<float4 *>(latS + (size_t)r * ldl + q * 4) = v;
    }
    for (int t = tid; t < R * 4; t += THREADS) {
      const int r = t >> 2, o = t & 3;
      doS[t] = (r < nrows && o < O) ? dout[(row0 + r) * O + o] : 0.f;
    }
    __syncthreads();

    // ================= forward recompute: park relu(net_i), relu(h_i) in the scratch slot ===================
    float net[RM][CN];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
      for (int c = 0; c < CN; ++c) net[i][c] = 0.f;
    gemm_nn<RM, CN>(net, latS, ldl, r0, a.wc_t, wld, C, c0, H);
    add_bias(net, a.bc, c0);
    for (int blk = 0; blk < nb; ++blk) {
      gemm_nn<RM, CN>(net, latS, ldl, r0, a.wc_t + (size_t)(blk + 1) * H, wld, C, c0, H);
      add_bias(net, a.bc + (size_t)(blk + 1) * H, c0);
      __syncthreads();
      store_tile(net, bufA, LDX, r0, c0, true);
      store_tile(net, slot + (size_t)(2 * blk) * R * H, H, r0, c0, true);
      __syncthreads();
      float hacc[RM][CN];
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int c = 0; c < CN; ++c) hacc[i][c] = 0.f;
      gemm_nn<RM, CN>(hacc, bufA, LDX, r0, a.w0_t + (size_t)blk * H * H, H, H, c0, H);
      add_bias(hacc, a.b0 + (size_t)blk * H, c0);
      __syncthreads();
      store_tile(hacc, bufA, LDX, r0, c0, true);
      store_tile(hacc, slot + (size_t)(2 * blk + 1) * R * H, H, r0, c0, true);
      __syncthreads();
      gemm_nn<RM, CN>(net, bufA, LDX, r0, a.w1_t + (size_t)blk * H * H, H, H, c0, H);
      add_bias(net, a.b1 + (size_t)blk * H, c0);
    }
    __syncthreads();
    // x_last = relu(net_final) -> bufA (also the mask of d_net)
    store_tile(net, bufA, LDX, r0, c0, true);
    __syncthreads();

    // ================= backward ==========================================================================
    // fc_out: d_wo[k][o] += sum_r x[r][k]*dout[r][o]; d_bo[o] += sum_r dout[r][o]; dnet = (dout*Wo^T) * [x>0]
    if (tid < H) {
      float s[4] = {0.f, 0.f, 0.f, 0.f};
      for (int r = 0; r < R; ++r) {
        const float x = bufA[(size_t)r * LDX + tid];
        const float4 d = *reinterpret_cast<const float4 *>(doS + r * 4);
        s[0] = fmaf(x, d.x, s[0]); s[1] = fmaf(x, d.y, s[1]); s[2] = fmaf(x, d.z, s[2]); s[3] = fmaf(x, d.w, s[3]);
      }
      for (int o = 0; o < O; ++o) atomicAdd(g.d_wo_t + (size_t)tid * O + o, s[o]);
    } else if (tid < H + O) {
      const int o = tid - H;
      float s = 0.f;
      for (int r = 0; r < R; ++r) s += doS[r * 4 + o];
      atomicAdd(g.d_bo + o, s);
    }
    float dnet[RM][CN];
#pragma unroll
    for (int i = 0; i < RM; ++i) {
      const float4 d = *reinterpret_cast<const float4 *>(doS + (r0 + i) * 4);
      const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int c = 0; c < CN; ++c) {
        float s = 0.f;
        for (int o = 0; o < O; ++o) s = fmaf(dv[o], __ldg(a.wo_t + (size_t)(c0 + c) * O + o), s);
        dnet[i][c] = bufA[(size_t)(r0 + i) * LDX + c0 + c] > 0.f ? s : 0.f;
      }
    }
    float dlat[LRM][LCN];
#pragma unroll
    for (int i = 0; i < LRM; ++i)
#pragma unroll
      for (int c = 0; c < LCN; ++c) dlat[i][c] = 0.f;

    for (int blk = nb - 1; blk >= 0; --blk) {
      __syncthreads();
      // bufB = dnet ; bufA = y = relu(h_blk)
      store_tile(dnet, bufB, LDX, r0, c0, false);
      for (int t = tid; t < R * (H / 4); t += THREADS) {
        const int r = t / (H / 4), q = t - r * (H / 4);
        *reinterpret_cast<float4 *>(bufA + (size_t)r * LDX + q * 4) =
            *reinterpret_cast<const float4 *>(slot + ((size_t)(2 * blk + 1) * R + r) * H + q * 4);
      }
      __syncthreads();
      // fc_1: d_b1, d_w1t[k][c] += y^T dnet ; dy = dnet*W1 ; dhh = dy*[y>0]
      colsum_atomic(bufB, g.d_b1 + (size_t)blk * H, tid);
      gemm_tn_atomic(bufA, LDX, H, bufB, g.d_w1_t + (size_t)blk * H * H, H, tx, ty);
      float dh[RM][CN];
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int c = 0; c < CN; ++c) dh[i][c] = 0.f;
      gemm_nn<RM, CN>(dh, bufB, LDX, r0, wu.w1 + (size_t)blk * H * H, H, H, c0, H);
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int c = 0; c < CN; ++c)
          if (!(bufA[(size_t)(r0 + i) * LDX + c0 + c] > 0.f)) dh[i][c] = 0.f;
      __syncthreads();
      // bufC = dhh ; bufA = x = relu(n_blk)
      store_tile(dh, bufC, LDX, r0, c0, false);
      for (int t = tid; t < R * (H / 4); t += THREADS) {
        const int r = t / (H / 4), q = t - r * (H / 4);
        *reinterpret_cast<float4 *>(bufA + (size_t)r * LDX + q * 4) =
            *reinterpret_cast<const float4 *>(slot + ((size_t)(2 * blk) * R + r) * H + q * 4);
      }
      __syncthreads();
      // fc_0: d_b0, d_w0t += x^T dhh ; dx = dhh*W0 ; dn = dnet + dx*[x>0]
      colsum_atomic(bufC, g.d_b0 + (size_t)blk * H, tid);
      gemm_tn_atomic(bufA, LDX, H, bufC, g.d_w0_t + (size_t)blk * H * H, H, tx, ty);
      float dx[RM][CN];
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int c = 0; c < CN; ++c) dx[i][c] = 0.f;
      gemm_nn<RM, CN>(dx, bufC, LDX, r0, wu.w0 + (size_t)blk * H * H, H, H, c0, H);
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int c = 0; c < CN; ++c)
          if (bufA[(size_t)(r0 + i) * LDX + c0 + c] > 0.f) dnet[i][c] += dx[i][c];
      __syncthreads();
      // fc_c[blk]: pre_{blk+1} gets dn. bufB = dn ; d_bc, d_wct[:, (blk+1)H..] += lat^T dn ; dlat += dn*Wc[(blk+1)H..]
      store_tile(dnet, bufB, LDX, r0, c0, false);
      __syncthreads();
      colsum_atomic(bufB, g.d_bc + (size_t)(blk + 1) * H, tid);
      gemm_tn_atomic(latS, ldl, C, bufB, g.d_wc_t + (size_t)(blk + 1) * H, wld, tx, ty);
      gemm_nn<LRM, LCN>(dlat, bufB, LDX, lr0, wu.wc + (size_t)(blk + 1) * H * C, C, H, lc0, C);
    }
    // init_enc
    __syncthreads();
    store_tile(dnet, bufB, LDX, r0, c0, false);
    __syncthreads();
    colsum_atomic(bufB, g.d_bc, tid);
    gemm_tn_atomic(latS, ldl, C, bufB, g.d_wc_t, wld, tx, ty);
    gemm_nn<LRM, LCN>(dlat, bufB, LDX, lr0, wu.wc, C, H, lc0, C);
#pragma unroll
    for (int i = 0; i < LRM; ++i) {
      const int r = lr0 + i;
      if (r < nrows) {
#pragma unroll
        for (int c = 0; c < LCN; c += 4)
          if (lc0 + c < C)
            *reinterpret_cast<float4 *>(g.d_lat + (size_t)(row0 + r) * C + lc0 + c) =
                make_float4(dlat[i][c], dlat[i][c + 1], dlat[i][c + 2], dlat[i][c + 3]);
      }
    }
  }
}

static int grid_size(long long tiles) {
  const long long sms = num_sms();
  return (int)(tiles < sms ? tiles : sms);
}

}  // namespace tailb
}  // namespace nsdp

using namespace nsdp;

// workspace layout: [wc (1+n)H*C][w0 n*H*H][w1 n*H*H] un-transposed copies, then the per-CTA scratch slots
static size_t tail_bwd_weight_floats(const nsdp_tail_args *a) {
  return (size_t)(1 + a->n_blocks) * tailb::H * a->C + 2 * (size_t)a->n_blocks * tailb::H * tailb::H;
}

namespace nsdp {
size_t tail_bwd_tc_workspace_bytes(const nsdp_tail_args *a);
int tail_bwd_tc_dispatch(const nsdp_tail_args *a, const float *dout, const nsdp_tail_grads *g, void *workspace,
                         size_t ws_bytes, cudaStream_t st, bool *handled);
}

extern "C" size_t nsdp_resnet_tail_bwd_workspace_bytes(const nsdp_tail_args *a) {
  if (!a || a->R <= 0) return 0;
  if (a->impl != 1) {
    const size_t tc = nsdp::tail_bwd_tc_workspace_bytes(a);
    if (tc) return tc;
  }
  const long long tiles = ceil_div((long long)a->R, (long long)tailb::R);
  return sizeof(float) * (tail_bwd_weight_floats(a) + (size_t)tailb::grid_size(tiles) * tailb::kSlotFloats);
}

namespace nsdp {
// dst[c*rows + r] = src[r*cols + c]
__global__ void transpose_kernel(const float *__restrict__ src, float *__restrict__ dst, int rows, int cols, int src_ld) {
  const long long n = (long long)rows * cols;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t / rows), r = (int)(t - (long long)c * rows);
    dst[t] = src[(size_t)r * src_ld + c];
  }
}
}  // namespace nsdp

extern "C" int nsdp_resnet_tail_bwd_f32(const nsdp_tail_args *a, const float *d_out, const nsdp_tail_grads *g,
                                        void *workspace, size_t workspace_bytes, void *stream) {
  if (!a || !d_out || !g || !a->lat || !a->wc_t || !a->bc || !a->wo_t || !a->bo) return NSDP_ERR_INVALID_ARGUMENT;
  if (!g->d_lat || !g->d_wc_t || !g->d_bc || !g->d_wo_t || !g->d_bo) return NSDP_ERR_INVALID_ARGUMENT;
  if (a->n_blocks > 0 && (!a->w0_t || !a->b0 || !a->w1_t || !a->b1 || !g->d_w0_t || !g->d_b0 || !g->d_w1_t || !g->d_b1))
    return NSDP_ERR_INVALID_ARGUMENT;
  if (a->R <= 0 || a->C <= 0 || a->O <= 0 || a->n_blocks < 0) return NSDP_ERR_INVALID_ARGUMENT;
  if (a->H != tailb::H || a->C % 4 != 0 || a->C > 256 || a->O > 4) return NSDP_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < nsdp_resnet_tail_bwd_workspace_bytes(a)) return NSDP_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  if (a->impl != 1) {
    bool handled = false;
    int rc = tail_bwd_tc_dispatch(a, d_out, g, workspace, workspace_bytes, st, &handled);
    if (handled) return rc;
    if (a->impl == 2) return NSDP_ERR_UNSUPPORTED;
  }
  const int H = tailb::H, C = a->C, nb = a->n_blocks;
  float *ws = (float *)workspace;
  float *wc = ws;                                   // ((1+n)H, C)  = transpose of wc_t (C, (1+n)H)
  float *w0 = wc + (size_t)(1 + nb) * H * C;        // (n, H, H)    = per-block transpose of w0_t
  float *w1 = w0 + (size_t)nb * H * H;
  float *scratch = w1 + (size_t)nb * H * H;
  transpose_kernel<<<64, 256, 0, st>>>(a->wc_t, wc, C, (1 + nb) * H, (1 + nb) * H);
  for (int i = 0; i < nb; ++i) {
    transpose_kernel<<<16, 256, 0, st>>>(a->w0_t + (size_t)i * H * H, w0 + (size_t)i * H * H, H, H, H);
    transpose_kernel<<<16, 256, 0, st>>>(a->w1_t + (size_t)i * H * H, w1 + (size_t)i * H * H, H, H, H);
  }
  int rc = check_launch();
  if (rc != NSDP_OK) return rc;
  const size_t smem = tailb::smem_bytes(C);
  cudaError_t e = cudaFuncSetAttribute(tailb::resnet_tail_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_rc(e);
  const long long tiles = ceil_div((long long)a->R, (long long)tailb::R);
  tailb::Weights wu{wc, w0, w1};
  tailb::resnet_tail_bwd_kernel<<<tailb::grid_size(tiles), tailb::THREADS, smem, st>>>(*a, wu, d_out, *g, scratch, tiles);
  return check_launch();
}

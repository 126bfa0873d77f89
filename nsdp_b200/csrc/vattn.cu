// Fused vector ("point-transformer") attention over neighbourhoods — fp32 CUDA-core version.
//
// Replaces the pair-level torch-op chain of TransformerBlock (model/encoder/blocks.py:104-126),
// TransformerSetAbstraction (:290-308) and CrossTransformerBlock (model/decoder/blocks.py:62-91): in the
// reference every one of gather / Linear / ReLU / Linear / sub / add / Linear / ReLU / Linear / softmax /
// einsum is its own kernel with a [B, M, K, D] fp32 round trip through HBM (2.56 GB per tensor for the
// decoder at B=8, Q=50k). Here a CTA owns a tile of R = (points per tile) x (neighbours) pair rows and runs
// the whole chain on chip; the only HBM traffic is the per-point tables and the [B, M, D] result.
//
// Math (see include/nsdp_b200.h, nsdp_vattn_args): the caller folds the linear algebra that does not
// depend on the pair (W' = Wg0*Wd2, Q' = Wg0*Wq*x, K' = Wg0*Wk*x, biases) so the pair level needs three
// D x D products, two of which share the SAME left operand h = relu(Wd0*rel + b), a K=3 layer that is
// recomputed on the fly in registers instead of being stored:
//     g = relu(h*W' + P)   ->   a = g*Wg2   ->   w = softmax_rows(a)   ->   out = sum_rows w*(V + h*Wd2)
// One [R][D] shared-memory buffer is reused for g, a, w and w*(V+delta) in turn.
//
// This is the numerically-straight fp32 path (parity reference for the tensor-core kernels).
#include <math.h>

#include "common.cuh"

namespace nsdp {

template <int TX_, int CN_, int TY_, int RM_>
struct VCfg {
  static constexpr int TX = TX_, CN = CN_, TY = TY_, RM = RM_;
  static constexpr int DP = TX * CN;      // padded channel count handled by the thread grid
  static constexpr int R = TY * RM;       // pair rows per tile
  static constexpr int THREADS = TX * TY;
  static constexpr int LD = DP + 4;       // row pitch of the activation buffer (floats)
  static constexpr size_t smem_bytes() {
    return sizeof(float) * ((size_t)R * LD + (size_t)R * 4 + (size_t)DP * 4) + sizeof(int) * (size_t)R * 2;
  }
};

struct RowRef {
  int c;  // flattened centre index b*M+i, or -1 for an inactive row
  int n;  // flattened source index b*N+j, or -(b+1) for the global row
};

// rel4[r] = (rx, ry, rz, flag): flag 1 -> h = relu(wd0*rel + bd0); flag 0 -> h = 0 (global / inactive rows)
template <class C>
__device__ __forceinline__ void tile_rows_setup(const nsdp_vattn_args &a, long long tile, int krows, int tp,
                                                float4 *rel4, RowRef *rows) {
  const long long BM = (long long)a.B * a.M;
  for (int r = threadIdx.x; r < C::R; r += C::THREADS) {
    const int p = r / krows, t = r - p * krows;
    const long long ci = tile * tp + p;
    RowRef rr;
    rr.c = -1;
    rr.n = 0;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p < tp && ci < BM) {
      const int b = (int)(ci / a.M);
      rr.c = (int)ci;
      if (t < a.K) {
        const int j = a.idx ? a.idx[ci * a.K + t] : t;
        rr.n = b * a.N + j;
        const float *xc = a.xyz_c + ci * 3;
        const float *xn = a.xyz_n + (size_t)rr.n * 3;
        v.x = a.sign * (xc[0] - xn[0]);
        v.y = a.sign * (xc[1] - xn[1]);
        v.z = a.sign * (xc[2] - xn[2]);
        v.w = 1.f;
      } else {
        rr.n = -(b + 1);
      }
    }
    rel4[r] = v;
    rows[r] = rr;
  }
}

// acc[i][c] = sum_kk h(row_i, kk) * wt[kk][c0 + c], h recomputed from rel on the fly.
template <class C>
__device__ __forceinline__ void gemm_h(float (&acc)[C::RM][C::CN], const float4 (&rel)[C::RM],
                                       const float4 *__restrict__ wd0s, const float *__restrict__ wt, int D, int c0) {
#pragma unroll
  for (int i = 0; i < C::RM; ++i)
#pragma unroll
    for (int c = 0; c < C::CN; ++c) acc[i][c] = 0.f;
  const bool col_ok = c0 < D;  // D % 4 == 0 and CN % 4 == 0: a thread's float4 groups are all-in or all-out
#pragma unroll 2
  for (int kk = 0; kk < D; ++kk) {
    const float4 w0 = wd0s[kk];
    float h[C::RM];
#pragma unroll
    for (int i = 0; i < C::RM; ++i) {
      const float pre = fmaf(w0.x, rel[i].x, fmaf(w0.y, rel[i].y, fmaf(w0.z, rel[i].z, w0.w)));
      h[i] = rel[i].w * fmaxf(pre, 0.f);
    }
    float w[C::CN];
#pragma unroll
    for (int c = 0; c < C::CN; c += 4) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col_ok && c0 + c < D) t = ldg4(wt + (size_t)kk * D + c0 + c);
      w[c] = t.x; w[c + 1] = t.y; w[c + 2] = t.z; w[c + 3] = t.w;
    }
#pragma unroll
    for (int i = 0; i < C::RM; ++i)
#pragma unroll
      for (int c = 0; c < C::CN; ++c) acc[i][c] = fmaf(h[i], w[c], acc[i][c]);
  }
}

// acc[i][c] = sum_kk buf[row_i][kk] * wt[kk][c0 + c]
template <class C>
__device__ __forceinline__ void gemm_smem(float (&acc)[C::RM][C::CN], const float *__restrict__ buf, int r0,
                                          const float *__restrict__ wt, int D, int c0) {
#pragma unroll
  for (int i = 0; i < C::RM; ++i)
#pragma unroll
    for (int c = 0; c < C::CN; ++c) acc[i][c] = 0.f;
  for (int kk = 0; kk < D; kk += 4) {
    float4 av[C::RM];
#pragma unroll
    for (int i = 0; i < C::RM; ++i) av[i] = *reinterpret_cast<const float4 *>(buf + (size_t)(r0 + i) * C::LD + kk);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float w[C::CN];
#pragma unroll
      for (int c = 0; c < C::CN; c += 4) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 + c < D) t = ldg4(wt + (size_t)(kk + u) * D + c0 + c);
        w[c] = t.x; w[c + 1] = t.y; w[c + 2] = t.z; w[c + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < C::RM; ++i) {
        const float a = u == 0 ? av[i].x : (u == 1 ? av[i].y : (u == 2 ? av[i].z : av[i].w));
#pragma unroll
        for (int c = 0; c < C::CN; ++c) acc[i][c] = fmaf(a, w[c], acc[i][c]);
      }
    }
  }
}

template <class C>
__global__ void __launch_bounds__(C::THREADS, 1) vattn_fwd_kernel(const nsdp_vattn_args a, float *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *buf = reinterpret_cast<float *>(smem_raw);                  // [R][LD]
  float4 *rel4 = reinterpret_cast<float4 *>(buf + (size_t)C::R * C::LD);  // [R]
  float4 *wd0s = rel4 + C::R;                                        // [DP]
  RowRef *rows = reinterpret_cast<RowRef *>(wd0s + C::DP);           // [R]

  const int D = a.D;
  const int krows = a.K + (a.has_global ? 1 : 0);
  const int tp = C::R / krows;
  const int tid = threadIdx.x;
  const int tx = tid % C::TX, ty = tid / C::TX;
  const int c0 = tx * C::CN;
  const int r0 = ty * C::RM;
  const long long tile = blockIdx.x;

  for (int kk = tid; kk < C::DP; kk += C::THREADS) {
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kk < D) w = make_float4(a.wd0[kk * 3 + 0], a.wd0[kk * 3 + 1], a.wd0[kk * 3 + 2], a.bd0[kk]);
    wd0s[kk] = w;
  }
  tile_rows_setup<C>(a, tile, krows, tp, rel4, rows);
  __syncthreads();

  float4 rel[C::RM];
  RowRef rr[C::RM];
#pragma unroll
  for (int i = 0; i < C::RM; ++i) {
    rel[i] = rel4[r0 + i];
    rr[i] = rows[r0 + i];
  }

  float acc[C::RM][C::CN];
  // ---- g = relu(h*W' + P) ---------------------------------------------------------------------------
  gemm_h<C>(acc, rel, wd0s, a.wpt, D, c0);
#pragma unroll
  for (int i = 0; i < C::RM; ++i) {
#pragma unroll
    for (int c = 0; c < C::CN; c += 4) {
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      const int col = c0 + c;
      if (rr[i].c >= 0 && col < D) {
        float4 p;
        if (rr[i].n >= 0) {
          p = ldg4(a.pc + col);
          if (a.qp) {
            const float4 q = ldg4(a.qp + (size_t)rr[i].c * D + col);
            p.x += q.x; p.y += q.y; p.z += q.z; p.w += q.w;
          }
          if (a.kp) {
            const float4 k = ldg4(a.kp + (size_t)rr[i].n * D + col);
            p.x -= k.x; p.y -= k.y; p.z -= k.z; p.w -= k.w;
          }
        } else {
          p = ldg4(a.gq + (size_t)(-rr[i].n - 1) * D + col);
        }
        g.x = fmaxf(acc[i][c] + p.x, 0.f);
        g.y = fmaxf(acc[i][c + 1] + p.y, 0.f);
        g.z = fmaxf(acc[i][c + 2] + p.z, 0.f);
        g.w = fmaxf(acc[i][c + 3] + p.w, 0.f);
      }
      *reinterpret_cast<float4 *>(buf + (size_t)(r0 + i) * C::LD + col) = g;
    }
  }
  __syncthreads();
  // ---- a = g*Wg2 ------------------------------------------------------------------------------------
  gemm_smem<C>(acc, buf, r0, a.wg2t, D, c0);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < C::RM; ++i)
#pragma unroll
    for (int c = 0; c < C::CN; c += 4)
      *reinterpret_cast<float4 *>(buf + (size_t)(r0 + i) * C::LD + c0 + c) =
          make_float4(acc[i][c], acc[i][c + 1], acc[i][c + 2], acc[i][c + 3]);
  __syncthreads();
  // ---- w = softmax over the rows of each point, per channel (in place) --------------------------------
  for (int item = tid; item < tp * D; item += C::THREADS) {
    const int p = item / D, c = item - p * D;
    float *col = buf + (size_t)(p * krows) * C::LD + c;
    float mx = -INFINITY;
    for (int t = 0; t < krows; ++t) mx = fmaxf(mx, col[(size_t)t * C::LD]);
    float sum = 0.f;
    for (int t = 0; t < krows; ++t) {
      const float e = expf(col[(size_t)t * C::LD] - mx);
      col[(size_t)t * C::LD] = e;
      sum += e;
    }
    const float inv = 1.f / sum;
    for (int t = 0; t < krows; ++t) col[(size_t)t * C::LD] *= inv;
  }
  __syncthreads();
  // ---- out = sum_rows w * (V + h*Wd2) ----------------------------------------------------------------
  gemm_h<C>(acc, rel, wd0s, a.wd2t, D, c0);
#pragma unroll
  for (int i = 0; i < C::RM; ++i) {
#pragma unroll
    for (int c = 0; c < C::CN; c += 4) {
      const int col = c0 + c;
      float4 *slot = reinterpret_cast<float4 *>(buf + (size_t)(r0 + i) * C::LD + col);
      float4 w = *slot;
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rr[i].c >= 0 && col < D) {
        if (rr[i].n >= 0) {
          s = ldg4(a.vc + col);
          if (a.vp) {
            const float4 v = ldg4(a.vp + (size_t)rr[i].n * D + col);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
          }
          s.x += acc[i][c]; s.y += acc[i][c + 1]; s.z += acc[i][c + 2]; s.w += acc[i][c + 3];
        } else {
          s = ldg4(a.gv + (size_t)(-rr[i].n - 1) * D + col);
        }
      }
      w.x *= s.x; w.y *= s.y; w.z *= s.z; w.w *= s.w;
      *slot = w;
    }
  }
  __syncthreads();
  const long long BM = (long long)a.B * a.M;
  for (int item = tid; item < tp * D; item += C::THREADS) {
    const int p = item / D, c = item - p * D;
    const long long ci = tile * tp + p;
    if (ci >= BM) continue;
    const float *col = buf + (size_t)(p * krows) * C::LD + c;
    float sum = 0.f;
    for (int t = 0; t < krows; ++t) sum += col[(size_t)t * C::LD];
    out[ci * D + c] = sum;
  }
}

template <class C>
static int launch_vattn_fwd(const nsdp_vattn_args &a, float *out, cudaStream_t st) {
  const int krows = a.K + (a.has_global ? 1 : 0);
  if (krows > C::R) return NSDP_ERR_UNSUPPORTED;
  const int tp = C::R / krows;
  const long long tiles = ceil_div((long long)a.B * a.M, (long long)tp);
  if (tiles > 0x7fffffffll) return NSDP_ERR_UNSUPPORTED;
  auto kern = vattn_fwd_kernel<C>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes());
    if (e != cudaSuccess) return cuda_rc(e);
    attr_set = true;
  }
  kern<<<(unsigned)tiles, C::THREADS, C::smem_bytes(), st>>>(a, out);
  return check_launch();
}

using VCfg120 = VCfg<30, 4, 16, 8>;   // D <= 120: 128 rows, 480 threads
using VCfg128 = VCfg<32, 4, 16, 8>;   // D <= 128
using VCfg200 = VCfg<25, 8, 20, 4>;   // D <= 200: 80 rows (10 decoder queries x 8), 500 threads
using VCfg256 = VCfg<32, 8, 16, 7>;   // D <= 256: 112 rows (full attention over 100 anchors), 512 threads

int vattn_validate(const nsdp_vattn_args *a) {
  if (!a || !a->xyz_c || !a->xyz_n || !a->wd0 || !a->bd0 || !a->wd2t || !a->wpt || !a->wg2t || !a->pc || !a->vc)
    return NSDP_ERR_INVALID_ARGUMENT;
  if (a->B <= 0 || a->M <= 0 || a->N <= 0 || a->K <= 0 || a->D <= 0) return NSDP_ERR_INVALID_ARGUMENT;
  if (!a->idx && a->K != a->N) return NSDP_ERR_INVALID_ARGUMENT;
  if (a->has_global && (!a->gq || !a->gv)) return NSDP_ERR_INVALID_ARGUMENT;
  if (a->D % 4 != 0 || a->D > 256) return NSDP_ERR_UNSUPPORTED;
  if ((long long)a->B * a->M >= (1ll << 31) || (long long)a->B * a->N >= (1ll << 31)) return NSDP_ERR_UNSUPPORTED;
  return NSDP_OK;
}

}  // namespace nsdp

extern "C" int nsdp_vattn_fwd_f32(const nsdp_vattn_args *args, float *out, void *stream) {
  using namespace nsdp;
  int rc = vattn_validate(args);
  if (rc != NSDP_OK) return rc;
  if (!out) return NSDP_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const int krows = args->K + (args->has_global ? 1 : 0);
  const int D = args->D;
  if (D <= 120 && krows <= VCfg120::R) return launch_vattn_fwd<VCfg120>(*args, out, st);
  if (D <= 128 && krows <= VCfg128::R) return launch_vattn_fwd<VCfg128>(*args, out, st);
  if (D <= 200 && krows <= VCfg200::R) return launch_vattn_fwd<VCfg200>(*args, out, st);
  return launch_vattn_fwd<VCfg256>(*args, out, st);
}

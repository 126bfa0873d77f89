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
#include "vattn_common.cuh"

namespace nsdp {

template <class C>
__global__ void __launch_bounds__(C::THREADS, 1) vattn_fwd_kernel(const nsdp_vattn_args a, float *__restrict__ out,
                                                                   float *__restrict__ stats) {
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
    if (stats) {  // softmax statistics for the backward kernel: w = exp(a - mx) * inv, row by row
      const long long ci = tile * tp + p;
      if (ci < (long long)a.B * a.M) {
        stats[ci * D + c] = mx;
        stats[((long long)a.B * a.M + ci) * D + c] = inv;
      }
    }
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
static int launch_vattn_fwd(const nsdp_vattn_args &a, float *out, float *stats, cudaStream_t st) {
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
  kern<<<(unsigned)tiles, C::THREADS, C::smem_bytes(), st>>>(a, out, stats);
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

namespace nsdp {
int vattn_fwd_tc_dispatch(const nsdp_vattn_args *args, float *out, float *stats, void *workspace, size_t ws_bytes,
                          cudaStream_t st, bool *handled);
}

extern "C" int nsdp_vattn_fwd_f32(const nsdp_vattn_args *args, float *out, float *stats, void *workspace,
                                  size_t workspace_bytes, void *stream) {
  using namespace nsdp;
  int rc = vattn_validate(args);
  if (rc != NSDP_OK) return rc;
  if (!out) return NSDP_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  if (args->impl != 1) {
    bool handled = false;
    rc = vattn_fwd_tc_dispatch(args, out, stats, workspace, workspace_bytes, st, &handled);
    if (handled) return rc;
    if (args->impl == 2) return NSDP_ERR_UNSUPPORTED;
  }
  const int krows = args->K + (args->has_global ? 1 : 0);
  const int D = args->D;
  if (D <= 120 && krows <= VCfg120::R) return launch_vattn_fwd<VCfg120>(*args, out, stats, st);
  if (D <= 128 && krows <= VCfg128::R) return launch_vattn_fwd<VCfg128>(*args, out, stats, st);
  if (D <= 200 && krows <= VCfg200::R) return launch_vattn_fwd<VCfg200>(*args, out, stats, st);
  return launch_vattn_fwd<VCfg256>(*args, out, stats, st);
}

// Fused neural-field MLP over query rows — fp32 CUDA-core version + the C-ABI entry points (nsdp_mlp_args).
//
// BASELINE.json configs[3] / SURVEY.md §8 C4: the decoder-only microbenchmark of the pattern the reference's decoder
// applies to every spatial sample (model/decoder/crosstransformer_decoder.py:63-69: a stack of nn.Linear + ReLU on
// [B*Q, width] rows). The reference runs it as one cuBLAS GEMM + one elementwise kernel per layer with a [rows, W]
// round trip through HBM each; here a CTA keeps a tile of rows on chip through the whole stack.
//
// This file is the numerically-straight fp32 path (`impl = 1`, any W % 4 == 0 up to 256): the in-repo reference the
// tcgen05 kernel (fused_mlp_tc.cu) is tested against, and the path for widths the tensor-core kernel is not
// instantiated for.
#include "common.cuh"

namespace nsdp {
namespace mlp {

constexpr int ROWS = 16;       // rows per CTA
constexpr int THREADS = 256;   // thread t owns output columns t, t + 256, ... of every row of the tile
constexpr int MAXW = 256;

// y[r][n] = relu?(bias[n] + sum_k x[r][k] * wt[k * ldw + n]) for the ROWS rows of the tile; x, y in shared memory
__device__ __forceinline__ void layer(const float *__restrict__ x, int kdim, const float *__restrict__ wt, int ldw,
                                      const float *__restrict__ bias, int ndim, bool relu, float *__restrict__ y,
                                      int ldy) {
  for (int n = threadIdx.x; n < ndim; n += THREADS) {
    float acc[ROWS];
    const float b = __ldg(bias + n);
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[r] = b;
    for (int k = 0; k < kdim; ++k) {
      const float w = __ldg(wt + (size_t)k * ldw + n);
#pragma unroll
      for (int r = 0; r < ROWS; ++r) acc[r] = fmaf(x[r * MAXW + k], w, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) y[r * ldy + n] = relu ? fmaxf(acc[r], 0.f) : acc[r];
  }
}

__global__ void __launch_bounds__(THREADS) fused_mlp_fwd_kernel(const nsdp_mlp_args a, float *__restrict__ out) {
  __shared__ float buf[2][ROWS * MAXW];
  const long long row0 = (long long)blockIdx.x * ROWS;
  const int nrows = (int)min((long long)ROWS, (long long)a.R - row0);
  for (int t = threadIdx.x; t < ROWS * a.Cin; t += THREADS) {
    const int r = t / a.Cin, c = t - r * a.Cin;
    buf[0][r * MAXW + c] = r < nrows ? __ldg(a.x + (size_t)(row0 + r) * a.Cin + c) : 0.f;
  }
  __syncthreads();
  layer(buf[0], a.Cin, a.w_in_t, a.W, a.b_in, a.W, true, buf[1], MAXW);
  __syncthreads();
  int cur = 1;
  for (int l = 0; l < a.n_hidden; ++l) {
    layer(buf[cur], a.W, a.w_h_t + (size_t)l * a.W * a.W, a.W, a.b_h + (size_t)l * a.W, a.W, true, buf[cur ^ 1], MAXW);
    __syncthreads();
    cur ^= 1;
  }
  layer(buf[cur], a.W, a.w_out_t, a.O, a.b_out, a.O, false, buf[cur ^ 1], MAXW);
  __syncthreads();
  for (int t = threadIdx.x; t < nrows * a.O; t += THREADS) {
    const int r = t / a.O, o = t - r * a.O;
    out[(row0 + r) * a.O + o] = buf[cur ^ 1][r * MAXW + o];
  }
}

}  // namespace mlp

size_t mlp_bwd_tc_workspace_bytes(const nsdp_mlp_args *a);
int mlp_bwd_tc_dispatch(const nsdp_mlp_args *a, const float *dout, const nsdp_mlp_grads *g, void *workspace, size_t ws_bytes,
                        cudaStream_t st);
size_t mlp_tc_workspace_bytes(const nsdp_mlp_args *a);
int mlp_tc_dispatch(const nsdp_mlp_args *a, float *out, void *workspace, size_t ws_bytes, cudaStream_t st, bool *handled);
}  // namespace nsdp

static int mlp_validate(const nsdp_mlp_args *a) {
  if (!a || !a->x || !a->w_in_t || !a->b_in || !a->w_out_t || !a->b_out) return NSDP_ERR_INVALID_ARGUMENT;
  if (a->n_hidden > 0 && (!a->w_h_t || !a->b_h)) return NSDP_ERR_INVALID_ARGUMENT;
  if (a->R <= 0 || a->Cin <= 0 || a->W <= 0 || a->O <= 0 || a->n_hidden < 0) return NSDP_ERR_INVALID_ARGUMENT;
  if (a->Cin > 4 || a->O > 4 || a->W > nsdp::mlp::MAXW || a->W % 4 != 0) return NSDP_ERR_UNSUPPORTED;
  return NSDP_OK;
}

extern "C" size_t nsdp_fused_mlp_fwd_workspace_bytes(const nsdp_mlp_args *a) {
  if (mlp_validate(a) != NSDP_OK || a->impl == 1) return 0;
  return nsdp::mlp_tc_workspace_bytes(a);
}

extern "C" int nsdp_fused_mlp_fwd_f32(const nsdp_mlp_args *a, float *out, void *workspace, size_t workspace_bytes,
                                      void *stream) {
  using namespace nsdp;
  int rc = mlp_validate(a);
  if (rc != NSDP_OK) return rc;
  if (!out) return NSDP_ERR_INVALID_ARGUMENT;
  if (a->impl != 1) {
    bool handled = false;
    rc = mlp_tc_dispatch(a, out, workspace, workspace_bytes, (cudaStream_t)stream, &handled);
    if (handled) return rc;
    if (a->impl == 2) return NSDP_ERR_UNSUPPORTED;
  }
  const long long tiles = ceil_div((long long)a->R, (long long)mlp::ROWS);
  mlp::fused_mlp_fwd_kernel<<<(unsigned)tiles, mlp::THREADS, 0, (cudaStream_t)stream>>>(*a, out);
  return check_launch();
}

extern "C" size_t nsdp_fused_mlp_bwd_workspace_bytes(const nsdp_mlp_args *a) {
  if (mlp_validate(a) != NSDP_OK) return 0;
  return nsdp::mlp_bwd_tc_workspace_bytes(a);
}

extern "C" int nsdp_fused_mlp_bwd_f32(const nsdp_mlp_args *a, const float *d_out, const nsdp_mlp_grads *g, void *workspace,
                                      size_t workspace_bytes, void *stream) {
  int rc = mlp_validate(a);
  if (rc != NSDP_OK) return rc;
  if (!d_out || !g || !g->d_w_in_t || !g->d_b_in || !g->d_w_out_t || !g->d_b_out || !g->d_w_h_t || !g->d_b_h)
    return NSDP_ERR_INVALID_ARGUMENT;
  return nsdp::mlp_bwd_tc_dispatch(a, d_out, g, workspace, workspace_bytes, (cudaStream_t)stream);
}

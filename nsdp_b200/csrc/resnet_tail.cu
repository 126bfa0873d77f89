// Fused decoder tail (ResNet-FC) — fp32 CUDA-core version.
//
// Replaces model/decoder/crosstransformer_decoder.py:63-69 + ResnetBlockFC (model/decoder/blocks.py:133-142):
// 17 cuBLAS GEMMs + elementwise kernels in the reference, each with a [B*Q, 128] round trip through HBM.
// Here a CTA keeps a tile of R = 128 query rows on chip for the whole stack: `lat` is staged once in shared
// memory (it is re-used by init_enc and all five fc_c layers), `net` lives in registers, and one [R][H]
// shared buffer carries relu(net) / relu(h) between the two GEMMs of a block. HBM traffic per query is
// C*4 bytes in, O*4 bytes out; weights (634 KB fp32) stream from L2.
#include "common.cuh"

namespace nsdp {

namespace tail {
constexpr int H = 128;
constexpr int TX = 32, CN = 4, TY = 16, RM = 8;
constexpr int R = TY * RM;  // 128
constexpr int THREADS = TX * TY;
constexpr int LDX = H + 4;

__host__ __device__ inline int lat_ld(int C) { return C + 4; }
inline size_t smem_bytes(int C) { return sizeof(float) * ((size_t)R * lat_ld(C) + (size_t)R * LDX); }

// acc[i][c] += sum_kk buf[(r0+i)*ld + kk] * wt[kk*wld + c0 + c]
__device__ __forceinline__ void gemm_acc(float (&acc)[RM][CN], const float *__restrict__ buf, int ld, int r0,
                                         const float *__restrict__ wt, int wld, int kdim, int c0) {
  for (int kk = 0; kk < kdim; kk += 4) {
    float4 av[RM];
#pragma unroll
    for (int i = 0; i < RM; ++i) av[i] = *reinterpret_cast<const float4 *>(buf + (size_t)(r0 + i) * ld + kk);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float4 w = ldg4(wt + (size_t)(kk + u) * wld + c0);
#pragma unroll
      for (int i = 0; i < RM; ++i) {
        const float a = u == 0 ? av[i].x : (u == 1 ? av[i].y : (u == 2 ? av[i].z : av[i].w));
        acc[i][0] = fmaf(a, w.x, acc[i][0]);
        acc[i][1] = fmaf(a, w.y, acc[i][1]);
        acc[i][2] = fmaf(a, w.z, acc[i][2]);
        acc[i][3] = fmaf(a, w.w, acc[i][3]);
      }
    }
  }
}

__device__ __forceinline__ void add_bias(float (&acc)[RM][CN], const float *__restrict__ bias, int c0) {
  const float4 b = ldg4(bias + c0);
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    acc[i][0] += b.x; acc[i][1] += b.y; acc[i][2] += b.z; acc[i][3] += b.w;
  }
}

__device__ __forceinline__ void store_relu(const float (&acc)[RM][CN], float *__restrict__ xbuf, int r0, int c0) {
#pragma unroll
  for (int i = 0; i < RM; ++i)
    *reinterpret_cast<float4 *>(xbuf + (size_t)(r0 + i) * LDX + c0) =
        make_float4(fmaxf(acc[i][0], 0.f), fmaxf(acc[i][1], 0.f), fmaxf(acc[i][2], 0.f), fmaxf(acc[i][3], 0.f));
}

__global__ void __launch_bounds__(THREADS, 1) resnet_tail_fwd_kernel(const nsdp_tail_args a, float *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C = a.C;
  const int ldl = lat_ld(C);
  float *lat = reinterpret_cast<float *>(smem_raw);  // [R][ldl]
  float *xbuf = lat + (size_t)R * ldl;               // [R][LDX]
  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int c0 = tx * CN, r0 = ty * RM;
  const long long row0 = (long long)blockIdx.x * R;
  const int nrows = (int)min((long long)R, (long long)a.R - row0);

  // stage the lat tile (rows beyond the end are zero-filled so the arithmetic stays finite)
  const int c4 = C / 4;
  for (int t = tid; t < R * c4; t += THREADS) {
    const int r = t / c4, q = t - r * c4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < nrows) v = ldg4(a.lat + (size_t)(row0 + r) * C + q * 4);
    *reinterpret_cast<float4 *>(lat + (size_t)r * ldl + q * 4) = v;
  }
  __syncthreads();

  const int wld = (1 + a.n_blocks) * H;
  float net[RM][CN];
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int c = 0; c < CN; ++c) net[i][c] = 0.f;
  gemm_acc(net, lat, ldl, r0, a.wc_t, wld, C, c0);
  add_bias(net, a.bc, c0);

  for (int blk = 0; blk < a.n_blocks; ++blk) {
    gemm_acc(net, lat, ldl, r0, a.wc_t + (size_t)(blk + 1) * H, wld, C, c0);
    add_bias(net, a.bc + (size_t)(blk + 1) * H, c0);
    __syncthreads();  // previous readers of xbuf are done
    store_relu(net, xbuf, r0, c0);
    __syncthreads();
    float hacc[RM][CN];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
      for (int c = 0; c < CN; ++c) hacc[i][c] = 0.f;
    gemm_acc(hacc, xbuf, LDX, r0, a.w0_t + (size_t)blk * H * H, H, H, c0);
    add_bias(hacc, a.b0 + (size_t)blk * H, c0);
    __syncthreads();
    store_relu(hacc, xbuf, r0, c0);
    __syncthreads();
    gemm_acc(net, xbuf, LDX, r0, a.w1_t + (size_t)blk * H * H, H, H, c0);
    add_bias(net, a.b1 + (size_t)blk * H, c0);
  }
  __syncthreads();
  store_relu(net, xbuf, r0, c0);
  __syncthreads();
  const int O = a.O;
  for (int t = tid; t < nrows * O; t += THREADS) {
    const int r = t / O, o = t - r * O;
    float s = a.bo[o];
    const float *x = xbuf + (size_t)r * LDX;
    for (int kk = 0; kk < H; ++kk) s = fmaf(x[kk], __ldg(a.wo_t + (size_t)kk * O + o), s);
    out[(row0 + r) * O + o] = s;
  }
}
}  // namespace tail
}  // namespace nsdp

namespace nsdp {
size_t tail_tc_workspace_bytes(const nsdp_tail_args *a);
int tail_tc_dispatch(const nsdp_tail_args *a, float *out, void *workspace, size_t ws_bytes, cudaStream_t st, bool *handled);
}

static int tail_validate(const nsdp_tail_args *a) {
  using namespace nsdp;
  if (!a || !a->lat || !a->wc_t || !a->bc || !a->wo_t || !a->bo) return NSDP_ERR_INVALID_ARGUMENT;
  if (a->n_blocks > 0 && (!a->w0_t || !a->b0 || !a->w1_t || !a->b1)) return NSDP_ERR_INVALID_ARGUMENT;
  if (a->R <= 0 || a->C <= 0 || a->O <= 0 || a->n_blocks < 0) return NSDP_ERR_INVALID_ARGUMENT;
  if (a->H != tail::H || a->C % 4 != 0 || a->C > 256 || a->O > 4) return NSDP_ERR_UNSUPPORTED;
  return NSDP_OK;
}

extern "C" size_t nsdp_resnet_tail_fwd_workspace_bytes(const nsdp_tail_args *a) {
  if (tail_validate(a) != NSDP_OK || a->impl == 1) return 0;
  return nsdp::tail_tc_workspace_bytes(a);
}

extern "C" int nsdp_resnet_tail_fwd_f32(const nsdp_tail_args *a, float *out, void *workspace, size_t workspace_bytes,
                                        void *stream) {
  using namespace nsdp;
  int rc = tail_validate(a);
  if (rc != NSDP_OK) return rc;
  if (!out) return NSDP_ERR_INVALID_ARGUMENT;
  if (a->impl != 1) {
    bool handled = false;
    rc = tail_tc_dispatch(a, out, workspace, workspace_bytes, (cudaStream_t)stream, &handled);
    if (handled) return rc;
    if (a->impl == 2) return NSDP_ERR_UNSUPPORTED;
  }
  const size_t smem = tail::smem_bytes(a->C);
  cudaError_t e = cudaFuncSetAttribute(tail::resnet_tail_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_rc(e);
  const long long tiles = ceil_div((long long)a->R, (long long)tail::R);
  tail::resnet_tail_fwd_kernel<<<(unsigned)tiles, tail::THREADS, smem, (cudaStream_t)stream>>>(*a, out);
  return check_launch();
}

// The remaining pointnet2_ops._ext operators as sm_100a kernels.
//
// Reference: pointnet2_ops_lib/pointnet2_ops/_ext-src/src/{sampling_gpu.cu:8-57, ball_query_gpu.cu:9-44,
// group_points_gpu.cu:8-64, interpolate_gpu.cu:9-143}. The reference launches ONE block per batch element
// (<= 512 threads), so at most B of the 148 SMs ever work; these are pure HBM-bound gathers/scans, so the
// kernels below flatten (batch, channel, point) into one grid sized from the problem, keep the innermost
// (contiguous) index on threadIdx.x for coalesced writes, and stage the scanned cloud in shared memory.
// Index results are bit-identical to the reference (same FMA contraction, same comparison order).
#include "common.cuh"

namespace nsdp {

constexpr int kThreads = 256;

// out[b,c,j] = points[b,c,idx[b,j]]
__global__ void gather_points_kernel(const float *__restrict__ points, const int32_t *__restrict__ idx, int C, int N,
                                     int M, long long total, float *__restrict__ out) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(t % M);
    const long long bc = t / M;
    const long long b = bc / C;
    const int a = idx[b * M + j];
    out[t] = points[bc * N + a];
  }
}

__global__ void gather_points_grad_kernel(const float *__restrict__ grad_out, const int32_t *__restrict__ idx, int C,
                                          int N, int M, long long total, float *__restrict__ grad_points) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(t % M);
    const long long bc = t / M;
    const long long b = bc / C;
    const int a = idx[b * M + j];
    atomicAdd(grad_points + bc * N + a, grad_out[t]);
  }
}

// out[b,c,j,k] = points[b,c,idx[b,j,k]]
__global__ void group_points_kernel(const float *__restrict__ points, const int32_t *__restrict__ idx, int C, int N,
                                    long long MK, long long total, float *__restrict__ out) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long jk = t % MK;
    const long long bc = t / MK;
    const long long b = bc / C;
    const int a = idx[b * MK + jk];
    out[t] = points[bc * N + a];
  }
}

__global__ void group_points_grad_kernel(const float *__restrict__ grad_out, const int32_t *__restrict__ idx, int C,
                                         int N, long long MK, long long total, float *__restrict__ grad_points) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long jk = t % MK;
    const long long bc = t / MK;
    const long long b = bc / C;
    const int a = idx[b * MK + jk];
    atomicAdd(grad_points + bc * N + a, grad_out[t]);
  }
}

constexpr int kScanChunk = 2048;  // points per shared-memory tile (24 KB)

// First `nsample` points in index order with d2 < r2; row pre-filled with the first hit.
// grid = (ceil(M / T), B); one thread per centre; the cloud streams through shared memory.
__global__ void __launch_bounds__(kThreads)
ball_query_kernel(const float *__restrict__ new_xyz, const float *__restrict__ xyz, int N, int M, float radius2,
                  int nsample, int32_t *__restrict__ out) {
  __shared__ float tile[kScanChunk * 3];
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = j < M;
  const float *__restrict__ p = xyz + (size_t)b * N * 3;
  float cx = 0.f, cy = 0.f, cz = 0.f;
  int32_t *o = nullptr;
  if (active) {
    const float *c = new_xyz + ((size_t)b * M + j) * 3;
    cx = c[0]; cy = c[1]; cz = c[2];
    o = out + ((size_t)b * M + j) * nsample;
    for (int l = 0; l < nsample; ++l) o[l] = 0;
  }
  int cnt = 0;
  for (int base = 0; base < N; base += kScanChunk) {
    const int n = min(kScanChunk, N - base);
    __syncthreads();
    for (int t = threadIdx.x; t < n * 3; t += blockDim.x) tile[t] = p[(size_t)base * 3 + t];
    __syncthreads();
    if (active && cnt < nsample) {
      for (int t = 0; t < n && cnt < nsample; ++t) {
        const float dx = cx - tile[t * 3 + 0], dy = cy - tile[t * 3 + 1], dz = cz - tile[t * 3 + 2];
        const float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
        if (d2 < radius2) {
          const int k = base + t;
          if (cnt == 0)
            for (int l = 0; l < nsample; ++l) o[l] = k;
          o[cnt] = k;
          ++cnt;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads)
three_nn_kernel(const float *__restrict__ unknown, const float *__restrict__ known, int n, int m,
                float *__restrict__ dist2, int32_t *__restrict__ idx) {
  __shared__ float tile[kScanChunk * 3];
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = j < n;
  const float *__restrict__ kn = known + (size_t)b * m * 3;
  float ux = 0.f, uy = 0.f, uz = 0.f;
  if (active) {
    const float *u = unknown + ((size_t)b * n + j) * 3;
    ux = u[0]; uy = u[1]; uz = u[2];
  }
  // running best kept in double initialised to 1e40, strict '<' — exactly interpolate_gpu.cu:27-49
  double best1 = 1e40, best2 = 1e40, best3 = 1e40;
  int b1 = 0, b2 = 0, b3 = 0;
  for (int base = 0; base < m; base += kScanChunk) {
    const int cnt = min(kScanChunk, m - base);
    __syncthreads();
    for (int t = threadIdx.x; t < cnt * 3; t += blockDim.x) tile[t] = kn[(size_t)base * 3 + t];
    __syncthreads();
    if (active) {
      for (int t = 0; t < cnt; ++t) {
        const float dx = ux - tile[t * 3 + 0], dy = uy - tile[t * 3 + 1], dz = uz - tile[t * 3 + 2];
        const float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
        const int k = base + t;
        if (d < best1) {
          best3 = best2; b3 = b2; best2 = best1; b2 = b1; best1 = d; b1 = k;
        } else if (d < best2) {
          best3 = best2; b3 = b2; best2 = d; b2 = k;
        } else if (d < best3) {
          best3 = d; b3 = k;
        }
      }
    }
  }
  if (active) {
    float *dd = dist2 + ((size_t)b * n + j) * 3;
    int32_t *ii = idx + ((size_t)b * n + j) * 3;
    dd[0] = (float)best1; dd[1] = (float)best2; dd[2] = (float)best3;
    ii[0] = b1; ii[1] = b2; ii[2] = b3;
  }
}

// out[b,c,j] = sum_t points[b,c,idx[b,j,t]] * weight[b,j,t]
__global__ void three_interpolate_kernel(const float *__restrict__ points, const int32_t *__restrict__ idx,
                                         const float *__restrict__ weight, int C, int m, int n, long long total,
                                         float *__restrict__ out) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(t % n);
    const long long bc = t / n;
    const long long b = bc / C;
    const float *w = weight + (b * n + j) * 3;
    const int32_t *ii = idx + (b * n + j) * 3;
    const float *p = points + bc * m;
    // same contraction nvcc gives the reference's `p1*w1 + p2*w2 + p3*w3`: FMUL p2*w2, FFMA p1*w1+., FFMA p3*w3+.
    out[t] = __fmaf_rn(p[ii[2]], w[2], __fmaf_rn(p[ii[0]], w[0], __fmul_rn(p[ii[1]], w[1])));
  }
}

__global__ void three_interpolate_grad_kernel(const float *__restrict__ grad_out, const int32_t *__restrict__ idx,
                                              const float *__restrict__ weight, int C, int n, int m, long long total,
                                              float *__restrict__ grad_points) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(t % n);
    const long long bc = t / n;
    const long long b = bc / C;
    const float *w = weight + (b * n + j) * 3;
    const int32_t *ii = idx + (b * n + j) * 3;
    float *g = grad_points + bc * m;
    const float go = grad_out[t];
    atomicAdd(g + ii[0], go * w[0]);
    atomicAdd(g + ii[1], go * w[1]);
    atomicAdd(g + ii[2], go * w[2]);
  }
}

static unsigned flat_grid(long long total) {
  long long blocks = ceil_div(total, (long long)kThreads);
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace nsdp

using namespace nsdp;

extern "C" int nsdp_gather_points_f32(const float *points, const int32_t *idx, int B, int C, int N, int M, float *out,
                                      void *stream) {
  if (!points || !idx || !out || B <= 0 || C <= 0 || N <= 0 || M < 0) return NSDP_ERR_INVALID_ARGUMENT;
  if (M == 0) return NSDP_OK;
  const long long total = (long long)B * C * M;
  gather_points_kernel<<<flat_grid(total), kThreads, 0, (cudaStream_t)stream>>>(points, idx, C, N, M, total, out);
  return check_launch();
}

extern "C" int nsdp_gather_points_grad_f32(const float *grad_out, const int32_t *idx, int B, int C, int N, int M,
                                           float *grad_points, void *stream) {
  if (!grad_out || !idx || !grad_points || B <= 0 || C <= 0 || N <= 0 || M < 0) return NSDP_ERR_INVALID_ARGUMENT;
  if (M == 0) return NSDP_OK;
  const long long total = (long long)B * C * M;
  gather_points_grad_kernel<<<flat_grid(total), kThreads, 0, (cudaStream_t)stream>>>(grad_out, idx, C, N, M, total,
                                                                                      grad_points);
  return check_launch();
}

extern "C" int nsdp_group_points_f32(const float *points, const int32_t *idx, int B, int C, int N, int M, int K,
                                     float *out, void *stream) {
  if (!points || !idx || !out || B <= 0 || C <= 0 || N <= 0 || M < 0 || K < 0) return NSDP_ERR_INVALID_ARGUMENT;
  const long long MK = (long long)M * K;
  if (MK == 0) return NSDP_OK;
  const long long total = (long long)B * C * MK;
  group_points_kernel<<<flat_grid(total), kThreads, 0, (cudaStream_t)stream>>>(points, idx, C, N, MK, total, out);
  return check_launch();
}

extern "C" int nsdp_group_points_grad_f32(const float *grad_out, const int32_t *idx, int B, int C, int N, int M, int K,
                                          float *grad_points, void *stream) {
  if (!grad_out || !idx || !grad_points || B <= 0 || C <= 0 || N <= 0 || M < 0 || K < 0)
    return NSDP_ERR_INVALID_ARGUMENT;
  const long long MK = (long long)M * K;
  if (MK == 0) return NSDP_OK;
  const long long total = (long long)B * C * MK;
  group_points_grad_kernel<<<flat_grid(total), kThreads, 0, (cudaStream_t)stream>>>(grad_out, idx, C, N, MK, total,
                                                                                     grad_points);
  return check_launch();
}

extern "C" int nsdp_ball_query_f32(const float *new_xyz, const float *xyz, int B, int N, int M, float radius,
                                   int nsample, int32_t *out_idx, void *stream) {
  if (!new_xyz || !xyz || !out_idx || B <= 0 || N <= 0 || M <= 0 || nsample <= 0) return NSDP_ERR_INVALID_ARGUMENT;
  if (B > 65535) return NSDP_ERR_UNSUPPORTED;
  dim3 grid((unsigned)ceil_div(M, kThreads), (unsigned)B);
  ball_query_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(new_xyz, xyz, N, M, radius * radius, nsample, out_idx);
  return check_launch();
}

extern "C" int nsdp_three_nn_f32(const float *unknown, const float *known, int B, int n, int m, float *dist2,
                                 int32_t *out_idx, void *stream) {
  if (!unknown || !known || !dist2 || !out_idx || B <= 0 || n <= 0 || m <= 0) return NSDP_ERR_INVALID_ARGUMENT;
  if (B > 65535) return NSDP_ERR_UNSUPPORTED;
  dim3 grid((unsigned)ceil_div(n, kThreads), (unsigned)B);
  three_nn_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(unknown, known, n, m, dist2, out_idx);
  return check_launch();
}

extern "C" int nsdp_three_interpolate_f32(const float *points, const int32_t *idx, const float *weight, int B, int C,
                                          int m, int n, float *out, void *stream) {
  if (!points || !idx || !weight || !out || B <= 0 || C <= 0 || m <= 0 || n <= 0) return NSDP_ERR_INVALID_ARGUMENT;
  const long long total = (long long)B * C * n;
  three_interpolate_kernel<<<flat_grid(total), kThreads, 0, (cudaStream_t)stream>>>(points, idx, weight, C, m, n, total,
                                                                                     out);
  return check_launch();
}

extern "C" int nsdp_three_interpolate_grad_f32(const float *grad_out, const int32_t *idx, const float *weight, int B,
                                               int C, int n, int m, float *grad_points, void *stream) {
  if (!grad_out || !idx || !weight || !grad_points || B <= 0 || C <= 0 || m <= 0 || n <= 0)
    return NSDP_ERR_INVALID_ARGUMENT;
  const long long total = (long long)B * C * n;
  three_interpolate_grad_kernel<<<flat_grid(total), kThreads, 0, (cudaStream_t)stream>>>(grad_out, idx, weight, C, n, m,
                                                                                          total, grad_points);
  return check_launch();
}

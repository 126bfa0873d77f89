// ElementwiseMLP of the point-transformer encoder as fused fp32 kernels (reference: model/encoder/blocks.py:137-159):
//
//     out = bn3( x + relu( bn2( conv2( relu( bn1( conv1 x ) ) ) ) ) )        x: (R, C) rows, conv = 1x1 Conv1d = Linear
//
// The reference runs 2 cuDNN convolutions, 3 cuDNN batch-norms (each: statistics + transform), 2 ReLUs, an add and the
// permutes between (B, n, C) and (B, C, n): ~12 launches forward and ~30 backward for a few hundred KB of data — pure
// launch latency. Here the forward is 4 launches and the backward 6:
//
//   forward   L1  t1 = x W1^T + b1                              (+ column sums of t1, t1^2 in the epilogue)
//             L2  t2 = relu(bn1(t1)) W2^T + b2                  (bn1 + ReLU applied while the operand tile is loaded)
//             E3  s  = x + relu(bn2(t2))                        (+ column sums of s, s^2)
//             E4  out = bn3(s)                                  (+ running-stat updates of the three BatchNorms)
//   backward  B1  column sums of dout, dout * s_hat                                   -> d gamma3, d beta3
//             B2  ds = bn3'(dout);  dx = ds;  dh2 = ds * [bn2(t2) > 0]                (+ sums for bn2')
//             B3  da1 = (bn2'(dh2) W2) * [bn1(t1) > 0]                                (+ sums for bn1')
//             B4  dW2 = bn2'(dh2)^T relu(bn1(t1)),  db2
//             B5  dx += bn1'(da1) W1
//             B6  dW1 = bn1'(da1)^T x,  db1;  BatchNorm parameter gradients
//
// Everything is plain fp32 FMA on CUDA cores (exact fp32 semantics, like the cuBLAS SGEMM the reference's convolution
// ends in; the matrices are [<=4000 x 256] x [256 x 256]: 0.5 GFLOP, tens of microseconds), batch statistics accumulate
// in fp64 (sum, sum of squares -> biased variance for normalisation, unbiased for running_var, as nn.BatchNorm1d).
// Training mode uses batch statistics, eval mode the running ones; both modes are differentiable.
#include "common.cuh"

namespace nsdp {
namespace emlp {

constexpr int BM = 64, BN = 64, BK = 16, THREADS = 256;
constexpr int MAXC = 256;

// where the (mean, 1/std) of one BatchNorm come from
struct BnSrc {
  const double *sums;   // training: [2C] column sums of x and x^2 over `rows` rows
  const float *rm, *rv; // eval: running mean / var
  const float *gamma, *beta;
  int training, rows;
  float eps;
};

__device__ __forceinline__ void bn_stats(const BnSrc &s, int c, int C, float &mean, float &istd) {
  if (s.training) {
    const double m = s.sums[c] / (double)s.rows;
    double v = s.sums[C + c] / (double)s.rows - m * m;
    v = v < 0.0 ? 0.0 : v;
    mean = (float)m;
    istd = (float)(1.0 / sqrt(v + (double)s.eps));
  } else {
    mean = s.rm[c];
    istd = 1.0f / sqrtf(s.rv[c] + s.eps);
  }
}

// Per-column coefficients kept in shared memory by the kernels below.
//   forward transform   y = x * fs + fo                 (fs = gamma * istd, fo = beta - mean * fs)
//   backward transform  dx = bs * (dy - c1 - x_hat * c2),  x_hat = (x - mean) * istd,  c1 = sum(dy)/R, c2 = sum(dy x_hat)/R
struct Coef {
  float fs[MAXC], fo[MAXC], mean[MAXC], istd[MAXC], c1[MAXC], c2[MAXC];
};

__device__ __forceinline__ void load_fwd_coef(Coef &k, const BnSrc &s, int C) {
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float m, i;
    bn_stats(s, c, C, m, i);
    k.mean[c] = m;
    k.istd[c] = i;
    k.fs[c] = s.gamma[c] * i;
    k.fo[c] = s.beta[c] - m * k.fs[c];
  }
}
// bsums: [2C] column sums of dy and dy * x_hat (fp64); in eval mode the batch terms vanish
__device__ __forceinline__ void load_bwd_coef(Coef &k, const BnSrc &s, const double *bsums, int C) {
  load_fwd_coef(k, s, C);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    k.c1[c] = s.training ? (float)(bsums[c] / (double)s.rows) : 0.f;
    k.c2[c] = s.training ? (float)(bsums[C + c] / (double)s.rows) : 0.f;
  }
}
__device__ __forceinline__ float bn_fwd(const Coef &k, int c, float x) { return fmaf(x, k.fs[c], k.fo[c]); }
__device__ __forceinline__ float bn_bwd(const Coef &k, int c, float dy, float x) {
  const float xh = (x - k.mean[c]) * k.istd[c];
  return k.fs[c] * (dy - k.c1[c] - xh * k.c2[c]);
}

// ---------------------------------------------------------------------------------------------------------------------
// 64 x 64 output tile, 256 threads, 4 x 4 per thread; operands staged k-major in shared memory
// ---------------------------------------------------------------------------------------------------------------------
struct Tile {
  float a[BK][BM + 4];
  float b[BK][BN + 4];
};

__device__ __forceinline__ void tile_fma(const Tile &t, int ty, int tx, float (&acc)[4][4]) {
#pragma unroll
  for (int kk = 0; kk < BK; ++kk) {
    const float4 av = *reinterpret_cast<const float4 *>(&t.a[kk][ty * 4]);
    const float4 bv = *reinterpret_cast<const float4 *>(&t.b[kk][tx * 4]);
    const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
  }
}

// column sums of a 64 x 64 tile held as 4 x 4 per thread: v[i][j] -> sums[col] (+= over rows), optional second moment
__device__ __forceinline__ void tile_colsums(float (*red)[BN], int ty, int tx, const float (&p)[4], double *dst, int n0, int C) {
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) red[ty][tx * 4 + j] = p[j];
  __syncthreads();
  if (threadIdx.x < BN) {
    float s = 0.f;
#pragma unroll
    for (int y = 0; y < 16; ++y) s += red[y][threadIdx.x];
    if (n0 + (int)threadIdx.x < C) atomicAdd(dst + n0 + threadIdx.x, (double)s);
  }
}

// L1 / L2:  Y[r][n] = sum_k A'[r][k] W[n][k] + bias[n],  A' = X or relu(bn(X));  epilogue: column sums of Y and Y^2
template <bool PRO>
__global__ void __launch_bounds__(THREADS)
linear_nt_kernel(const float *__restrict__ X, const float *__restrict__ W, const float *__restrict__ bias,
                 float *__restrict__ Y, int R, int C, BnSrc pro, double *__restrict__ sums_out) {
  __shared__ Tile t;
  __shared__ Coef k;
  __shared__ float red[16][BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  if (PRO) load_fwd_coef(k, pro, C);
  float acc[4][4] = {};
  for (int k0 = 0; k0 < C; k0 += BK) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * THREADS, row = e >> 4, kk = e & 15;
      const int r = m0 + row, c = k0 + kk;
      float v = 0.f;
      if (r < R && c < C) {
        v = X[(size_t)r * C + c];
        if (PRO) v = fmaxf(bn_fwd(k, c, v), 0.f);
      }
      t.a[kk][row] = v;
      const int n = n0 + row;
      t.b[kk][row] = (n < C && c < C) ? W[(size_t)n * C + c] : 0.f;
    }
    __syncthreads();
    tile_fma(t, ty, tx, acc);
  }
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n0 + tx * 4 + j;
    const float b = (bias && n < C) ? bias[n] : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = m0 + ty * 4 + i;
      const float y = acc[i][j] + b;
      if (r < R && n < C) {
        Y[(size_t)r * C + n] = y;
        s1[j] += y;
        s2[j] = fmaf(y, y, s2[j]);
      }
    }
  }
  if (sums_out) {
    tile_colsums(red, ty, tx, s1, sums_out, n0, C);
    tile_colsums(red, ty, tx, s2, sums_out + C, n0, C);
  }
}

// B3 / B5:  O[r][k] = sum_n A'[r][n] W[n][k],  A' = bn'(DY; T)  (BatchNorm backward applied on load)
//   EPI 0 (B3): O *= [bn_m(TM)[r][k] > 0]; write; column sums of O and O * tm_hat (for the next BatchNorm backward)
//   EPI 1 (B5): O is ADDED to the output (dx already holds the residual branch)
template <int EPI>
__global__ void __launch_bounds__(THREADS)
linear_nn_bwd_kernel(const float *__restrict__ DY, const float *__restrict__ T, const float *__restrict__ W,
                     float *__restrict__ O, int R, int C, BnSrc bn, const double *__restrict__ bsums,
                     const float *__restrict__ TM, BnSrc bnm, double *__restrict__ sums_out) {
  __shared__ Tile t;
  __shared__ Coef k;
  __shared__ Coef km;
  __shared__ float red[16][BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;   // n0 indexes OUTPUT columns (k)
  load_bwd_coef(k, bn, bsums, C);
  if (EPI == 0) load_fwd_coef(km, bnm, C);
  float acc[4][4] = {};
  for (int k0 = 0; k0 < C; k0 += BK) {     // k0 runs over the contraction index n
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * THREADS;
      {
        const int row = e >> 4, kk = e & 15, r = m0 + row, c = k0 + kk;
        float v = 0.f;
        if (r < R && c < C) v = bn_bwd(k, c, DY[(size_t)r * C + c], T[(size_t)r * C + c]);
        t.a[kk][row] = v;
      }
      {
        const int kk = e >> 6, col = e & 63, n = k0 + kk, c = n0 + col;   // W rows are contiguous along the output index
        t.b[kk][col] = (n < C && c < C) ? W[(size_t)n * C + c] : 0.f;
      }
    }
    __syncthreads();
    tile_fma(t, ty, tx, acc);
  }
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = n0 + tx * 4 + j;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = m0 + ty * 4 + i;
      if (r < R && c < C) {
        const size_t o = (size_t)r * C + c;
        if (EPI == 0) {
          const float tm = TM[o];
          const float v = bn_fwd(km, c, tm) > 0.f ? acc[i][j] : 0.f;
          O[o] = v;
          s1[j] += v;
          s2[j] = fmaf(v, (tm - km.mean[c]) * km.istd[c], s2[j]);
        } else {
          O[o] += acc[i][j];
        }
      }
    }
  }
  if (EPI == 0) {
    tile_colsums(red, ty, tx, s1, sums_out, n0, C);
    tile_colsums(red, ty, tx, s2, sums_out + C, n0, C);
  }
}

// B4 / B6:  dW[n][k] += sum_r P[r][n] Q[r][k],  db[n] += sum_r P[r][n],  P = bn'(DY; T),  Q = relu(bn_q(XQ)) or XQ.
// Grid: (n tiles, k tiles, row splits); partial tiles are added atomically.
template <bool QPRO>
__global__ void __launch_bounds__(THREADS)
weight_tn_kernel(const float *__restrict__ DY, const float *__restrict__ T, const float *__restrict__ XQ,
                 float *__restrict__ dW, float *__restrict__ db, int R, int C, int rows_per_split, BnSrc bn,
                 const double *__restrict__ bsums, BnSrc bnq) {
  __shared__ Tile t;
  __shared__ Coef k;
  __shared__ Coef kq;
  __shared__ float red[16][BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.x * BM, k0c = blockIdx.y * BN;
  const int r_begin = blockIdx.z * rows_per_split, r_end = min(R, r_begin + rows_per_split);
  load_bwd_coef(k, bn, bsums, C);
  if (QPRO) load_fwd_coef(kq, bnq, C);
  float acc[4][4] = {};
  float bsum = 0.f;   // thread (kk = tid >> 6, col = tid & 63) accumulates P[.][n0 + col] for the bias gradient
  for (int r0 = r_begin; r0 < r_end; r0 += BK) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * THREADS, kk = e >> 6, col = e & 63, r = r0 + kk;
      float p = 0.f, q = 0.f;
      if (r < r_end) {
        const int cn = n0 + col, ck = k0c + col;
        if (cn < C) p = bn_bwd(k, cn, DY[(size_t)r * C + cn], T[(size_t)r * C + cn]);
        if (ck < C) {
          q = XQ[(size_t)r * C + ck];
          if (QPRO) q = fmaxf(bn_fwd(kq, ck, q), 0.f);
        }
      }
      t.a[kk][col] = p;
      t.b[kk][col] = q;
      bsum += p;
    }
    __syncthreads();
    tile_fma(t, ty, tx, acc);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = k0c + tx * 4 + j;
      if (n < C && c < C) atomicAdd(dW + (size_t)n * C + c, acc[i][j]);
    }
  }
  if (db && blockIdx.y == 0) {
    // element e = tid + i*256 -> col = tid & 63 for every i: thread tid owns column (tid & 63), 4 threads per column
    __syncthreads();
    red[tid >> 6][tid & 63] = bsum;
    __syncthreads();
    if (tid < BN && n0 + tid < C) atomicAdd(db + n0 + tid, red[0][tid] + red[1][tid] + red[2][tid] + red[3][tid]);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// elementwise passes: thread = column, block = a contiguous chunk of rows (coalesced rows, per-column register sums)
// ---------------------------------------------------------------------------------------------------------------------
// E3: s = x + relu(bn2(t2)); sums of s, s^2
__global__ void __launch_bounds__(MAXC)
residual_kernel(const float *__restrict__ X, const float *__restrict__ T2, float *__restrict__ S, int R, int C,
                int rows_per_block, BnSrc bn2, double *__restrict__ sums_out) {
  const int c = threadIdx.x;
  if (c >= C) return;
  float mean, istd;
  bn_stats(bn2, c, C, mean, istd);
  const float fs = bn2.gamma[c] * istd, fo = bn2.beta[c] - mean * fs;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(R, r0 + rows_per_block);
  float s1 = 0.f, s2 = 0.f;
  for (int r = r0; r < r1; ++r) {
    const size_t o = (size_t)r * C + c;
    const float v = X[o] + fmaxf(fmaf(T2[o], fs, fo), 0.f);
    S[o] = v;
    s1 += v;
    s2 = fmaf(v, v, s2);
  }
  if (sums_out && r1 > r0) {
    atomicAdd(sums_out + c, (double)s1);
    atomicAdd(sums_out + C + c, (double)s2);
  }
}

struct Running {
  float *rm[3], *rv[3];
  long long *nbt[3];
  float momentum;
};

// E4: out = bn3(s); block 0 also folds the batch statistics of the three BatchNorms into their running buffers
__global__ void __launch_bounds__(MAXC)
bn_out_kernel(const float *__restrict__ S, float *__restrict__ OUT, int R, int C, int rows_per_block, BnSrc bn3,
              const double *__restrict__ sums_all /* [3][2C] */, Running run) {
  const int c = threadIdx.x;
  if (c >= C) return;
  float mean, istd;
  bn_stats(bn3, c, C, mean, istd);
  const float fs = bn3.gamma[c] * istd, fo = bn3.beta[c] - mean * fs;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(R, r0 + rows_per_block);
  for (int r = r0; r < r1; ++r) {
    const size_t o = (size_t)r * C + c;
    OUT[o] = fmaf(S[o], fs, fo);
  }
  if (blockIdx.x == 0 && bn3.training) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double *su = sums_all + (size_t)i * 2 * C;
      const double m = su[c] / (double)R;
      double v = su[C + c] / (double)R - m * m;
      v = v < 0.0 ? 0.0 : v;
      const double unbiased = R > 1 ? v * (double)R / (double)(R - 1) : v;
      run.rm[i][c] = (1.f - run.momentum) * run.rm[i][c] + run.momentum * (float)m;
      run.rv[i][c] = (1.f - run.momentum) * run.rv[i][c] + run.momentum * (float)unbiased;
      if (c == 0 && run.nbt[i]) *run.nbt[i] += 1;
    }
  }
}

// B1: sums of dout and dout * s_hat
__global__ void __launch_bounds__(MAXC)
bwd_sums_kernel(const float *__restrict__ DOUT, const float *__restrict__ S, int R, int C, int rows_per_block, BnSrc bn3,
                double *__restrict__ bsums) {
  const int c = threadIdx.x;
  if (c >= C) return;
  float mean, istd;
  bn_stats(bn3, c, C, mean, istd);
  const int r0 = blockIdx.x * rows_per_block, r1 = min(R, r0 + rows_per_block);
  float s1 = 0.f, s2 = 0.f;
  for (int r = r0; r < r1; ++r) {
    const size_t o = (size_t)r * C + c;
    const float d = DOUT[o];
    s1 += d;
    s2 = fmaf(d, (S[o] - mean) * istd, s2);
  }
  if (r1 > r0) {
    atomicAdd(bsums + c, (double)s1);
    atomicAdd(bsums + C + c, (double)s2);
  }
}

// B2: ds = bn3'(dout); dx = ds; dh2 = ds * [bn2(t2) > 0]; sums of dh2 and dh2 * t2_hat
__global__ void __launch_bounds__(MAXC)
bwd_residual_kernel(const float *__restrict__ DOUT, const float *__restrict__ S, const float *__restrict__ T2,
                    float *__restrict__ DX, float *__restrict__ DH2, int R, int C, int rows_per_block, BnSrc bn3,
                    const double *__restrict__ bsums3, BnSrc bn2, double *__restrict__ bsums2) {
  const int c = threadIdx.x;
  if (c >= C) return;
  float m3, i3, m2, i2;
  bn_stats(bn3, c, C, m3, i3);
  bn_stats(bn2, c, C, m2, i2);
  const float bs3 = bn3.gamma[c] * i3;
  const float c1 = bn3.training ? (float)(bsums3[c] / (double)R) : 0.f, c2 = bn3.training ? (float)(bsums3[C + c] / (double)R) : 0.f;
  const float fs2 = bn2.gamma[c] * i2, fo2 = bn2.beta[c] - m2 * fs2;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(R, r0 + rows_per_block);
  float s1 = 0.f, s2 = 0.f;
  for (int r = r0; r < r1; ++r) {
    const size_t o = (size_t)r * C + c;
    const float ds = bs3 * (DOUT[o] - c1 - (S[o] - m3) * i3 * c2);
    DX[o] = ds;
    const float t2 = T2[o];
    const float dh = fmaf(t2, fs2, fo2) > 0.f ? ds : 0.f;
    DH2[o] = dh;
    s1 += dh;
    s2 = fmaf(dh, (t2 - m2) * i2, s2);
  }
  if (r1 > r0) {
    atomicAdd(bsums2 + c, (double)s1);
    atomicAdd(bsums2 + C + c, (double)s2);
  }
}

// BatchNorm parameter gradients from the fp64 sums: d gamma = sum(dy x_hat), d beta = sum(dy)
__global__ void bn_param_grads_kernel(const double *__restrict__ bsums /* [3][2C]: bn1, bn2, bn3 */, int C,
                                      float *dg1, float *db1, float *dg2, float *db2, float *dg3, float *db3) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float *dg[3] = {dg1, dg2, dg3}, *dbt[3] = {db1, db2, db3};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (dbt[i]) dbt[i][c] = (float)bsums[(size_t)i * 2 * C + c];
    if (dg[i]) dg[i][c] = (float)bsums[(size_t)i * 2 * C + C + c];
  }
}

static BnSrc bn_src(const nsdp_emlp_args &a, int which, const double *stats) {
  BnSrc s;
  s.sums = stats ? stats + (size_t)which * 2 * a.C : nullptr;
  s.rm = a.running_mean[which];
  s.rv = a.running_var[which];
  s.gamma = a.bn_weight[which];
  s.beta = a.bn_bias[which];
  s.training = a.training;
  s.rows = a.R;
  s.eps = a.eps;
  return s;
}

static int rows_per_block(int R) {
  const int blocks = 2 * num_sms();
  int rpb = ceil_div(R, blocks);
  return rpb < 4 ? 4 : rpb;
}

static bool args_ok(const nsdp_emlp_args *a) {
  if (!a || !a->x || !a->w1 || !a->w2 || a->R <= 0 || a->C <= 0) return false;
  for (int i = 0; i < 3; ++i) {
    if (!a->bn_weight[i] || !a->bn_bias[i]) return false;
    if (!a->training && (!a->running_mean[i] || !a->running_var[i])) return false;
  }
  return true;
}

}  // namespace emlp
}  // namespace nsdp

extern "C" size_t nsdp_emlp_stats_bytes(const nsdp_emlp_args *a) {
  return a && a->C > 0 ? (size_t)6 * a->C * sizeof(double) : 0;
}

extern "C" int nsdp_emlp_fwd_f32(const nsdp_emlp_args *a, float *out, float *t1, float *t2, float *s, double *stats,
                                 void *stream) {
  using namespace nsdp;
  using namespace nsdp::emlp;
  if (!args_ok(a) || !out || !t1 || !t2 || !s || !stats) return NSDP_ERR_INVALID_ARGUMENT;
  if (a->C > MAXC) return NSDP_ERR_UNSUPPORTED;
  for (int i = 0; i < 3; ++i)   // track_running_stats=False is a configuration the model never uses
    if (!a->running_mean[i] || !a->running_var[i]) return NSDP_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int R = a->R, C = a->C;
  cudaError_t e = cudaMemsetAsync(stats, 0, nsdp_emlp_stats_bytes(a), st);
  if (e != cudaSuccess) return cuda_rc(e);
  const dim3 grid((unsigned)ceil_div(R, BM), (unsigned)ceil_div(C, BN));
  const BnSrc b1 = bn_src(*a, 0, stats), b2 = bn_src(*a, 1, stats), b3 = bn_src(*a, 2, stats);
  linear_nt_kernel<false><<<grid, THREADS, 0, st>>>(a->x, a->w1, a->b1, t1, R, C, b1, a->training ? stats : nullptr);
  int rc = check_launch();
  if (rc != NSDP_OK) return rc;
  linear_nt_kernel<true><<<grid, THREADS, 0, st>>>(t1, a->w2, a->b2, t2, R, C, b1, a->training ? stats + 2 * C : nullptr);
  rc = check_launch();
  if (rc != NSDP_OK) return rc;
  const int rpb = rows_per_block(R);
  const unsigned blocks = (unsigned)ceil_div(R, rpb);
  residual_kernel<<<blocks, MAXC, 0, st>>>(a->x, t2, s, R, C, rpb, b2, a->training ? stats + 4 * C : nullptr);
  rc = check_launch();
  if (rc != NSDP_OK) return rc;
  Running run;
  for (int i = 0; i < 3; ++i) {
    run.rm[i] = a->running_mean[i];
    run.rv[i] = a->running_var[i];
    run.nbt[i] = a->num_batches_tracked[i];
  }
  run.momentum = a->momentum;
  bn_out_kernel<<<blocks, MAXC, 0, st>>>(s, out, R, C, rpb, b3, stats, run);
  return check_launch();
}

extern "C" size_t nsdp_emlp_bwd_workspace_bytes(const nsdp_emlp_args *a) {
  if (!a || a->R <= 0 || a->C <= 0) return 0;
  // dh2, da1 (R x C each) + fp64 sums for the three BatchNorm backwards
  return (size_t)2 * a->R * a->C * sizeof(float) + (size_t)6 * a->C * sizeof(double);
}

extern "C" int nsdp_emlp_bwd_f32(const nsdp_emlp_args *a, const float *t1, const float *t2, const float *s,
                                 const double *stats, const float *d_out, const nsdp_emlp_grads *g, void *workspace,
                                 size_t workspace_bytes, void *stream) {
  using namespace nsdp;
  using namespace nsdp::emlp;
  if (!args_ok(a) || !t1 || !t2 || !s || !stats || !d_out || !g || !g->d_x || !g->d_w1 || !g->d_w2)
    return NSDP_ERR_INVALID_ARGUMENT;
  if (a->C > MAXC) return NSDP_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < nsdp_emlp_bwd_workspace_bytes(a)) return NSDP_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int R = a->R, C = a->C;
  double *bsums = (double *)workspace;                 // [3][2C]: bn1, bn2, bn3
  float *dh2 = (float *)(bsums + (size_t)6 * C);
  float *da1 = dh2 + (size_t)R * C;
  cudaError_t e = cudaMemsetAsync(bsums, 0, (size_t)6 * C * sizeof(double), st);
  if (e != cudaSuccess) return cuda_rc(e);
  // weight / bias gradients are accumulated atomically: the caller passes ZEROED buffers (or ones to accumulate into)
  const BnSrc b1 = bn_src(*a, 0, stats), b2 = bn_src(*a, 1, stats), b3 = bn_src(*a, 2, stats);
  const int rpb = rows_per_block(R);
  const unsigned blocks = (unsigned)ceil_div(R, rpb);
  int rc;
  bwd_sums_kernel<<<blocks, MAXC, 0, st>>>(d_out, s, R, C, rpb, b3, bsums + 4 * C);
  if ((rc = check_launch()) != NSDP_OK) return rc;
  bwd_residual_kernel<<<blocks, MAXC, 0, st>>>(d_out, s, t2, g->d_x, dh2, R, C, rpb, b3, bsums + 4 * C, b2, bsums + 2 * C);
  if ((rc = check_launch()) != NSDP_OK) return rc;
  const dim3 grid((unsigned)ceil_div(R, BM), (unsigned)ceil_div(C, BN));
  linear_nn_bwd_kernel<0><<<grid, THREADS, 0, st>>>(dh2, t2, a->w2, da1, R, C, b2, bsums + 2 * C, t1, b1, bsums);
  if ((rc = check_launch()) != NSDP_OK) return rc;
  const int tiles = ceil_div(C, BM) * ceil_div(C, BN);
  int splits = ceil_div(2 * num_sms(), tiles);
  int rps = ceil_div(ceil_div(R, splits), BK) * BK;
  if (rps < 4 * BK) rps = 4 * BK;
  splits = ceil_div(R, rps);
  const dim3 wgrid((unsigned)ceil_div(C, BM), (unsigned)ceil_div(C, BN), (unsigned)splits);
  weight_tn_kernel<true><<<wgrid, THREADS, 0, st>>>(dh2, t2, t1, g->d_w2, g->d_b2, R, C, rps, b2, bsums + 2 * C, b1);
  if ((rc = check_launch()) != NSDP_OK) return rc;
  linear_nn_bwd_kernel<1><<<grid, THREADS, 0, st>>>(da1, t1, a->w1, g->d_x, R, C, b1, bsums, nullptr, b1, nullptr);
  if ((rc = check_launch()) != NSDP_OK) return rc;
  weight_tn_kernel<false><<<wgrid, THREADS, 0, st>>>(da1, t1, a->x, g->d_w1, g->d_b1, R, C, rps, b1, bsums, b1);
  if ((rc = check_launch()) != NSDP_OK) return rc;
  bn_param_grads_kernel<<<(unsigned)ceil_div(C, 128), 128, 0, st>>>(bsums, C, g->d_bn_weight[0], g->d_bn_bias[0],
                                                                    g->d_bn_weight[1], g->d_bn_bias[1], g->d_bn_weight[2],
                                                                    g->d_bn_bias[2]);
  return check_launch();
}

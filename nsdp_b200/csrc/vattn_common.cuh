// Device-side building blocks shared by the vector-attention forward and backward kernels.
#pragma once
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


int vattn_validate(const nsdp_vattn_args *a);

}  // namespace nsdp

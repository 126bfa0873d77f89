// Backward of the fused vector attention (fp32 CUDA-core version).
//
// The reference relies on autograd, which keeps ~12 [B, M, K, D] fp32 activations alive per attention block
// (17 GB for one decoder pass at B=4, Q=50k). Here NOTHING of size [pairs, D] is stored: a CTA
// takes a tile of R consecutive pair rows (flattened over (centre, neighbour); tiles need not align to
// centres because the softmax statistics (max, 1/sum) per (centre, channel) were saved by the forward),
// recomputes h, g, a, w, s on chip, and pushes the gradient back through the chain. Three [R][D]
// shared-memory buffers are recycled through the roles  g -> dgp,  ds,  da -> h -> dpre.
//
// With w = exp(a - mx) * inv,  s = vc + vp + h*Wd2 (or gv),  out = sum_rows w*s:
//     ds = w*dout                      da = ds * (s - out)
//     dWg2t += g^T da                  dg = da*Wg2^T ; dgp = dg * [g > 0]
//     dWpt  += h^T dgp                 dWd2t += h^T ds
//     dh = dgp*Wp^T + ds*Wd2^T         dpre = dh * [h > 0]
//     dWd0 += dpre^T rel, dbd0 += sum dpre, drel = dpre*Wd0
// Per-point gradients (d_qp, d_kp, d_vp, d_xyz_*) are scattered with fp32 atomics (order-nondeterministic,
// like the reference's own index_put/atomicAdd backward); weight gradients are reduced per tile on chip and
// added to the global accumulators with one atomic per element per tile.
#include "vattn_common.cuh"

namespace nsdp {

template <int TX_, int CN_, int TY_, int RM_, int KM_>
struct BCfg : VCfg<TX_, CN_, TY_, RM_> {
  using Base = VCfg<TX_, CN_, TY_, RM_>;
  static constexpr int KM = KM_;  // rows of a [D x D] weight-gradient tile owned by a thread per pass
  static constexpr size_t smem_bytes() {
    return sizeof(float) * (3 * (size_t)Base::R * Base::LD + (size_t)Base::R * 4 + (size_t)Base::DP * 4 +
                            2 * (size_t)Base::DP) + sizeof(int) * (size_t)Base::R * 2;
  }
};

// rows are flattened: rho = tile*R + r -> (centre = rho / krows, t = rho % krows)
template <class C>
__device__ __forceinline__ void bwd_rows_setup(const nsdp_vattn_args &a, long long tile, int krows, float4 *rel4,
                                               RowRef *rows) {
  const long long total = (long long)a.B * a.M * krows;
  for (int r = threadIdx.x; r < C::R; r += C::THREADS) {
    const long long rho = tile * C::R + r;
    RowRef rr;
    rr.c = -1;
    rr.n = 0;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rho < total) {
      const long long ci = rho / krows;
      const int t = (int)(rho - ci * krows);
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

// acc[i][c] += sum_kk buf[row_i][kk] * wt[kk][c0 + c]   (no reset)
template <class C>
__device__ __forceinline__ void gemm_smem_acc(float (&acc)[C::RM][C::CN], const float *__restrict__ buf, int r0,
                                              const float *__restrict__ wt, int D, int c0) {
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
        const float x = u == 0 ? av[i].x : (u == 1 ? av[i].y : (u == 2 ? av[i].z : av[i].w));
#pragma unroll
        for (int c = 0; c < C::CN; ++c) acc[i][c] = fmaf(x, w[c], acc[i][c]);
      }
    }
  }
}

// out[k][c] += sum_r A[r][k] * Bm[r][c]  over the tile's R rows; one atomic per element per tile.
template <class C>
__device__ __forceinline__ void gemm_tn_atomic(const float *__restrict__ A, const float *__restrict__ Bm,
                                               float *__restrict__ out, int D, int tx, int ty) {
  if (!out) return;
  const int c0 = tx * C::CN;
  for (int kbase = 0; kbase < D; kbase += C::TY * C::KM) {
    const int k0 = kbase + ty * C::KM;
    if (k0 >= D) continue;  // warp-divergent only at the ragged end
    float acc[C::KM][C::CN];
#pragma unroll
    for (int j = 0; j < C::KM; ++j)
#pragma unroll
      for (int c = 0; c < C::CN; ++c) acc[j][c] = 0.f;
#pragma unroll 2
    for (int r = 0; r < C::R; ++r) {
      float av[C::KM];
#pragma unroll
      for (int j = 0; j < C::KM; j += 2) {
        const float2 t = *reinterpret_cast<const float2 *>(A + (size_t)r * C::LD + k0 + j);
        av[j] = t.x; av[j + 1] = t.y;
      }
      float bv[C::CN];
#pragma unroll
      for (int c = 0; c < C::CN; c += 4) {
        const float4 t = *reinterpret_cast<const float4 *>(Bm + (size_t)r * C::LD + c0 + c);
        bv[c] = t.x; bv[c + 1] = t.y; bv[c + 2] = t.z; bv[c + 3] = t.w;
      }
#pragma unroll
      for (int j = 0; j < C::KM; ++j)
#pragma unroll
        for (int c = 0; c < C::CN; ++c) acc[j][c] = fmaf(av[j], bv[c], acc[j][c]);
    }
#pragma unroll
    for (int j = 0; j < C::KM; ++j) {
      const int k = k0 + j;
      if (k < D) {
#pragma unroll
        for (int c = 0; c < C::CN; ++c)
          if (c0 + c < D) atomicAdd(out + (size_t)k * D + c0 + c, acc[j][c]);
      }
    }
  }
}

template <class C>
__global__ void __launch_bounds__(C::THREADS, 1)
vattn_bwd_kernel(const nsdp_vattn_args a, const float *__restrict__ out, const float *__restrict__ stats,
                 const float *__restrict__ dout, const nsdp_vattn_grads g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *bufG = reinterpret_cast<float *>(smem_raw);        // g -> dgp
  float *bufX = bufG + (size_t)C::R * C::LD;                // ds
  float *bufY = bufX + (size_t)C::R * C::LD;                // da -> h -> dpre
  float4 *rel4 = reinterpret_cast<float4 *>(bufY + (size_t)C::R * C::LD);
  float4 *wd0s = rel4 + C::R;
  float *red_a = reinterpret_cast<float *>(wd0s + C::DP);   // [DP] per-channel reductions
  float *red_b = red_a + C::DP;                             // [DP]
  RowRef *rows = reinterpret_cast<RowRef *>(red_b + C::DP);

  const int D = a.D;
  const int krows = a.K + (a.has_global ? 1 : 0);
  const int tid = threadIdx.x;
  const int tx = tid % C::TX, ty = tid / C::TX;
  const int c0 = tx * C::CN, r0 = ty * C::RM;
  const long long tile = blockIdx.x;
  const long long BM = (long long)a.B * a.M;

  for (int kk = tid; kk < C::DP; kk += C::THREADS) {
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kk < D) w = make_float4(a.wd0[kk * 3 + 0], a.wd0[kk * 3 + 1], a.wd0[kk * 3 + 2], a.bd0[kk]);
    wd0s[kk] = w;
    red_a[kk] = 0.f;
    red_b[kk] = 0.f;
  }
  bwd_rows_setup<C>(a, tile, krows, rel4, rows);
  __syncthreads();

  float4 rel[C::RM];
  RowRef rr[C::RM];
#pragma unroll
  for (int i = 0; i < C::RM; ++i) {
    rel[i] = rel4[r0 + i];
    rr[i] = rows[r0 + i];
  }

  float acc[C::RM][C::CN];
  // ---- recompute g = relu(h*W' + P) -> bufG ------------------------------------------------------------
  gemm_h<C>(acc, rel, wd0s, a.wpt, D, c0);
#pragma unroll
  for (int i = 0; i < C::RM; ++i) {
#pragma unroll
    for (int c = 0; c < C::CN; c += 4) {
      float4 gg = make_float4(0.f, 0.f, 0.f, 0.f);
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
        gg.x = fmaxf(acc[i][c] + p.x, 0.f);
        gg.y = fmaxf(acc[i][c + 1] + p.y, 0.f);
        gg.z = fmaxf(acc[i][c + 2] + p.z, 0.f);
        gg.w = fmaxf(acc[i][c + 3] + p.w, 0.f);
      }
      *reinterpret_cast<float4 *>(bufG + (size_t)(r0 + i) * C::LD + col) = gg;
    }
  }
  __syncthreads();
  // ---- a = g*Wg2 ; w = exp(a - mx)*inv (kept in registers) ---------------------------------------------------
  gemm_smem<C>(acc, bufG, r0, a.wg2t, D, c0);
  float wsm[C::RM][C::CN];
#pragma unroll
  for (int i = 0; i < C::RM; ++i) {
#pragma unroll
    for (int c = 0; c < C::CN; c += 4) {
      const int col = c0 + c;
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rr[i].c >= 0 && col < D) {
        const float4 mx = ldg4(stats + (size_t)rr[i].c * D + col);
        const float4 iv = ldg4(stats + ((size_t)BM + rr[i].c) * D + col);
        w.x = expf(acc[i][c] - mx.x) * iv.x;
        w.y = expf(acc[i][c + 1] - mx.y) * iv.y;
        w.z = expf(acc[i][c + 2] - mx.z) * iv.z;
        w.w = expf(acc[i][c + 3] - mx.w) * iv.w;
      }
      wsm[i][c] = w.x; wsm[i][c + 1] = w.y; wsm[i][c + 2] = w.z; wsm[i][c + 3] = w.w;
    }
  }
  // ---- s = V + h*Wd2 ; ds = w*dout -> bufX ; da = ds*(s - out) -> bufY ; scatter d_vp / d_gv / d_vc ----------
  gemm_h<C>(acc, rel, wd0s, a.wd2t, D, c0);
  {
    float vc_part[C::CN];
#pragma unroll
    for (int c = 0; c < C::CN; ++c) vc_part[c] = 0.f;
#pragma unroll
    for (int i = 0; i < C::RM; ++i) {
#pragma unroll
      for (int c = 0; c < C::CN; c += 4) {
        const int col = c0 + c;
        float4 ds = make_float4(0.f, 0.f, 0.f, 0.f), da = ds;
        if (rr[i].c >= 0 && col < D) {
          const bool glob = rr[i].n < 0;
          float4 s;
          if (!glob) {
            s = ldg4(a.vc + col);
            if (a.vp) {
              const float4 v = ldg4(a.vp + (size_t)rr[i].n * D + col);
              s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
            s.x += acc[i][c]; s.y += acc[i][c + 1]; s.z += acc[i][c + 2]; s.w += acc[i][c + 3];
          } else {
            s = ldg4(a.gv + (size_t)(-rr[i].n - 1) * D + col);
          }
          const float4 go = ldg4(dout + (size_t)rr[i].c * D + col);
          const float4 o = ldg4(out + (size_t)rr[i].c * D + col);
          ds.x = wsm[i][c] * go.x; ds.y = wsm[i][c + 1] * go.y; ds.z = wsm[i][c + 2] * go.z; ds.w = wsm[i][c + 3] * go.w;
          da.x = ds.x * (s.x - o.x); da.y = ds.y * (s.y - o.y); da.z = ds.z * (s.z - o.z); da.w = ds.w * (s.w - o.w);
          if (!glob) {
            if (g.d_vp) {
              float *dst = g.d_vp + (size_t)rr[i].n * D + col;
              atomicAdd(dst, ds.x); atomicAdd(dst + 1, ds.y); atomicAdd(dst + 2, ds.z); atomicAdd(dst + 3, ds.w);
            }
            vc_part[c] += ds.x; vc_part[c + 1] += ds.y; vc_part[c + 2] += ds.z; vc_part[c + 3] += ds.w;
          } else if (g.d_gv) {
            float *dst = g.d_gv + (size_t)(-rr[i].n - 1) * D + col;
            atomicAdd(dst, ds.x); atomicAdd(dst + 1, ds.y); atomicAdd(dst + 2, ds.z); atomicAdd(dst + 3, ds.w);
          }
        }
        *reinterpret_cast<float4 *>(bufX + (size_t)(r0 + i) * C::LD + col) = ds;
        *reinterpret_cast<float4 *>(bufY + (size_t)(r0 + i) * C::LD + col) = da;
      }
    }
    if (g.d_vc) {
#pragma unroll
      for (int c = 0; c < C::CN; ++c)
        if (c0 + c < D) atomicAdd(&red_a[c0 + c], vc_part[c]);
    }
  }
  __syncthreads();
  // ---- dWg2t += g^T da ---------------------------------------------------------------------------------------
  gemm_tn_atomic<C>(bufG, bufY, g.d_wg2t, D, tx, ty);
  // ---- dg = da*Wg2^T ; dgp = dg*[g>0] -> bufG (in place) ; scatter d_qp / d_kp / d_gq / d_pc --------------------------
#pragma unroll
  for (int i = 0; i < C::RM; ++i)
#pragma unroll
    for (int c = 0; c < C::CN; ++c) acc[i][c] = 0.f;
  gemm_smem_acc<C>(acc, bufY, r0, a.wg2, D, c0);
  __syncthreads();  // every reader of bufG (g) and bufY (da) is done
  {
    float pc_part[C::CN];
#pragma unroll
    for (int c = 0; c < C::CN; ++c) pc_part[c] = 0.f;
#pragma unroll
    for (int i = 0; i < C::RM; ++i) {
#pragma unroll
      for (int c = 0; c < C::CN; c += 4) {
        const int col = c0 + c;
        float4 *slot = reinterpret_cast<float4 *>(bufG + (size_t)(r0 + i) * C::LD + col);
        const float4 gg = *slot;
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rr[i].c >= 0 && col < D) {
          d.x = gg.x > 0.f ? acc[i][c] : 0.f;
          d.y = gg.y > 0.f ? acc[i][c + 1] : 0.f;
          d.z = gg.z > 0.f ? acc[i][c + 2] : 0.f;
          d.w = gg.w > 0.f ? acc[i][c + 3] : 0.f;
          if (rr[i].n >= 0) {
            if (g.d_qp) {
              float *dst = g.d_qp + (size_t)rr[i].c * D + col;
              atomicAdd(dst, d.x); atomicAdd(dst + 1, d.y); atomicAdd(dst + 2, d.z); atomicAdd(dst + 3, d.w);
            }
            if (g.d_kp) {
              float *dst = g.d_kp + (size_t)rr[i].n * D + col;
              atomicAdd(dst, -d.x); atomicAdd(dst + 1, -d.y); atomicAdd(dst + 2, -d.z); atomicAdd(dst + 3, -d.w);
            }
            pc_part[c] += d.x; pc_part[c + 1] += d.y; pc_part[c + 2] += d.z; pc_part[c + 3] += d.w;
          } else if (g.d_gq) {
            float *dst = g.d_gq + (size_t)(-rr[i].n - 1) * D + col;
            atomicAdd(dst, d.x); atomicAdd(dst + 1, d.y); atomicAdd(dst + 2, d.z); atomicAdd(dst + 3, d.w);
          }
        }
        *slot = d;
      }
    }
    if (g.d_pc) {
#pragma unroll
      for (int c = 0; c < C::CN; ++c)
        if (c0 + c < D) atomicAdd(&red_b[c0 + c], pc_part[c]);
    }
  }
  // ---- h -> bufY (da is dead) -----------------------------------------------------------------------------------
#pragma unroll
  for (int i = 0; i < C::RM; ++i) {
#pragma unroll
    for (int c = 0; c < C::CN; ++c) {
      const int col = c0 + c;
      const float4 w0 = wd0s[col];
      const float pre = fmaf(w0.x, rel[i].x, fmaf(w0.y, rel[i].y, fmaf(w0.z, rel[i].z, w0.w)));
      bufY[(size_t)(r0 + i) * C::LD + col] = (col < D) ? rel[i].w * fmaxf(pre, 0.f) : 0.f;
    }
  }
  __syncthreads();
  if (g.d_vc)
    for (int c = tid; c < D; c += C::THREADS) atomicAdd(g.d_vc + c, red_a[c]);
  if (g.d_pc)
    for (int c = tid; c < D; c += C::THREADS) atomicAdd(g.d_pc + c, red_b[c]);
  // ---- dWpt += h^T dgp ; dWd2t += h^T ds ---------------------------------------------------------------------------
  gemm_tn_atomic<C>(bufY, bufG, g.d_wpt, D, tx, ty);
  gemm_tn_atomic<C>(bufY, bufX, g.d_wd2t, D, tx, ty);
  // ---- dh = dgp*Wp^T + ds*Wd2^T ; dpre = dh*[h>0] -> bufY (in place) ---------------------------------------------------
  const bool need_dpre = g.d_wd0 || g.d_bd0 || g.d_xyz_c || g.d_xyz_n;
  if (!need_dpre) return;
#pragma unroll
  for (int i = 0; i < C::RM; ++i)
#pragma unroll
    for (int c = 0; c < C::CN; ++c) acc[i][c] = 0.f;
  gemm_smem_acc<C>(acc, bufG, r0, a.wp, D, c0);
  gemm_smem_acc<C>(acc, bufX, r0, a.wd2, D, c0);
  __syncthreads();  // all readers of bufY (h) in gemm_tn_atomic are done
#pragma unroll
  for (int i = 0; i < C::RM; ++i) {
#pragma unroll
    for (int c = 0; c < C::CN; ++c) {
      float *slot = bufY + (size_t)(r0 + i) * C::LD + c0 + c;
      *slot = (*slot > 0.f) ? acc[i][c] : 0.f;
    }
  }
  __syncthreads();
  // ---- dWd0, dbd0: one thread per hidden channel ----------------------------------------------------------------
  if (g.d_wd0 || g.d_bd0) {
    for (int kk = tid; kk < D; kk += C::THREADS) {
      float sx = 0.f, sy = 0.f, sz = 0.f, sb = 0.f;
      for (int r = 0; r < C::R; ++r) {
        const float d = bufY[(size_t)r * C::LD + kk];
        const float4 rl = rel4[r];
        sx = fmaf(d, rl.x, sx); sy = fmaf(d, rl.y, sy); sz = fmaf(d, rl.z, sz); sb += d;
      }
      if (g.d_wd0) {
        atomicAdd(g.d_wd0 + kk * 3 + 0, sx); atomicAdd(g.d_wd0 + kk * 3 + 1, sy); atomicAdd(g.d_wd0 + kk * 3 + 2, sz);
      }
      if (g.d_bd0) atomicAdd(g.d_bd0 + kk, sb);
    }
  }
  // ---- drel = dpre*Wd0 -> d_xyz_c / d_xyz_n: one warp per row ---------------------------------------------------------
  if (g.d_xyz_c || g.d_xyz_n) {
    // only FULL warps take rows: the block size need not be a multiple of 32 (BCfg200 has 500 threads) and the
    // shuffle reduction below is undefined on a partial warp
    const int warp = tid >> 5, lane = tid & 31, nwarps = C::THREADS >> 5;
    for (int r = warp; r < C::R && warp < nwarps; r += nwarps) {
      const RowRef ref = rows[r];
      if (ref.c < 0 || ref.n < 0) continue;  // warp-uniform
      float sx = 0.f, sy = 0.f, sz = 0.f;
      for (int kk = lane; kk < D; kk += 32) {
        const float d = bufY[(size_t)r * C::LD + kk];
        const float4 w0 = wd0s[kk];
        sx = fmaf(d, w0.x, sx); sy = fmaf(d, w0.y, sy); sz = fmaf(d, w0.z, sz);
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, off);
        sy += __shfl_xor_sync(0xffffffffu, sy, off);
        sz += __shfl_xor_sync(0xffffffffu, sz, off);
      }
      if (lane == 0) {
        // rel = sign*(xc - xn)
        if (g.d_xyz_c) {
          float *dst = g.d_xyz_c + (size_t)ref.c * 3;
          atomicAdd(dst, a.sign * sx); atomicAdd(dst + 1, a.sign * sy); atomicAdd(dst + 2, a.sign * sz);
        }
        if (g.d_xyz_n) {
          float *dst = g.d_xyz_n + (size_t)ref.n * 3;
          atomicAdd(dst, -a.sign * sx); atomicAdd(dst + 1, -a.sign * sy); atomicAdd(dst + 2, -a.sign * sz);
        }
      }
    }
  }
}

template <class C>
static int launch_vattn_bwd(const nsdp_vattn_args &a, const float *out, const float *stats, const float *dout,
                            const nsdp_vattn_grads &g, cudaStream_t st) {
  const int krows = a.K + (a.has_global ? 1 : 0);
  const long long tiles = ceil_div((long long)a.B * a.M * krows, (long long)C::R);
  if (tiles > 0x7fffffffll) return NSDP_ERR_UNSUPPORTED;
  auto kern = vattn_bwd_kernel<C>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes());
  if (e != cudaSuccess) return cuda_rc(e);
  kern<<<(unsigned)tiles, C::THREADS, C::smem_bytes(), st>>>(a, out, stats, dout, g);
  return check_launch();
}

using BCfg120 = BCfg<30, 4, 16, 8, 8>;    // D <= 120: 128 rows
using BCfg128 = BCfg<32, 4, 16, 8, 8>;    // D <= 128
using BCfg200 = BCfg<25, 8, 20, 4, 10>;   // D <= 200: 80 rows
using BCfg256 = BCfg<32, 8, 16, 4, 8>;    // D <= 256: 64 rows

}  // namespace nsdp

namespace nsdp {
size_t vattn_bwd_tc_workspace_bytes(const nsdp_vattn_args *a);
int vattn_bwd_tc_dispatch(const nsdp_vattn_args *a, const float *out, const float *stats, const float *dout,
                          const nsdp_vattn_grads *g, void *workspace, size_t ws_bytes, cudaStream_t st, bool *handled);
}

extern "C" size_t nsdp_vattn_bwd_workspace_bytes(const nsdp_vattn_args *args) {
  if (!args || nsdp::vattn_validate(args) != NSDP_OK || args->impl == 1) return 0;
  return nsdp::vattn_bwd_tc_workspace_bytes(args);
}

extern "C" int nsdp_vattn_bwd_f32(const nsdp_vattn_args *args, const float *out, const float *stats, const float *d_out,
                                  const nsdp_vattn_grads *grads, void *workspace, size_t workspace_bytes, void *stream) {
  using namespace nsdp;
  int rc = vattn_validate(args);
  if (rc != NSDP_OK) return rc;
  if (!out || !stats || !d_out || !grads) return NSDP_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  if (args->impl != 1 && grads->d_wd2t && grads->d_wpt && grads->d_wg2t) {
    bool handled = false;
    rc = vattn_bwd_tc_dispatch(args, out, stats, d_out, grads, workspace, workspace_bytes, st, &handled);
    if (handled) return rc;
    if (args->impl == 2) return NSDP_ERR_UNSUPPORTED;
  }
  if (!args->wd2 || !args->wp || !args->wg2) return NSDP_ERR_INVALID_ARGUMENT;
  const int D = args->D;
  if (D <= 120) return launch_vattn_bwd<BCfg120>(*args, out, stats, d_out, *grads, st);
  if (D <= 128) return launch_vattn_bwd<BCfg128>(*args, out, stats, d_out, *grads, st);
  if (D <= 200) return launch_vattn_bwd<BCfg200>(*args, out, stats, d_out, *grads, st);
  return launch_vattn_bwd<BCfg256>(*args, out, stats, d_out, *grads, st);
}

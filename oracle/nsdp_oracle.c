/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the integer/index kernels on NSDP's TDNet hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; nothing under nsdp_b200/ may.
 *
 * Parity status: the reference ships NO tests or golden vectors for these kernels
 * (SURVEY.md §4, §8c), so the restatement is pinned the other way round: on the GPU box
 * the reference's own CUDA sources are compiled in place into oracle/_ref/ (see
 * oracle/build_ref.py) and tests/test_gpu_ref_ext.py checks this file against that
 * binary, index for index.
 *
 * Build (see oracle/build.py):  gcc -O2 -ffp-contract=off -shared -fPIC ...
 * -ffp-contract=off is REQUIRED: every fused multiply-add below is an explicit fmaf()
 * that mirrors a contraction nvcc performs in the reference kernel; the compiler must
 * not add or remove any.
 *
 * Every function cites the reference file:line (paths relative to /root/reference/) it
 * follows.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* pointnet2_ops_lib/pointnet2_ops/_ext-src/include/cuda_utils.h:13-19 (opt_n_threads):
 * pow_2 = (int)(log(work)/log(2)) evaluated in double; clamp to [1, 512]. */
int nsdp_oracle_opt_n_threads(int work_size) {
  const int pow_2 = (int)(log((double)work_size) / log(2.0));
  int t = 1 << pow_2;
  if (t > 512) t = 512;
  if (t < 1) t = 1;
  return t;
}

/*
 * Furthest point sampling — literal simulation of the CUDA block:
 *   kernel   pointnet2_ops_lib/pointnet2_ops/_ext-src/src/sampling_gpu.cu:69-173
 *   __update pointnet2_ops_lib/pointnet2_ops/_ext-src/src/sampling_gpu.cu:59-65
 *   host     pointnet2_ops_lib/pointnet2_ops/_ext-src/src/sampling.cpp:66-87 (temp = 1e10, int32 out)
 *
 * One "thread" t of a block of BS threads walks points t, t+BS, ...; the per-thread
 * (best, besti) pairs go through the same BS/2 ... 1 shared-memory tree, with the same
 * strict '>' at both levels, so the tie-break is reproduced by construction rather than
 * by formula. Arithmetic follows the FMA contraction nvcc emits for sm_100a
 * (FMUL y*y ; FFMA x*x+. ; FFMA z*z+.), see SURVEY.md Appendix B.
 *
 * xyz: (B, N, 3) float32, out: (B, m) int32. Returns 0.
 */
int nsdp_oracle_fps(const float *xyz, int B, int N, int m, int32_t *out) {
  if (m <= 0) return 0;
  const int BS = nsdp_oracle_opt_n_threads(N);
  float *temp = (float *)malloc(sizeof(float) * (size_t)N);
  float *dists = (float *)malloc(sizeof(float) * (size_t)BS);
  int *dists_i = (int *)malloc(sizeof(int) * (size_t)BS);
  if (!temp || !dists || !dists_i) return -1;
  for (int b = 0; b < B; ++b) {
    const float *p = xyz + (size_t)b * N * 3;
    int32_t *idx = out + (size_t)b * m;
    for (int k = 0; k < N; ++k) temp[k] = 1e10f; /* sampling.cpp:74-76 */
    for (int j = 0; j < m; ++j) idx[j] = 0;      /* torch::zeros, sampling.cpp:70-72 */
    int old = 0;                                 /* sampling_gpu.cu:85-86 */
    for (int j = 1; j < m; ++j) {
      const float x1 = p[old * 3 + 0], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
      for (int t = 0; t < BS; ++t) { /* per-thread strided pass, :93-110 */
        int besti = 0;
        float best = -1.0f;
        for (int k = t; k < N; k += BS) {
          const float x2 = p[k * 3 + 0], y2 = p[k * 3 + 1], z2 = p[k * 3 + 2];
          const float mag = fmaf(z2, z2, fmaf(x2, x2, y2 * y2));
          if ((double)mag <= 1e-3) continue; /* :100, compared in double */
          const float dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
          const float d = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
          const float d2 = fminf(d, temp[k]);
          temp[k] = d2;
          besti = d2 > best ? k : besti;
          best = d2 > best ? d2 : best;
        }
        dists[t] = best;
        dists_i[t] = besti;
      }
      for (int s = BS / 2; s >= 1; s >>= 1) { /* tree, :115-168 */
        for (int t = 0; t < s; ++t) {
          const float v1 = dists[t], v2 = dists[t + s];
          const int i1 = dists_i[t], i2 = dists_i[t + s];
          dists[t] = fmaxf(v1, v2);
          dists_i[t] = v2 > v1 ? i2 : i1;
        }
      }
      old = dists_i[0];
      idx[j] = old;
    }
  }
  free(temp);
  free(dists);
  free(dists_i);
  return 0;
}

/*
 * k nearest neighbours — restates model/utils.py:39-55 (square_distance: sum over the
 * last dim of (src - dst)**2, i.e. ((dx*dx + dy*dy) + dz*dz) with separately rounded
 * squares, no FMA) followed by `dists.argsort()[:, :, :k]`
 * (model/encoder/blocks.py:101-102, 287-288; model/decoder/blocks.py:50-52).
 * torch's default argsort is unstable, so tie order is undefined in the reference; the
 * restatement (and the product kernel) define it as "lowest index first".
 *
 * query: (B, M, 3), ref: (B, N, 3) -> idx: (B, M, k) int32, d2 (nullable): (B, M, k).
 */
typedef struct {
  float d;
  int32_t i;
} nsdp_pair;

static int pair_less(float da, int32_t ia, float db, int32_t ib) {
  return (da < db) || (da == db && ia < ib);
}

int nsdp_oracle_knn(const float *query, const float *ref, int B, int M, int N, int k,
                    int32_t *idx, float *d2out) {
  if (k > N || k <= 0) return -2;
  nsdp_pair *heap = (nsdp_pair *)malloc(sizeof(nsdp_pair) * (size_t)k);
  if (!heap) return -1;
  for (int b = 0; b < B; ++b) {
    const float *q = query + (size_t)b * M * 3;
    const float *r = ref + (size_t)b * N * 3;
    for (int i = 0; i < M; ++i) {
      const float qx = q[i * 3], qy = q[i * 3 + 1], qz = q[i * 3 + 2];
      int cnt = 0;
      for (int j = 0; j < N; ++j) {
        const float dx = qx - r[j * 3], dy = qy - r[j * 3 + 1], dz = qz - r[j * 3 + 2];
        const float xx = dx * dx, yy = dy * dy, zz = dz * dz;
        const float d = (xx + yy) + zz;
        /* sorted insertion, ascending (d, index) */
        if (cnt < k || pair_less(d, j, heap[cnt - 1].d, heap[cnt - 1].i)) {
          int pos = cnt < k ? cnt : k - 1;
          while (pos > 0 && pair_less(d, j, heap[pos - 1].d, heap[pos - 1].i)) {
            heap[pos] = heap[pos - 1];
            --pos;
          }
          heap[pos].d = d;
          heap[pos].i = j;
          if (cnt < k) ++cnt;
        }
      }
      for (int t = 0; t < k; ++t) {
        idx[((size_t)b * M + i) * k + t] = heap[t].i;
        if (d2out) d2out[((size_t)b * M + i) * k + t] = heap[t].d;
      }
    }
  }
  free(heap);
  return 0;
}

/*
 * Ball query — pointnet2_ops_lib/pointnet2_ops/_ext-src/src/ball_query_gpu.cu:9-44:
 * first `nsample` points (in index order) with d2 < radius^2, the row pre-filled with the
 * first hit; rows with no hit stay 0 (torch::zeros, ball_query.cpp). The distance uses
 * the contraction nvcc emits for `a*a + b*b + c*c`: fmaf(c,c, fmaf(a,a, b*b)).
 * new_xyz: (B, M, 3), xyz: (B, N, 3) -> idx: (B, M, nsample) int32.
 */
int nsdp_oracle_ball_query(const float *new_xyz, const float *xyz, int B, int N, int M,
                           float radius, int nsample, int32_t *idx) {
  const float radius2 = radius * radius;
  for (int b = 0; b < B; ++b) {
    const float *c = new_xyz + (size_t)b * M * 3;
    const float *p = xyz + (size_t)b * N * 3;
    int32_t *o = idx + (size_t)b * M * nsample;
    for (int j = 0; j < M; ++j) {
      for (int l = 0; l < nsample; ++l) o[j * nsample + l] = 0;
      const float nx = c[j * 3], ny = c[j * 3 + 1], nz = c[j * 3 + 2];
      int cnt = 0;
      for (int k = 0; k < N && cnt < nsample; ++k) {
        const float dx = nx - p[k * 3], dy = ny - p[k * 3 + 1], dz = nz - p[k * 3 + 2];
        const float d2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
        if (d2 < radius2) {
          if (cnt == 0)
            for (int l = 0; l < nsample; ++l) o[j * nsample + l] = k;
          o[j * nsample + cnt] = k;
          ++cnt;
        }
      }
    }
  }
  return 0;
}

/*
 * three_nn — pointnet2_ops_lib/pointnet2_ops/_ext-src/src/interpolate_gpu.cu:9-59:
 * running best three kept in double, strict '<' (earliest index wins ties), float d with
 * the same FMA contraction as above. unknown: (B, n, 3), known: (B, m, 3).
 */
int nsdp_oracle_three_nn(const float *unknown, const float *known, int B, int n, int m,
                         float *dist2, int32_t *idx) {
  for (int b = 0; b < B; ++b) {
    const float *u = unknown + (size_t)b * n * 3;
    const float *kn = known + (size_t)b * m * 3;
    for (int j = 0; j < n; ++j) {
      const float ux = u[j * 3], uy = u[j * 3 + 1], uz = u[j * 3 + 2];
      double best1 = 1e40, best2 = 1e40, best3 = 1e40;
      int b1 = 0, b2 = 0, b3 = 0;
      for (int k = 0; k < m; ++k) {
        const float dx = ux - kn[k * 3], dy = uy - kn[k * 3 + 1], dz = uz - kn[k * 3 + 2];
        const float d = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
        if (d < best1) {
          best3 = best2; b3 = b2; best2 = best1; b2 = b1; best1 = d; b1 = k;
        } else if (d < best2) {
          best3 = best2; b3 = b2; best2 = d; b2 = k;
        } else if (d < best3) {
          best3 = d; b3 = k;
        }
      }
      float *dd = dist2 + ((size_t)b * n + j) * 3;
      int32_t *ii = idx + ((size_t)b * n + j) * 3;
      dd[0] = (float)best1; dd[1] = (float)best2; dd[2] = (float)best3;
      ii[0] = b1; ii[1] = b2; ii[2] = b3;
    }
  }
  return 0;
}

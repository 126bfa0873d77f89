// Furthest point sampling for sm_100a — cluster-resident, bit-exact with the reference kernel.
//
// Replaces pointnet2_ops_lib/pointnet2_ops/_ext-src/src/sampling_gpu.cu:69-229 (+ sampling.cpp:66-87).
//
// Reference design: one 512-thread block per cloud; every iteration re-reads the whole cloud and a
// global `temp` array (12+8 B per point through L1/L2) and runs a 9-level __syncthreads tree.
// Here: the cloud and its running min-distance live in REGISTERS for all m iterations (P points per
// thread), spread over a thread-block cluster when one CTA's register file is too small; the
// arg-max is two redux.sync + one ballot per warp, one shared-memory exchange per CTA and one DSMEM
// exchange + cluster barrier per iteration. The winner carries its coordinates through the
// reduction, so the serial dependency never touches global memory.
//
// Bit-exactness contract:
//   * distance:  fmaf(dz,dz, fmaf(dx,dx, dy*dy)); the nvcc contraction of the reference expression,
//     pinned with intrinsics so that no compiler flag can change it;
//   * skip rule: points with !( (double)|p|^2 > 1e-3 )... precisely `mag <= 1e-3` in double are never
//     updated nor selected (sampling_gpu.cu:100); evaluated once, it is iteration-invariant;
//   * tie-break among equal maxima: the reference's per-thread strided pass (strict >, lowest k wins)
//     followed by its BS/2..1 tree (strict >, lower slot wins) selects the candidate minimising
//     (bitreverse_{log2 BS}(k mod BS), k div BS), BS = opt_n_threads(N) (cuda_utils.h:15-19). That
//     pair is packed into a 32-bit tie key; the arg-max runs on (distance bits + 1, ~tie key).
//   * no eligible point at all -> index 0 (threads contribute (-1, 0) in the reference).
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#ifndef NSDP_FPS_VARIANT_DEFAULT
#define NSDP_FPS_VARIANT_DEFAULT 0
#endif

#include "common.cuh"

namespace cg = cooperative_groups;

namespace nsdp {

struct __align__(16) FpsSlot {
  unsigned hi, lo;
  float x, y, z;
  unsigned pad0, pad1, pad2;
};

constexpr int kFpsMaxCluster = 16;

__device__ __forceinline__ void fps_pick(unsigned hi, unsigned lo, float x, float y, float z, unsigned &whi,
                                         unsigned &wlo, float &wx, float &wy, float &wz) {
  // warp arg-max on (hi, lo); lanes that lose still participate
  const unsigned full = 0xffffffffu;
  whi = __reduce_max_sync(full, hi);
  const unsigned cand = (hi == whi) ? lo : 0u;
  wlo = __reduce_max_sync(full, cand);
  const unsigned owners = __ballot_sync(full, hi == whi && lo == wlo);
  const int src = __ffs(owners) - 1;
  wx = __shfl_sync(full, x, src);
  wy = __shfl_sync(full, y, src);
  wz = __shfl_sync(full, z, src);
}

template <int P, int T>
__global__ void __launch_bounds__(T, 1)
fps_kernel(const float *__restrict__ xyz, int N, int m, int log2_bs, int C, int32_t *__restrict__ out) {
  constexpr int NW = T / 32;
  __shared__ FpsSlot wslot[2][NW];
  __shared__ FpsSlot cslot[2][kFpsMaxCluster];

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (C > 1) ? (int)cluster.block_rank() : 0;
  const int b = blockIdx.x / C;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const float *__restrict__ p = xyz + (size_t)b * N * 3;
  int32_t *__restrict__ o = out + (size_t)b * m;

  float px[P], py[P], pz[P], temp[P];
  unsigned inv[P];  // ~tiekey for eligible points, 0 otherwise
  const unsigned bs_mask = (1u << log2_bs) - 1u;
#pragma unroll
  for (int s = 0; s < P; ++s) {
    const int k = (s * C + rank) * T + tid;
    const bool valid = k < N;
    float x = 0.f, y = 0.f, z = 0.f;
    if (valid) {
      x = p[k * 3 + 0];
      y = p[k * 3 + 1];
      z = p[k * 3 + 2];
    }
    const float mag = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
    const bool elig = valid && !((double)mag <= 1e-3);
    const unsigned rev = log2_bs ? (__brev((unsigned)k & bs_mask) >> (32 - log2_bs)) : 0u;
    const unsigned tie = (rev << 20) | ((unsigned)k >> log2_bs);
    px[s] = x; py[s] = y; pz[s] = z;
    temp[s] = 1e10f;
    inv[s] = elig ? ~tie : 0u;
  }
  const float p0x = p[0], p0y = p[1], p0z = p[2];
  float ox = p0x, oy = p0y, oz = p0z;
  if (rank == 0 && tid == 0) o[0] = 0;

  for (int j = 1; j < m; ++j) {
    const int par = j & 1;
    unsigned bhi = 0u, blo = 0u;
    float bx = 0.f, by = 0.f, bz = 0.f;
#pragma unroll
    for (int s = 0; s < P; ++s) {
      const float dx = px[s] - ox, dy = py[s] - oy, dz = pz[s] - oz;
      const float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
      const float t = fminf(d, temp[s]);
      if (inv[s]) {
        temp[s] = t;
        const unsigned hi = __float_as_uint(t) + 1u;
        if (hi > bhi || (hi == bhi && inv[s] > blo)) {
          bhi = hi; blo = inv[s];
          bx = px[s]; by = py[s]; bz = pz[s];
        }
      }
    }
    unsigned whi, wlo;
    float wx, wy, wz;
    fps_pick(bhi, blo, bx, by, bz, whi, wlo, wx, wy, wz);
    if (lane == 0) {
      FpsSlot s;
      s.hi = whi; s.lo = wlo; s.x = wx; s.y = wy; s.z = wz; s.pad0 = s.pad1 = s.pad2 = 0u;
      wslot[par][warp] = s;
    }
    __syncthreads();
    {
      unsigned hi = 0u, lo = 0u;
      float x = 0.f, y = 0.f, z = 0.f;
      if (lane < NW) {
        const FpsSlot s = wslot[par][lane];
        hi = s.hi; lo = s.lo; x = s.x; y = s.y; z = s.z;
      }
      fps_pick(hi, lo, x, y, z, whi, wlo, wx, wy, wz);
    }
    if (C > 1) {
      if (warp == 0 && lane < C) {
        FpsSlot s;
        s.hi = whi; s.lo = wlo; s.x = wx; s.y = wy; s.z = wz; s.pad0 = s.pad1 = s.pad2 = 0u;
        FpsSlot *remote = cluster.map_shared_rank(&cslot[par][rank], lane);
        *remote = s;
      }
      cluster.sync();
      unsigned hi = 0u, lo = 0u;
      float x = 0.f, y = 0.f, z = 0.f;
      if (lane < C) {
        const FpsSlot s = cslot[par][lane];
        hi = s.hi; lo = s.lo; x = s.x; y = s.y; z = s.z;
      }
      fps_pick(hi, lo, x, y, z, whi, wlo, wx, wy, wz);
    }
    int win = 0;
    if (whi != 0u) {
      const unsigned tie = ~wlo;
      const unsigned rev = log2_bs ? (__brev(tie >> 20) >> (32 - log2_bs)) : 0u;
      win = (int)(((tie & 0xFFFFFu) << log2_bs) | rev);
      ox = wx; oy = wy; oz = wz;
    } else {
      ox = p0x; oy = p0y; oz = p0z;
    }
    if (rank == 0 && tid == 0) o[j] = win;
  }
  // keep every CTA of the cluster alive until no peer can still write into its shared memory
  if (C > 1) cluster.sync();
}

template <int P, int T>
static int launch_fps(const float *xyz, int B, int N, int m, int log2_bs, int C, int32_t *out, cudaStream_t st) {
  auto kern = fps_kernel<P, T>;
  if (C > 8) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return cuda_rc(e);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * C));
  cfg.blockDim = dim3(T);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, xyz, N, m, log2_bs, C, out);
  if (e != cudaSuccess) return cuda_rc(e);
  return check_launch();
}

// cuda_utils.h:15-19 — the reference block size decides the tie-break order, so it is reproduced
// with the same double-precision log quotient.
static int ref_block_log2(int n) {
  int pow_2 = (int)(log((double)n) / log(2.0));
  if (pow_2 > 9) pow_2 = 9;
  if (pow_2 < 0) pow_2 = 0;
  return pow_2;
}

}  // namespace nsdp

extern "C" int nsdp_fps_f32(const float *xyz, int B, int N, int m, int32_t *out_idx, void *stream) {
  using namespace nsdp;
  if (!xyz || !out_idx || B <= 0 || N <= 0 || m < 0) return NSDP_ERR_INVALID_ARGUMENT;
  if (m == 0) return NSDP_OK;
  if ((long long)N >= (1ll << 28)) return NSDP_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int lb = ref_block_log2(N);
  if (N <= 128) return launch_fps<1, 128>(xyz, B, N, m, lb, 1, out_idx, st);
  if (N <= 256) return launch_fps<2, 128>(xyz, B, N, m, lb, 1, out_idx, st);
  if (N <= 512) return launch_fps<4, 128>(xyz, B, N, m, lb, 1, out_idx, st);
  if (N <= 1024) return launch_fps<2, 512>(xyz, B, N, m, lb, 1, out_idx, st);
  if (N <= 2048) return launch_fps<4, 512>(xyz, B, N, m, lb, 1, out_idx, st);
  if (N <= 4096) {
    // per-iteration cost at this size is synchronisation latency, not arithmetic: fewer, fatter warps win (A/B on B200,
    // tools/microbench_fps.py; NSDP_FPS_VARIANT picks the others)
    static const int variant = [] { const char *e = getenv("NSDP_FPS_VARIANT"); return e ? atoi(e) : NSDP_FPS_VARIANT_DEFAULT; }();
    switch (variant) {
      case 1: return launch_fps<16, 256>(xyz, B, N, m, lb, 1, out_idx, st);
      case 2: return launch_fps<32, 128>(xyz, B, N, m, lb, 1, out_idx, st);
      case 3: return launch_fps<4, 512>(xyz, B, N, m, lb, 2, out_idx, st);
      case 4: return launch_fps<2, 512>(xyz, B, N, m, lb, 4, out_idx, st);
      case 5: return launch_fps<8, 128>(xyz, B, N, m, lb, 4, out_idx, st);
      default: return launch_fps<8, 512>(xyz, B, N, m, lb, 1, out_idx, st);
    }
  }
  if (N <= 8192) return launch_fps<16, 512>(xyz, B, N, m, lb, 1, out_idx, st);
  int C = 2;
  while (C < kFpsMaxCluster && (long long)C * 8192 < N) C *= 2;
  if ((long long)C * 8192 < N) return NSDP_ERR_UNSUPPORTED;
  const int per = ceil_div(N, C * 512);
  if (per <= 4) return launch_fps<4, 512>(xyz, B, N, m, lb, C, out_idx, st);
  if (per <= 8) return launch_fps<8, 512>(xyz, B, N, m, lb, C, out_idx, st);
  return launch_fps<16, 512>(xyz, B, N, m, lb, C, out_idx, st);
}

// Adam step of torch.optim.Adam (amsgrad = False, maximize = False) over MANY parameter tensors in two launches.
//
// torch's fused Adam (multi_tensor_apply) walks the parameter list in launches of a few dozen tensors and gives every
// thread block one 64 K-element chunk: for this model (505 tensors, 4.5 M parameters) that is 7 launches of ~70 blocks,
// each block streaming 1.8 MB by itself -- 0.38 ms per step, 15x the time the 126 MB of traffic need. Here ALL tensor
// pointers travel in ONE kernel-parameter block (CUDA 12.1+: up to 32 764 bytes; 5 pointers + 2 ints per tensor) and the
// grid has one block per 4096-element chunk, found by binary search over the chunk prefix sums.
//
// Numerics follow torch/optim/_functional / FusedAdamKernel for capturable = True: the step count is a float32 tensor per
// parameter (kept, so that optimizer.state_dict() stays interchangeable with the reference's), incremented by a first tiny
// kernel; then with s = step:
//   g' = g + weight_decay * p;  m = m + (g' - m) * (1 - beta1);  v = beta2 * v + (1 - beta2) * g' * g'
//   p  = p - (lr / (1 - beta1^s)) * m / (sqrt(v) / sqrt(1 - beta2^s) + eps)
#include <math.h>

#include "common.cuh"

namespace nsdp {
namespace adam {

constexpr int MAX_T = 640;
constexpr int CHUNK = 4096;
constexpr int THREADS = 256;

struct Table {
  float *p[MAX_T];
  const float *g[MAX_T];
  float *m[MAX_T];
  float *v[MAX_T];
  float *step[MAX_T];
  int n[MAX_T];
  int chunk_begin[MAX_T + 1];
  int ntensors;
  double lr, beta1, beta2, eps, weight_decay;   // Python doubles, as torch holds them: 1 - beta and beta^step are formed in double
};
static_assert(sizeof(Table) <= 32764, "kernel parameter block");

__global__ void bump_steps_kernel(const __grid_constant__ Table t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < t.ntensors) *t.step[i] += 1.f;
}

__global__ void __launch_bounds__(THREADS) adam_kernel(const __grid_constant__ Table t) {
  // tensor of this block's chunk: last index with chunk_begin <= blockIdx.x
  int lo = 0, hi = t.ntensors;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (t.chunk_begin[mid] <= (int)blockIdx.x) lo = mid; else hi = mid;
  }
  const int ti = lo;
  const int base = ((int)blockIdx.x - t.chunk_begin[ti]) * CHUNK;
  const int n = t.n[ti];
  float *__restrict__ p = t.p[ti];
  const float *__restrict__ g = t.g[ti];
  float *__restrict__ m = t.m[ti];
  float *__restrict__ v = t.v[ti];
  const double s = (double)*t.step[ti];
  const double bc1 = 1.0 - pow(t.beta1, s), bc2 = 1.0 - pow(t.beta2, s);
  const float step_size = (float)(t.lr / bc1), bc2_sqrt = (float)sqrt(bc2);
  const float omb1 = (float)(1.0 - t.beta1), omb2 = (float)(1.0 - t.beta2), b2 = (float)t.beta2;
  const float wd = (float)t.weight_decay, eps = (float)t.eps;
  auto update = [&](float &pp, float gg, float &mm, float &vv) {
    gg = fmaf(wd, pp, gg);
    mm = fmaf(gg - mm, omb1, mm);
    vv = fmaf(omb2 * gg, gg, b2 * vv);
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pp -= step_size * mm / denom;
  };
  const int end = base + CHUNK < n ? base + CHUNK : n;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  if (vec) {
    for (int i = base + 4 * (int)threadIdx.x; i + 3 < end; i += 4 * THREADS) {
      float4 p4 = *reinterpret_cast<float4 *>(p + i), m4 = *reinterpret_cast<float4 *>(m + i), v4 = *reinterpret_cast<float4 *>(v + i);
      const float4 g4 = *reinterpret_cast<const float4 *>(g + i);
      update(p4.x, g4.x, m4.x, v4.x);
      update(p4.y, g4.y, m4.y, v4.y);
      update(p4.z, g4.z, m4.z, v4.z);
      update(p4.w, g4.w, m4.w, v4.w);
      *reinterpret_cast<float4 *>(p + i) = p4;
      *reinterpret_cast<float4 *>(m + i) = m4;
      *reinterpret_cast<float4 *>(v + i) = v4;
    }
    const int tail = base + ((end - base) & ~3);
    for (int i = tail + (int)threadIdx.x; i < end; i += THREADS) update(p[i], g[i], m[i], v[i]);
  } else {
    for (int i = base + (int)threadIdx.x; i < end; i += THREADS) update(p[i], g[i], m[i], v[i]);
  }
}

}  // namespace adam
}  // namespace nsdp

extern "C" int nsdp_adam_step_f32(int ntensors, float *const *p, const float *const *g, float *const *m, float *const *v,
                                  float *const *step, const long long *numel, double lr, double beta1, double beta2, double eps,
                                  double weight_decay, void *stream) {
  using namespace nsdp;
  using namespace nsdp::adam;
  if (ntensors < 0 || (ntensors > 0 && (!p || !g || !m || !v || !step || !numel))) return NSDP_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  for (int t0 = 0; t0 < ntensors; t0 += MAX_T) {
    Table t;
    const int nt = ntensors - t0 < MAX_T ? ntensors - t0 : MAX_T;
    long long chunks = 0;
    for (int i = 0; i < nt; ++i) {
      const long long n = numel[t0 + i];
      if (n <= 0 || n > 0x7fffffffll || !p[t0 + i] || !g[t0 + i] || !m[t0 + i] || !v[t0 + i] || !step[t0 + i])
        return NSDP_ERR_INVALID_ARGUMENT;
      t.p[i] = p[t0 + i]; t.g[i] = g[t0 + i]; t.m[i] = m[t0 + i]; t.v[i] = v[t0 + i]; t.step[i] = step[t0 + i];
      t.n[i] = (int)n;
      t.chunk_begin[i] = (int)chunks;
      chunks += (n + CHUNK - 1) / CHUNK;
      if (chunks > 0x7fffffffll) return NSDP_ERR_UNSUPPORTED;
    }
    t.chunk_begin[nt] = (int)chunks;
    t.ntensors = nt;
    t.lr = lr; t.beta1 = beta1; t.beta2 = beta2; t.eps = eps; t.weight_decay = weight_decay;
    bump_steps_kernel<<<(nt + 127) / 128, 128, 0, st>>>(t);
    int rc = check_launch();
    if (rc != NSDP_OK) return rc;
    adam_kernel<<<(unsigned)chunks, THREADS, 0, st>>>(t);
    rc = check_launch();
    if (rc != NSDP_OK) return rc;
  }
  return NSDP_OK;
}

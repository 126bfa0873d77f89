// Brute-force k-nearest-neighbour selection for sm_100a.
//
// Replaces the reference's `square_distance(q, ref).argsort()[:, :, :k]`
// (model/utils.py:39-55; model/encoder/blocks.py:101-102, 287-288; model/decoder/blocks.py:50-52), which
// materialises a [B,M,N,3] temporary and a [B,M,N] distance matrix and fully sorts every row.
// Here nothing of size M*N ever exists: one thread owns one query and keeps its k best (distance, index)
// pairs sorted in registers while the reference cloud streams through shared memory (every lane reads
// the same reference point -> one broadcast LDS per coordinate). For large N the reference range is
// split over S threads per query (more parallelism than M alone provides) and a second tiny kernel
// merges the S sorted partial lists.
//
// Insertion into the sorted register list costs ~3 instructions per slot and — one query per thread — the whole warp pays
// for it whenever ANY lane has a candidate, which early in a scan is almost every step (P = min(1, 32 k / t) at step t: at
// the model's 4096 x 4096 / k = 10 site insertion, not distance evaluation, was 4/5 of the kernel). Candidates that beat
// the thread's threshold (its k-th best distance as of the last flush) are therefore BUFFERED in shared memory and the lists
// are updated in batches, when some lane's buffer fills up: the same insertions in the same order (a stale threshold only
// admits extra candidates, which the insertion re-checks), a fraction of the divergent passes.
//
// Parity contract: distance = ((dx*dx + dy*dy) + dz*dz) with dx = query - ref, every operation rounded
// separately (no FMA) exactly like torch's elementwise kernels; order = ascending (distance, index), i.e.
// the stable order torch's unstable argsort leaves undefined on ties. Index tensors are int32.
#include <float.h>

#include <stdlib.h>

#include "common.cuh"

namespace nsdp {

constexpr int kKnnThreads = 128;
constexpr int kKnnChunk = 1024;  // reference points staged per shared-memory tile (12 KB)
constexpr int kKnnBuf = 8;       // buffered candidates per thread (8 KB of shared memory per block)
constexpr int kKnnStep = 4;      // reference points between two "is some buffer nearly full" votes

template <int KMAX>
__device__ __forceinline__ void knn_insert(float (&dist)[KMAX], int (&id)[KMAX], float d, int j) {
#pragma unroll
  for (int i = KMAX - 1; i > 0; --i) {
    const bool shift = d < dist[i - 1];
    const bool here = !shift && d < dist[i];
    dist[i] = shift ? dist[i - 1] : (here ? d : dist[i]);
    id[i] = shift ? id[i - 1] : (here ? j : id[i]);
  }
  if (d < dist[0]) {
    dist[0] = d;
    id[0] = j;
  }
}

// grid = (ceil(M / T), S, B)
template <int KMAX>
__global__ void __launch_bounds__(kKnnThreads)
knn_scan_kernel(const float *__restrict__ query, const float *__restrict__ ref, int M, int N, int k, int S,
                int split_len, int32_t *__restrict__ out_idx, float *__restrict__ out_d2,
                float *__restrict__ part_d, int32_t *__restrict__ part_i) {
  __shared__ float tile[kKnnChunk * 3];
  __shared__ float buf_d[kKnnBuf][kKnnThreads];
  __shared__ int buf_i[kKnnBuf][kKnnThreads];
  const int b = blockIdx.z, s = blockIdx.y;
  const int qi = blockIdx.x * kKnnThreads + threadIdx.x;
  const bool active = qi < M;
  const float *__restrict__ r = ref + (size_t)b * N * 3;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (active) {
    const float *q = query + ((size_t)b * M + qi) * 3;
    qx = q[0]; qy = q[1]; qz = q[2];
  }
  float dist[KMAX];
  int id[KMAX];
#pragma unroll
  for (int i = 0; i < KMAX; ++i) {
    dist[i] = FLT_MAX;
    id[i] = 0x7fffffff;
  }
  int nbuf = 0;              // candidates waiting in this thread's column of the buffer
  float tau = FLT_MAX;       // k-th best distance as of the last flush
  auto flush = [&]() {
    const int most = __reduce_max_sync(0xffffffffu, nbuf);
    for (int c = 0; c < most; ++c) {
      if (c < nbuf) {
        const float d = buf_d[c][threadIdx.x];
        if (d < dist[KMAX - 1]) knn_insert<KMAX>(dist, id, d, buf_i[c][threadIdx.x]);
      }
    }
    nbuf = 0;
    tau = dist[KMAX - 1];
  };
  // +inf distances must still be selectable when N is tiny: FLT_MAX sentinels lose against any finite d,
  // and genuine inf/NaN distances are never inserted (same as "sorted last"). A query or cloud with non-finite
  // coordinates (a diverging canonicaliser in FlowArbitrary feeds network OUTPUTS into this search) therefore leaves
  // slots unfilled: they are reported as index min(slot, N-1) with a NaN distance, so downstream gathers stay in bounds
  // and the NaNs propagate numerically, as they would through the reference's argsort.
  const int j0 = s * split_len;
  const int j1 = min(N, j0 + split_len);
  for (int base = j0; base < j1; base += kKnnChunk) {
    const int cnt = min(kKnnChunk, j1 - base);
    __syncthreads();
    for (int t = threadIdx.x; t < cnt * 3; t += kKnnThreads) tile[t] = r[(size_t)base * 3 + t];
    __syncthreads();
    // whole warps run the loop (inactive lanes never produce candidates): the flush votes are warp-wide
    for (int t0 = 0; t0 < cnt; t0 += kKnnStep) {
      if (__any_sync(0xffffffffu, nbuf > kKnnBuf - kKnnStep)) flush();
#pragma unroll
      for (int u = 0; u < kKnnStep; ++u) {
        const int t = t0 + u;
        if (t < cnt) {
          const float dx = __fsub_rn(qx, tile[t * 3 + 0]);
          const float dy = __fsub_rn(qy, tile[t * 3 + 1]);
          const float dz = __fsub_rn(qz, tile[t * 3 + 2]);
          const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
          if (active && d < tau) {
            buf_d[nbuf][threadIdx.x] = d;
            buf_i[nbuf][threadIdx.x] = base + t;
            ++nbuf;
          }
        }
      }
    }
  }
  flush();
  if (!active) return;
  if (S == 1) {
    int32_t *oi = out_idx + ((size_t)b * M + qi) * k;
#pragma unroll
    for (int i = 0; i < KMAX; ++i)
      if (i < k) oi[i] = id[i] < N ? id[i] : min(i, N - 1);   // unfilled slot (non-finite distances): a VALID index, see below
    if (out_d2) {
      float *od = out_d2 + ((size_t)b * M + qi) * k;
#pragma unroll
      for (int i = 0; i < KMAX; ++i)
        if (i < k) od[i] = id[i] < N ? dist[i] : __int_as_float(0x7fc00000);
    }
  } else {
    const size_t off = (((size_t)b * M + qi) * S + s) * k;
#pragma unroll
    for (int i = 0; i < KMAX; ++i)
      if (i < k) {
        part_d[off + i] = dist[i];
        part_i[off + i] = id[i];
      }
  }
}

// S-way merge of sorted partial lists; one thread per query.
__global__ void knn_merge_kernel(const float *__restrict__ part_d, const int32_t *__restrict__ part_i, long long BM,
                                 int k, int S, int N, int32_t *__restrict__ out_idx, float *__restrict__ out_d2) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= BM) return;
  const float *pd = part_d + (size_t)q * S * k;
  const int32_t *pi = part_i + (size_t)q * S * k;
  unsigned char head[64];
  for (int s = 0; s < S; ++s) head[s] = 0;
  for (int t = 0; t < k; ++t) {
    float bd = FLT_MAX;
    int bi = 0x7fffffff, bs = 0;
    for (int s = 0; s < S; ++s) {
      if (head[s] >= k) continue;
      const float d = pd[s * k + head[s]];
      const int i = pi[s * k + head[s]];
      if (d < bd || (d == bd && i < bi)) {
        bd = d; bi = i; bs = s;
      }
    }
    head[bs]++;
    out_idx[(size_t)q * k + t] = bi < N ? bi : min(t, N - 1);
    if (out_d2) out_d2[(size_t)q * k + t] = bi < N ? bd : __int_as_float(0x7fc00000);
  }
}

static int knn_splits(int B, int M, int N) {
  static const int forced = [] { const char *e = getenv("NSDP_KNN_SPLITS"); return e ? atoi(e) : 0; }();
  if (forced >= 1 && forced <= 64 && N / forced >= 64) return forced;
  const long long queries = (long long)B * M;
  const long long want = 148ll * 1024;  // ~8 warps per SMSP of scanning threads
  long long S = (want + queries - 1) / queries;
  const long long max_by_n = N / 512 > 0 ? N / 512 : 1;
  if (S > max_by_n) S = max_by_n;
  if (S > 64) S = 64;
  if (S < 1) S = 1;
  return (int)S;
}

}  // namespace nsdp

extern "C" size_t nsdp_knn_workspace_bytes(int B, int M, int N, int k) {
  if (B <= 0 || M <= 0 || N <= 0 || k <= 0) return 0;
  const int S = nsdp::knn_splits(B, M, N);
  if (S == 1) return 0;
  return (size_t)B * M * S * k * (sizeof(float) + sizeof(int32_t));
}

extern "C" int nsdp_knn_f32(const float *query, const float *ref, int B, int M, int N, int k, int32_t *out_idx,
                            float *out_d2, void *workspace, size_t workspace_bytes, void *stream) {
  using namespace nsdp;
  if (!query || !ref || !out_idx || B <= 0 || M <= 0 || N <= 0 || k <= 0 || k > N) return NSDP_ERR_INVALID_ARGUMENT;
  if (k > 64 || B > 65535) return NSDP_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int S = knn_splits(B, M, N);
  float *part_d = nullptr;
  int32_t *part_i = nullptr;
  if (S > 1) {
    const size_t need = nsdp_knn_workspace_bytes(B, M, N, k);
    if (!workspace || workspace_bytes < need) return NSDP_ERR_WORKSPACE;
    part_d = (float *)workspace;
    part_i = (int32_t *)(part_d + (size_t)B * M * S * k);
  }
  const int split_len = ceil_div(N, S);
  dim3 grid((unsigned)ceil_div(M, kKnnThreads), (unsigned)S, (unsigned)B);
  if (k <= 8)
    knn_scan_kernel<8><<<grid, kKnnThreads, 0, st>>>(query, ref, M, N, k, S, split_len, out_idx, out_d2, part_d, part_i);
  else if (k <= 16)
    knn_scan_kernel<16><<<grid, kKnnThreads, 0, st>>>(query, ref, M, N, k, S, split_len, out_idx, out_d2, part_d, part_i);
  else if (k <= 32)
    knn_scan_kernel<32><<<grid, kKnnThreads, 0, st>>>(query, ref, M, N, k, S, split_len, out_idx, out_d2, part_d, part_i);
  else
    knn_scan_kernel<64><<<grid, kKnnThreads, 0, st>>>(query, ref, M, N, k, S, split_len, out_idx, out_d2, part_d, part_i);
  int rc = check_launch();
  if (rc != NSDP_OK) return rc;
  if (S > 1) {
    const long long BM = (long long)B * M;
    knn_merge_kernel<<<(unsigned)ceil_div(BM, 128ll), 128, 0, st>>>(part_d, part_i, BM, k, S, N, out_idx, out_d2);
    rc = check_launch();
  }
  return rc;
}

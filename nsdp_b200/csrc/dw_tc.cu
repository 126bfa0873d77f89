// Weight-gradient products on tcgen05:  out[m][n] += sum_r X[r][m] * Y[r][n]   ("TN" GEMM, reduction over pair rows).
//
// The backward chain kernels (vattn_bwd_tc.cu, resnet_tail_bwd_tc.cu) cannot keep the d x d weight-gradient
// accumulators on chip next to their own TMEM accumulators (3 x 208 x 208 fp32 = 520 KB per CTA for the decoder), so
// they STAGE the operand tiles (bf16 hi/lo, 4 B/element) and this kernel reduces them: every CTA takes one job and a
// contiguous range of row tiles (split-K), accumulates a [256 x N] fp32 partial entirely in TMEM (2 M-tiles, up to
// 2 x 256 columns) and adds it to the global gradient once at the end.
//
// Staged layout of a [128 rows x W cols] tile ("k-step major", written by the chain kernels' epilogues):
//     byte(r, c) = (r / 16) * (2 * W * 32) + [hi: 0 | lo: W * 32] + (c / 8) * 256 + (r % 16) * 16 + (c % 8) * 2
// i.e. per k-step (16 rows) one contiguous [hi slab][lo slab] pair; inside a slab the 8 contiguous elements run
// along the OUTPUT dimension, so the slab is consumed directly as an MN-major operand (LBO = 128, SBO = 256).
#include <stdlib.h>

#include "common.cuh"
#include "dw_tc.cuh"
#include "stage_f16.cuh"
#include "umma.cuh"

namespace nsdp {
namespace dwtc {

using namespace umma;

constexpr int MAX_STAGES = 8;   // ring depth is chosen per launch: as many stages as fit in shared memory (HBM latency)
constexpr int THREADS = 6 * 32;  // warp 0 producer, warp 1 MMA, warps 2..5 flush
constexpr int MAX_JOBS = 24;
constexpr int MAX_CHUNKS = 64;

struct Params {
  Job jobs[MAX_JOBS];
  int njobs;
  int cta_begin[MAX_JOBS + 1];   // CTAs [cta_begin[j], cta_begin[j+1]) split the tile range of job j
  int nchunks;                   // > 0: chunk-aligned mode, CTA = (chunk blockIdx / njobs, job blockIdx % njobs)
  int chunk_t0[MAX_CHUNKS + 1];
  int chunk_shape[MAX_CHUNKS];
  int stages;              // ring depth (<= MAX_STAGES)
  int group;               // k-steps (of 16 rows) per ring stage: 1, 2, 4 or 8. > 1 only when every job streams exactly what
                           // is staged (no skipped lo slabs), so that consecutive k-steps of a tile are contiguous in memory:
                           // fewer, larger bulk copies and barrier round trips per byte (fp16 staging halved the bytes per
                           // k-step and left the reduction issue-bound at 48 % of the HBM rate; measured, profiles/)
  uint32_t ones_off;       // offset of the constant one-hot operand (column sums by MMA) inside dynamic shared memory
  int terms;               // 3: xh*yh + xl*yh + xh*yl (fp32-grade) ; 2: xh*yh + xl*yh ; 1: xh*yh (plain bf16 operands)
  int *err;
};

__global__ void __launch_bounds__(THREADS, 1) dw_tc_kernel(const Params p) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[MAX_STAGES], empty[MAX_STAGES], acc_done;
  const int STAGES = p.stages;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int jid = 0;
  long long t0, t1;
  Job job;
  if (p.nchunks > 0) {
    const int c = (int)blockIdx.x / p.njobs;
    jid = (int)blockIdx.x - c * p.njobs;
    job = p.jobs[jid];
    t0 = p.chunk_t0[c];
    t1 = p.chunk_t0[c + 1];
    job.out += (long long)p.chunk_shape[c] * job.out_shape_stride;
  } else {
    while (jid + 1 < p.njobs && (int)blockIdx.x >= p.cta_begin[jid + 1]) ++jid;
    const int splits = p.cta_begin[jid + 1] - p.cta_begin[jid], split = (int)blockIdx.x - p.cta_begin[jid];
    job = p.jobs[jid];
    const long long jt = job.t1 - job.t0;
    t0 = job.t0 + jt * split / splits;
    t1 = job.t0 + jt * (split + 1) / splits;
  }
  const uint32_t xs = (uint32_t)job.wx * 32, ys = (uint32_t)job.wy * 32;  // slab bytes
  const uint32_t xk = job.x_lo ? 2 * xs : xs, yk = job.y_lo ? 2 * ys : ys;   // k-step pitch inside a staged tile
  const bool x_lo = job.x_lo && p.terms >= 2, y_lo = job.y_lo && p.terms >= 3;   // which lo slabs are streamed at all
  const uint32_t xb = x_lo ? 2 * xs : xs, yb = y_lo ? 2 * ys : ys;
  const int G = p.group;
  const uint32_t stage_bytes = (uint32_t)G * (xb + yb);   // one ring stage: G k-steps of X, then G k-steps of Y
  const int mtiles = job.wx > 128 ? 2 : 1;
  // Column sums (bias gradients) by the tensor core: colsum[n] = sum_r ONES[r][0] * Y[r][n] with the constant operand
  // ONES[r][m] = (m == 0), accumulated in a spare TMEM region whose row 0 is flushed at the end. The warp-level path
  // below only remains for the per-batch sums (gsum) and for shapes without a spare region.
  const int cs_col = job.wy <= 128 ? 128 : (mtiles == 1 ? 256 : -1);
  const bool mma_cs = job.colsum != nullptr && job.gsum == nullptr && cs_col >= 0;
  unsigned char *ones = smem + p.ones_off;
  if (warp >= 2) {
    const int t = tid - 64;   // 128 threads x 32 bytes
    *reinterpret_cast<uint4 *>(ones + t * 32) = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4 *>(ones + t * 32 + 16) = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
    if (t < 8) {   // rows 2t, 2t+1 of column group 0: element (r, 0) = 1.0 (bf16, or fp16 for fp16-staged jobs)
      const uint32_t one = job.f16 ? 0x00003C00u : 0x00003F80u;
      *reinterpret_cast<uint32_t *>(ones + t * 32) = one;
      *reinterpret_cast<uint32_t *>(ones + t * 32 + 16) = one;
    }
    fence_async_smem();
  }

  // the four flush warps also read the slabs in the ring only for warp-level column sums (bias gradients without a spare
  // TMEM region, per-batch sums); otherwise they stay out of the ring protocol altogether
  const bool ring_warps = (job.colsum != nullptr && !mma_cs) || job.gsum != nullptr;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], ring_warps ? 5 : 1);  // MMA commit (+ the four column-sum warps)
    }
    mbar_init(&acc_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const bool has_work = t1 > t0;

  if (warp == 0) {
    // Producer: PL lanes issue the bulk copies, lane l serving k-steps l, l + PL, ... (one thread sustains only about
    // one copy per ~500 cycles: the wait / expect_tx / issue chain is serial; measured with tools/micro/stream_bw.cu)
    const int PL = STAGES >= 4 ? STAGES / 2 : 1;
    if (lane < PL) {
      const int spt = 8 / G;                      // stages per tile
      const long long total = (t1 - t0) * spt;
      for (long long it = lane; it < total; it += PL) {
        const long long t = t0 + it / spt;
        const int ks = (int)(it % spt) * G;       // first k-step of the stage
        const int s = (int)(it % STAGES);
        const uint32_t ph = (uint32_t)(it / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1, p.err);
        mbar_arrive_expect_tx(&full[s], stage_bytes);
        unsigned char *dst = smem + (size_t)s * stage_bytes;
        // G > 1 implies xk == xb and yk == yb: the G k-steps are one contiguous run
        bulk_g2s(dst, job.x + (size_t)t * 8 * xk + (size_t)ks * xk, (uint32_t)G * xb, &full[s]);
        bulk_g2s(dst + (size_t)G * xb, job.y + (size_t)t * 8 * yk + (size_t)ks * yk, (uint32_t)G * yb, &full[s]);
      }
    }
  } else if (warp == 1) {
    // MMA issuer: whole warp runs the loop and the waits, one elected lane issues (elect_one: straight UTCHMMA issue);
    // descriptors advance by adds, the ring position is a running counter
    if (has_work) {
      // kind::f16 instruction descriptor: operand formats in bits [7,10) / [10,13), 1 = bf16, 0 = fp16
      const uint32_t idesc = job.f16 ? (idesc_bf16_mn(128, job.wy) & ~((1u << 7) | (1u << 10))) : idesc_bf16_mn(128, job.wy);
      const uint64_t x0 = smem_desc(smem_u32(smem), 128, 256);
      const uint64_t ones_desc = smem_desc(smem_u32(ones), 128, 256);
      const uint32_t stage16 = stage_bytes >> 4;
      uint32_t slot = 0, slot_phase = 0;
      bool first = true;
      const long long total = (t1 - t0) * (8 / G);
      for (long long it = 0; it < total; ++it) {
        mbar_wait(&full[slot], slot_phase, p.err);
        tc_fence_after();
        if (elect_one()) {
          for (int gk = 0; gk < G; ++gk) {
            const uint64_t xh0 = x0 + (uint64_t)slot * stage16 + (uint64_t)gk * (xb >> 4);
            const uint64_t yh = x0 + (uint64_t)slot * stage16 + (uint64_t)G * (xb >> 4) + (uint64_t)gk * (yb >> 4), yl = yh + (ys >> 4);
            const bool acc = !first || gk > 0;
            for (int j = 0; j < mtiles; ++j) {
              // M-tile j = output rows [128j, 128j+128) = column groups 16j.. of the X slab (4096 bytes further)
              const uint64_t xh = xh0 + (uint64_t)j * (4096 >> 4), xl = xh + (xs >> 4);
              const uint32_t d = tmem_base + j * 256;
              mma_bf16(d, xh, yh, idesc, acc);
              if (x_lo) mma_bf16(d, xl, yh, idesc, true);
              if (y_lo) mma_bf16(d, xh, yl, idesc, true);
            }
            if (mma_cs) {
              mma_bf16(tmem_base + cs_col, ones_desc, yh, idesc, acc);
              if (y_lo) mma_bf16(tmem_base + cs_col, ones_desc, yl, idesc, true);
            }
          }
          mma_commit(&empty[slot]);
        }
        first = false;
        if (++slot == (uint32_t)STAGES) { slot = 0; slot_phase ^= 1; }
      }
      if (elect_one()) mma_commit(&acc_done);
    }
  } else if (has_work) {
    if (ring_warps) {
      // Column sums of Y straight from the staged slabs as they pass through shared memory. Lane = column group of 8;
      // the four warps split the 16 rows of a k-step. All rows -> bias gradients (colsum); a periodic subset of rows
      // (row % g_kr == g_kr - 1, g_kr a power of two) -> per-batch sums (gsum, the decoder's global-token rows).
      const int w4 = warp - 2;
      const float osc = job.gmax ? 1.f / stage16::scale_from_max(*job.gmax) : 1.f;
      float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      float gs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      int gb = -1;   // batch the running gs[] belongs to
      const bool mine = lane * 8 < job.wy;
      const bool want_cs = job.colsum != nullptr && !mma_cs, want_gs = job.gsum != nullptr;
      const unsigned kr_mask = want_gs ? (unsigned)job.g_kr - 1u : 0u;
      const int kr_shift = want_gs ? 31 - __clz(job.g_kr) : 0;
      auto flush_gs = [&]() {
        if (gb >= 0 && mine) {
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (lane * 8 + u < job.nv) atomicAdd(job.gsum + (size_t)gb * job.nv + lane * 8 + u, gs[u] * osc);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) gs[u] = 0.f;
      };
      uint32_t it = 0;
      for (long long t = t0; t < t1; ++t) {
        for (int ks0 = 0; ks0 < 8; ks0 += G, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full[s], ph, p.err);
          for (int gk = 0; gk < G; ++gk) {
          const int ks = ks0 + gk;
          if ((want_cs || want_gs) && mine) {
            const unsigned char *yh = smem + (size_t)s * stage_bytes + (size_t)G * xb + (size_t)gk * yb + (size_t)lane * 256;
            const unsigned char *yl = yh + ys;
            const unsigned row0 = (unsigned)(t * 128) + (unsigned)(ks * 16);   // row inside the segment (< 2^31)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int rr = w4 + 4 * q;
              const bool grow = want_gs && (((row0 + rr) & kr_mask) == kr_mask);
              if (want_cs || grow) {
                const uint4 h4 = *reinterpret_cast<const uint4 *>(yh + rr * 16);
                const uint4 l4 = y_lo ? *reinterpret_cast<const uint4 *>(yl + rr * 16) : make_uint4(0u, 0u, 0u, 0u);
                const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
                float v[8];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  if (job.f16) {
                    const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&hw[u]));
                    v[2 * u] = f.x;
                    v[2 * u + 1] = f.y;
                  } else {
                    v[2 * u] = __uint_as_float(hw[u] << 16) + __uint_as_float(lw[u] << 16);
                    v[2 * u + 1] = __uint_as_float(hw[u] & 0xffff0000u) + __uint_as_float(lw[u] & 0xffff0000u);
                  }
                }
                if (want_cs) {
#pragma unroll
                  for (int u = 0; u < 8; ++u) cs[u] += v[u];
                }
                if (grow) {
                  const int b = (int)(((unsigned)job.g_centre0 + ((row0 + rr) >> kr_shift)) / (unsigned)job.g_M);
                  if (b != gb) {
                    flush_gs();
                    gb = b;
                  }
#pragma unroll
                  for (int u = 0; u < 8; ++u) gs[u] += v[u];
                }
              }
            }
          }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[s]);
        }
      }
      if (want_gs) flush_gs();
      if (want_cs && mine) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (lane * 8 + u < job.nv) atomicAdd(job.colsum + lane * 8 + u, cs[u] * osc);
      }
    }
    // flush: TMEM lane = output row m, columns = n
    const float osc = job.gmax ? 1.f / stage16::scale_from_max(*job.gmax) : 1.f;
    const int quarter = warp & 3;
    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
    mbar_wait(&acc_done, 0, p.err);
    tc_fence_after();
    for (int j = 0; j < mtiles; ++j) {
      const int m = j * 128 + quarter * 32 + lane;
      for (int n0 = 0; n0 < job.wy; n0 += 8) {
        float v[8];
        tmem_ld8(trow + j * 256 + n0, v);
        if (m < job.mv) {
          float *dst = job.out + (size_t)m * job.ldo + n0;
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (n0 + u < job.nv) atomicAdd(dst + u, v[u] * osc);
        }
      }
    }
    if (mma_cs && quarter == 0) {   // row 0 of the column-sum accumulator
      for (int n0 = 0; n0 < job.wy; n0 += 8) {
        float v[8];
        tmem_ld8(trow + cs_col + n0, v);
        if (lane == 0) {
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (n0 + u < job.nv) atomicAdd(job.colsum + n0 + u, v[u] * osc);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace dwtc

// Precision of the weight-gradient reduction (NSDP_DW_TERMS = 1 | 2 | 3, default 3 = fp32-grade bf16x3).
static int dw_terms() {
  static const int t = [] {
    const char *e = getenv("NSDP_DW_TERMS");
    const int v = e ? atoi(e) : 3;
    return v < 1 || v > 3 ? 3 : v;
  }();
  return t;
}

// dynamic shared memory the kernel may use: 227 KB minus its static part (barriers) with some slack. The function attribute
// is raised to this ONCE: a per-launch value would be baked differently into every captured graph node (and a profiler that
// replays single nodes then launches them with whatever value was set last).
constexpr size_t kSmemBudget = 227 * 1024 - 1024;
static int set_smem_limit() {
  static const cudaError_t e =
      cudaFuncSetAttribute(dwtc::dw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget);
  return e == cudaSuccess ? NSDP_OK : cuda_rc(e);
}

// k-steps per ring stage (Params::group)
static int pick_group(const dwtc::Job *jobs, int njobs, int terms, size_t max_stage) {
  static const int forced = [] { const char *e = getenv("NSDP_DW_GROUP"); return e ? atoi(e) : 0; }();
  for (int i = 0; i < njobs; ++i) {
    const bool x_streamed_lo = jobs[i].x_lo && terms >= 2, y_streamed_lo = jobs[i].y_lo && terms >= 3;
    if ((jobs[i].x_lo && !x_streamed_lo) || (jobs[i].y_lo && !y_streamed_lo)) return 1;   // pitch != streamed bytes
  }
  int g = 1;
  while (g < 8 && (size_t)(2 * g) * max_stage * 4 <= kSmemBudget - 2 * 4096) g *= 2;       // keep >= 4 stages in flight
  if (forced == 1 || forced == 2 || forced == 4 || forced == 8) g = forced < g ? forced : g;
  return g;
}

// Launches the reduction for `njobs` jobs that all span `tiles` row tiles.
int dw_tc_launch(const dwtc::Job *jobs, int njobs, long long tiles, int *err, cudaStream_t st) {
  using namespace dwtc;
  if (njobs <= 0 || njobs > MAX_JOBS || tiles <= 0) return NSDP_ERR_INVALID_ARGUMENT;
  Params p;
  p.terms = dw_terms();
  size_t max_stage = 0;
  double cost[MAX_JOBS], total = 0.0;
  for (int i = 0; i < njobs; ++i) {
    Job j = jobs[i];
    if (j.wx % 16 || j.wy % 16 || j.wx > 256 || j.wy > 256) return NSDP_ERR_UNSUPPORTED;
    if (j.t1 < 0) { j.t0 = 0; j.t1 = tiles; }
    if (j.t1 <= j.t0) return NSDP_ERR_INVALID_ARGUMENT;
    p.jobs[i] = j;
    // bytes one k-step brings into shared memory (the reduction is bandwidth-bound: CTAs are dealt out by bytes)
    const size_t stage = (size_t)32 * (j.wx * ((j.x_lo && p.terms >= 2) ? 2 : 1) + j.wy * ((j.y_lo && p.terms >= 3) ? 2 : 1));
    max_stage = stage > max_stage ? stage : max_stage;
    cost[i] = (double)(stage + 8192) * (double)(j.t1 - j.t0);   // + a fixed per-k-step share (handoffs, MMA issue)
    total += cost[i];
  }
  p.njobs = njobs;
  p.nchunks = 0;
  p.group = pick_group(p.jobs, njobs, p.terms, max_stage);
  // CTAs per job: proportional to bytes, at least 1, at most the job's tile count; about one CTA per SM in total
  const int budget = num_sms() > njobs ? num_sms() : njobs;
  int ctas = 0;
  p.cta_begin[0] = 0;
  for (int i = 0; i < njobs; ++i) {
    long long n = (long long)(cost[i] / total * budget);
    const long long jt = p.jobs[i].t1 - p.jobs[i].t0;
    if (n < 1) n = 1;
    if (n > jt) n = jt;
    ctas += (int)n;
    p.cta_begin[i + 1] = ctas;
  }
  p.err = err;
  // The ring is as deep as shared memory allows (+ slack for the M-tile-1 overrun): the operand tiles stream from HBM
  // (several microseconds of latency under load), 4 stages left the SMs waiting
  max_stage *= (size_t)p.group;
  int stages = (int)((kSmemBudget - 2 * 4096) / max_stage);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) return NSDP_ERR_UNSUPPORTED;
  static const int forced = [] { const char *e = getenv("NSDP_DW_STAGES"); return e ? atoi(e) : 0; }();
  if (forced >= 2 && forced < stages) stages = forced;
  p.stages = stages;
  p.ones_off = (uint32_t)((size_t)stages * max_stage + 4096);
  const size_t smem = (size_t)stages * max_stage + 2 * 4096;
  const int rc0 = set_smem_limit();
  if (rc0 != NSDP_OK) return rc0;
  dw_tc_kernel<<<ctas, THREADS, smem, st>>>(p);
  return check_launch();
}

int dw_tc_launch_chunked(const dwtc::Job *jobs, int njobs, long long tiles, const long long *bounds, int nshapes, int *err,
                         cudaStream_t st) {
  using namespace dwtc;
  if (njobs <= 0 || njobs > MAX_JOBS || tiles <= 0 || nshapes <= 0) return NSDP_ERR_INVALID_ARGUMENT;
  static const bool off = [] { const char *e = getenv("NSDP_DW_NO_CHUNKS"); return e && atoi(e) != 0; }();
  if (off || nshapes > MAX_CHUNKS / 2 || tiles > 0x7fffffffll) return NSDP_ERR_UNSUPPORTED;
  Params p;
  p.terms = dw_terms();
  size_t max_stage = 0;
  for (int i = 0; i < njobs; ++i) {
    Job j = jobs[i];
    if (j.wx % 16 || j.wy % 16 || j.wx > 256 || j.wy > 256 || j.gsum) return NSDP_ERR_UNSUPPORTED;
    j.t0 = 0; j.t1 = tiles;
    p.jobs[i] = j;
    const size_t stage = (size_t)32 * (j.wx * ((j.x_lo && p.terms >= 2) ? 2 : 1) + j.wy * ((j.y_lo && p.terms >= 3) ? 2 : 1));
    max_stage = stage > max_stage ? stage : max_stage;
  }
  p.njobs = njobs;
  p.group = pick_group(p.jobs, njobs, p.terms, max_stage);
  // about one CTA per SM in total: P chunks, dealt out to the shapes by length (at least one each), uniform inside a shape
  int P = num_sms() / njobs;
  if (P < nshapes) P = nshapes;
  if (P > MAX_CHUNKS) P = MAX_CHUNKS;
  int nc = 0;
  p.chunk_t0[0] = 0;
  for (int s = 0; s < nshapes; ++s) {
    const long long len = bounds[s + 1] - bounds[s];
    if (len <= 0) return NSDP_ERR_INVALID_ARGUMENT;
    long long n = (len * P + tiles / 2) / tiles;
    if (n < 1) n = 1;
    if (n > len) n = len;
    if (nc + n > MAX_CHUNKS) n = MAX_CHUNKS - nc;
    if (n < 1) return NSDP_ERR_UNSUPPORTED;
    for (long long c = 0; c < n; ++c) {
      p.chunk_shape[nc] = s;
      p.chunk_t0[nc + 1] = (int)(bounds[s] + len * (c + 1) / n);
      ++nc;
    }
  }
  p.nchunks = nc;
  p.err = err;
  max_stage *= (size_t)p.group;
  int stages = (int)((kSmemBudget - 2 * 4096) / max_stage);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) return NSDP_ERR_UNSUPPORTED;
  p.stages = stages;
  p.ones_off = (uint32_t)((size_t)stages * max_stage + 4096);
  const size_t smem = (size_t)stages * max_stage + 2 * 4096;
  const int rc0 = set_smem_limit();
  if (rc0 != NSDP_OK) return rc0;
  dw_tc_kernel<<<nc * njobs, THREADS, smem, st>>>(p);
  return check_launch();
}

}  // namespace nsdp

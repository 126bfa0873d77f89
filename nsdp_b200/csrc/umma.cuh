// Hand-written tcgen05 / TMEM / mbarrier / bulk-copy primitives for sm_100a (inline PTX, no CUTLASS).
//
// Conventions used by every tensor-core kernel in this library:
//   * operands are bf16, K-major, in the canonical NO-SWIZZLE ("interleave") shared-memory layout: a core matrix
//     is 8 rows x 8 elements = 8 x 16 B = 128 contiguous bytes; core matrices are laid out
//         byte(row, k) = (k / 8) * LBO + (row / 8) * SBO + (row % 8) * 16 + (k % 8) * 2
//     with SBO = 128 (row groups contiguous) and LBO = rows * 16 (one slab of all rows per 8-wide k chunk).
//     A thread that owns one row writes one 16-byte chunk per k chunk: 32 lanes -> 512 contiguous bytes, no
//     bank conflicts; the tensor core reads whole 128-byte core matrices.
//   * fp32 operands are split x = hi + lo (both bf16); a product uses three MMAs hi*hi + lo*hi + hi*lo
//     accumulated in fp32 in TMEM ("bf16x3"): ~2^-17 relative error per term, i.e. fp32-grade for this model
//     (plain bf16 misses the 1e-4 flow tolerance by 40x).
//   * accumulators: M = 128 rows <-> the 128 TMEM lanes, N columns of fp32.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace nsdp {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
#ifdef NSDP_MBAR_TESTWAIT
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"   // pure polling: never suspends the thread
#else
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
#endif
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must never hang the GPU (a hung box is a strike). mbarrier.try_wait may itself block
// for a hardware-defined interval, so the bound is on TIME (%globaltimer): 2 s. A wait that times out must not be
// walked past (the operands / TMEM accumulators behind it are not ready): the error word is set for post-mortems and
// the kernel TRAPS, so the launch fails with a sticky CUDA error that the next check() / synchronize raises on the
// host instead of silently corrupting outputs and gradients.
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, int *err = nullptr) {
  if (mbar_try_wait(bar, parity)) return;
  const unsigned long long t0 = global_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 63u) == 0u) {
      const unsigned long long dt = global_ns() - t0;
      const bool flagged = err && *reinterpret_cast<volatile int *>(err) != 0;
      if (dt > 2000000000ull || (flagged && dt > 20000ull)) {
        if (err) atomicExch(err, 1);
        __trap();
      }
    }
  }
}

// Polling variant for the single MMA-issuing warp: mbarrier.test_wait never suspends the thread, so the issuer reacts to a
// finished operand / landed weight slot a few hundred cycles sooner than with try_wait. Only ONE warp per CTA may poll like
// this: 16+ polling worker warps measurably slow the tensor pipe's shared-memory operand reads down.
__device__ __forceinline__ bool mbar_test_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// pure polling, always (no suspension: reacts within tens of cycles). For the ONE warp per CTA that sits on a latency-
// critical relay (CTA-pair kernels: "my half of the slot has landed" -> leader), never for many warps at once.
__device__ __forceinline__ void mbar_spin(uint64_t *bar, uint32_t parity, int *err = nullptr) {
  if (mbar_test_wait(bar, parity)) return;
  const unsigned long long t0 = global_ns();
  uint32_t spins = 0;
  while (!mbar_test_wait(bar, parity)) {
    if ((++spins & 1023u) == 0u) {
      const unsigned long long dt = global_ns() - t0;
      const bool flagged = err && *reinterpret_cast<volatile int *>(err) != 0;
      if (dt > 2000000000ull || (flagged && dt > 20000ull)) {
        if (err) atomicExch(err, 1);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void mbar_wait_poll(uint64_t *bar, uint32_t parity, int *err = nullptr) {
#ifndef NSDP_ISSUER_POLL   // measured on B200: no gain over try_wait (A/B, bench step 38.5 vs 39.4 ms within noise), so off
  mbar_wait(bar, parity, err);
#else
  if (mbar_test_wait(bar, parity)) return;
  const unsigned long long t0 = global_ns();
  uint32_t spins = 0;
  while (!mbar_test_wait(bar, parity)) {
    if ((++spins & 255u) == 0u) {
      const unsigned long long dt = global_ns() - t0;
      const bool flagged = err && *reinterpret_cast<volatile int *>(err) != 0;
      if (dt > 2000000000ull || (flagged && dt > 20000ull)) {
        if (err) atomicExch(err, 1);
        __trap();
      }
    }
  }
#endif
}

// ---- bulk async copy global -> shared (TMA engine, 1-D, completes on an mbarrier) -------------------------------
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// generic-proxy writes (st.shared) -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM ---------------------------------------------------------------------------------------------------
// Must be executed by one full warp. ncols: power of two in [32, 512]. The base address lands in *dst_smem.
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// True in exactly one lane of a fully converged warp. Guarding tcgen05.mma / tcgen05.commit with this (instead of
// `lane == 0`) lets ptxas emit them straight: under a plain divergent condition every warp-uniform instruction is
// wrapped in an ELECT/branch loop over the active lanes.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- descriptors --------------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp): K-major, SWIZZLE_NONE.
//   bits [0,14)  start address >> 4          bits [16,30) leading byte offset >> 4 (between the two k chunks)
//   bits [32,46) stride byte offset >> 4 (between 8-row groups)   bits [46,48) version = 1   bits [61,64) layout = 0
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16, bf16 x bf16 -> fp32, both operands K-major.
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Same, but both operands MN-major (bits 15 / 16): used for "transposed" products D[m][n] = sum_r X[r][m] * Y[r][n]
// where X and Y are the SAME canonical tiles (row r, column m) the K-major kernels write: seen as an MN-major operand
// the 8 contiguous elements run along M, the 8 rows of a core matrix along K, so the descriptor takes
// LBO = 128 (next 8 rows) and SBO = rows*16 (next 8 columns).
__host__ __device__ constexpr uint32_t idesc_bf16_mn(int M, int N) {
  return idesc_bf16(M, N) | (1u << 15) | (1u << 16);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// Arrives on the mbarrier once every previously issued tcgen05.mma of this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM -> registers: 32 lanes x 32 consecutive fp32 columns (one row per thread of the warp) -------------------------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- registers -> TMEM: 32 lanes x 8 consecutive 32-bit columns (one row per thread); tmem_st_wait() before anyone reads ---
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// raw 32-bit variant of tmem_ld8 (parked operand words)
__device__ __forceinline__ void tmem_ld8u(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// several loads in flight, ONE wait: tmem_ld8_nowait(...) x n, then tmem_ld_wait() before the registers are read
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
// Scheduling fence for registers: the compiler may not move uses of r[] above this point (asm volatile statements keep
// their order), so a fully unrolled "produce chunk, hand it over" loop is not turned into "all the math, then all the
// hand-overs".
__device__ __forceinline__ void pin8(uint32_t (&r)[8]) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


// ---- CTA pair (cluster of two CTAs, tcgen05 cta_group::2): conventions pinned by nsdp_selftest_umma2 -------------------
//   * both CTAs allocate / free TMEM with the cta_group::2 forms and get the same address; CTA c's TMEM holds rows
//     [128 c, 128 c + 128) of the M = 256 accumulator, all N columns;
//   * the leader (cluster rank 0) issues the MMAs; ONE descriptor pair addresses the same shared-memory offsets in both
//     CTAs: each holds its own 128 rows of A and its own N/2 rows of B;
//   * tcgen05.commit with the multicast mask arrives on the mbarrier at the same offset in BOTH CTAs;
//   * threads of the peer CTA arrive on the leader's mbarriers through their cluster-space address (mapa), the leader waits
//     with cluster-scope acquire.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// cluster-space address of the same shared-memory location in CTA `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// default semantics (release at CTA scope), as CUTLASS' ClusterBarrier::arrive(cta_id): a cluster-scope release on every
// hand-over made the pair kernel twice as slow (measured); what the arrival publishes is this CTA's own shared memory
// (fence.proxy.async before it), read by this CTA's own tensor core
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded like mbar_wait; for barriers that receive arrivals from the peer CTA
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity, int *err = nullptr) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const unsigned long long t0 = global_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if ((++spins & 63u) == 0u) {
      const unsigned long long dt = global_ns() - t0;
      const bool flagged = err && *reinterpret_cast<volatile int *>(err) != 0;
      if (dt > 2000000000ull || (flagged && dt > 20000ull)) {
        if (err) atomicExch(err, 1);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t *dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// arrives on the mbarrier at this offset in BOTH CTAs once every previously issued MMA of this thread has completed
__device__ __forceinline__ void mma_commit_pair(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

// ---- fp32 -> (hi, lo) bf16 split, two values packed per 32-bit word (low half = first element) ----------------------------
__device__ __forceinline__ void split2(float x0, float x1, uint32_t &hi, uint32_t &lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  const float r0 = x0 - __low2float(h), r1 = x1 - __high2float(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(r0, r1);
  hi = *reinterpret_cast<const uint32_t *>(&h);
  lo = *reinterpret_cast<const uint32_t *>(&l);
}

// byte offset of element (row, k) inside a canonical no-swizzle K-major operand of `rows` rows
__host__ __device__ constexpr uint32_t canon_off(int rows, int row, int k) {
  return (uint32_t)((k >> 3) * rows * 16 + (row >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2);
}

}  // namespace umma
}  // namespace nsdp

#ifdef NSDP_TRACE
// timeline of CTA 0 (debug builds only): (event id, clock64) pairs appended to a buffer whose address comes from the
// environment (NSDP_TRACE_PTR, a device pointer to >= 64 KB of zeroed memory; word 0 = number of events)
#define TR(id)                                                                              \
  do {                                                                                      \
    if (trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0) {                              \
      const unsigned long long n_ = atomicAdd(trace, 1ull);                                 \
      if (n_ < 4000) { trace[1 + 2 * n_] = (unsigned long long)(id); trace[2 + 2 * n_] = (unsigned long long)clock64(); } \
    }                                                                                       \
  } while (0)
#else
#define TR(id) do { } while (0)
#endif



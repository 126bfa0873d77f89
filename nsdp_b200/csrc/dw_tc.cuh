// Job description of the tensor-core weight-gradient reduction (dw_tc.cu).
#pragma once
#include <cuda_runtime.h>

namespace nsdp {
namespace dwtc {

struct Job {
  const unsigned char *x;  // staged tiles of X (k-step-major bf16 hi/lo, see dw_tc.cu), tile stride = 512 * wx bytes
  const unsigned char *y;  // staged tiles of Y
  float *out;              // [mv][ldo], accumulated atomically: out[m][n] += sum_r X[r][m] * Y[r][n]
  int wx, wy;              // padded widths (multiples of 16, <= 256)
  int mv, nv, ldo;         // valid extent / leading dimension of `out`
  float *colsum;           // optional [nv]: colsum[n] += sum_r Y[r][n]  (bias gradients), or nullptr
  // optional per-batch sums over a periodic subset of the rows (the decoder's global-token rows):
  //   gsum[b][n] += sum over rows r with r % g_kr == g_kr - 1 of Y[r][n],  b = (g_centre0 + row / g_kr) / g_M,
  // where `row` counts from the first row of the job's first tile. nullptr disables it.
  float *gsum;
  int g_kr, g_M;
  long long g_centre0;
  // staged precision per operand: 1 = [hi slab][lo slab] per k-step (bf16x2, fp32-grade), 0 = hi slab only (plain bf16;
  // used for exact operands such as one-hot tiles and for reductions over millions of rows)
  int x_lo = 1, y_lo = 1;
  // tile range of this job inside the staging buffers (tile index relative to x / y); t1 < 0 = all tiles of the launch
  long long t0 = 0, t1 = -1;
  // f16 != 0: both operands are staged as ONE fp16 slab per k-step (csrc/stage_f16.cuh; x_lo = y_lo = 0 then) and the
  // product is a single f16 x f16 MMA. gmax != nullptr: device word holding the float bits of max|g| from which the
  // producer derived its power-of-two gradient scale; every result of the job is multiplied by its inverse.
  int f16 = 0;
  const unsigned *gmax = nullptr;
  // chunked launches only: `out` advances by this many floats per shape index of the CTA's chunk (per-shape outputs)
  long long out_shape_stride = 0;
};

}  // namespace dwtc

int dw_tc_launch(const dwtc::Job *jobs, int njobs, long long tiles, int *err, cudaStream_t st);

// Chunk-aligned variant: the tile range [0, tiles) is cut into the SAME chunks for every job (chunk boundaries never cross
// the shape boundaries `bounds[0] = 0 < ... < bounds[nshapes] = tiles`), one CTA per (chunk, job), CTAs of a chunk adjacent
// in launch order. Jobs that read the same staged operand (the H tiles feed two weight gradients, every dY tile one weight
// and one table gradient, `lat` six of the tail's products) then start on it together and part of the later reads hit L2
// (measured: DRAM reads -6 %; the jobs drift apart because they stream at different rates, and pacing them through a
// progress board in global memory cost more than it saved). Returns NSDP_ERR_UNSUPPORTED when the segment holds too many shapes (use dw_tc_launch then).
int dw_tc_launch_chunked(const dwtc::Job *jobs, int njobs, long long tiles, const long long *bounds, int nshapes, int *err,
                         cudaStream_t st);

}  // namespace nsdp

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
};

}  // namespace dwtc

int dw_tc_launch(const dwtc::Job *jobs, int njobs, long long tiles, int *err, cudaStream_t st);

}  // namespace nsdp

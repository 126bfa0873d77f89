/*
 * nsdp_b200 — C ABI of libnsdp_b200.so: the B200 (sm_100a) kernels behind NSDP's TDNet hot path.
 *
 * This is the drop-in boundary. Part 1 replaces, entry point for entry point, what the
 * reference's pybind11 module `pointnet2_ops._ext` binds
 * (pointnet2_ops_lib/pointnet2_ops/_ext-src/src/bindings.cpp:6-19); part 2 is the fused replacement of
 * the torch-op chains in model/encoder/blocks.py and model/decoder/blocks.py.
 *
 * Conventions (all entry points):
 *   - plain pointers + sizes; every pointer is DEVICE memory unless noted, fp32 / int32, contiguous,
 *     16-byte aligned; `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *   - the callee never allocates, never synchronises, never calls exit(); kernels are enqueued on
 *     `stream` of the CURRENT device (the caller holds the device guard — the reference's _ext has none,
 *     a latent multi-GPU bug noted in SURVEY.md §2.2);
 *   - returns NSDP_OK (0) or a negative nsdp_status; nsdp_strerror() names it;
 *   - outputs are fully overwritten unless documented as "accumulates";
 *   - thread-safe, no global mutable state except a lazily initialised kernel-attribute cache.
 */
#ifndef NSDP_B200_H_
#define NSDP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  NSDP_OK = 0,
  NSDP_ERR_INVALID_ARGUMENT = -1, /* null pointer, non-positive size, k > N ... */
  NSDP_ERR_UNSUPPORTED = -2,      /* shape outside what the kernels are instantiated for */
  NSDP_ERR_CUDA = -3,             /* launch failed; nsdp_last_cuda_error() has the cudaError_t */
  NSDP_ERR_WORKSPACE = -4         /* workspace missing or too small */
} nsdp_status;

const char *nsdp_strerror(int status);
int nsdp_last_cuda_error(void);   /* cudaError_t of the last failed launch on this thread */
const char *nsdp_version(void);
/* Arch the library was compiled for, e.g. "sm_100a". */
const char *nsdp_build_arch(void);

/* ====================================================================================================
 * Part 1 — pointnet2_ops._ext replacements
 * ==================================================================================================== */

/* furthest_point_sampling(points (B,N,3) f32, nsamples) -> (B,nsamples) i32
 * replaces sampling.h:6 / sampling.cpp:66-87 / sampling_gpu.cu:69-229. Bit-exact with the reference
 * kernel including its tie-break. `out_idx` (B,m) is fully written.
 * No workspace: the running min-distances live in registers of a thread-block cluster. */
int nsdp_fps_f32(const float *xyz, int B, int N, int m, int32_t *out_idx, void *stream);

/* gather_points(points (B,C,N), idx (B,M)) -> (B,C,M)            sampling.h:4, sampling_gpu.cu:8-30 */
int nsdp_gather_points_f32(const float *points, const int32_t *idx, int B, int C, int N, int M,
                           float *out, void *stream);
/* gather_points_grad(grad_out (B,C,M), idx (B,M), n) -> (B,C,N)  sampling.h:5, sampling_gpu.cu:34-57
 * `grad_points` must be zero-filled by the caller (the reference's host wrapper does torch::zeros);
 * ACCUMULATES with atomics. */
int nsdp_gather_points_grad_f32(const float *grad_out, const int32_t *idx, int B, int C, int N, int M,
                                float *grad_points, void *stream);

/* ball_query(new_xyz (B,M,3), xyz (B,N,3), radius, nsample) -> (B,M,nsample) i32
 * ball_query.h:4-5, ball_query_gpu.cu:9-44. Rows without any hit are all 0 (reference: torch::zeros). */
int nsdp_ball_query_f32(const float *new_xyz, const float *xyz, int B, int N, int M, float radius,
                        int nsample, int32_t *out_idx, void *stream);

/* group_points(points (B,C,N), idx (B,M,K)) -> (B,C,M,K)         group_points.h:4, group_points_gpu.cu:8-28 */
int nsdp_group_points_f32(const float *points, const int32_t *idx, int B, int C, int N, int M, int K,
                          float *out, void *stream);
/* group_points_grad(grad_out (B,C,M,K), idx, n) -> (B,C,N)       group_points.h:5, group_points_gpu.cu:43-64
 * ACCUMULATES into caller-zeroed `grad_points`. */
int nsdp_group_points_grad_f32(const float *grad_out, const int32_t *idx, int B, int C, int N, int M,
                               int K, float *grad_points, void *stream);

/* three_nn(unknown (B,n,3), known (B,m,3)) -> dist2 (B,n,3) f32, idx (B,n,3) i32
 * interpolate.h:6, interpolate_gpu.cu:9-59 (squared distances; the Python side takes the sqrt). */
int nsdp_three_nn_f32(const float *unknown, const float *known, int B, int n, int m, float *dist2,
                      int32_t *out_idx, void *stream);
/* three_interpolate(points (B,C,m), idx (B,n,3), weight (B,n,3)) -> (B,C,n)  interpolate_gpu.cu:72-101 */
int nsdp_three_interpolate_f32(const float *points, const int32_t *idx, const float *weight, int B,
                               int C, int m, int n, float *out, void *stream);
/* three_interpolate_grad(grad_out (B,C,n), idx, weight, m) -> (B,C,m)        interpolate_gpu.cu:116-143
 * ACCUMULATES into caller-zeroed `grad_points`. */
int nsdp_three_interpolate_grad_f32(const float *grad_out, const int32_t *idx, const float *weight,
                                    int B, int C, int n, int m, float *grad_points, void *stream);

/* ====================================================================================================
 * Part 2 — fused replacements of the torch-op chains (no _ext equivalent in the reference)
 * ==================================================================================================== */

/* k nearest neighbours: replaces `square_distance(q, ref).argsort()[:, :, :k]`
 * (model/utils.py:39-55; model/encoder/blocks.py:101-102, 287-288; model/decoder/blocks.py:50-52).
 * query (B,M,3), ref (B,N,3) -> out_idx (B,M,k) i32 ascending by (distance, index); out_d2 (B,M,k) or
 * NULL. Distance = ((dx*dx + dy*dy) + dz*dz), separately rounded, like torch. The [M,N] matrix is never
 * materialised. `workspace` is only needed when nsdp_knn_workspace_bytes() > 0 (reference range split
 * over several threads per query for large N). 1 <= k <= min(N, 64). */
size_t nsdp_knn_workspace_bytes(int B, int M, int N, int k);
int nsdp_knn_f32(const float *query, const float *ref, int B, int M, int N, int k, int32_t *out_idx,
                 float *out_d2, void *workspace, size_t workspace_bytes, void *stream);

/* Vector ("point-transformer") attention core over neighbourhoods — the pair-level part of
 * TransformerBlock (model/encoder/blocks.py:104-126), TransformerSetAbstraction (:290-308) and
 * CrossTransformerBlock (model/decoder/blocks.py:62-91). For centre i and neighbour j = idx[i][t]:
 *     rel   = sign * (xyz_c[i] - xyz_n[j])
 *     h     = relu(wd0 * rel + bd0)                         (3 -> D)
 *     dlt   = wd2 * h                                       (bias folded into vc by the caller)
 *     g     = relu(wp * h + pc + qp[i] - kp[j])             (wp = Wgamma0 * Wdelta2, folded by the caller)
 *     a     = wg2 * g                                       (its bias cancels in the softmax)
 *     w     = softmax over t (per channel)
 *     out[i]= sum_t w * (vc + vp[j] + dlt)
 * With has_global != 0 one extra row per centre takes part in the softmax: rel-free (h = 0), with
 * g = relu(gq[b]) and value gv[b] (the decoder's global token, decoder/blocks.py:64-75).
 * Weight matrices are passed TRANSPOSED (K-major): wd2t[kk*D + c] = Wdelta2[c][kk] etc.
 * idx == NULL means "every centre attends to all N source points" (group_all, blocks.py:96-99).
 * qp/kp/vp may be NULL (treated as zeros: the pos_only block). D <= 256, K <= 128 (K <= 127 with global).
 */
typedef struct {
  const float *xyz_c; /* (B,M,3) */
  const float *xyz_n; /* (B,N,3) */
  const int32_t *idx; /* (B,M,K) or NULL */
  const float *qp;    /* (B,M,D) or NULL */
  const float *kp;    /* (B,N,D) or NULL */
  const float *vp;    /* (B,N,D) or NULL */
  const float *gq;    /* (B,D) pre-activation of the global row, or NULL */
  const float *gv;    /* (B,D) value of the global row, or NULL */
  const float *wd0;   /* (D,3) row-major as in nn.Linear(3,D).weight */
  const float *bd0;   /* (D) */
  const float *wd2t;  /* (D,D) transposed */
  const float *wpt;   /* (D,D) transposed */
  const float *wg2t;  /* (D,D) transposed */
  const float *pc;    /* (D) */
  const float *vc;    /* (D) */
  const float *wd2;   /* (D,D) un-transposed copies, read by the BACKWARD kernel only (NULL for forward) */
  const float *wp;    /* (D,D) */
  const float *wg2;   /* (D,D) */
  int B, M, N, K, D;
  int has_global;
  float sign;
  int impl;           /* 0 = auto (tensor-core kernel when the shape is instantiated, else CUDA cores),
                         1 = force the fp32 CUDA-core kernel, 2 = require the tcgen05 kernel */
  void *saved;        /* optional opaque forward -> backward buffer of nsdp_vattn_saved_bytes() bytes (device memory), or
                         NULL. When the forward call gets one (together with `stats`), it keeps the operand tiles and
                         pre-softmax values of the pair level there and the backward call that receives the SAME buffer
                         skips the recomputation of the forward chain. */
  size_t saved_bytes; /* size of `saved` */
} nsdp_vattn_args;

/* `stats` (2,B,M,D) or NULL: when given, the per-(centre, channel) softmax max and 1/sum are stored for the
 * backward kernel. */
size_t nsdp_vattn_fwd_workspace_bytes(const nsdp_vattn_args *args); /* packed bf16 hi/lo weight image (tcgen05 path) */
/* bytes of the optional forward -> backward buffer `args->saved` (0: this shape / implementation keeps nothing and the
 * backward recomputes the chain, which is what the reference-sized memory footprint asks for) */
size_t nsdp_vattn_saved_bytes(const nsdp_vattn_args *args);
int nsdp_vattn_fwd_f32(const nsdp_vattn_args *args, float *out /* (B,M,D) */, float *stats, void *workspace,
                       size_t workspace_bytes, void *stream);

/* Backward of nsdp_vattn_fwd_f32. Recomputes the forward chain tile by tile from the inputs, the forward
 * result `out` and the softmax statistics `stats`; no [pairs, D] activation is ever stored.
 * All gradient buffers ACCUMULATE and must be zero-filled (or hold a running sum) on entry; any of
 * them may be NULL to skip it (d_xyz_* are NULL at every level but the first, where the reference's
 * anchors are detached). */
typedef struct {
  float *d_qp;    /* (B,M,D) */
  float *d_kp;    /* (B,N,D) */
  float *d_vp;    /* (B,N,D) */
  float *d_gq;    /* (B,D) */
  float *d_gv;    /* (B,D) */
  float *d_wd0;   /* (D,3) */
  float *d_bd0;   /* (D) */
  float *d_wd2t;  /* (D,D) */
  float *d_wpt;   /* (D,D) */
  float *d_wg2t;  /* (D,D) */
  float *d_pc;    /* (D) */
  float *d_vc;    /* (D) */
  float *d_xyz_c; /* (B,M,3) */
  float *d_xyz_n; /* (B,N,3) */
} nsdp_vattn_grads;

size_t nsdp_vattn_bwd_workspace_bytes(const nsdp_vattn_args *args);
int nsdp_vattn_bwd_f32(const nsdp_vattn_args *args, const float *out /* (B,M,D) */,
                       const float *stats /* (2,B,M,D) */, const float *d_out /* (B,M,D) */,
                       const nsdp_vattn_grads *grads, void *workspace, size_t workspace_bytes,
                       void *stream);

/* Decoder ResNet-FC tail (model/decoder/crosstransformer_decoder.py:63-69 with ResnetBlockFC,
 * model/decoder/blocks.py:133-142), fused over row tiles:
 *     net = init(lat); for i < n_blocks: net += fc_c[i](lat); net += fc_1[i](relu(fc_0[i](relu(net))));
 *     out = fc_out(relu(net))
 * lat (R, C) with C <= 256; hidden H == 128; out (R, O) with O <= 4. Weights are passed transposed
 * (K-major) and concatenated:
 *     wc_t  (C, (1+n_blocks)*H): columns [0,H) = init_enc, then fc_c[0..n)      bc ((1+n_blocks)*H)
 *     w0_t  (n_blocks, H, H), b0 (n_blocks, H); w1_t (n_blocks, H, H), b1 (n_blocks, H)
 *     wo_t  (H, O), bo (O)
 */
typedef struct {
  const float *lat;
  const float *wc_t, *bc;
  const float *w0_t, *b0, *w1_t, *b1;
  const float *wo_t, *bo;
  int R, C, H, O, n_blocks;
  int impl;  /* 0 = auto (tcgen05 kernel when instantiated), 1 = fp32 CUDA-core kernel, 2 = require tcgen05 */
} nsdp_tail_args;

size_t nsdp_resnet_tail_fwd_workspace_bytes(const nsdp_tail_args *args); /* packed bf16 hi/lo weights (tcgen05 path) */
int nsdp_resnet_tail_fwd_f32(const nsdp_tail_args *args, float *out /* (R,O) */, void *workspace,
                             size_t workspace_bytes, void *stream);

/* Backward of nsdp_resnet_tail_fwd_f32 (recomputes the activations tile by tile). Gradient buffers have the
 * layouts of the corresponding (transposed, concatenated) forward arguments, ACCUMULATE, and must be
 * zero-filled on entry; d_lat (R,C) is fully overwritten. */
typedef struct {
  float *d_lat;
  float *d_wc_t, *d_bc;
  float *d_w0_t, *d_b0, *d_w1_t, *d_b1;
  float *d_wo_t, *d_bo;
} nsdp_tail_grads;

size_t nsdp_resnet_tail_bwd_workspace_bytes(const nsdp_tail_args *args);
int nsdp_resnet_tail_bwd_f32(const nsdp_tail_args *args, const float *d_out /* (R,O) */,
                             const nsdp_tail_grads *grads, void *workspace, size_t workspace_bytes,
                             void *stream);

/* Plain fused neural-field MLP over query rows (BASELINE.json configs[3], SURVEY.md §8b "nsdp_fused_mlp_fwd": the
 * decoder-only microbenchmark of the pattern model/decoder/crosstransformer_decoder.py:63-69 applies to every
 * spatial sample — one small weight stack, every row independent):
 *     h = relu(x W_in + b_in);  for l < n_hidden: h = relu(h W_l + b_l);  out = h W_out + b_out
 * i.e. 2 + n_hidden nn.Linear layers with ReLU between them. x (R, Cin), Cin <= 4; out (R, O), O <= 4;
 * width W <= 256 and a multiple of 4 (tcgen05 kernel: W in {16, 32, 64, 128, 256}, 1 <= n_hidden <= 7; everything else,
 * and impl == 1, runs the fp32 CUDA-core kernel). Weights are passed transposed (K-major), like nsdp_tail_args:
 *     w_in_t (Cin, W), b_in (W);  w_h_t (n_hidden, W, W) with w_h_t[l][k][n] = W_l[n][k], b_h (n_hidden, W);
 *     w_out_t (W, O), b_out (O).
 * The activations never leave the SM: HBM traffic is 4 (Cin + O) bytes per row; the packed weights stream from L2.
 * reuse_packed != 0: `workspace` still holds the packed weight image written by an earlier call with the same
 * weights and shapes (inference: weights are constant), so the packing kernel is skipped. */
typedef struct {
  const float *x;
  const float *w_in_t, *b_in;
  const float *w_h_t, *b_h;
  const float *w_out_t, *b_out;
  int R, Cin, W, O, n_hidden;
  int impl;          /* 0 = auto, 1 = fp32 CUDA-core kernel, 2 = require tcgen05 */
  int reuse_packed;
} nsdp_mlp_args;

size_t nsdp_fused_mlp_fwd_workspace_bytes(const nsdp_mlp_args *args);
int nsdp_fused_mlp_fwd_f32(const nsdp_mlp_args *args, float *out /* (R,O) */, void *workspace, size_t workspace_bytes,
                           void *stream);

/* Backward of nsdp_fused_mlp_fwd_f32 (recomputes the activations tile by tile; what autograd does to the same nn.Linear /
 * ReLU stack). Gradient buffers have the layouts of the corresponding (transposed) forward arguments, ACCUMULATE and must
 * be zero-filled on entry; d_x (R, Cin) is fully overwritten and may be NULL. tcgen05 kernel only (W in {16, 32, 64, 128,
 * 256}, 1 <= n_hidden <= 7): the chain kernel stages the operand tiles of every layer (bf16 hi/lo) segment by segment and
 * the split-K reduction kernel turns them into weight / bias gradients. */
typedef struct {
  float *d_x;
  float *d_w_in_t, *d_b_in;
  float *d_w_h_t, *d_b_h;
  float *d_w_out_t, *d_b_out;
} nsdp_mlp_grads;

size_t nsdp_fused_mlp_bwd_workspace_bytes(const nsdp_mlp_args *args);
int nsdp_fused_mlp_bwd_f32(const nsdp_mlp_args *args, const float *d_out /* (R,O) */, const nsdp_mlp_grads *grads,
                           void *workspace, size_t workspace_bytes, void *stream);

/* ElementwiseMLP of the point-transformer encoder (model/encoder/blocks.py:137-159):
 *     out = bn3( x + relu( bn2( conv2( relu( bn1( conv1 x ) ) ) ) ) )
 * over rows x (R, C) = the (B, n, C) feature tensor; conv1 / conv2 are the kernel-size-1 Conv1d layers (weights (C, C)
 * = [out][in], biases (C)), bn1..bn3 nn.BatchNorm1d(C). Replaces the reference's permute -> 2 cuDNN convolutions ->
 * 3 cuDNN batch-norms -> 2 ReLUs -> add chain (~12 launches forward, ~30 backward) by 4 + 7 fused fp32 kernels
 * (csrc/emlp.cu). training != 0: batch statistics (biased variance) and in-place running-stat updates exactly as
 * nn.BatchNorm1d (momentum, unbiased running_var, num_batches_tracked += 1); training == 0: running statistics.
 * C <= 256. The forward also returns what the backward needs: t1 = conv1(x), t2 = conv2(relu(bn1 t1)), s = x + relu(bn2 t2)
 * (R, C each) and `stats`, nsdp_emlp_stats_bytes() bytes of fp64 column sums (sum, sum of squares of t1, t2, s). */
typedef struct {
  const float *x;                       /* (R, C) */
  const float *w1, *b1, *w2, *b2;       /* conv1 / conv2: (C, C) [out][in], (C) */
  const float *bn_weight[3], *bn_bias[3];
  float *running_mean[3], *running_var[3];     /* updated in place when training */
  long long *num_batches_tracked[3];           /* int64 scalars, incremented when training (may be NULL) */
  int R, C;
  int training;
  float momentum, eps;
} nsdp_emlp_args;

size_t nsdp_emlp_stats_bytes(const nsdp_emlp_args *args);
int nsdp_emlp_fwd_f32(const nsdp_emlp_args *args, float *out, float *t1, float *t2, float *s, double *stats, void *stream);

/* Backward: d_x (R, C) is overwritten; d_w1 / d_w2 (C, C) and d_b1 / d_b2 (C) ACCUMULATE (pass zero-filled buffers);
 * d_bn_weight / d_bn_bias (C each) are overwritten and may be NULL. Both modes (batch / running statistics). */
typedef struct {
  float *d_x;
  float *d_w1, *d_b1, *d_w2, *d_b2;
  float *d_bn_weight[3], *d_bn_bias[3];
} nsdp_emlp_grads;

size_t nsdp_emlp_bwd_workspace_bytes(const nsdp_emlp_args *args);
int nsdp_emlp_bwd_f32(const nsdp_emlp_args *args, const float *t1, const float *t2, const float *s, const double *stats,
                      const float *d_out, const nsdp_emlp_grads *grads, void *workspace, size_t workspace_bytes,
                      void *stream);

/* Staging format of the operand tiles that the decoder attention / decoder tail backward hand to the weight-gradient
 * reduction through HBM (the dominant HBM traffic of a training step; csrc/stage_f16.cuh):
 *   1 = fp16 (default): 2 B / element, one MMA per product, gradient tiles scaled by a power of two derived from a sample
 *       of max|d_out|; weight / table gradients carry a relative error of ~2e-4 (11-bit operands);
 *   0 = bf16 hi + lo: 4 B / element, three MMAs per product, fp32-grade (~1e-5) at twice the traffic.
 * Process-wide; returns the previous format; any other value only queries. Environment override at first use:
 * NSDP_STAGE_FMT=fp16 | bf16x2. The encoder's attention blocks and nsdp_fused_mlp_bwd_f32 always use bf16 hi + lo. */
int nsdp_set_stage_format(int fmt);

/* Weight / bias gradient of `nn.Linear` with a narrow input (K <= 8 channels) on many rows, the backward of the encoder's
 * input layer `enc_sdf` (model/encoder/pointransformer.py:25, :97) and of the projections folded through it:
 *   dW (N,K) += dy (R,N)^T x (R,K);   db (N) += column sums of dy (db may be NULL).
 * All row-major contiguous fp32; dW / db are ACCUMULATED into (zero them first for a plain gradient). */
int nsdp_linear_narrow_dw_f32(const float *x, const float *dy, long long R, int K, int N, float *dW, float *db, void *stream);

/* One optimizer step of torch.optim.Adam (amsgrad = False, maximize = False, capturable layout: the step count is a
 * float32 device scalar per parameter) over `ntensors` parameter tensors in two launches: replaces `optimizer.step()` of the
 * reference's train_on_batch_* (model/deformation_networks.py:72, model/flow_arbitrary.py:45; optimizer built by
 * model/__init__.py:21-40). p / g / m / v / step: HOST arrays of device pointers (parameter, gradient, exp_avg, exp_avg_sq,
 * step), numel: host array of element counts. All tensors contiguous fp32. Update rule of torch's fused Adam:
 *   step += 1; g' = g + weight_decay p; m += (g' - m)(1 - beta1); v = beta2 v + (1 - beta2) g'^2;
 *   p -= lr / (1 - beta1^step) * m / (sqrt(v) / sqrt(1 - beta2^step) + eps);
 * the hyper-parameters are doubles (as torch holds them): 1 - beta and beta^step are formed in double, the rest in fp32. */
int nsdp_adam_step_f32(int ntensors, float *const *p, const float *const *g, float *const *m, float *const *v,
                       float *const *step, const long long *numel, double lr, double beta1, double beta2, double eps,
                       double weight_decay, void *stream);

/* Hardware self-test of the tcgen05 / TMEM conventions the tensor-core kernels rely on:
 * D (128,N) = A (128,K) * B (N,K)^T in bf16 (split == 0) or bf16x3 split precision (split != 0), single CTA.
 * N % 16 == 0, 16 <= N <= 256, K % 16 == 0. *err (device int) is set to 1 if an mbarrier wait timed out.
 * mn == 0: operands K-major as stated. mn != 0: A is given as X (K,128), B as Y (K,N) and D = X^T * Y, the tiles
 * being consumed as MN-major operands (the weight-gradient kernels' use of activation tiles). */
int nsdp_selftest_umma(const float *A, const float *B, float *D, int N, int K, int split, int mn, int *err,
                       void *stream);

/* The same product on a CTA PAIR (tcgen05 cta_group::2, a cluster of two CTAs, one MMA stream issued by the leader):
 * D (256,N) = A (256,K) * B (N,K)^T; CTA c holds rows [128c, 128c+128) of A and rows [N/2 c, N/2 c + N/2) of B.
 * N % 16 == 0, 32 <= N <= 256, K % 16 == 0. Pins the pair conventions (allocation in both CTAs, M = 256 instruction
 * descriptor, per-CTA halves of B behind one descriptor, multicast commit) before a kernel relies on them. */
int nsdp_selftest_umma2(const float *A, const float *B, float *D, int N, int K, int split, int *err, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NSDP_B200_H_ */

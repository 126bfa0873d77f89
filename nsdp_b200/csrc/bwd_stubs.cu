// Placeholder entry points for the backward kernels (replaced by vattn_bwd.cu / resnet_tail_bwd.cu).
#include "common.cuh"
extern "C" size_t nsdp_vattn_bwd_workspace_bytes(const nsdp_vattn_args *) { return 0; }
extern "C" int nsdp_vattn_bwd_f32(const nsdp_vattn_args *, const float *, const nsdp_vattn_grads *, void *, size_t,
                                  void *) { return NSDP_ERR_UNSUPPORTED; }
extern "C" size_t nsdp_resnet_tail_bwd_workspace_bytes(const nsdp_tail_args *) { return 0; }
extern "C" int nsdp_resnet_tail_bwd_f32(const nsdp_tail_args *, const float *, const nsdp_tail_grads *, void *, size_t,
                                        void *) { return NSDP_ERR_UNSUPPORTED; }

// Status / version entry points of the C ABI (include/nsdp_b200.h).
#include "common.cuh"

namespace nsdp {
thread_local int g_last_cuda_error = 0;
}

extern "C" {

const char *nsdp_strerror(int status) {
  switch (status) {
    case NSDP_OK: return "ok";
    case NSDP_ERR_INVALID_ARGUMENT: return "invalid argument";
    case NSDP_ERR_UNSUPPORTED: return "unsupported shape for the compiled kernels";
    case NSDP_ERR_CUDA: return "CUDA launch/runtime error (see nsdp_last_cuda_error)";
    case NSDP_ERR_WORKSPACE: return "workspace missing or too small";
    default: return "unknown nsdp status";
  }
}

int nsdp_last_cuda_error(void) { return nsdp::g_last_cuda_error; }

const char *nsdp_version(void) { return "nsdp_b200 0.1.0"; }

const char *nsdp_build_arch(void) { return "sm_100a"; }

}  // extern "C"

"""TEST INFRASTRUCTURE — compiles the REFERENCE's own pointnet2_ops CUDA extension, from the
sources where they lie under /root/reference (never copied), for sm_100a, into oracle/_ref/.

Why: the reference's FPS / ball-query / grouping / 3-NN kernels are CUDA-only
(sampling.cpp:82-84 "CPU not supported") and its build scripts pin an arch list nvcc 12.9
rejects (pointnet2_ops_lib/setup.py:19). Compiling the unmodified sources directly with the
arch overridden gives a real-reference binary that `-m gpu` tests use on the B200 box to pin
both oracle/nsdp_oracle.c and the product kernels, index for index.

Recipe = torch.utils.cpp_extension.load on the reference's 5 .cpp + 4 .cu files (not the
reference's own setup.py). Output: oracle/_ref/nsdp_ref_pointnet2_ext.so (git-ignored, NOT
gpurun-ignored). /root/reference does not exist on the GPU box: there the prebuilt .so is
simply imported (see oracle/ref_ext.py).
"""
import glob
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/pointnet2_ops_lib/pointnet2_ops/_ext-src"
OUT_DIR = os.path.join(HERE, "_ref")
NAME = "nsdp_ref_pointnet2_ext"


def build(verbose: bool = False):
    if not os.path.isdir(REF_SRC):
        return None
    out = os.path.join(OUT_DIR, NAME + ".so")
    srcs = sorted(glob.glob(os.path.join(REF_SRC, "src", "*.cpp")) +
                  glob.glob(os.path.join(REF_SRC, "src", "*.cu")))
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in srcs):
        return out
    os.makedirs(OUT_DIR, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    from torch.utils.cpp_extension import load
    load(NAME, sources=srcs, extra_include_paths=[os.path.join(REF_SRC, "include")],
         extra_cflags=["-O3"], extra_cuda_cflags=["-O3"], with_cuda=True,
         build_directory=OUT_DIR, verbose=verbose, is_python_module=False)
    return out


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))

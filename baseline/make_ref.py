"""Recipe for baseline/_ref/: a byte-for-byte copy of the parts of the reference that bench.py's reference arm runs.

    python baseline/make_ref.py

Copies (never edits)  /root/reference/model/**                      -> baseline/_ref/model/
                      /root/reference/{train,test,run}.py, dataset/*.py, utils/*.py -> baseline/_ref/ (the unchanged callers)
                      /root/reference/pointnet2_ops_lib/pointnet2_ops/*.py -> baseline/_ref/pointnet2_ops_lib/pointnet2_ops/
baseline/_ref/ is git-ignored (reference sources never enter this repository's history) but NOT gpurun-ignored, so it
travels to the GPU box, where /root/reference does not exist. The one thing the copy cannot contain is a CPU build of
`pointnet2_ops._ext` (the reference's FPS kernel is CUDA-only, sampling.cpp:82-84): baseline/ref_loader.py supplies it.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
DST = os.path.join(HERE, "_ref")


def build(verbose: bool = False):
    if not os.path.isdir(os.path.join(REF, "model")):
        return DST if os.path.isdir(os.path.join(DST, "model")) else None
    pairs = []
    for root, _dirs, files in os.walk(os.path.join(REF, "model")):
        for f in files:
            if f.endswith(".py"):
                src = os.path.join(root, f)
                pairs.append((src, os.path.join(DST, os.path.relpath(src, REF))))
    # the caller scripts and their host-side helpers: tests/test_gpu_reference_scripts.py runs train.py / test.py UNCHANGED
    # through `python -m nsdp_b200.launch` (the GPU box has no /root/reference)
    for script in ("train.py", "test.py", "run.py"):
        pairs.append((os.path.join(REF, script), os.path.join(DST, script)))
    for sub in ("dataset", "utils"):
        for f in os.listdir(os.path.join(REF, sub)):
            if f.endswith(".py"):
                pairs.append((os.path.join(REF, sub, f), os.path.join(DST, sub, f)))
    p2 = os.path.join(REF, "pointnet2_ops_lib", "pointnet2_ops")
    for f in os.listdir(p2):
        if f.endswith(".py"):
            pairs.append((os.path.join(p2, f), os.path.join(DST, "pointnet2_ops_lib", "pointnet2_ops", f)))
    for src, dst in pairs:
        if os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        if verbose:
            print("copied", os.path.relpath(dst, HERE))
    return DST


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))

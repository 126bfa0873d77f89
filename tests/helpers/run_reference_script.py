"""Runs an UNCHANGED reference script (baseline/_ref/train.py ...) on the REFERENCE model: binds the compiled
`pointnet2_ops._ext` the reference expects (baseline/ref_loader.py: the reference's own CUDA extension on a GPU, the C
restatement of its FPS kernel on the CPU) and hands over to the script. The counterpart of `python -m nsdp_b200.launch`,
used by the harness tests to produce the reference's own numbers for the same command line."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from baseline import ref_loader  # noqa: E402

ref_loader.load()
script = sys.argv[1]
sys.argv = sys.argv[1:]
sys.path.insert(0, os.path.dirname(os.path.abspath(script)))
runpy.run_path(script, run_name="__main__")

"""bench.py's contract on a machine without a GPU: the reference arm (CPU restatement of the path on the host cores) prints
exactly one JSON line with the keys the driver reads; under torchrun only rank 0 prints; our arm refuses to run without CUDA."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e,
                          timeout=600)


def test_reference_arm_prints_one_contract_line():
    res = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                   # stdout carries the JSON line and nothing else
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "query-points/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("query-points/sec (TDNet fwd+bwd")
    assert d["value"] > 0 and d["steps"] == 1 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "configs[1]" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    # "reference" where baseline/_ref (the unmodified reference model, baseline/make_ref.py) travelled with the repo
    want_kind = "reference" if os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "model")) else "port"
    assert cb["kind"] == want_kind and cb["cores"] >= 1 and cb["value"] == d["value"] and "4096" in cb["sample"]
    assert d["warmup"] == 0 and d["config"]["per_gpu_batch"] == 8 and 1 <= d["config"]["sample_shapes_per_step"] <= 8
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_rank0_under_torchrun_env_needs_no_process_group():
    """Under torchrun (N > 1) rank 0 alone runs the reference arm; the other ranks have already exited, so it must not try
    to join a process group (round-1 bug: it went through nsdp_b200.model.build_model -> dist.maybe_init_from_env)."""
    res = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--queries", "2048", "--surface", "1024"],
               env={"RANK": "0", "LOCAL_RANK": "0", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29591"})
    assert res.returncode == 0, res.stderr[-2000:]
    d = json.loads([l for l in res.stdout.splitlines() if l.strip()][0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0


def test_reference_arm_does_not_import_the_product_model():
    code = ("import sys, runpy; sys.argv=['bench.py','--impl','reference','--steps','1','--warmup','0','--queries','2048','--surface','1024'];"
            "runpy.run_path(%r, run_name='__main__');"
            "bad=[m for m in sys.modules if m.startswith(('nsdp_b200.model','nsdp_b200.ops','nsdp_b200.dist','nsdp_b200._lib'))];"
            "sys.stderr.write('BAD=%%r\\n' %% bad)") % os.path.join(ROOT, "bench.py")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "BAD=[]" in res.stderr, res.stderr[-500:]


def test_reference_arm_is_silent_on_other_ranks():
    res = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert res.returncode == 0 and res.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a machine WITHOUT a GPU")
def test_our_arm_has_no_cpu_fallback():
    res = _run(["--steps", "1", "--warmup", "0"])
    assert res.returncode != 0 and "no CPU fallback" in (res.stderr + res.stdout)

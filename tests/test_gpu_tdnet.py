"""End-to-end GPU parity of the product model (nsdp_b200.model, CUDA kernels through the C ABI) against
(a) the golden vectors produced by the live reference (tests/golden/) and (b) the CPU oracle."""
import numpy as np
import pytest
import torch

from nsdp_b200 import synth
from nsdp_b200.model import build_model
from oracle import tdnet_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4  # BASELINE.json north_star: flow L2 error vs reference < 1e-4 (mean per-point L2)


def _mean_l2(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64), axis=-1).mean())


def _model(schemas, mtype):
    model, *_ = build_model(synth.make_config(mtype), device=DEV)
    model.load_state_dict(synth.named_state_dict([(k, s) for k, s in schemas[mtype]], seed=0))
    return model


@pytest.mark.parametrize("mtype", ["forward", "backward"])
def test_c1_forward_against_reference_golden(golden, schemas, mtype):
    model = _model(schemas, mtype).eval()
    batch = synth.forward_batch(1, 1024, 2048, seed=1234, fp16_grid=False)
    with torch.no_grad():
        surf = batch["surface_samples_inputs"].to(DEV)
        enc = model.encode(surf)
        out = model(batch["space_samples_src"].to(DEV), surf)
    np.testing.assert_array_equal(enc["anchors"].cpu().numpy(), golden[f"c1_{mtype}_anchors"])  # FPS: bit-exact
    np.testing.assert_allclose(enc["z"].cpu().numpy(), golden[f"c1_{mtype}_z"], atol=1e-4, rtol=1e-3)
    err = _mean_l2(out.cpu().numpy(), golden[f"c1_{mtype}_flow"])
    assert err < TOL, err


def test_flow_arbitrary_eval_against_reference_golden(golden, schemas):
    model = _model(schemas, "arbitrary").eval()
    b = synth.forward_batch(1, 1024, 512, seed=77, fp16_grid=False)
    s = b["surface_samples_inputs"].to(DEV)
    src = s[:, :, 0:3].contiguous()
    with torch.no_grad():
        space_c = model.model_canonicalize(b["space_samples_src"].to(DEV), src)
        surf_c = model.model_canonicalize(src, src)
        assert _mean_l2(space_c.cpu().numpy(), golden["arb_eval_space_src2cano"]) < TOL
        assert _mean_l2(surf_c.cpu().numpy(), golden["arb_eval_surface_src2cano"]) < TOL
        # stage 2 teacher-forced with the reference's stage-1 coordinates (see tests/test_oracle_golden.py)
        inp = torch.cat([torch.from_numpy(golden["arb_eval_surface_src2cano"]).to(DEV), s[:, :, 3:6], s[:, :, 6:7]], -1)
        out_tf = model.model_deform(torch.from_numpy(golden["arb_eval_space_src2cano"]).to(DEV), inp.contiguous())
        assert _mean_l2(out_tf.cpu().numpy(), golden["arb_eval_flow"]) < TOL
        out = model(b["space_samples_src"].to(DEV), src, s[:, :, 3:6], s[:, :, 6:7])
        assert _mean_l2(out.cpu().numpy(), golden["arb_eval_flow"]) < 5e-3


def test_forward_against_oracle_multi_shape_fp16_grid(schemas):
    """B=3 shapes on the fp16 grid (real data layout): anchors bit-exact, flow within tolerance of the CPU oracle."""
    model = _model(schemas, "forward").eval()
    batch = synth.forward_batch(3, 1500, 1000, seed=42, fp16_grid=True)
    cfg = synth.make_config("forward")["model"]
    sd = synth.named_state_dict([(k, s) for k, s in schemas["forward"]], seed=0)
    trace = {}
    with torch.no_grad():
        want = orc.tdnet_forward(sd, "", batch["space_samples_src"], batch["surface_samples_inputs"], cfg, False, trace=trace)
        surf = batch["surface_samples_inputs"].to(DEV)
        enc = model.encode(surf)
        got = model.decode(batch["space_samples_src"].to(DEV), enc)
    assert torch.equal(enc["anchors"].cpu(), trace["anchors"])
    assert _mean_l2(got.cpu().numpy(), want.numpy()) < TOL

"""The product modules must expose the reference's state_dict schema so reference checkpoints
(forward.pt / backward.pt / arbitrary.pt, model_%05d) load unchanged. Schema fixture generated from the live
reference by tests/golden/make_golden.py."""
import io

import pytest
import torch

from nsdp_b200 import synth
from nsdp_b200.model import build_model, optimizer_factory


@pytest.mark.parametrize("mtype", ["forward", "backward", "arbitrary"])
def test_state_dict_schema_equals_reference(schemas, mtype):
    model, train_fn, val_fn, test_fn = build_model(synth.make_config(mtype))
    ours = [[k, list(v.shape)] for k, v in model.state_dict().items()]
    assert ours == schemas[mtype]
    assert callable(train_fn) and callable(val_fn) and callable(test_fn)


def test_param_counts():
    fwd, *_ = build_model(synth.make_config("forward"))
    bwd, *_ = build_model(synth.make_config("backward"))
    arb, *_ = build_model(synth.make_config("arbitrary"))
    n = lambda m: sum(p.numel() for p in m.parameters())
    assert n(fwd) == 4492267 and n(bwd) == 4491667 and n(arb) == 8983934


def test_checkpoint_containers(tmp_path, schemas):
    """raw state_dict and {'model_state_dict': ...} are both accepted (model/__init__.py:85-92)."""
    cfg = synth.make_config("forward")
    sd = synth.named_state_dict([(k, s) for k, s in schemas["forward"]], seed=3)
    raw, wrapped = tmp_path / "raw.pt", tmp_path / "wrapped.pt"
    torch.save(sd, raw)
    torch.save({"model_state_dict": sd, "epoch": 7}, wrapped)
    for path in (raw, wrapped):
        model, *_ = build_model(cfg, weight_file=str(path))
        assert torch.equal(model.state_dict()["decoder.fc_out.weight"], sd["decoder.fc_out.weight"])
    arb_cfg = synth.make_config("arbitrary")
    model, *_ = build_model(arb_cfg, weight_forward_file=str(raw))
    assert torch.equal(model.model_deform.state_dict()["decoder.fc_out.weight"], sd["decoder.fc_out.weight"])


def test_optimizer_factory():
    model, *_ = build_model(synth.make_config("forward"))
    sched, opt = optimizer_factory({"optimizer": "Adam", "lr": 5e-4, "lr_step": 200, "lr_decay": 0.1}, model.parameters())
    assert isinstance(opt, torch.optim.Adam) and opt.param_groups[0]["lr"] == 5e-4
    assert abs(sched.get_learning_rate(400) - 5e-6) < 1e-12
    sched, opt = optimizer_factory({"optimizer": "SGD", "lr": 0.1}, model.parameters())
    assert isinstance(opt, torch.optim.SGD) and opt.param_groups[0]["momentum"] == 0.9
    with pytest.raises(NotImplementedError):
        optimizer_factory({"optimizer": "LBFGS"}, model.parameters())

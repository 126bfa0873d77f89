"""Host-side bookkeeping of nsdp_b200/graph.py that needs no GPU: what the cache of captured steps is keyed on."""
import gc

import torch

from nsdp_b200 import graph


def test_signature_depends_on_shapes_mode_trainable_set_and_optimizer():
    model = torch.nn.Sequential(torch.nn.Linear(3, 4), torch.nn.Linear(4, 2))
    opt = torch.optim.Adam(model.parameters())
    data = {"a": torch.zeros(2, 5, 3), "b": torch.zeros(2, 7, 3), "unused": torch.zeros(1)}
    sig = graph._signature(model, opt, data, ("a", "b"))
    assert sig == graph._signature(model, opt, dict(data, unused=torch.zeros(9)), ("a", "b"))      # only the keys the step reads
    assert sig != graph._signature(model, opt, dict(data, a=torch.zeros(2, 6, 3)), ("a", "b"))
    assert sig != graph._signature(model.eval(), opt, data, ("a", "b"))
    model.train()
    model[1].weight.requires_grad_(False)                                                         # a frozen sub-network
    assert sig != graph._signature(model, opt, data, ("a", "b"))
    model[1].weight.requires_grad_(True)
    assert sig != graph._signature(model, torch.optim.Adam(model.parameters()), data, ("a", "b"))


def test_entry_is_dropped_when_another_optimizer_reuses_the_id():
    model = torch.nn.Linear(2, 2)
    opt = torch.optim.Adam(model.parameters())
    e = graph._Entry(opt)
    assert not e.stale(opt)
    other = torch.optim.Adam(model.parameters())
    assert e.stale(other)
    del opt
    gc.collect()
    assert e.stale(other)                   # the dead optimizer's entry never matches a live one, whatever its id()
    assert not graph._Entry().stale(other)  # forward-only entries hold no optimizer


def test_learning_rate_and_weight_decay_are_part_of_the_capture_key():
    model = torch.nn.Linear(2, 2)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    before = graph._lrs(opt)
    opt.param_groups[0]["lr"] = 1e-4                        # model/learningrate.py adjust_learning_rate (train.py:188)
    assert graph._lrs(opt) != before
    opt.param_groups[0]["lr"] = torch.tensor(1e-4)          # a tensor lr lives on the device: no re-capture needed
    assert graph._lrs(opt)[0] is None


def test_cpu_inputs_never_reach_the_graph_machinery():
    model = torch.nn.Linear(3, 3).eval()
    x = torch.ones(2, 3)
    with torch.no_grad():
        out = graph.graphed_forward(model, (x,), lambda t: model(t))     # plain call: no CUDA query, no capture
    assert torch.equal(out, model(x))
    opt = torch.optim.Adam(model.parameters())
    assert not graph._usable(model, opt, {"x": x}, ("x",))

"""Model factory with the reference's API (model/__init__.py:10-118):

    build_model(config, weight_file=None, weight_forward_file=None, weight_backward_file=None, device="cpu")
        -> (model, train_on_batch, validate_on_batch, test_on_batch)
    optimizer_factory(config["training"], parameters) -> (lr_schedule, optimizer)

Checkpoints are either a raw state_dict or a dict holding one under "model_state_dict"
(model/__init__.py:85-92); parameter/buffer names and shapes equal the reference's.
"""
from __future__ import annotations

import torch

from nsdp_b200.model.deformation_networks import (Deformation_Networks, test_on_batch_with_cano,
                                                  train_on_batch_with_cano, validate_on_batch_with_cano)
from nsdp_b200.model.flow_arbitrary import (FlowArbitrary, test_on_batch_with_arbitrary,
                                            train_on_batch_with_arbitrary, validate_on_batch_with_arbitrary)
from nsdp_b200.model.learningrate import StepLearningRateSchedule


def optimizer_factory(config, parameters):
    name = config.get("optimizer", "Adam")
    schedule = StepLearningRateSchedule({
        "type": "step",
        "initial": config.get("lr", 1e-3),
        "interval": config.get("lr_step", 100),
        "factor": config.get("lr_decay", 0.1),
    })
    parameters = list(parameters)
    group = {"params": parameters, "lr": schedule.get_learning_rate(0),
             "weight_decay": config.get("weight_decay", 0.0)}
    if name == "SGD":
        group["momentum"] = config.get("momentum", 0.9)
        return schedule, torch.optim.SGD([group])
    if name == "Adam":
        # same update rule and state_dict layout as the reference's torch.optim.Adam; on the GPU the whole step is two
        # launches of the library (nsdp_b200/optim.py, csrc/adam.cu); capturable (step count on the device) so that the
        # whole training step can be replayed as a CUDA graph (nsdp_b200/graph.py)
        fused = len(parameters) > 0 and all(p.is_cuda for p in parameters)
        if fused:
            from nsdp_b200.optim import Adam
            return schedule, Adam([group])
        return schedule, torch.optim.Adam([group])
    raise NotImplementedError()


def _load_weights(module, path, device):
    print("Loading weight file from {}".format(path))
    blob = torch.load(path, map_location=device)
    if isinstance(blob, dict) and "model_state_dict" in blob:
        blob = blob["model_state_dict"]
    module.load_state_dict(blob)


_BATCH_FNS = {
    "cano": (train_on_batch_with_cano, validate_on_batch_with_cano, test_on_batch_with_cano),
    "arbitrary": (train_on_batch_with_arbitrary, validate_on_batch_with_arbitrary, test_on_batch_with_arbitrary),
}


def build_model(config, weight_file=None, weight_forward_file=None, weight_backward_file=None, device="cpu"):
    model_type = config["model"]["type"]
    if model_type in ("forward", "backward"):
        model = Deformation_Networks(config, no_input_corr=(model_type == "backward"))
        fns = _BATCH_FNS["cano"]
    elif model_type == "arbitrary":
        canonicalize = Deformation_Networks(config, no_input_corr=True)
        deform = Deformation_Networks(config, no_input_corr=False)
        if weight_forward_file is not None:
            _load_weights(deform, weight_forward_file, device)
        if weight_backward_file is not None:
            _load_weights(canonicalize, weight_backward_file, device)
        model = FlowArbitrary(config, canonicalize, deform)
        fns = _BATCH_FNS["arbitrary"]
    else:
        raise NotImplementedError()
    if weight_file is not None:
        _load_weights(model, weight_file, device)
    model.to(device)
    nsdp_dist.maybe_init_from_env(model, device)
    return (model, *fns)


from nsdp_b200 import dist as nsdp_dist  # noqa: E402  (after the public names, mirrors the reference's import order)

"""Host-side LR bookkeeping with the reference's names (model/learningrate.py:6-45), used by train.py:12-13."""
from __future__ import annotations


def print_num_parameters(model):
    total = sum(p.numel() for p in model.parameters())
    trainable = sum(p.numel() for p in model.parameters() if p.requires_grad)
    print(f"Number of parameters in {model.__class__.__name__}:  {trainable} / {total}")


class LearningRateSchedule:
    def get_learning_rate(self, epoch):
        raise NotImplementedError


class StepLearningRateSchedule(LearningRateSchedule):
    """lr = initial * factor ** (epoch // interval)  (model/learningrate.py:17-25)."""

    def __init__(self, specs):
        print(specs)
        self.initial = specs["initial"]
        self.interval = specs["interval"]
        self.factor = specs["factor"]

    def get_learning_rate(self, epoch):
        return self.initial * (self.factor ** (epoch // self.interval))


def adjust_learning_rate(lr_schedules, optimizer, epoch):
    per_group = isinstance(lr_schedules, list)
    for i, group in enumerate(optimizer.param_groups):
        sched = lr_schedules[i] if per_group else lr_schedules
        group["lr"] = sched.get_learning_rate(epoch)


def get_learning_rates(optimizer):
    return [group["lr"] for group in optimizer.param_groups]


def print_learning_rates(optimizer):
    print("".join(" | " + str(group["lr"]) for group in optimizer.param_groups))


def weights_init(m):
    """`model.apply(weights_init)`: Xavier-uniform weights and zero biases for convolutions and Linear layers, identity affine
    for BatchNorm layers (model/learningrate.py:50-60)."""
    import torch
    if isinstance(m, (torch.nn.modules.conv._ConvNd, torch.nn.Linear)):
        torch.nn.init.xavier_uniform_(m.weight)
        if m.bias is not None:
            torch.nn.init.zeros_(m.bias)
    elif isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
        torch.nn.init.ones_(m.weight)
        torch.nn.init.zeros_(m.bias)


def clamp_gradient(model, clip):
    """Clip every gradient element of `model` into [-clip, clip] (model/learningrate.py:62-64)."""
    import torch
    torch.nn.utils.clip_grad_value_(list(model.parameters()), clip)

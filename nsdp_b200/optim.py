"""torch.optim.Adam with the step as ONE kernel over all parameter tensors (csrc/adam.cu).

`optimizer_factory` (nsdp_b200/model/__init__.py, mirror of /root/reference/model/__init__.py:21-40) returns this class for
CUDA models. It IS a torch.optim.Adam: same constructor arguments, same param_groups, same per-parameter state
(`step` float32 scalar on the device, `exp_avg`, `exp_avg_sq`), so `optimizer.state_dict()` / `load_state_dict()` exchange
checkpoints (`opt_%05d`, utils/checkpoints.py) with the reference's optimizer, and CUDA-graph capture works as with torch's
capturable fused Adam. Only `step()` differs: torch's fused implementation needs 7 launches of ~70 thread blocks for this
model's 505 tensors (0.38 ms per training step), the library call below 2 launches of ~1100 blocks.
Anything the kernel does not cover (several param groups with different hyper-parameters are fine; amsgrad, maximize,
non-fp32 / non-contiguous / sparse tensors, CPU tensors, a closure) goes through torch's own step."""
from __future__ import annotations

import ctypes as C

import torch

from nsdp_b200 import _lib


class Adam(torch.optim.Adam):
    def __init__(self, params, **kwargs):
        kwargs.setdefault("fused", True)
        kwargs.setdefault("capturable", True)
        super().__init__(params, **kwargs)

    def _fast_group(self, group) -> bool:
        if group.get("amsgrad") or group.get("maximize") or group.get("differentiable") or not group.get("capturable"):
            return False
        if torch.is_tensor(group["lr"]):
            return False
        for p in group["params"]:
            g = p.grad
            if g is None:
                continue
            if not (p.is_cuda and p.dtype == torch.float32 and g.dtype == torch.float32 and not g.is_sparse
                    and p.is_contiguous() and g.is_contiguous() and g.device == p.device):
                return False
        return True

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None or not all(self._fast_group(g) for g in self.param_groups):
            return super().step(closure)
        from nsdp_b200 import ops
        stream = torch.cuda.current_stream().cuda_stream
        plans = []        # every group is validated BEFORE the first launch: falling back to torch half-way would step twice
        for group in self.param_groups:
            ps, gs, ms, vs, ss, ns = [], [], [], [], [], []
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if len(st) == 0:          # same lazy initialisation as torch.optim.Adam._init_group (capturable / fused)
                    st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                m, v, s = st["exp_avg"], st["exp_avg_sq"], st["step"]
                if not (torch.is_tensor(s) and s.is_cuda and s.dtype == torch.float32 and m.is_contiguous() and v.is_contiguous()):
                    return super().step(closure)      # e.g. a state loaded from a non-capturable optimizer: torch converts it
                ps.append(p.data_ptr()); gs.append(p.grad.data_ptr()); ms.append(m.data_ptr()); vs.append(v.data_ptr())
                ss.append(s.data_ptr()); ns.append(p.numel())
            if ps:
                plans.append((group, ps, gs, ms, vs, ss, ns))
        for group, ps, gs, ms, vs, ss, ns in plans:
            n = len(ps)
            arr = C.c_void_p * n
            beta1, beta2 = group["betas"]
            rc = _lib.lib().nsdp_adam_step_f32(n, arr(*ps), arr(*gs), arr(*ms), arr(*vs), arr(*ss), (C.c_longlong * n)(*ns),
                                               float(group["lr"]), float(beta1), float(beta2), float(group["eps"]),
                                               float(group["weight_decay"]), stream)
            _lib.check(rc, "nsdp_adam_step_f32")
            ops._count(2)
        return None

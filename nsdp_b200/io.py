"""The data formats and the staging step either side of the hot path.

* On-disk samples: `surface_points.npz` / `flow.npz`-style archives with fp16 arrays `points` (and `normals`), read exactly
  like the reference's `dataset/utils.py:8-17` (`load_npz_surface_flow`, `load_npz_space_flow`: fp16 -> fp32).
* Checkpoints: `model_%05d` / `opt_%05d` / `modelbest_%05d_%f` files holding `torch.save`d state_dicts, the names and the
  resume rule of `utils/checkpoints.py:8-74` — files written by either implementation load in the other (the
  state_dict schema is identical, tests/test_schema.py).
* `DeviceStager`: what `train.py:192-193` does with a blocking `.to(device)` per tensor on the compute stream —
  here the `default_collate`d dict of batch i+1 is copied from REUSED pinned host buffers on a side stream while
  batch i computes; fp16 arrays cross the bus as fp16 and are widened on the device.

Plumbing only: nothing here computes on the CPU what the CUDA kernels compute.
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, Iterator, Optional

import numpy as np
import torch


# ---------------------------------------------------------------------------------------------------
# sample files (dataset/utils.py:8-17)
# ---------------------------------------------------------------------------------------------------
def load_npz_surface_flow(path: str):
    """-> (points (N,3) float32, normals (N,3) float32); the arrays are stored as fp16
    (preprocess/generate_dataset_deform4d_surfaceflow.py:74-79)."""
    with np.load(path) as d:
        return d["points"].astype(np.float32), d["normals"].astype(np.float32)


def load_npz_space_flow(path: str):
    """-> points (Q,3) float32 (stored fp16, preprocess/generate_dataset_deform4d_spaceflow.py:106-112)."""
    with np.load(path) as d:
        return d["points"].astype(np.float32)


def save_npz_surface_flow(path: str, points, normals) -> None:
    np.savez(path, points=np.asarray(points).astype(np.float16), normals=np.asarray(normals).astype(np.float16))


def save_npz_space_flow(path: str, points) -> None:
    np.savez(path, points=np.asarray(points).astype(np.float16))


# ---------------------------------------------------------------------------------------------------
# checkpoints (utils/checkpoints.py:8-74)
# ---------------------------------------------------------------------------------------------------
def save_checkpoints(epoch: int, model, optimizer, experiment_directory: str) -> None:
    torch.save(model.state_dict(), os.path.join(experiment_directory, "model_{:05d}".format(epoch)))
    torch.save(optimizer.state_dict(), os.path.join(experiment_directory, "opt_{:05d}".format(epoch)))


def load_checkpoints(model, optimizer, experiment_directory: str, args, device) -> None:
    """Resume from the highest-numbered `model_*` that has a matching `opt_*`; sets args.continue_from_epoch."""
    ids = [int(f[6:]) for f in os.listdir(experiment_directory) if f.startswith("model_")]
    if not ids:
        return
    last = max(ids)
    model_path = os.path.join(experiment_directory, "model_{:05d}".format(last))
    opt_path = os.path.join(experiment_directory, "opt_{:05d}".format(last))
    if not (os.path.exists(model_path) and os.path.exists(opt_path)):
        return
    model.load_state_dict(torch.load(model_path, map_location=device))
    optimizer.load_state_dict(torch.load(opt_path, map_location=device))
    args.continue_from_epoch = last + 1


def save_best_checkpoints(epoch: int, model, experiment_directory: str, val_loss: float) -> None:
    torch.save(model.state_dict(), os.path.join(experiment_directory, "modelbest_{:05d}_{:03f}".format(epoch, val_loss)))


def load_best_checkpoints(model, experiment_directory: str, args, device) -> None:
    ids = sorted(f[10:] for f in os.listdir(experiment_directory) if f.startswith("modelbest_"))
    if not ids:
        return
    epoch, val_loss = int(ids[-1][0:5]), float(ids[-1][6:])
    path = os.path.join(experiment_directory, "modelbest_{:05d}_{:03f}".format(epoch, val_loss))
    if not os.path.exists(path):
        return
    model.load_state_dict(torch.load(path, map_location=device))
    args.continue_from_epoch = epoch + 1
    args.best_val_loss = val_loss


# ---------------------------------------------------------------------------------------------------
# host -> device staging
# ---------------------------------------------------------------------------------------------------
class DeviceStager:
    """Iterates over `loader` (any iterable of dicts of CPU tensors, e.g. a DataLoader with `default_collate`) and yields
    the same dicts on `device`, one batch AHEAD of the consumer:

        for sample in DeviceStager(train_loader, device):
            loss = train_on_batch(model, optimizer, sample, config)

    Each of the `depth` slots owns pinned host buffers that are re-used for every batch of the same shapes (a fresh
    `pin_memory()` per batch costs a cudaHostAlloc); the copy runs on a side stream, the consumer's stream waits on the
    slot's event, the device tensors are tied to the consumer's stream (`record_stream`), and a slot's pinned buffers are
    only overwritten once the previous copy out of them has completed. fp16
    tensors stay fp16 on the wire and become fp32 on the device (`widen_half=True`), halving the bytes of the on-disk
    sample format. Requires a CUDA device: there is no CPU path to stage for."""

    def __init__(self, loader: Iterable[Dict[str, torch.Tensor]], device, depth: int = 2, widen_half: bool = True):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceStager: CPU not supported (nsdp_b200 has no CPU path)")
        self.loader, self.depth, self.widen_half = loader, max(2, depth), widen_half
        self.copy_stream = torch.cuda.Stream(self.device)
        self._pinned = [dict() for _ in range(self.depth)]
        self._ready = [torch.cuda.Event() for _ in range(self.depth)]
        self._used = [False] * self.depth
        self.h2d_bytes = 0

    def __len__(self):
        return len(self.loader)

    def _stage(self, slot: int, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        if self._used[slot]:
            self._ready[slot].synchronize()   # the previous copy out of this slot's pinned buffers has finished
        self._used[slot] = True
        out = {}
        with torch.cuda.stream(self.copy_stream):
            for k, v in batch.items():
                if not torch.is_tensor(v):
                    out[k] = v
                    continue
                buf = self._pinned[slot].get(k)
                if buf is None or buf.shape != v.shape or buf.dtype != v.dtype:
                    buf = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
                    self._pinned[slot][k] = buf
                buf.copy_(v)
                d = buf.to(self.device, non_blocking=True)
                self.h2d_bytes += buf.numel() * buf.element_size()
                if self.widen_half and d.dtype == torch.float16:
                    d = d.float()
                out[k] = d
            self._ready[slot].record(self.copy_stream)
        return out

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        it = iter(self.loader)
        pending = []   # (slot, device dict), oldest first
        slot = 0

        def fill():
            nonlocal slot
            try:
                batch = next(it)
            except StopIteration:
                return False
            pending.append((slot, self._stage(slot, batch)))
            slot = (slot + 1) % self.depth
            return True

        for _ in range(self.depth - 1):
            if not fill():
                break
        while pending:
            s, dev_batch = pending.pop(0)
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(self._ready[s])
            for v in dev_batch.values():
                if torch.is_tensor(v):
                    v.record_stream(cur)
            fill()                       # next batch's copy is in flight while the consumer works on this one
            yield dev_batch

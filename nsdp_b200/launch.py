"""Run an UNCHANGED reference script (train.py / test.py / run.py) on top of nsdp_b200:

    python -m nsdp_b200.launch /path/to/NSDP/train.py config.yaml [script args...]
    python -m torch.distributed.run --nproc-per-node 8 -m nsdp_b200.launch /path/to/NSDP/train.py config.yaml ...

The scripts do `from model import build_model, optimizer_factory`, `import model.learningrate` and (through the
model files) `import pointnet2_ops_lib.pointnet2_ops.pointnet2_utils` (train.py:12-13, model/encoder/blocks.py:15).
Those names are pre-bound in sys.modules to the nsdp_b200 mirrors, so the script's own `model/` directory is never
imported. Under torchrun every rank pins itself to its GPU via CUDA_VISIBLE_DEVICES so that the script's hard-wired
`cuda:0` (train.py:74-77) is the rank's device; build_model() then joins the process group and the train_on_batch_*
functions all-reduce gradients (nsdp_b200/dist.py). For `train.py` every rank also gets its own `--seed` (base + rank:
the script builds its own shuffled DataLoader without a sampler, train.py:121-127, so distinct seeds give distinct
batches) and ranks > 0 their own `experiment.out_dir` (a rank-suffixed copy of the YAML config), otherwise all ranks
would race on the same `model_%05d` / `opt_%05d` files (utils/checkpoints.py:35-43).
"""
from __future__ import annotations

import os
import runpy
import sys
import types


def install_aliases() -> None:
    import nsdp_b200.model as model
    import nsdp_b200.pointnet2_ops as p2
    sys.modules["model"] = model
    for sub in ("deformation_networks", "flow_arbitrary", "learningrate", "utils", "encoder", "decoder"):
        sys.modules[f"model.{sub}"] = __import__(f"nsdp_b200.model.{sub}", fromlist=["_"])
    sys.modules["pointnet2_ops"] = p2
    sys.modules["pointnet2_ops._ext"] = p2._ext
    sys.modules["pointnet2_ops.pointnet2_utils"] = p2.pointnet2_utils
    sys.modules["pointnet2_ops.pointnet2_modules"] = p2.pointnet2_modules
    lib = types.ModuleType("pointnet2_ops_lib")
    lib.pointnet2_ops = p2
    lib.__path__ = []
    sys.modules["pointnet2_ops_lib"] = lib
    sys.modules["pointnet2_ops_lib.pointnet2_ops"] = p2
    sys.modules["pointnet2_ops_lib.pointnet2_ops.pointnet2_utils"] = p2.pointnet2_utils
    sys.modules["pointnet2_ops_lib.pointnet2_ops.pointnet2_modules"] = p2.pointnet2_modules


def mirror_checkpoints(shared_dir: str, private_dir: str) -> None:
    """Resume must give every replica the SAME weights, optimizer state and epoch (utils/checkpoints.py:8-31 picks the
    newest model_%05d/opt_%05d and modelbest_* of the directory it is pointed at): a rank's private directory only
    receives its WRITES; what it READS at start-up are links to rank 0's checkpoint files, replacing whatever an earlier
    run with another world size left there. Rank 0 writes nothing before the first epoch ends, so all ranks list the
    same files. (nsdp_b200.dist.sync_training_state additionally re-broadcasts rank 0's state on the first step.)"""
    os.makedirs(private_dir, exist_ok=True)
    is_ckpt = lambda f: f.startswith(("model_", "opt_", "modelbest_"))   # noqa: E731
    for f in os.listdir(private_dir):
        if is_ckpt(f):
            os.remove(os.path.join(private_dir, f))
    if os.path.isdir(shared_dir):
        for f in os.listdir(shared_dir):
            if is_ckpt(f):
                os.symlink(os.path.abspath(os.path.join(shared_dir, f)), os.path.join(private_dir, f))


def rewrite_argv_for_rank(argv, rank: int, world: int, scratch_dir: str):
    """argv = [train.py, config.yaml, ...] -> the same with `--seed base+rank` and, for rank > 0, a copy of the config whose
    `experiment.out_dir` is `<out_dir>/rank<r>`. Other scripts and single-process runs are returned unchanged."""
    argv = list(argv)
    if world <= 1 or not argv or os.path.basename(argv[0]) != "train.py":
        return argv
    base = 27                                            # train.py:45 default
    if "--seed" in argv[:-1]:
        i = argv.index("--seed")
        base = int(argv[i + 1])
        del argv[i:i + 2]
    else:
        for i, tok in enumerate(argv):
            if tok.startswith("--seed="):
                base = int(tok.split("=", 1)[1])
                del argv[i]
                break
    argv += ["--seed", str(base + rank)]
    if rank > 0:
        pos = [i for i, tok in enumerate(argv[1:], 1) if not tok.startswith("-") and (i == 1 or not argv[i - 1].startswith("--"))]
        if pos:
            import yaml
            ci = pos[0]
            with open(argv[ci]) as f:
                cfg = yaml.safe_load(f)
            shared = os.path.join(str(cfg["experiment"]["out_dir"]), str(cfg["experiment"].get("name", "")))
            cfg["experiment"]["out_dir"] = os.path.join(str(cfg["experiment"]["out_dir"]), f"rank{rank}")
            mirror_checkpoints(shared, os.path.join(cfg["experiment"]["out_dir"], str(cfg["experiment"].get("name", ""))))
            os.makedirs(scratch_dir, exist_ok=True)
            out = os.path.join(scratch_dir, f"rank{rank}_" + os.path.basename(argv[ci]))
            with open(out, "w") as f:
                yaml.safe_dump(cfg, f)
            argv[ci] = out
    return argv


def device_for_local_rank(local_rank: int, visible) -> str:
    """The CUDA_VISIBLE_DEVICES value that makes the script's hard-wired `cuda:0` this rank's GPU: the local_rank-th entry of
    the devices the job was given (indices or UUIDs), or the plain index when the job sees every GPU."""
    ids = [v.strip() for v in (visible or "").split(",") if v.strip()]
    if not ids:
        return str(local_rank)
    if local_rank >= len(ids):
        raise SystemExit(f"nsdp_b200.launch: LOCAL_RANK={local_rank} but CUDA_VISIBLE_DEVICES lists {len(ids)} device(s)")
    return ids[local_rank]


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m nsdp_b200.launch <script.py> [args...]")
    local_rank = os.environ.get("LOCAL_RANK")
    if local_rank is not None and "NSDP_B200_KEEP_VISIBLE" not in os.environ:
        os.environ["CUDA_VISIBLE_DEVICES"] = device_for_local_rank(int(local_rank), os.environ.get("CUDA_VISIBLE_DEVICES"))
    install_aliases()
    argv = rewrite_argv_for_rank(argv, int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
                                 os.path.join(os.environ.get("TMPDIR", "/tmp"), f"nsdp_b200_launch_{os.getpid()}"))
    script = argv[0]
    sys.argv = argv
    sys.path.insert(0, os.path.dirname(os.path.abspath(script)))
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()

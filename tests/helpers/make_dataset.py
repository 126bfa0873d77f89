"""Writes a tiny DeformingThings4D-shaped dataset + split lists + a YAML config to disk, in the formats the reference's
Deform4DFlow_Dataset reads (dataset/dataset_deform4d_flow.py:36-172, dataset/utils.py:8-26): per frame directory
`orig_to_gaps.txt` (4x4), `surface_points.npz` {points, normals: float16 (N,3)}, `flow.npz` {points: float16 (Q,3)},
`mesh_orig.obj`; every array is in material correspondence across the frames of a sequence (index i = the same point)."""
import os

import numpy as np
import yaml


def _surface(rng, n):
    v = rng.standard_normal((n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    th, ph = np.arccos(np.clip(v[:, 2], -1, 1)), np.arctan2(v[:, 1], v[:, 0])
    return v * (0.35 * (1 + 0.3 * np.sin(3 * th) * np.cos(2 * ph)))[:, None], v


def _deform(p, t):
    return p + 0.04 * t * np.stack([np.sin(4 * p[:, 1] + t), np.cos(3 * p[:, 2] - t), np.sin(5 * p[:, 0] + 2 * t)], 1)


def _octa_sphere():
    v = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], float)
    f = np.array([[0, 2, 4], [2, 1, 4], [1, 3, 4], [3, 0, 4], [2, 0, 5], [1, 2, 5], [3, 1, 5], [0, 3, 5]])
    for _ in range(3):                                  # 3 subdivisions: 258 vertices, 512 faces
        mid, nf, v = {}, [], list(map(tuple, v))

        def m(a, b):
            key = (min(a, b), max(a, b))
            if key not in mid:
                p = (np.array(v[a]) + np.array(v[b])) / 2
                v.append(tuple(p / np.linalg.norm(p)))
                mid[key] = len(v) - 1
            return mid[key]
        for a, b, c in f:
            ab, bc, ca = m(a, b), m(b, c), m(c, a)
            nf += [[a, ab, ca], [b, bc, ab], [c, ca, bc], [ab, bc, ca]]
        v, f = np.array(v), np.array(nf)
    return v * 0.35, f


def write(root, n_surf=1500, n_space=2400, frames=(0, 3, 6, 9), model_type="forward", epochs=2, batch_size=2,
          num_surf_samples=1024, num_space_samples=1536):
    rng = np.random.default_rng(7)
    data_dir, split_dir, out_dir = (os.path.join(root, d) for d in ("data", "splits", "out"))
    surf, nrm = _surface(rng, n_surf)
    base, bn = _surface(rng, n_space)
    space = base + bn * rng.uniform(-1, 1, (n_space, 1)) * 0.05
    verts, faces = _octa_sphere()
    for seq in ("bear_idle", "bear_run"):
        for fr in frames:
            d = os.path.join(data_dir, seq, "%04d" % fr)
            os.makedirs(d, exist_ok=True)
            t = 0.0 if seq == "bear_idle" else fr / 3.0
            np.savetxt(os.path.join(d, "orig_to_gaps.txt"), np.eye(4))
            np.savez(os.path.join(d, "surface_points.npz"), points=_deform(surf, t).astype(np.float16), normals=nrm.astype(np.float16))
            np.savez(os.path.join(d, "flow.npz"), points=_deform(space, t).astype(np.float16))
            with open(os.path.join(d, "mesh_orig.obj"), "w") as f:
                for v in _deform(verts, t):
                    f.write("v %.6f %.6f %.6f\n" % tuple(v))
                for a in faces + 1:
                    f.write("f %d %d %d\n" % tuple(a))
    os.makedirs(os.path.join(split_dir, "deform4d"), exist_ok=True)
    for name, seqs in (("identity_seen", ["bear_idle"]), ("train_seen", ["bear_run"]), ("test_unseen_motions", ["bear_run"])):
        with open(os.path.join(split_dir, "deform4d", name + ".lst"), "w") as f:
            f.write("\n".join(seqs) + "\n")
    from nsdp_b200 import synth
    cfg = {
        "experiment": {"out_dir": out_dir, "name": "harness"},
        "data": {"type": "deform4d", "dataset_dir": data_dir, "split_dir": split_dir, "interval": 3,
                 "arbitrary": model_type == "arbitrary", "inverse": model_type == "backward", "fix_coord_system": False,
                 "num_surf_samples": num_surf_samples, "num_space_samples": num_space_samples, "partial_range": 0.1,
                 "noise_level": 0.0, "partial_shape_ratio": 1.0, "norm_params_file": "orig_to_gaps.txt",
                 "surface_flow_file": "surface_points.npz", "space_flow_file": "flow.npz", "mesh_file": "mesh_orig.obj"},
        "model": synth.make_config(model_type)["model"],
        "training": {"iden_split": "identity_seen", "motion_split": "train_seen", "load_mesh": False, "num_sampled_pairs": -1,
                     "epochs": epochs, "save_frequency": 1, "batch_size": batch_size, "optimizer": "Adam", "lr": 5e-4,
                     "lr_step": 200, "lr_decay": 0.1, "weight_decay": 0.0},
        "validation": {"iden_split": "identity_seen", "motion_split": "test_unseen_motions", "load_mesh": False,
                       "num_sampled_pairs": -1, "frequency": 1, "batch_size": 2},
        "test": {"iden_split": "identity_seen", "motion_split": "test_unseen_motions", "load_mesh": True, "num_sampled_pairs": 2,
                 "batch_size": 1, "generate_mesh": False, "mesh_folder": "meshes", "mesh_format": "ply",
                 "generate_pointcloud": False, "pointcloud_folder": "pointclouds", "pointcloud_format": "ply",
                 "weight_file": os.path.join(out_dir, "harness", "model_%05d" % (epochs - 1))},
        "logger": {"type": "none", "project": "NSDP"},
    }
    path = os.path.join(root, "harness.yaml")
    with open(path, "w") as f:
        yaml.safe_dump(cfg, f)
    return path, cfg

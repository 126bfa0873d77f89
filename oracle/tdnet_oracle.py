"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

CPU restatement (torch tensors on the CPU, fp32 or fp64) of the reference's TDNet forward
path, written functionally over a flat ``state_dict`` so it shares no module code with
either the reference or the product. Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.

Parity status: the reference has no tests / golden vectors for this path (SURVEY.md §4,
§8c). This restatement is pinned against the LIVE reference instead:
``tests/golden/make_golden.py`` imports ``/root/reference/model`` in the authoring container,
runs it on seeded inputs + name-keyed seeded weights and commits the outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` re-checks this file against those vectors
everywhere (the GPU box has no ``/root/reference``).

Every function cites the reference file:line (relative to /root/reference/) it follows.
Autograd works through everything here (plain torch ops), which is how gradient parity
is checked.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        from oracle.build import build
        _LIB = ctypes.CDLL(build())
    return _LIB


# --------------------------------------------------------------------------------------
# index kernels (C restatement, oracle/nsdp_oracle.c)
# --------------------------------------------------------------------------------------
def fps(xyz: torch.Tensor, m: int) -> torch.Tensor:
    """sampling_gpu.cu:69-173 + sampling.cpp:66-87 -> (B, m) int32.
    CUDA tensors (bench.py --impl reference --ref-device cuda only): the reference's own kernel, rebuilt by
    oracle/build_ref.py."""
    if xyz.is_cuda:
        from oracle import ref_ext
        ext = ref_ext.load()
        if ext is None:
            raise RuntimeError("oracle/_ref is not built: the GPU run of the reference path needs the reference's FPS kernel")
        return ext.furthest_point_sampling(xyz.contiguous(), m)
    x = np.ascontiguousarray(xyz.detach().cpu().to(torch.float32).numpy())
    B, N, _ = x.shape
    out = np.zeros((B, m), dtype=np.int32)
    rc = _lib().nsdp_oracle_fps(x.ctypes.data_as(ctypes.c_void_p), B, N, m,
                                out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return torch.from_numpy(out)


def knn(query: torch.Tensor, ref: torch.Tensor, k: int, return_d2: bool = False):
    """model/utils.py:39-55 + argsort()[:, :, :k] (encoder/blocks.py:101-102) -> (B, M, k) int32,
    ties resolved lowest-index-first. CUDA tensors (GPU timing of the reference op chain only): the reference's own torch
    formula, materialising [B, M, N, 3] and sorting every row."""
    if query.is_cuda:
        d2 = torch.sum((query[:, :, None] - ref[:, None]) ** 2, dim=-1)
        idx = d2.argsort()[:, :, :k]
        return (idx, torch.gather(d2, 2, idx)) if return_d2 else idx
    q = np.ascontiguousarray(query.detach().cpu().to(torch.float32).numpy())
    r = np.ascontiguousarray(ref.detach().cpu().to(torch.float32).numpy())
    B, M, _ = q.shape
    N = r.shape[1]
    out = np.zeros((B, M, k), dtype=np.int32)
    d2 = np.zeros((B, M, k), dtype=np.float32)
    rc = _lib().nsdp_oracle_knn(q.ctypes.data_as(ctypes.c_void_p), r.ctypes.data_as(ctypes.c_void_p),
                                B, M, N, k, out.ctypes.data_as(ctypes.c_void_p),
                                d2.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0, rc
    if return_d2:
        return torch.from_numpy(out), torch.from_numpy(d2)
    return torch.from_numpy(out)


def ball_query(new_xyz: torch.Tensor, xyz: torch.Tensor, radius: float, nsample: int) -> torch.Tensor:
    """ball_query_gpu.cu:9-44 -> (B, M, nsample) int32."""
    c = np.ascontiguousarray(new_xyz.detach().cpu().to(torch.float32).numpy())
    p = np.ascontiguousarray(xyz.detach().cpu().to(torch.float32).numpy())
    B, M, _ = c.shape
    N = p.shape[1]
    out = np.zeros((B, M, nsample), dtype=np.int32)
    rc = _lib().nsdp_oracle_ball_query(c.ctypes.data_as(ctypes.c_void_p), p.ctypes.data_as(ctypes.c_void_p),
                                       B, N, M, ctypes.c_float(radius), nsample,
                                       out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return torch.from_numpy(out)


def three_nn(unknown: torch.Tensor, known: torch.Tensor):
    """interpolate_gpu.cu:9-59 -> (dist2 (B, n, 3) f32, idx (B, n, 3) int32)."""
    u = np.ascontiguousarray(unknown.detach().cpu().to(torch.float32).numpy())
    kn = np.ascontiguousarray(known.detach().cpu().to(torch.float32).numpy())
    B, n, _ = u.shape
    m = kn.shape[1]
    d2 = np.zeros((B, n, 3), dtype=np.float32)
    idx = np.zeros((B, n, 3), dtype=np.int32)
    rc = _lib().nsdp_oracle_three_nn(u.ctypes.data_as(ctypes.c_void_p), kn.ctypes.data_as(ctypes.c_void_p),
                                     B, n, m, d2.ctypes.data_as(ctypes.c_void_p),
                                     idx.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return torch.from_numpy(d2), torch.from_numpy(idx)


# numpy restatements of the pure gather/scatter ops (group_points_gpu.cu, sampling_gpu.cu:8-57,
# interpolate_gpu.cu:72-143); layouts are the reference's channel-major (B, C, N).
def gather_points(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """sampling_gpu.cu:8-20: out[b,c,j] = points[b,c,idx[b,j]]."""
    return torch.gather(points, 2, idx.long()[:, None, :].expand(-1, points.shape[1], -1))


def group_points(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """group_points_gpu.cu:8-28: out[b,c,j,k] = points[b,c,idx[b,j,k]]."""
    B, C, N = points.shape
    _, M, K = idx.shape
    flat = idx.long().reshape(B, 1, M * K).expand(-1, C, -1)
    return torch.gather(points, 2, flat).reshape(B, C, M, K)


def three_interpolate(points: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """interpolate_gpu.cu:72-101: out[b,c,j] = sum_t points[b,c,idx[b,j,t]] * weight[b,j,t]
    (evaluated left to right like the kernel: (p1*w1 + p2*w2) + p3*w3 with FMA contraction
    left to the compiler -> compared with a tolerance, not bit-exactly)."""
    g = group_points(points, idx)  # B, C, n, 3
    return (g * weight[:, None]).sum(-1)


# --------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------
def _lin(sd, p, x, bias=True):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"] if bias else None)


def _mlp2(sd, p, x):
    """nn.Sequential(Linear, ReLU, Linear) — e.g. encoder/blocks.py:69-79."""
    return _lin(sd, p + ".2", F.relu(_lin(sd, p + ".0", x)))


def _conv1(sd, p, x_bnc):
    """nn.Conv1d(dim, dim, 1) applied to a (B, n, C) tensor (the reference permutes to (B, C, n),
    encoder/blocks.py:158, 300)."""
    return F.linear(x_bnc, sd[p + ".weight"][:, :, 0], sd[p + ".bias"])


def _bn(sd, p, x_bnc, training: bool):
    """nn.BatchNorm1d over (B, C, n) == statistics over B*n per channel (encoder/blocks.py:132,159).
    Train mode updates running stats in `sd` in place exactly like the module does
    (momentum 0.1, unbiased running_var, num_batches_tracked += 1)."""
    x = x_bnc.permute(0, 2, 1)
    if training:
        sd[p + ".num_batches_tracked"] += 1
    y = F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                     sd[p + ".bias"], training, 0.1, 1e-5)
    return y.permute(0, 2, 1)


def index_points(points, idx):
    """model/utils.py:58-70."""
    raw = idx.shape
    flat = idx.reshape(raw[0], -1).long()
    res = torch.gather(points, 1, flat[..., None].expand(-1, -1, points.shape[-1]))
    return res.reshape(*raw, -1)


def _vattn(sd, p_delta, p_gamma, q, k_g, v_g, rel_xyz, pos_only=False):
    """Vector attention core shared by encoder/blocks.py:114-126 and :295-308:
    delta = MLP_delta(rel); a = MLP_gamma(q - k + delta); w = softmax over neighbours (dim=-2);
    out = sum_j w * (v + delta)."""
    pos = _mlp2(sd, p_delta, rel_xyz)
    if pos_only:
        attn = _mlp2(sd, p_gamma, pos)
        val = pos
    else:
        attn = _mlp2(sd, p_gamma, q[:, :, None] - k_g + pos)
        val = v_g + pos
    attn = F.softmax(attn, dim=-2)
    return (attn * val).sum(dim=2)


# --------------------------------------------------------------------------------------
# encoder blocks
# --------------------------------------------------------------------------------------
def transformer_block(sd, p, xyz, feats, k, pos_only=False, group_all=False, training=False,
                      knn_override=None):
    """TransformerBlock.forward, model/encoder/blocks.py:86-134."""
    B, n, _ = xyz.shape
    with torch.no_grad():
        if group_all:
            idx = torch.arange(n, device=xyz.device).view(1, 1, n).expand(B, n, n)
        elif knn_override is not None:
            idx = knn_override
        else:
            idx = knn(xyz, xyz, k).long()
    knn_xyz = index_points(xyz, idx)
    rel = xyz[:, :, None] - knn_xyz  # centre - neighbour, :114
    if pos_only:
        res = _vattn(sd, p + ".fc_delta", p + ".fc_gamma", None, None, None, rel, pos_only=True)
    else:
        q = _lin(sd, p + ".w_qs", feats, bias=False)
        kk = index_points(_lin(sd, p + ".w_ks", feats, bias=False), idx)
        vv = index_points(_lin(sd, p + ".w_vs", feats, bias=False), idx)
        res = _vattn(sd, p + ".fc_delta", p + ".fc_gamma", q, kk, vv, rel) + feats
    return _bn(sd, p + ".bn", res, training)


def elementwise_mlp(sd, p, x, training=False):
    """ElementwiseMLP.forward, model/encoder/blocks.py:153-159."""
    h = F.relu(_bn(sd, p + ".bn1", _conv1(sd, p + ".conv1", x), training))
    h = F.relu(_bn(sd, p + ".bn2", _conv1(sd, p + ".conv2", h), training))
    return _bn(sd, p + ".bn3", x + h, training)


def transformer_set_abstraction(sd, p, xyz, feats, npoint, k, training=False):
    """TransformerSetAbstraction.forward, model/encoder/blocks.py:270-314."""
    with torch.no_grad():
        fps_idx = fps(xyz, npoint).long()
        new_xyz = index_points(xyz, fps_idx)
        idx = knn(new_xyz, xyz, k).long()
    q = index_points(_lin(sd, p + ".w_qs", feats, bias=False), fps_idx)
    kk = index_points(_lin(sd, p + ".w_ks", feats, bias=False), idx)
    vv = index_points(_lin(sd, p + ".w_vs", feats, bias=False), idx)
    rel = index_points(xyz, idx) - new_xyz[:, :, None]  # neighbour - centre, :295
    pos = _mlp2(sd, p + ".fc_delta1", rel)
    attn = F.softmax(_mlp2(sd, p + ".fc_gamma1", q[:, :, None] - kk + pos), dim=-2)
    res1 = (attn * (vv + pos)).sum(dim=2)
    res1 = res1 + _conv1(sd, p + ".conv2", F.relu(_bn(sd, p + ".bn1", _conv1(sd, p + ".conv1", res1), training)))
    res1 = _bn(sd, p + ".bnorm0", res1, training)
    q2 = _lin(sd, p + ".w_qs2", res1, bias=False)
    kk2 = index_points(_lin(sd, p + ".w_ks2", feats, bias=False), idx)
    vv2 = index_points(_lin(sd, p + ".w_vs2", feats, bias=False), idx)
    attn2 = F.softmax(_mlp2(sd, p + ".fc_gamma2", q2[:, :, None] - kk2 + pos), dim=-2)
    res2 = (attn2 * (vv2 + pos)).sum(dim=2)
    out = _bn(sd, p + ".bnorm1", res1 + res2, training)
    out = out + index_points(feats, fps_idx)
    out = _bn(sd, p + ".bnorm2", out, training)
    return new_xyz, out


def encoder(sd, p, x, cfg, has_features, training=False, trace: Optional[dict] = None):
    """PointTransformerEncoder.forward, model/encoder/pointransformer.py:87-140.
    cfg = the YAML ``encoder_kwargs``."""
    npts = cfg["npoints_per_layer"]
    nn_, nn_red = cfg["nneighbor"], cfg["nneighbor_reduced"]
    d_t, d_r = cfg["d_transformer"], cfg["d_reduced"]
    if has_features:
        feats = _lin(sd, p + ".enc_sdf", x[:, :, 3:])
        xyz = x[:, :, :3].contiguous()
        feats = transformer_block(sd, p + ".transformer_begin", xyz, feats, nn_red, training=training)
    else:
        xyz = x
        feats = transformer_block(sd, p + ".transformer_begin", xyz, None, nn_red, pos_only=True,
                                  training=training)
    if trace is not None:
        trace["begin"] = feats
    for i in range(len(npts) - 1):
        k_sa = min(nn_, npts[i])
        k_tb = min(nn_, npts[i + 1])
        xyz, feats = transformer_set_abstraction(sd, f"{p}.transition_downs.{i}.sa", xyz, feats,
                                                 npts[i + 1], k_sa, training)
        if trace is not None:
            trace[f"sa{i}"] = feats
            trace[f"xyz{i}"] = xyz
        feats = elementwise_mlp(sd, f"{p}.elementwise_extras.{i}", feats, training)
        feats = transformer_block(sd, f"{p}.transformer_downs.{i}", xyz, feats, k_tb, training=training)
        if i == 0 and d_r != d_t:
            feats = _lin(sd, p + ".fc1", feats)
        feats = elementwise_mlp(sd, f"{p}.elementwise.{i}", feats, training)
        if trace is not None:
            trace[f"level{i}"] = feats
    for i in range(cfg["nfinal_transformers"]):
        feats = transformer_block(sd, f"{p}.final_transformers.{i}", xyz, feats, 2 * nn_,
                                  group_all=cfg.get("full_SA", False), training=training)
        feats = elementwise_mlp(sd, f"{p}.final_elementwise.{i}", feats, training)
    z = _mlp2(sd, p + ".fc_middle", feats.max(dim=1)[0])
    return {"z": z, "anchors": xyz, "anchor_feats": feats}


# --------------------------------------------------------------------------------------
# decoder
# --------------------------------------------------------------------------------------
def _note_kink(kink, pre, rows_dim0=2):
    """Test aid: per query, the smallest |ReLU pre-activation| seen so far, relative to that layer's rms (a gradient is
    discontinuous where a pre-activation crosses 0, so parity tests exclude queries sitting on a kink on BOTH sides)."""
    if kink is None:
        return
    with torch.no_grad():
        m = pre.detach().abs() / pre.detach().pow(2).mean().sqrt().clamp_min(1e-30)
        m = m.reshape(m.shape[0], m.shape[1], -1).min(dim=-1)[0]
        kink["margin"] = m if "margin" not in kink else torch.minimum(kink["margin"], m)


def cross_transformer_block(sd, p, xyz_q, z, anchors, anchor_feats, nneigh, kink=None):
    """CrossTransformerBlock.forward, model/decoder/blocks.py:48-95 (2-D lat_rep branch,
    reduce_dim=True, separate_delta=True — numerically the same delta used twice)."""
    with torch.no_grad():
        idx = knn(xyz_q, anchors, nneigh).long()
    B, Q, _ = xyz_q.shape
    q = _lin(sd, p + ".w_qs", z, bias=False)[:, None, None, :]           # B,1,1,d
    kg = _lin(sd, p + ".w_k_global", z, bias=False)[:, None, None, :].expand(-1, Q, -1, -1)
    vg = _lin(sd, p + ".w_v_global", z, bias=False)[:, None, None, :].expand(-1, Q, -1, -1)
    kk = torch.cat([index_points(_lin(sd, p + ".w_ks", anchor_feats, bias=False), idx), kg], dim=2)
    vv = torch.cat([index_points(_lin(sd, p + ".w_vs", anchor_feats, bias=False), idx), vg], dim=2)
    rel = xyz_q[:, :, None] - index_points(anchors, idx)
    pos = _mlp2(sd, p + ".fc_delta", rel)
    _note_kink(kink, _lin(sd, p + ".fc_delta.0", rel))
    pos = torch.cat([pos, torch.zeros(B, Q, 1, pos.shape[-1], dtype=pos.dtype, device=pos.device)], dim=2)
    _note_kink(kink, _lin(sd, p + ".fc_gamma.0", q - kk + pos))
    attn = F.softmax(_mlp2(sd, p + ".fc_gamma", q - kk + pos), dim=-2)
    return (attn * (vv + pos)).sum(dim=2)


def decoder(sd, p, xyz_q, enc, cfg, kink=None):
    """CrossTransformerDecoder.forward, model/decoder/crosstransformer_decoder.py:45-70 with
    ResnetBlockFC (model/decoder/blocks.py:133-142). cfg = the YAML ``decoder_kwargs``."""
    lat = cross_transformer_block(sd, p + ".ct1", xyz_q, enc["z"], enc["anchors"], enc["anchor_feats"],
                                  cfg.get("nneigh", 7), kink=kink)
    net = _lin(sd, p + ".init_enc", lat)
    for i in range(cfg.get("n_blocks", 5)):
        net = net + _lin(sd, f"{p}.fc_c.{i}", lat)
        _note_kink(kink, net)
        h = _lin(sd, f"{p}.blocks.{i}.fc_0", F.relu(net))
        _note_kink(kink, h)
        net = net + _lin(sd, f"{p}.blocks.{i}.fc_1", F.relu(h))
    _note_kink(kink, net)
    return _lin(sd, p + ".fc_out", F.relu(net))


# --------------------------------------------------------------------------------------
# whole networks
# --------------------------------------------------------------------------------------
def tdnet_forward(sd: Dict[str, torch.Tensor], prefix: str, points, surface, model_cfg, no_input_corr,
                  training=False, trace=None, kink=None):
    """Deformation_Networks.forward, model/deformation_networks.py:43-60. `prefix` is '' for a bare
    TDNet or 'model_deform.' / 'model_canonicalize.' inside FlowArbitrary."""
    if no_input_corr:
        enc = encoder(sd, prefix + "encoder", surface[:, :, 0:3].contiguous(), model_cfg["encoder_kwargs"],
                      has_features=False, training=training, trace=trace)
    else:
        enc = encoder(sd, prefix + "encoder", surface, model_cfg["encoder_kwargs"], has_features=True,
                      training=training, trace=trace)
    if trace is not None:
        trace.update({"z": enc["z"], "anchors": enc["anchors"], "anchor_feats": enc["anchor_feats"]})
    return decoder(sd, prefix + "decoder", points, enc, model_cfg["decoder_kwargs"], kink=kink)


def flow_arbitrary_forward(sd, space_src, surf_src, surf_tgt, mask, model_cfg, training=False):
    """FlowArbitrary.forward, model/flow_arbitrary.py:15-27 (canonicalise twice, then deform)."""
    space_c = tdnet_forward(sd, "model_canonicalize.", space_src, surf_src, model_cfg, True, training)
    surf_c = tdnet_forward(sd, "model_canonicalize.", surf_src, surf_src, model_cfg, True, training)
    inp = torch.cat([surf_c, surf_tgt, mask], dim=-1).contiguous()
    return tdnet_forward(sd, "model_deform.", space_c, inp, model_cfg, False, training)


def l2_loss(pred, gt):
    """compute_l2_error, model/utils.py:8-11."""
    return torch.mean((pred - gt).pow(2).sum(dim=2) / 2.0)

"""GPU parity of the index kernels (FPS, k-NN, ball query, 3-NN, gather/group/interpolate) through the C ABI:
bit-exact against (a) the C restatement oracle/nsdp_oracle.c and (b) the REFERENCE's own CUDA extension
compiled from /root/reference into oracle/_ref/ (the real pointnet2_ops kernels, rebuilt for sm_100a)."""
import numpy as np
import pytest
import torch

from nsdp_b200 import ops, synth
from oracle import ref_ext
from oracle import tdnet_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _clouds():
    rng = np.random.default_rng(0)
    lattice = rng.integers(-4, 5, size=(2, 1500, 3)).astype(np.float32) * 0.125     # massive exact ties + duplicates
    near0 = rng.uniform(-0.05, 0.05, size=(2, 700, 3)).astype(np.float32)          # many |p|^2 <= 1e-3 points
    near0[:, ::3] *= 10
    allskip = rng.uniform(-0.01, 0.01, size=(1, 300, 3)).astype(np.float32)        # every point skipped
    return {
        "bumpy_fp32_4096": (synth.surface_cloud(3, 4096, seed=1, fp16_grid=False), 500),
        "bumpy_fp16_4096": (synth.surface_cloud(3, 4096, seed=2, fp16_grid=True), 500),
        "bumpy_fp16_5000": (synth.surface_cloud(2, 5000, seed=3, fp16_grid=True), 500),
        "bumpy_fp16_500": (synth.surface_cloud(4, 500, seed=4, fp16_grid=True), 100),
        "tiny_37": (synth.surface_cloud(2, 37, seed=5, fp16_grid=True), 20),
        "single_point": (synth.surface_cloud(2, 1, seed=6), 1),
        "n_2": (synth.surface_cloud(2, 2, seed=6), 2),
        "lattice_ties": (torch.from_numpy(lattice), 400),
        "near_origin_skip": (torch.from_numpy(near0), 128),
        "all_skipped": (torch.from_numpy(allskip), 16),
        "uniform_1000": (torch.rand(2, 1000, 3) - 0.5, 250),
        "n_8192": (synth.surface_cloud(1, 8192, seed=7, fp16_grid=True), 300),
        "cluster_20000": (synth.surface_cloud(2, 20000, seed=8, fp16_grid=True), 600),
        "cluster_100k_fp16": (synth.surface_cloud(1, 100000, seed=9, fp16_grid=True), 1024),
    }


@pytest.mark.parametrize("name", list(_clouds().keys()))
def test_fps_bit_exact(name):
    xyz, m = _clouds()[name]
    got = ops.furthest_point_sampling(xyz.to(DEV).contiguous(), m).cpu()
    want = orc.fps(xyz, m)
    assert got.dtype == torch.int32 and tuple(got.shape) == (xyz.shape[0], m)
    assert torch.equal(got, want), f"{name}: first mismatch at {(got != want).nonzero()[:3].tolist()}"
    ref = ref_ext.load()
    if ref is not None:
        theirs = ref.furthest_point_sampling(xyz.to(DEV).contiguous(), m).cpu()
        assert torch.equal(theirs, want), f"{name}: C oracle differs from the reference CUDA kernel"
        assert torch.equal(got, theirs)


def test_fps_full_c5_size_against_reference_kernel():
    """BASELINE config #5: 100k -> 4096 (fp16-grid cloud => real arg-max ties). The reference kernel itself is
    the checker at this size (the C oracle needs ~1 s per 1000 samples here)."""
    ref = ref_ext.load()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    xyz = synth.surface_cloud(1, 100000, seed=11, fp16_grid=True).to(DEV)
    got = ops.furthest_point_sampling(xyz, 4096)
    theirs = ref.furthest_point_sampling(xyz, 4096)
    assert torch.equal(got, theirs)
    assert got[0].unique().numel() == 4096  # idempotence-style property: no point is picked twice


def test_fps_rejects_cpu_and_bad_dtype():
    with pytest.raises(RuntimeError):
        ops.furthest_point_sampling(torch.rand(1, 10, 3), 4)
    with pytest.raises(RuntimeError):
        ops.furthest_point_sampling(torch.rand(1, 10, 3, device=DEV).double(), 4)
    with pytest.raises(RuntimeError):
        ops.furthest_point_sampling(torch.rand(1, 10, 6, device=DEV)[:, :, :3], 4)  # non-contiguous


@pytest.mark.parametrize("M,N,k", [(4096, 4096, 10), (500, 4096, 16), (500, 500, 16), (100, 500, 16), (100, 100, 16),
                                   (5000, 100, 7), (37, 37, 37), (64, 3000, 64), (300, 20000, 32), (1, 1, 1)])
def test_knn_matches_oracle(M, N, k):
    torch.manual_seed(M * 7 + N)
    ref_pts = torch.rand(2, N, 3) - 0.5
    q = ref_pts[:, :M].clone() if M <= N else torch.rand(2, M, 3) - 0.5
    if M == 5000:
        q = torch.rand(2, M, 3) - 0.5
    got, d2 = ops.knn(q.to(DEV), ref_pts.to(DEV), k, return_d2=True)
    want, wd2 = orc.knn(q, ref_pts, k, return_d2=True)
    assert torch.equal(got.cpu(), want)
    assert torch.equal(d2.cpu(), wd2)  # distances bit-identical (same association, no FMA)


@pytest.mark.parametrize("k", [16, 32, 64])
def test_knn_full_c5_size(k):
    """BASELINE.json configs[4]: 4096 queries (the FPS picks) against a 100 000-point cloud, k = 16 / 32 / 64; indices and
    distances bit-exact against the C oracle (B = 1 keeps the scalar oracle at a few seconds)."""
    xyz = synth.surface_cloud(1, 100000, seed=11, fp16_grid=False)
    q = xyz[:, torch.randperm(100000, generator=torch.Generator().manual_seed(k))[:4096]].contiguous()
    got, d2 = ops.knn(q.to(DEV), xyz.to(DEV), k, return_d2=True)
    want, wd2 = orc.knn(q, xyz, k, return_d2=True)
    assert torch.equal(got.cpu(), want)
    assert torch.equal(d2.cpu(), wd2)


def test_knn_ties_are_lowest_index_first():
    pts = (torch.randint(-3, 4, (2, 600, 3)).float() * 0.25)
    got, d2 = ops.knn(pts.to(DEV), pts.to(DEV), 16, return_d2=True)
    want, wd2 = orc.knn(pts, pts, 16, return_d2=True)
    assert torch.equal(got.cpu(), want)
    assert torch.equal(d2.cpu(), wd2)
    # against torch's own (unstable) path: equal distance multisets on every row
    dist = ((pts[:, :, None] - pts[:, None]) ** 2).sum(-1)
    assert torch.equal(dist.sort(dim=-1)[0][:, :, :16], d2.cpu())


def test_knn_matches_reference_idiom_on_tie_free_rows():
    """square_distance + argsort()[:, :, :k] exactly as the reference spells it (model/utils.py:55,
    encoder/blocks.py:101-102), on the CPU in fp32."""
    torch.manual_seed(3)
    x = torch.rand(2, 800, 3) - 0.5
    dist = torch.sum((x[:, :, None] - x[:, None]) ** 2, dim=-1)
    ref_idx = dist.argsort()[:, :, :10]
    got = ops.knn(x.to(DEV), x.to(DEV), 10).cpu().long()
    assert torch.equal(got, ref_idx)


@pytest.mark.parametrize("radius,nsample", [(0.1, 16), (0.2, 32), (0.02, 8), (5.0, 64)])
def test_ball_query(radius, nsample):
    xyz = synth.surface_cloud(3, 3000, seed=21, fp16_grid=True)
    centres = xyz[:, ::7].contiguous()
    got = ops.ball_query(centres.to(DEV), xyz.to(DEV), radius, nsample).cpu()
    want = orc.ball_query(centres, xyz, radius, nsample)
    assert torch.equal(got, want)
    ref = ref_ext.load()
    if ref is not None:
        assert torch.equal(ref.ball_query(centres.to(DEV), xyz.to(DEV), radius, nsample).cpu(), got)


def test_three_nn_and_interpolate():
    known = synth.surface_cloud(2, 700, seed=31, fp16_grid=True)
    unknown = synth.surface_cloud(2, 2500, seed=32, fp16_grid=True)
    d2, idx = ops.three_nn(unknown.to(DEV), known.to(DEV))
    wd2, widx = orc.three_nn(unknown, known)
    assert torch.equal(idx.cpu(), widx) and torch.equal(d2.cpu(), wd2)
    feats = torch.randn(2, 19, 700)
    w = torch.rand(2, 2500, 3)
    w = w / w.sum(-1, keepdim=True)
    out = ops.three_interpolate(feats.to(DEV), idx, w.to(DEV))
    torch.testing.assert_close(out.cpu(), orc.three_interpolate(feats, widx, w), atol=1e-6, rtol=1e-6)
    go = torch.randn(2, 19, 2500)
    gin = ops.three_interpolate_grad(go.to(DEV), idx, w.to(DEV), 700).cpu()
    f2 = feats.clone().requires_grad_(True)
    (orc.three_interpolate(f2, widx, w) * go).sum().backward()
    torch.testing.assert_close(gin, f2.grad, atol=1e-4, rtol=1e-4)
    ref = ref_ext.load()
    if ref is not None:
        rd2, ridx = ref.three_nn(unknown.to(DEV), known.to(DEV))
        assert torch.equal(ridx, idx) and torch.equal(rd2, d2)
        assert torch.equal(ref.three_interpolate(feats.to(DEV), idx, w.to(DEV)), out)  # same FMA contraction


def test_gather_and_group():
    feats = torch.randn(3, 23, 900)
    idx1 = torch.randint(0, 900, (3, 257), dtype=torch.int32)
    idx2 = torch.randint(0, 900, (3, 120, 16), dtype=torch.int32)
    g1 = ops.gather_points(feats.to(DEV), idx1.to(DEV))
    assert torch.equal(g1.cpu(), orc.gather_points(feats, idx1))
    g2 = ops.group_points(feats.to(DEV), idx2.to(DEV))
    assert torch.equal(g2.cpu(), orc.group_points(feats, idx2))
    go1, go2 = torch.randn(3, 23, 257), torch.randn(3, 23, 120, 16)
    f = feats.clone().requires_grad_(True)
    (orc.gather_points(f, idx1) * go1).sum().backward()
    torch.testing.assert_close(ops.gather_points_grad(go1.to(DEV), idx1.to(DEV), 900).cpu(), f.grad, atol=1e-5, rtol=1e-5)
    f = feats.clone().requires_grad_(True)
    (orc.group_points(f, idx2) * go2).sum().backward()
    torch.testing.assert_close(ops.group_points_grad(go2.to(DEV), idx2.to(DEV), 900).cpu(), f.grad, atol=1e-4, rtol=1e-4)
    ref = ref_ext.load()
    if ref is not None:
        assert torch.equal(ref.gather_points(feats.to(DEV), idx1.to(DEV)), g1)
        assert torch.equal(ref.group_points(feats.to(DEV), idx2.to(DEV)), g2)


# ---------------------------------------------------------------------------------------------------------------------
# vectors minted by the LIVE reference's own torch code (tests/golden/make_golden_index.py; inputs stored in the fixture)
# ---------------------------------------------------------------------------------------------------------------------
def _index_gold():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "index_reference.npz"))


def test_fps_equals_reference_torch_fps_vectors():
    """model/utils.py:73-93 `farthest_point_sample` with start index 0 — and, where oracle/_ref exists, the reference's CUDA
    kernel on the same stored clouds."""
    gold = _index_gold()
    names = sorted({k.split("::")[1] for k in gold.keys() if k.startswith("fps::")})
    assert len(names) == 6
    ref = ref_ext.load()
    for name in names:
        xyz = torch.from_numpy(gold[f"fps::{name}::xyz"]).to(DEV).contiguous()
        want = torch.from_numpy(gold[f"fps::{name}::idx"])
        got = ops.furthest_point_sampling(xyz, want.shape[1]).cpu()
        assert torch.equal(got, want), f"{name}: first mismatch at {(got != want).nonzero()[:3].tolist()}"
        if ref is not None:
            assert torch.equal(ref.furthest_point_sampling(xyz, want.shape[1]).cpu(), want), name


def test_knn_equals_reference_argsort_vectors():
    """square_distance(q, r).argsort()[:, :, :k] of the live reference (model/utils.py:39-55, encoder/blocks.py:101-102) on
    rows without distance ties: indices and distances bit-exact."""
    gold = _index_gold()
    names = sorted({k.split("::")[1] for k in gold.keys() if k.startswith("knn::")})
    assert len(names) == 4
    for name in names:
        q = torch.from_numpy(gold[f"knn::{name}::query"]).to(DEV).contiguous()
        r = torch.from_numpy(gold[f"knn::{name}::ref"]).to(DEV).contiguous()
        want = torch.from_numpy(gold[f"knn::{name}::idx"])
        got, d2 = ops.knn(q, r, want.shape[2], return_d2=True)
        assert torch.equal(got.cpu(), want), name
        assert torch.equal(d2.cpu(), torch.from_numpy(gold[f"knn::{name}::d2"])), name

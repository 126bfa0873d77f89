"""Pins the tcgen05 descriptor / canonical-layout / TMEM-lane conventions of nsdp_b200/csrc/umma.cuh on hardware."""
import pytest
import torch

from nsdp_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(N, K, split, seed=0, mn=0):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(K, 128, generator=g) if mn else torch.randn(128, K, generator=g)
    B = torch.randn(K, N, generator=g) if mn else torch.randn(N, K, generator=g)
    Ad, Bd = A.to(DEV), B.to(DEV)
    D = torch.zeros(128, N, device=DEV)
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    rc = _lib.lib().nsdp_selftest_umma(Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), N, K, split, mn, err.data_ptr(),
                                       torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "nsdp_selftest_umma")
    torch.cuda.synchronize()
    assert int(err.item()) == 0, "mbarrier wait timed out"
    return A, B, D.cpu()


@pytest.mark.parametrize("N,K", [(16, 16), (128, 64), (208, 128), (256, 128), (64, 256)])
def test_umma_bf16(N, K):
    A, B, D = _run(N, K, 0)
    want = A.bfloat16().double() @ B.bfloat16().double().t()
    assert (D.double() - want).abs().max().item() < 1e-3 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("N,K", [(128, 64), (208, 128), (256, 128)])
def test_umma_bf16x3_is_fp32_grade(N, K):
    A, B, D = _run(N, K, 1, seed=3)
    want = A.double() @ B.double().t()
    rel = ((D.double() - want).norm() / want.norm()).item()
    assert rel < 3e-5, rel   # plain bf16 is ~3e-3 here


@pytest.mark.parametrize("N,K", [(16, 16), (208, 128), (256, 64), (128, 128)])
def test_umma_mn_major_operands(N, K):
    """D = X^T Y with X (K,128), Y (K,N) stored as row tiles and consumed as MN-major operands."""
    X, Y, D = _run(N, K, 1, seed=7, mn=1)
    want = X.double().t() @ Y.double()
    rel = ((D.double() - want).norm() / want.norm()).item()
    assert rel < 3e-5, rel


def _run_pair(N, K, split, seed=0):
    g = torch.Generator().manual_seed(seed)
    A, B = torch.randn(256, K, generator=g), torch.randn(N, K, generator=g)
    Ad, Bd = A.to(DEV), B.to(DEV)
    D = torch.zeros(256, N, device=DEV)
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    rc = _lib.lib().nsdp_selftest_umma2(Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), N, K, split, err.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "nsdp_selftest_umma2")
    torch.cuda.synchronize()
    assert int(err.item()) == 0, "mbarrier wait timed out"
    return A, B, D.cpu()


@pytest.mark.parametrize("N,K", [(32, 16), (128, 64), (208, 128), (256, 128)])
def test_umma_cta_pair(N, K):
    """cta_group::2: one M = 256 product across two CTAs (each holds 128 rows of A and half of B's rows)."""
    A, B, D = _run_pair(N, K, 1, seed=11)
    want = A.double() @ B.double().t()
    rel = ((D.double() - want).norm() / want.norm()).item()
    assert rel < 3e-5, rel

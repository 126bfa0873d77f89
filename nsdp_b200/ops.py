"""torch-facing operators over the C ABI (include/nsdp_b200.h).

torch is plumbing here: it owns device memory and streams; every operator below passes raw device
pointers + the current CUDA stream into libnsdp_b200.so. CPU tensors are rejected loudly, exactly like
the reference extension ("CPU not supported", pointnet2_ops/_ext-src/src/sampling.cpp:82-84).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import EmlpArgs, EmlpGrads, MlpArgs, MlpGrads, TailArgs, TailGrads, VattnArgs, VattnGrads, check

LAUNCHES = 0  # number of libnsdp_b200 kernel-launching calls made (bench.py reports it)


def _count(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n


# 0 = auto (tcgen05 kernel where instantiated), 1 = fp32 CUDA-core kernels only, 2 = require tcgen05
VATTN_IMPL = int(__import__("os").environ.get("NSDP_B200_VATTN_IMPL", "0"))

TAIL_IMPL = int(__import__("os").environ.get("NSDP_B200_TAIL_IMPL", "0"))

# Keep the decoder attention's pair-level operand tiles from forward to backward (nsdp_vattn_args::saved): the backward then
# skips the recomputation of the forward chain. OFF by default: measured on B200 at the bench size (8 x 50 000 queries) the
# backward gains 0.65 ms but the forward's extra 10.6 GB of HBM writes cost 1.0 ms (and 10.6 GB of memory); recomputing on
# chip is cheaper than a round trip through HBM on this part.
SAVE_ACTIVATIONS = int(__import__("os").environ.get("NSDP_B200_SAVE_ACTIVATIONS", "0")) != 0

def set_stage_format(fmt: str) -> str:
    """'fp16' (default) or 'bf16x2': staging format of the decoder backward's weight-gradient operand tiles
    (nsdp_set_stage_format in include/nsdp_b200.h). Returns the previous format."""
    code = {"bf16x2": 0, "fp16": 1}[fmt]
    return ("bf16x2", "fp16")[_lib.lib().nsdp_set_stage_format(code)]


def get_stage_format() -> str:
    return ("bf16x2", "fp16")[_lib.lib().nsdp_set_stage_format(-1)]


TIMING = False     # when True every kernel call below is bracketed by CUDA events on the launching stream
_TIMED = []        # (name, start_event, end_event)


class _timed:
    """Brackets one C-ABI call with CUDA events on the current stream (only when ops.TIMING is on)."""

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if TIMING:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if TIMING:
            self.b.record()
            _TIMED.append((self.name, self.a, self.b))
        return False


def timing_summary(reset: bool = True) -> dict:
    """{kernel name: {"calls": n, "ms": total}} of the calls recorded since the last reset (synchronises)."""
    torch.cuda.synchronize()
    out = {}
    for name, a, b in _TIMED:
        e = out.setdefault(name, {"calls": 0, "ms": 0.0})
        e["calls"] += 1
        e["ms"] += a.elapsed_time(b)
    if reset:
        _TIMED.clear()
    return out


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _chk_f32(name: str, t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: CPU not supported (nsdp_b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be a float tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    return t


def _chk_i32(name: str, t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: CPU not supported (nsdp_b200 has no CPU path)")
    if t.dtype != torch.int32:
        raise RuntimeError(f"{name} must be an int tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    return t


# ---------------------------------------------------------------------------------------------------
# Part 1: pointnet2_ops._ext operators (same names / argument order as bindings.cpp:6-19)
# ---------------------------------------------------------------------------------------------------
def furthest_point_sampling(points: torch.Tensor, nsamples: int) -> torch.Tensor:
    """(B,N,3) f32 -> (B,nsamples) i32; sampling.cpp:66-87."""
    _chk_f32("points", points)
    B, N, _ = points.shape
    out = torch.empty((B, nsamples), dtype=torch.int32, device=points.device)
    with torch.cuda.device(points.device), _timed(f"fps_N{N}_m{nsamples}"):
        check(_lib.lib().nsdp_fps_f32(points.data_ptr(), B, N, int(nsamples), out.data_ptr(), _stream()), "nsdp_fps_f32")
    _count()
    return out


def gather_points(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """(B,C,N), (B,M) i32 -> (B,C,M); sampling.cpp:16-41."""
    _chk_f32("points", points)
    _chk_i32("idx", idx)
    B, Cc, N = points.shape
    M = idx.shape[1]
    out = torch.empty((B, Cc, M), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        check(_lib.lib().nsdp_gather_points_f32(points.data_ptr(), idx.data_ptr(), B, Cc, N, M, out.data_ptr(), _stream()),
              "nsdp_gather_points_f32")
    _count()
    return out


def gather_points_grad(grad_out: torch.Tensor, idx: torch.Tensor, n: int) -> torch.Tensor:
    _chk_f32("grad_out", grad_out)
    _chk_i32("idx", idx)
    B, Cc, M = grad_out.shape
    out = torch.zeros((B, Cc, n), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        check(_lib.lib().nsdp_gather_points_grad_f32(grad_out.data_ptr(), idx.data_ptr(), B, Cc, n, M, out.data_ptr(),
                                                     _stream()), "nsdp_gather_points_grad_f32")
    _count()
    return out


def ball_query(new_xyz: torch.Tensor, xyz: torch.Tensor, radius: float, nsample: int) -> torch.Tensor:
    _chk_f32("new_xyz", new_xyz)
    _chk_f32("xyz", xyz)
    B, M, _ = new_xyz.shape
    N = xyz.shape[1]
    out = torch.empty((B, M, nsample), dtype=torch.int32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        check(_lib.lib().nsdp_ball_query_f32(new_xyz.data_ptr(), xyz.data_ptr(), B, N, M, float(radius), int(nsample),
                                             out.data_ptr(), _stream()), "nsdp_ball_query_f32")
    _count()
    return out


def group_points(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    _chk_f32("points", points)
    _chk_i32("idx", idx)
    B, Cc, N = points.shape
    _, M, K = idx.shape
    out = torch.empty((B, Cc, M, K), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        check(_lib.lib().nsdp_group_points_f32(points.data_ptr(), idx.data_ptr(), B, Cc, N, M, K, out.data_ptr(),
                                               _stream()), "nsdp_group_points_f32")
    _count()
    return out


def group_points_grad(grad_out: torch.Tensor, idx: torch.Tensor, n: int) -> torch.Tensor:
    _chk_f32("grad_out", grad_out)
    _chk_i32("idx", idx)
    B, Cc, M, K = grad_out.shape
    out = torch.zeros((B, Cc, n), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        check(_lib.lib().nsdp_group_points_grad_f32(grad_out.data_ptr(), idx.data_ptr(), B, Cc, n, M, K, out.data_ptr(),
                                                    _stream()), "nsdp_group_points_grad_f32")
    _count()
    return out


def three_nn(unknowns: torch.Tensor, knows: torch.Tensor):
    _chk_f32("unknowns", unknowns)
    _chk_f32("knows", knows)
    B, n, _ = unknowns.shape
    m = knows.shape[1]
    dist2 = torch.empty((B, n, 3), dtype=torch.float32, device=unknowns.device)
    idx = torch.empty((B, n, 3), dtype=torch.int32, device=unknowns.device)
    with torch.cuda.device(unknowns.device):
        check(_lib.lib().nsdp_three_nn_f32(unknowns.data_ptr(), knows.data_ptr(), B, n, m, dist2.data_ptr(), idx.data_ptr(),
                                           _stream()), "nsdp_three_nn_f32")
    _count()
    return [dist2, idx]


def three_interpolate(points: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    _chk_f32("points", points)
    _chk_i32("idx", idx)
    _chk_f32("weight", weight)
    B, Cc, m = points.shape
    n = idx.shape[1]
    out = torch.empty((B, Cc, n), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        check(_lib.lib().nsdp_three_interpolate_f32(points.data_ptr(), idx.data_ptr(), weight.data_ptr(), B, Cc, m, n,
                                                    out.data_ptr(), _stream()), "nsdp_three_interpolate_f32")
    _count()
    return out


def three_interpolate_grad(grad_out: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor, m: int) -> torch.Tensor:
    _chk_f32("grad_out", grad_out)
    _chk_i32("idx", idx)
    _chk_f32("weight", weight)
    B, Cc, n = grad_out.shape
    out = torch.zeros((B, Cc, m), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        check(_lib.lib().nsdp_three_interpolate_grad_f32(grad_out.data_ptr(), idx.data_ptr(), weight.data_ptr(), B, Cc, n,
                                                         m, out.data_ptr(), _stream()), "nsdp_three_interpolate_grad_f32")
    _count()
    return out


# ---------------------------------------------------------------------------------------------------
# Part 2: fused hot-path operators
# ---------------------------------------------------------------------------------------------------
def knn(query: torch.Tensor, ref: torch.Tensor, k: int, return_d2: bool = False):
    """k nearest `ref` points of every `query` point, ascending (distance, index): (B,M,k) int32.
    Replaces square_distance(...).argsort()[:, :, :k] (model/utils.py:39-55, encoder/blocks.py:101-102)."""
    _chk_f32("query", query)
    _chk_f32("ref", ref)
    B, M, _ = query.shape
    N = ref.shape[1]
    out = torch.empty((B, M, k), dtype=torch.int32, device=query.device)
    d2 = torch.empty((B, M, k), dtype=torch.float32, device=query.device) if return_d2 else None
    L = _lib.lib()
    with torch.cuda.device(query.device):
        ws_bytes = L.nsdp_knn_workspace_bytes(B, M, N, int(k))
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=query.device) if ws_bytes else None
        with _timed(f"knn_M{M}_N{N}_k{k}"):
            check(L.nsdp_knn_f32(query.data_ptr(), ref.data_ptr(), B, M, N, int(k), out.data_ptr(), _p(d2), _p(ws),
                                 ws_bytes, _stream()), "nsdp_knn_f32")
    _count(2 if ws_bytes else 1)
    return (out, d2) if return_d2 else out


def _vattn_args(xyz_c, xyz_n, idx, qp, kp, vp, gq, gv, wd0, bd0, wd2t, wpt, wg2t, pc, vc, sign, wd2=None, wp=None,
                wg2=None) -> VattnArgs:
    B, M, _ = xyz_c.shape
    N = xyz_n.shape[1]
    D = wd2t.shape[0]
    K = idx.shape[2] if idx is not None else N
    a = VattnArgs()
    a.xyz_c, a.xyz_n, a.idx = _p(xyz_c), _p(xyz_n), _p(idx)
    a.qp, a.kp, a.vp, a.gq, a.gv = _p(qp), _p(kp), _p(vp), _p(gq), _p(gv)
    a.wd0, a.bd0, a.wd2t, a.wpt, a.wg2t, a.pc, a.vc = _p(wd0), _p(bd0), _p(wd2t), _p(wpt), _p(wg2t), _p(pc), _p(vc)
    a.wd2, a.wp, a.wg2 = _p(wd2), _p(wp), _p(wg2)
    a.B, a.M, a.N, a.K, a.D = B, M, N, K, D
    a.has_global = 1 if gq is not None else 0
    a.sign = float(sign)
    a.impl = VATTN_IMPL
    return a


class _VectorAttention(torch.autograd.Function):
    """out = nsdp_vattn_fwd_f32(...); backward = nsdp_vattn_bwd_f32, which recomputes the chain on chip from the
    inputs + the (B,M,D) result and softmax statistics — no [pairs, D] activation is saved."""

    @staticmethod
    def forward(ctx, xyz_c, xyz_n, idx, qp, kp, vp, gq, gv, wd0, bd0, wd2t, wpt, wg2t, pc, vc, sign, grad_enabled=True,
                wd2n=None, wpn=None, wg2n=None):
        tensors = dict(xyz_c=xyz_c, xyz_n=xyz_n, qp=qp, kp=kp, vp=vp, gq=gq, gv=gv, wd0=wd0, bd0=bd0, wd2t=wd2t,
                       wpt=wpt, wg2t=wg2t, pc=pc, vc=vc)
        for n, t in tensors.items():
            if t is not None:
                _chk_f32(n, t)
        if idx is not None:
            _chk_i32("idx", idx)
        a = _vattn_args(xyz_c, xyz_n, idx, qp, kp, vp, gq, gv, wd0, bd0, wd2t, wpt, wg2t, pc, vc, sign)
        out = torch.empty((a.B, a.M, a.D), dtype=torch.float32, device=xyz_c.device)
        # needs_input_grad ignores torch.no_grad() (parameters keep requires_grad=True) and Function.forward always runs
        # with grad mode off, so the caller's grad mode comes in as an argument: eval / validate / test.py forwards
        # must not allocate and write the (2,B,M,D) softmax statistics (640 MB per decoder call at the bench size)
        need_bwd = bool(grad_enabled) and any(ctx.needs_input_grad)
        stats = torch.empty((2, a.B, a.M, a.D), dtype=torch.float32, device=xyz_c.device) if need_bwd else None
        L = _lib.lib()
        saved = None
        with torch.cuda.device(xyz_c.device):
            if need_bwd and SAVE_ACTIVATIONS and a.impl != 1:
                sv_bytes = L.nsdp_vattn_saved_bytes(C.byref(a))
                if sv_bytes:
                    saved = torch.empty((sv_bytes,), dtype=torch.uint8, device=xyz_c.device)
                    a.saved, a.saved_bytes = saved.data_ptr(), sv_bytes
            ws_bytes = L.nsdp_vattn_fwd_workspace_bytes(C.byref(a)) if a.impl != 1 else 0
            ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=xyz_c.device) if ws_bytes else None
            with _timed(f"vattn_fwd_D{a.D}_K{a.K}_M{a.M}"):
                check(L.nsdp_vattn_fwd_f32(C.byref(a), out.data_ptr(), _p(stats), _p(ws), ws_bytes, _stream()),
                      "nsdp_vattn_fwd_f32")
        _count()
        ctx.sign = sign
        if need_bwd:
            ctx.save_for_backward(xyz_c, xyz_n, idx, qp, kp, vp, gq, gv, wd0, bd0, wd2t, wpt, wg2t, pc, vc, out, stats, saved,
                                  wd2n, wpn, wg2n)
        return out

    @staticmethod
    def backward(ctx, d_out):
        (xyz_c, xyz_n, idx, qp, kp, vp, gq, gv, wd0, bd0, wd2t, wpt, wg2t, pc, vc, out, stats, saved,
         wd2n, wpn, wg2n) = ctx.saved_tensors
        d_out = d_out.contiguous()
        # the data-gradient GEMMs need the un-transposed matrices as K-major operands: the caller's own (it transposed them
        # to make wd2t / wpt / wg2t) or three small transposes here
        wd2 = wd2n if wd2n is not None else wd2t.t().contiguous()
        wp = wpn if wpn is not None else wpt.t().contiguous()
        wg2 = wg2n if wg2n is not None else wg2t.t().contiguous()
        a = _vattn_args(xyz_c, xyz_n, idx, qp, kp, vp, gq, gv, wd0, bd0, wd2t, wpt, wg2t, pc, vc, ctx.sign, wd2, wp, wg2)
        if saved is not None and a.impl != 1:
            a.saved, a.saved_bytes = saved.data_ptr(), saved.numel()
        need = ctx.needs_input_grad

        # every gradient buffer is a view of ONE zero-filled allocation: one memset instead of ~14 per attention backward
        def z(t, flag):
            return t if (t is not None and flag) else None          # placeholder (shape donor), replaced by its view below

        # the tensor-core kernels always reduce the three d x d weight gradients (NSDP_ERR_INVALID_ARGUMENT on NULL): a
        # frozen network (requires_grad_(False), test-time optimisation of queries / latents) gets scratch buffers that
        # are dropped below
        wneed = (lambda i: True) if a.impl != 1 else (lambda i: need[i])

        g = dict(d_xyz_c=z(xyz_c, need[0]), d_xyz_n=z(xyz_n, need[1]), d_qp=z(qp, need[3]), d_kp=z(kp, need[4]),
                 d_vp=z(vp, need[5]), d_gq=z(gq, need[6]), d_gv=z(gv, need[7]), d_wd0=z(wd0, need[8]),
                 d_bd0=z(bd0, need[9]), d_wd2t=z(wd2t, wneed(10)), d_wpt=z(wpt, wneed(11)), d_wg2t=z(wg2t, wneed(12)),
                 d_pc=z(pc, need[13]), d_vc=z(vc, need[14]))
        sizes = {k: ((t.numel() + 3) // 4 * 4) for k, t in g.items() if t is not None}      # keep 16-byte alignment
        flat = torch.zeros((sum(sizes.values()),), dtype=torch.float32, device=d_out.device)
        off = 0
        for k, n in sizes.items():
            g[k] = flat[off:off + g[k].numel()].view(g[k].shape)
            off += n
        # xyz_c and xyz_n may be the SAME tensor (self attention): both gradients are returned and autograd
        # sums them.
        gs = VattnGrads()
        for n, t in g.items():
            setattr(gs, n, _p(t))
        L = _lib.lib()
        with torch.cuda.device(d_out.device):
            ws_bytes = L.nsdp_vattn_bwd_workspace_bytes(C.byref(a))
            ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=d_out.device) if ws_bytes else None
            with _timed(f"vattn_bwd_D{a.D}_K{a.K}_M{a.M}"):
                check(L.nsdp_vattn_bwd_f32(C.byref(a), out.data_ptr(), stats.data_ptr(), d_out.data_ptr(), C.byref(gs),
                                           _p(ws), ws_bytes, _stream()), "nsdp_vattn_bwd_f32")
        _count()
        return (g["d_xyz_c"], g["d_xyz_n"], None, g["d_qp"], g["d_kp"], g["d_vp"], g["d_gq"], g["d_gv"], g["d_wd0"],
                g["d_bd0"], g["d_wd2t"] if need[10] else None, g["d_wpt"] if need[11] else None,
                g["d_wg2t"] if need[12] else None, g["d_pc"], g["d_vc"], None, None, None, None, None)


def vector_attention(xyz_c, xyz_n, idx, qp, kp, vp, wd0, bd0, wd2t, wpt, wg2t, pc, vc, sign=1.0, gq=None, gv=None,
                     wd2n=None, wpn=None, wg2n=None):
    """Fused pair-level vector attention (see nsdp_vattn_args in include/nsdp_b200.h). `wd2n` / `wpn` / `wg2n`: optional
    contiguous, DETACHED copies of the three d x d matrices in their natural (un-transposed) layout, i.e. wd2t.t() etc.; a
    caller that built the transposed operands from them passes them along and saves the backward three transposes."""
    def nat(n, t):
        if n is None:
            return None
        if n.requires_grad or not n.is_contiguous() or n.shape != t.shape:
            raise ValueError("wd2n / wpn / wg2n: detached, contiguous, same shape as the transposed operand")
        return n
    return _VectorAttention.apply(xyz_c, xyz_n, idx, qp, kp, vp, gq, gv, wd0, bd0, wd2t, wpt, wg2t, pc, vc, float(sign),
                                  torch.is_grad_enabled(), nat(wd2n, wd2t), nat(wpn, wpt), nat(wg2n, wg2t))


def _tail_args(lat2d, wc_t, bc, w0_t, b0, w1_t, b1, wo_t, bo) -> TailArgs:
    a = TailArgs()
    a.lat, a.wc_t, a.bc = _p(lat2d), _p(wc_t), _p(bc)
    a.w0_t, a.b0, a.w1_t, a.b1 = _p(w0_t), _p(b0), _p(w1_t), _p(b1)
    a.wo_t, a.bo = _p(wo_t), _p(bo)
    a.R, a.C = lat2d.shape
    a.H = w0_t.shape[-1]
    a.O = wo_t.shape[1]
    a.n_blocks = w0_t.shape[0]
    a.impl = TAIL_IMPL
    return a


class _ResnetTail(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lat2d, wc_t, bc, w0_t, b0, w1_t, b1, wo_t, bo):
        for n, t in dict(lat=lat2d, wc_t=wc_t, bc=bc, w0_t=w0_t, b0=b0, w1_t=w1_t, b1=b1, wo_t=wo_t, bo=bo).items():
            _chk_f32(n, t)
        a = _tail_args(lat2d, wc_t, bc, w0_t, b0, w1_t, b1, wo_t, bo)
        out = torch.empty((a.R, a.O), dtype=torch.float32, device=lat2d.device)
        L = _lib.lib()
        with torch.cuda.device(lat2d.device):
            ws_bytes = L.nsdp_resnet_tail_fwd_workspace_bytes(C.byref(a))
            ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=lat2d.device) if ws_bytes else None
            with _timed("resnet_tail_fwd"):
                check(L.nsdp_resnet_tail_fwd_f32(C.byref(a), out.data_ptr(), _p(ws), ws_bytes, _stream()),
                      "nsdp_resnet_tail_fwd_f32")
        _count()
        ctx.save_for_backward(lat2d, wc_t, bc, w0_t, b0, w1_t, b1, wo_t, bo)
        return out

    @staticmethod
    def backward(ctx, d_out):
        lat2d, wc_t, bc, w0_t, b0, w1_t, b1, wo_t, bo = ctx.saved_tensors
        d_out = d_out.contiguous()
        a = _tail_args(lat2d, wc_t, bc, w0_t, b0, w1_t, b1, wo_t, bo)
        ws_ = (wc_t, bc, w0_t, b0, w1_t, b1, wo_t, bo)
        pad = [(t.numel() + 3) // 4 * 4 for t in ws_]
        flat = torch.zeros((sum(pad),), dtype=torch.float32, device=d_out.device)     # one memset for all parameter gradients
        offs = [sum(pad[:i]) for i in range(len(pad))]
        grads = [torch.empty_like(lat2d)] + [flat[o:o + t.numel()].view(t.shape) for o, t in zip(offs, ws_)]
        gs = TailGrads()
        for n, t in zip(("d_lat", "d_wc_t", "d_bc", "d_w0_t", "d_b0", "d_w1_t", "d_b1", "d_wo_t", "d_bo"), grads):
            setattr(gs, n, _p(t))
        L = _lib.lib()
        with torch.cuda.device(d_out.device):
            ws_bytes = L.nsdp_resnet_tail_bwd_workspace_bytes(C.byref(a))
            ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=d_out.device) if ws_bytes else None
            with _timed("resnet_tail_bwd"):
                check(L.nsdp_resnet_tail_bwd_f32(C.byref(a), d_out.data_ptr(), C.byref(gs), _p(ws), ws_bytes, _stream()),
                      "nsdp_resnet_tail_bwd_f32")
        _count()
        return tuple(grads)


def resnet_tail(lat2d, wc_t, bc, w0_t, b0, w1_t, b1, wo_t, bo):
    """Fused decoder ResNet-FC tail over rows: (R,C) -> (R,O). See nsdp_tail_args."""
    return _ResnetTail.apply(lat2d, wc_t, bc, w0_t, b0, w1_t, b1, wo_t, bo)


# ---------------------------------------------------------------------------------------------------
# ElementwiseMLP of the encoder (model/encoder/blocks.py:137-159): conv1 -> bn1 -> relu -> conv2 -> bn2 -> relu -> +x -> bn3
# ---------------------------------------------------------------------------------------------------
def _emlp_args(x2d, w1, b1, w2, b2, bns, training: bool) -> EmlpArgs:
    a = EmlpArgs()
    a.x, a.w1, a.b1, a.w2, a.b2 = _p(x2d), _p(w1), _p(b1), _p(w2), _p(b2)
    for i, bn in enumerate(bns):
        a.bn_weight[i], a.bn_bias[i] = _p(bn.weight), _p(bn.bias)
        a.running_mean[i], a.running_var[i] = _p(bn.running_mean), _p(bn.running_var)
        a.num_batches_tracked[i] = _p(bn.num_batches_tracked)
    a.R, a.C = x2d.shape
    a.training = 1 if training else 0
    a.momentum = float(bns[0].momentum)
    a.eps = float(bns[0].eps)
    return a


class _ElementwiseMLP(torch.autograd.Function):
    """forward = nsdp_emlp_fwd_f32 (4 kernels), backward = nsdp_emlp_bwd_f32 (7 kernels). The three BatchNorm modules are
    passed as a non-tensor argument: their running buffers are updated in place by the forward, like F.batch_norm does."""

    @staticmethod
    def forward(ctx, x2d, w1, b1, w2, b2, g1, be1, g2, be2, g3, be3, bns, training):
        for n, t in dict(x=x2d, w1=w1, b1=b1, w2=w2, b2=b2, g1=g1, be1=be1, g2=g2, be2=be2, g3=g3, be3=be3).items():
            _chk_f32(n, t)
        a = _emlp_args(x2d, w1, b1, w2, b2, bns, training)
        dev = x2d.device
        out, t1, t2, s = (torch.empty_like(x2d) for _ in range(4))
        L = _lib.lib()
        stats = torch.empty((L.nsdp_emlp_stats_bytes(C.byref(a)) // 8,), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev), _timed(f"emlp_fwd_R{a.R}_C{a.C}"):
            check(L.nsdp_emlp_fwd_f32(C.byref(a), out.data_ptr(), t1.data_ptr(), t2.data_ptr(), s.data_ptr(), stats.data_ptr(),
                                      _stream()), "nsdp_emlp_fwd_f32")
        _count()
        ctx.bns, ctx.training = bns, training
        ctx.save_for_backward(x2d, w1, b1, w2, b2, t1, t2, s, stats)
        return out

    @staticmethod
    def backward(ctx, d_out):
        x2d, w1, b1, w2, b2, t1, t2, s, stats = ctx.saved_tensors
        d_out = d_out.contiguous()
        a = _emlp_args(x2d, w1, b1, w2, b2, ctx.bns, ctx.training)
        Cc, dev = a.C, x2d.device
        d_x = torch.empty_like(x2d)
        flat = torch.zeros((2 * Cc * Cc + 8 * Cc,), dtype=torch.float32, device=dev)     # one memset for every gradient
        d_w1, d_w2 = flat[:Cc * Cc].view(Cc, Cc), flat[Cc * Cc:2 * Cc * Cc].view(Cc, Cc)
        small = flat[2 * Cc * Cc:].view(8, Cc)          # d_b1, d_b2, d_gamma1..3, d_beta1..3
        g = EmlpGrads()
        g.d_x, g.d_w1, g.d_w2, g.d_b1, g.d_b2 = d_x.data_ptr(), d_w1.data_ptr(), d_w2.data_ptr(), small[0].data_ptr(), small[1].data_ptr()
        for i in range(3):
            g.d_bn_weight[i], g.d_bn_bias[i] = small[2 + i].data_ptr(), small[5 + i].data_ptr()
        L = _lib.lib()
        with torch.cuda.device(dev):
            ws_bytes = L.nsdp_emlp_bwd_workspace_bytes(C.byref(a))
            ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
            with _timed(f"emlp_bwd_R{a.R}_C{a.C}"):
                check(L.nsdp_emlp_bwd_f32(C.byref(a), t1.data_ptr(), t2.data_ptr(), s.data_ptr(), stats.data_ptr(),
                                          d_out.data_ptr(), C.byref(g), ws.data_ptr(), ws_bytes, _stream()), "nsdp_emlp_bwd_f32")
        _count()
        return (d_x, d_w1, small[0], d_w2, small[1], small[2], small[5], small[3], small[6], small[4], small[7], None, None)


def elementwise_mlp(x, conv1, bn1, conv2, bn2, bn3):
    """bn3(x + relu(bn2(conv2(relu(bn1(conv1 x)))))) over the rows of x (B, n, C); the arguments are the nn.Conv1d /
    nn.BatchNorm1d modules of the reference's ElementwiseMLP (their parameters get gradients, their buffers are updated)."""
    B, n, Cc = x.shape
    for bn in (bn1, bn2, bn3):
        if bn.momentum is None or not bn.track_running_stats or not bn.affine:
            raise NotImplementedError("ElementwiseMLP kernel: BatchNorm1d with affine=True, track_running_stats=True and a "
                                      "fixed momentum (the reference's defaults, model/encoder/blocks.py:148-151)")
    training = bn1.training
    out = _ElementwiseMLP.apply(x.reshape(B * n, Cc).contiguous(), conv1.weight.squeeze(-1), conv1.bias,
                                conv2.weight.squeeze(-1), conv2.bias, bn1.weight, bn1.bias, bn2.weight, bn2.bias,
                                bn3.weight, bn3.bias, (bn1, bn2, bn3), training)
    return out.reshape(B, n, Cc)


# ---------------------------------------------------------------------------------------------------
# nn.Linear with a narrow input on many rows (the encoder's input layer and the projections folded through it)
# ---------------------------------------------------------------------------------------------------
class _LinearNarrow(torch.autograd.Function):
    """y = x W^T + b for x (R, K), K <= 8. Forward and d_x are ordinary GEMMs; the weight / bias gradient — a
    [N x R] x [R x K] product that cuBLAS runs at ~0.1 ms for 32 768 rows x 4 channels — is one coalesced pass over d_y
    (nsdp_linear_narrow_dw_f32)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return torch.nn.functional.linear(x, weight, bias)

    @staticmethod
    def backward(ctx, d_y):
        x, weight = ctx.saved_tensors
        d_y = d_y.contiguous()
        R, K = x.shape
        N = weight.shape[0]
        d_x = d_y @ weight if ctx.needs_input_grad[0] else None
        d_w = d_b = None
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            buf = torch.zeros(N * K + N, dtype=torch.float32, device=x.device)
            d_w = buf[:N * K].view(N, K)
            d_b = buf[N * K:] if ctx.has_bias else None
            with _timed(f"linear_narrow_dw_R{R}_K{K}_N{N}"):
                check(_lib.lib().nsdp_linear_narrow_dw_f32(x.data_ptr(), d_y.data_ptr(), R, K, N, d_w.data_ptr(),
                                                      d_b.data_ptr() if d_b is not None else None, _stream()),
                      "nsdp_linear_narrow_dw_f32")
            _count()
        return d_x, d_w, d_b


def linear(x, weight, bias=None):
    """torch.nn.functional.linear(x, weight, bias); rows x (<= 8 channels) inputs on the GPU take the narrow-input backward."""
    K = x.shape[-1]
    if (x.is_cuda and K <= 8 and x.dtype == torch.float32 and weight.dtype == torch.float32 and x.numel() // max(K, 1) >= 4096
            and torch.is_grad_enabled() and (weight.requires_grad or (bias is not None and bias.requires_grad))):
        out = _LinearNarrow.apply(x.reshape(-1, K).contiguous(), weight.contiguous(), bias)
        return out.reshape(*x.shape[:-1], weight.shape[0])
    return torch.nn.functional.linear(x, weight, bias)


# ---------------------------------------------------------------------------------------------------
# Plain fused neural-field MLP (BASELINE.json configs[3]; forward / inference only)
# ---------------------------------------------------------------------------------------------------
class FusedMLP:
    """h = relu(x W_in + b_in); n_hidden x { h = relu(h W_l + b_l) }; out = h W_out + b_out over rows (R, Cin) -> (R, O).

    Takes the weights in nn.Linear layout (out_features, in_features) — `w_in (W, Cin)`, `w_h (n_hidden, W, W)`,
    `w_out (O, W)` — transposes them once, and keeps the packed tensor-core weight image between calls (weights are
    constant at inference time), so a call is ONE kernel launch. See nsdp_mlp_args in include/nsdp_b200.h."""

    def __init__(self, w_in, b_in, w_h, b_h, w_out, b_out, impl: int = 0):
        for n, t in dict(w_in=w_in, b_in=b_in, w_h=w_h, b_h=b_h, w_out=w_out, b_out=b_out).items():
            _chk_f32(n, t)
        self.W, self.Cin = w_in.shape
        self.O = w_out.shape[0]
        self.n_hidden = w_h.shape[0]
        self.impl = impl
        self.w_in_t = w_in.t().contiguous()
        self.w_h_t = w_h.transpose(1, 2).contiguous()
        self.w_out_t = w_out.t().contiguous()
        self.b_in, self.b_h, self.b_out = b_in, b_h, b_out
        self._ws = None
        self._packed = False

    def refresh(self, w_in=None, b_in=None, w_h=None, b_h=None, w_out=None, b_out=None) -> None:
        """Call after the weights changed (optionally passing the new nn.Linear-layout tensors): the next call re-packs."""
        if w_in is not None:
            self.w_in_t = _chk_f32("w_in", w_in).t().contiguous()
        if w_h is not None:
            self.w_h_t = _chk_f32("w_h", w_h).transpose(1, 2).contiguous()
        if w_out is not None:
            self.w_out_t = _chk_f32("w_out", w_out).t().contiguous()
        self.b_in = self.b_in if b_in is None else _chk_f32("b_in", b_in)
        self.b_h = self.b_h if b_h is None else _chk_f32("b_h", b_h)
        self.b_out = self.b_out if b_out is None else _chk_f32("b_out", b_out)
        self._packed = False

    def _args(self, x: torch.Tensor) -> MlpArgs:
        a = MlpArgs()
        a.x, a.w_in_t, a.b_in = _p(x), _p(self.w_in_t), _p(self.b_in)
        a.w_h_t, a.b_h, a.w_out_t, a.b_out = _p(self.w_h_t), _p(self.b_h), _p(self.w_out_t), _p(self.b_out)
        a.R, a.Cin, a.W, a.O, a.n_hidden = x.shape[0], self.Cin, self.W, self.O, self.n_hidden
        a.impl = self.impl
        a.reuse_packed = 1 if self._packed else 0
        return a

    def __call__(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        _chk_f32("x", x)
        if x.dim() != 2 or x.shape[1] != self.Cin:
            raise RuntimeError(f"x must be (R, {self.Cin})")
        a = self._args(x)
        if out is None:
            out = torch.empty((a.R, a.O), dtype=torch.float32, device=x.device)
        elif _chk_f32("out", out).shape != (a.R, a.O) or out.device != x.device:
            raise RuntimeError(f"out must be a ({a.R}, {a.O}) tensor on {x.device}")
        L = _lib.lib()
        with torch.cuda.device(x.device):
            ws_bytes = L.nsdp_fused_mlp_fwd_workspace_bytes(C.byref(a))
            if ws_bytes and (self._ws is None or self._ws.numel() < ws_bytes or self._ws.device != x.device):
                self._ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=x.device)
                self._packed = False
                a.reuse_packed = 0
            with _timed("fused_mlp_fwd"):
                check(L.nsdp_fused_mlp_fwd_f32(C.byref(a), out.data_ptr(), _p(self._ws) if ws_bytes else None, ws_bytes,
                                               _stream()), "nsdp_fused_mlp_fwd_f32")
            self._packed = bool(ws_bytes)
        _count()
        return out


class _FusedMLPFn(torch.autograd.Function):
    """Differentiable form of the fused MLP: forward = nsdp_fused_mlp_fwd_f32, backward = nsdp_fused_mlp_bwd_f32 (which
    recomputes the activations on chip). Weights are taken TRANSPOSED (w_in_t (Cin,W), w_h_t (L,W,W), w_out_t (W,O)) so
    that the gradient buffers the kernel fills are the autograd results as they are."""

    @staticmethod
    def forward(ctx, x, w_in_t, b_in, w_h_t, b_h, w_out_t, b_out):
        for n, t in dict(x=x, w_in_t=w_in_t, b_in=b_in, w_h_t=w_h_t, b_h=b_h, w_out_t=w_out_t, b_out=b_out).items():
            _chk_f32(n, t)
        a = _mlp_args(x, w_in_t, b_in, w_h_t, b_h, w_out_t, b_out)
        out = torch.empty((a.R, a.O), dtype=torch.float32, device=x.device)
        L = _lib.lib()
        with torch.cuda.device(x.device):
            ws_bytes = L.nsdp_fused_mlp_fwd_workspace_bytes(C.byref(a))
            ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=x.device) if ws_bytes else None
            with _timed("fused_mlp_fwd"):
                check(L.nsdp_fused_mlp_fwd_f32(C.byref(a), out.data_ptr(), _p(ws), ws_bytes, _stream()), "nsdp_fused_mlp_fwd_f32")
        _count()
        ctx.save_for_backward(x, w_in_t, b_in, w_h_t, b_h, w_out_t, b_out)
        return out

    @staticmethod
    def backward(ctx, d_out):
        x, w_in_t, b_in, w_h_t, b_h, w_out_t, b_out = ctx.saved_tensors
        d_out = d_out.contiguous()
        a = _mlp_args(x, w_in_t, b_in, w_h_t, b_h, w_out_t, b_out)
        d_x = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        grads = [torch.zeros_like(t) for t in (w_in_t, b_in, w_h_t, b_h, w_out_t, b_out)]
        gs = MlpGrads()
        gs.d_x = _p(d_x)
        for n, t in zip(("d_w_in_t", "d_b_in", "d_w_h_t", "d_b_h", "d_w_out_t", "d_b_out"), grads):
            setattr(gs, n, _p(t))
        L = _lib.lib()
        with torch.cuda.device(d_out.device):
            ws_bytes = L.nsdp_fused_mlp_bwd_workspace_bytes(C.byref(a))
            ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=d_out.device) if ws_bytes else None
            with _timed("fused_mlp_bwd"):
                check(L.nsdp_fused_mlp_bwd_f32(C.byref(a), d_out.data_ptr(), C.byref(gs), _p(ws), ws_bytes, _stream()),
                      "nsdp_fused_mlp_bwd_f32")
        _count()
        return (d_x, *grads)


def _mlp_args(x, w_in_t, b_in, w_h_t, b_h, w_out_t, b_out) -> MlpArgs:
    a = MlpArgs()
    a.x, a.w_in_t, a.b_in = _p(x), _p(w_in_t), _p(b_in)
    a.w_h_t, a.b_h, a.w_out_t, a.b_out = _p(w_h_t), _p(b_h), _p(w_out_t), _p(b_out)
    a.R, a.Cin = x.shape
    a.W, a.O = w_out_t.shape
    a.n_hidden = w_h_t.shape[0]
    a.impl = 0
    a.reuse_packed = 0
    return a


def fused_mlp(x, w_in_t, b_in, w_h_t, b_h, w_out_t, b_out):
    """Differentiable fused MLP over rows (R, Cin) -> (R, O) with TRANSPOSED weights (see nsdp_mlp_args); training-time
    counterpart of `FusedMLP`."""
    return _FusedMLPFn.apply(x, w_in_t, b_in, w_h_t, b_h, w_out_t, b_out)

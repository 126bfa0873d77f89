"""TEST INFRASTRUCTURE: the kernel contracts of include/nsdp_b200.h restated as plain differentiable torch functions (any dtype,
CPU tensors), to stand in for the entry points of `nsdp_b200.ops` in the `-m "not gpu"` tests of the HOST logic
(tests/test_host_composition.py, tests/test_dist_gloo.py). The `-m gpu` kernel tests hold the CUDA kernels to the same
contracts; the product itself has no CPU path and never imports this file."""
import torch
import torch.nn.functional as F

from oracle import tdnet_oracle as orc


def contract_vector_attention(xyz_c, xyz_n, idx, qp, kp, vp, wd0, bd0, wd2t, wpt, wg2t, pc, vc, sign=1.0, gq=None, gv=None,
                              wd2n=None, wpn=None, wg2n=None):
    """nsdp_vattn_args (include/nsdp_b200.h, "Vector attention core over neighbourhoods")."""
    B, M, _ = xyz_c.shape
    N = xyz_n.shape[1]
    for nat, t in ((wd2n, wd2t), (wpn, wpt), (wg2n, wg2t)):        # the un-transposed copies must BE the transposes
        if nat is not None:
            assert not nat.requires_grad and torch.equal(nat, t.detach().t())
    if idx is None:
        idx = torch.arange(N).view(1, 1, N).expand(B, M, N)
    idx = idx.long()
    K = idx.shape[2]
    gather = lambda t: torch.gather(t, 1, idx.reshape(B, M * K, 1).expand(-1, -1, t.shape[-1])).reshape(B, M, K, -1)
    rel = sign * (xyz_c[:, :, None] - gather(xyz_n))
    h = F.relu(rel @ wd0.t() + bd0)
    dlt = h @ wd2t
    pre = h @ wpt + pc
    if qp is not None:
        pre = pre + qp[:, :, None]
    if kp is not None:
        pre = pre - gather(kp)
    a = F.relu(pre) @ wg2t
    val = vc + dlt
    if vp is not None:
        val = val + gather(vp)
    if gq is not None:
        a = torch.cat([a, (F.relu(gq) @ wg2t)[:, None, None, :].expand(-1, M, -1, -1)], dim=2)
        val = torch.cat([val, gv[:, None, None, :].expand(-1, M, -1, -1)], dim=2)
    return (torch.softmax(a, dim=2) * val).sum(dim=2)


def contract_resnet_tail(lat, wc_t, bc, w0_t, b0, w1_t, b1, wo_t, bo):
    """nsdp_tail_args: net = init(lat); for i: net += fc_c[i](lat); net += fc_1[i](relu(fc_0[i](relu(net)))); out = fc_out(relu(net))."""
    H = w0_t.shape[-1]
    pre = lat @ wc_t + bc
    net = pre[:, :H]
    for i in range(w0_t.shape[0]):
        net = net + pre[:, (1 + i) * H:(2 + i) * H]
        net = net + F.relu(F.relu(net) @ w0_t[i] + b0[i]) @ w1_t[i] + b1[i]
    return F.relu(net) @ wo_t + bo


def contract_elementwise_mlp(x, conv1, bn1, conv2, bn2, bn3):
    """nsdp_emlp_args: bn3(x + relu(bn2(conv2(relu(bn1(conv1 x)))))) over the rows of x, torch BatchNorm1d semantics."""
    B, n, C = x.shape
    rows = x.reshape(B * n, C)
    t1 = F.linear(rows, conv1.weight.squeeze(-1), conv1.bias)
    t2 = F.linear(F.relu(bn1(t1)), conv2.weight.squeeze(-1), conv2.bias)
    return bn3(rows + F.relu(bn2(t2))).reshape(B, n, C)




def install(setattr_fn=None):
    """Replace the ops entry points the model mirror calls. `setattr_fn(obj, name, value)`: pytest's monkeypatch.setattr, or
    None for a plain setattr (spawned worker processes). Index kernels -> the C oracle (bit-exact contract, pinned in
    tests/test_index_golden.py); fused kernels -> the header's formulas above."""
    from nsdp_b200 import ops
    put = setattr_fn or setattr
    put(ops, "vector_attention", contract_vector_attention)
    put(ops, "resnet_tail", contract_resnet_tail)
    put(ops, "elementwise_mlp", contract_elementwise_mlp)
    put(ops, "linear", F.linear)
    put(ops, "knn", lambda q, r, k, return_d2=False: orc.knn(q, r, k, return_d2=return_d2))
    put(ops, "furthest_point_sampling", lambda xyz, m: orc.fps(xyz, m))

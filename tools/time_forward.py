"""Ad-hoc per-stage CUDA-event timing of the C2 forward (B=8, N=4096, Q=50k). Scratch tool, not the bench."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nsdp_b200 import ops, synth
from nsdp_b200.model import build_model

dev = "cuda:0"
B, N, Q = 8, 4096, 50000
model, *_ = build_model(synth.make_config("forward"), device=dev)
schema = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
model.load_state_dict(synth.named_state_dict(schema, seed=0)); model.eval()
batch = {k: v.to(dev) for k, v in synth.forward_batch(B, N, Q, seed=1).items()}

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts)//2]

surf, qs = batch["surface_samples_inputs"], batch["space_samples_src"]
xyz = surf[:, :, :3].contiguous()
with torch.no_grad():
    print("fps 4096->500      ms", timeit(lambda: ops.furthest_point_sampling(xyz, 500)))
    x500 = xyz[:, :500].contiguous()
    print("fps 500->100       ms", timeit(lambda: ops.furthest_point_sampling(x500, 100)))
    print("knn 4096x4096 k10  ms", timeit(lambda: ops.knn(xyz, xyz, 10)))
    a100 = xyz[:, :100].contiguous()
    print("knn 50000x100 k7   ms", timeit(lambda: ops.knn(qs, a100, 7)))
    enc = model.encode(surf)
    print("encoder            ms", timeit(lambda: model.encode(surf)))
    print("  transformer_begin ms", timeit(lambda: model.encoder.transformer_begin(xyz, model.encoder.enc_sdf(surf[:, :, 3:]))))
    f256 = torch.randn(B, 100, 256, device=dev)
    print("  final_transformer ms", timeit(lambda: model.encoder.final_transformers[0](a100, f256)))
    print("decoder            ms", timeit(lambda: model.decode(qs, enc)))
    lat = model.decoder.ct1(qs, enc["z"], enc["anchors"], enc["anchor_feats"])
    print("  ct1 (attn)       ms", timeit(lambda: model.decoder.ct1(qs, enc["z"], enc["anchors"], enc["anchor_feats"])))
    w = model.decoder.packed_tail_weights()
    print("  tail             ms", timeit(lambda: ops.resnet_tail(lat.reshape(-1, 200), *w)))
    t = timeit(lambda: model(qs, surf))
    print("full forward       ms", t, " -> qpts/s", B * Q / t * 1e3)
x = synth.surface_cloud(1, 100000, seed=3).to(dev)
print("fps 100k->4096 B=1 ms", timeit(lambda: ops.furthest_point_sampling(x, 4096), n=3, warm=1))
for k in (16, 32, 64):
    q = x[:, :4096].contiguous()
    print(f"knn 4096x100k k{k}  ms", timeit(lambda: ops.knn(q, x, k), n=3, warm=1))
from oracle import ref_ext
ref = ref_ext.load()
if ref is not None:
    print("REF fps 4096->500  ms", timeit(lambda: ref.furthest_point_sampling(xyz, 500)))
    print("REF fps 100k->4096 ms", timeit(lambda: ref.furthest_point_sampling(x, 4096), n=3, warm=1))

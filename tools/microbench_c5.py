"""BASELINE.json configs[4] microbenchmark: FPS 100 000 -> 4096 and k-NN of the 4096 picks against the cloud (k = 16/32/64).
CUDA-event timing, 3 warm-up + 10 timed calls, median."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nsdp_b200 import ops, synth
DEV = "cuda:0"


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


out = {}
for B in (1, 8):
    xyz = synth.surface_cloud(B, 100000, seed=11, fp16_grid=True).to(DEV)
    ms = timed(lambda: ops.furthest_point_sampling(xyz, 4096))
    out[f"fps_B{B}_100k_to_4096_ms"] = ms
    out[f"fps_B{B}_point_updates_per_s"] = B * 4095 * 100000 / (ms * 1e-3)
    idx = ops.furthest_point_sampling(xyz, 4096).long()
    q = torch.gather(xyz, 1, idx[:, :, None].expand(-1, -1, 3)).contiguous()
    for k in (16, 32, 64):
        ms = timed(lambda: ops.knn(q, xyz, k))
        out[f"knn_B{B}_4096x100k_k{k}_ms"] = ms
        out[f"knn_B{B}_k{k}_distance_evals_per_s"] = B * 4096 * 100000 / (ms * 1e-3)
print(json.dumps(out, indent=1))

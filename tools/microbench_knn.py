"""k-NN at the model's largest site (8 x 4096 queries against the same 4096 points, k = 10) and the decoder site
(8 x 50 000 queries against 100 anchors, k = 7): time per call for the split count selected by NSDP_KNN_SPLITS."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nsdp_b200 import ops, synth
def bench(q, r, k, reps=30):
    for _ in range(3): ops.knn(q, r, k)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): ops.knn(q, r, k)
    b.record(); b.synchronize()
    return a.elapsed_time(b) / reps * 1e3
xyz = synth.surface_cloud(8, 4096, seed=3, fp16_grid=True).cuda()
sub = xyz[:, :500].contiguous()
print(f"splits {os.environ.get('NSDP_KNN_SPLITS', 'auto')}: 4096x4096 k10 {bench(xyz, xyz, 10):.1f} us, 500x4096 k16 {bench(sub, xyz, 16):.1f} us, "
      f"500x500 k16 {bench(sub, sub, 16):.1f} us")

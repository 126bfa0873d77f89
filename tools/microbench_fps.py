"""FPS at the model's first level (8 clouds x 4096 points -> 500): time per call and bit-exactness against the C oracle for
the launch variant selected by NSDP_FPS_VARIANT (csrc/fps.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nsdp_b200 import ops, synth
from oracle import tdnet_oracle as orc
xyz = synth.surface_cloud(8, 4096, seed=3, fp16_grid=True)
dev = xyz.cuda()
want = orc.fps(xyz, 500)
got = ops.furthest_point_sampling(dev, 500)
ok = bool(torch.equal(got.cpu(), want))
for _ in range(5):
    ops.furthest_point_sampling(dev, 500)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50):
    ops.furthest_point_sampling(dev, 500)
b.record(); b.synchronize()
print(f"variant {os.environ.get('NSDP_FPS_VARIANT', 'default')}: {a.elapsed_time(b) / 50 * 1e3:.1f} us per call, bit-exact vs oracle: {ok}")

"""Time md_topdown at BASELINE cfg2's size (8192 envs, 84 x 84 x 3 at +-30 m): CUDA events on the launch stream, after warm-up."""
import json

import torch

from metadrive_ped_b200 import BatchedMetaDriveEnv

env = BatchedMetaDriveEnv(8192, dict(num_scenarios=1000, start_seed=0))
env.reset()
a = torch.zeros((8192, 2), device="cuda")
a[:, 1] = 1.0
for _ in range(30):
    env.step(a)
out = torch.empty((8192, 84, 84, 3), device="cuda")
for _ in range(3):
    env.sim.topdown(84, 30.0, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    env.sim.topdown(84, 30.0, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
out2 = torch.empty((8192, 84, 84, 2), device="cuda")
for _ in range(3):
    env.sim.topdown(84, 30.0, out=out2, channels=2)
torch.cuda.synchronize()
f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
f0.record()
for _ in range(10):
    env.sim.topdown(84, 30.0, out=out2, channels=2)
f1.record()
torch.cuda.synchronize()
ms2 = f0.elapsed_time(f1) / 10
print(json.dumps(dict(kernel="k_topdown<2> (road_network + traffic_flow channels)", agents=8192, resolution=84, max_distance=30.0, ms=ms2,
                      images_per_s=8192 / ms2 * 1e3, written_gbs=out2.numel() * 4 / ms2 / 1e6)))
print(json.dumps(dict(kernel="k_topdown", agents=8192, resolution=84, max_distance=30.0, ms=ms, images_per_s=8192 / ms * 1e3,
                      written_gbs=out.numel() * 4 / ms / 1e6, nonzero_frac=float((out > 0).float().mean()))))

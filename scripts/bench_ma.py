"""BASELINE config 3 timing: MultiAgentRoundaboutEnv, 40 agents x 2048 envs, 240-beam lidar + crash checks, respawn on.
Reports valid agent transitions per second (device-resident, CUDA events, L2 flushed between steps)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metadrive_ped_b200 import BatchedMultiAgentEnv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=2048)
ap.add_argument("--agents", type=int, default=40)
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--burnin", type=int, default=100)
args = ap.parse_args()
env = BatchedMultiAgentEnv(args.envs, {"num_agents": args.agents,
                                       "vehicle_config": {"lidar": {"num_lasers": 240, "distance": 50}}})
sim = env.sim
env.reset()
g = torch.Generator(device="cuda").manual_seed(0)
seats = env.seats
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def actions():
    a = torch.rand((args.envs, seats, 2), generator=g, device="cuda")
    a[..., 0] = (a[..., 0] - 0.5) * 0.2   # small steering noise, throttle in [0, 1]
    return a.contiguous()


for _ in range(args.burnin):
    env.step(actions())
torch.cuda.synchronize()
K = args.steps
ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
valid = torch.zeros((), dtype=torch.int64, device="cuda")
sim.profile_begin(K)
for k in range(K):
    a = actions()
    flush.zero_()
    ev0[k].record()
    env.step(a)
    ev1[k].record()
    valid += ((sim.info_flags & 0x2000) != 0).sum()
torch.cuda.synchronize()
ms = np.array([ev0[k].elapsed_time(ev1[k]) for k in range(K)])
kms = sim.profile_end().mean(0)
v = float(valid.item())
print(json.dumps({"workload": "MultiAgentRoundaboutEnv %d agents x %d envs, 240-beam lidar, respawn on" % (args.agents, args.envs),
                  "agent_steps_per_sec": v / (ms.sum() * 1e-3), "seat_steps_per_sec": args.envs * seats * K / (ms.sum() * 1e-3),
                  "valid_agents_per_env_mean": v / K / args.envs, "ms_per_step": float(ms.mean()),
                  "kernel_ms": dict(zip(["k_pre", "k_dyn", "k_post", "k_lidar+respawn"], [float(x) for x in kms]))}))
env.close()

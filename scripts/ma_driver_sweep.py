"""GPU: live-agent fraction of the cfg3 workload under bench.ma_driver variants (sign / gain sweep)."""
import os, sys, itertools
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from metadrive_ped_b200.sim import BatchedSim
E = 256
lib, arrays, cfg = bench.build_world(E, 0, "cfg3")
dev = torch.device("cuda", 0)
for kh, kl, ka in [(2.2, 0.45, 0.4), (2.2, 0.45, -0.4), (2.2, -0.45, 0.4), (-2.2, 0.45, 0.4), (2.2, 0.45, 0.0), (-2.2, -0.45, 0.0), (2.2, -0.45, 0.0), (-2.2, 0.45, 0.0)]:
    sim = BatchedSim(arrays, cfg)
    sim.reset()
    g = torch.Generator(device=dev).manual_seed(0)
    def pol():
        obs, sd = sim.obs, 2
        herr = torch.asin((2.0 * obs[:, sd] - 1.0).clamp(-1.0, 1.0))
        lat = (2.0 * obs[:, sd + 6] - 1.0) * 2.25
        fwd, rhs = (2.0 * obs[:, sd + 7] - 1.0) * 50.0, (2.0 * obs[:, sd + 8] - 1.0) * 50.0
        near = (fwd * fwd + rhs * rhs) < 64.0
        fwd = torch.where(near, (2.0 * obs[:, sd + 12] - 1.0) * 50.0, fwd)
        rhs = torch.where(near, (2.0 * obs[:, sd + 13] - 1.0) * 50.0, rhs)
        ang = torch.atan2(-rhs, fwd.clamp_min(1.0))
        bend = torch.maximum(obs[:, sd + 9], obs[:, sd + 14])
        v = obs[:, sd + 1] * 81.0 - 1.0
        target = torch.where((bend > 0.0) & (bend < 0.3), 16.0, 28.0)
        u = torch.rand((obs.shape[0], 2), generator=g, device=obs.device) * 2.0 - 1.0
        steer = (kh * herr + kl * lat + ka * ang + 0.05 * u[:, 0]).clamp(-1.0, 1.0)
        thr = torch.where(v < target, 0.6, torch.where(v > target + 6.0, -0.3, 0.0)) + 0.05 * u[:, 1]
        return torch.stack([steer, thr.clamp(-1.0, 1.0)], 1).contiguous()
    valid = 0; arrive = 0; done = 0
    for t in range(600):
        sim.step(pol(), autoreset=True)
        if t >= 300:
            f = sim.info_flags
            valid += int(((f & 0x2000) != 0).sum()); arrive += int(((f & 0x800) != 0).sum())
            done += int(((sim.terminated | sim.truncated) != 0).sum())
    print("kh %5.2f kl %5.2f ka %5.2f  live %.3f  done/step/env %.3f arrive frac of done %.3f" % (kh, kl, ka, valid / (300 * E * 40), done / 300 / E, arrive / max(done, 1)), flush=True)
    sim.close()

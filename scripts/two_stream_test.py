"""Experiment: do two half-batches stepped on two streams overlap (latency-bound kernels of one half filling the idle issue
slots of the other)?  Compares one 8192-env handle with 2 x 4096 and 4 x 2048 handles on separate streams."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from metadrive_ped_b200.sim import BatchedSim

dev = torch.device("cuda", 0)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for parts in (1, 2, 4):
    E = 8192 // parts
    sims, acts, streams = [], [], []
    for p in range(parts):
        lib_, arrays, cfg = bench.build_world(E, p, "cfg2")
        sim = BatchedSim(arrays, cfg, device=0)
        b_arrays, b_cfg = lib_.build_world(list(range(len(lib_))), seed=0, **bench.bank_kw(lib_, "cfg2"))
        bank = BatchedSim(b_arrays, b_cfg, device=0)
        bank.reset(); sim.reset(); sim.attach_bank(bank, seed=1000 + p)
        sims.append((sim, bank))
        acts.append(torch.tensor([0.0, 1.0], device=dev).repeat(sim.n_agents, 1).contiguous())
        streams.append(torch.cuda.Stream(device=dev))
    for _ in range(150):
        for (sim, _), a in zip(sims, acts):
            sim.step(a, autoreset=True)
    torch.cuda.synchronize()
    K = 50
    ms = []
    for k in range(K):
        flush.zero_()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for (sim, _), a, st in zip(sims, acts, streams):
            st.wait_event(e0)
            with torch.cuda.stream(st):
                sim.step(a, autoreset=True)
            torch.cuda.current_stream().wait_stream(st)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms = np.array(ms)
    print("parts %d: %.4f ms per 8192-env step (median %.4f) -> %.2f M agent-steps/s" % (parts, ms.mean(), np.median(ms), 8192 / ms.mean() / 1e3))
    for sim, bank in sims:
        sim.close(); bank.close()

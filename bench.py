#!/usr/bin/env python
"""bench.py — agent-steps/s of the batched MetaDrive step with 240-beam lidar observations (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # ours (for N>1 launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the CPU arm, rank 0 only

Workload (configs[1] of BASELINE.json): `MetaDriveEnv(map=3, traffic_density=0.1, num_scenarios=1000)` — 3-block PG
maps with IDM traffic, 8192 environments per GPU cycling through the 1000 reference scenarios of the shipped
library, reference profiling protocol (examples/profile_metadrive.py:16-29): action [0, 1], finished envs reset
in place (on device).  A step = one env.step of every env: before_step (actuation, trigger, IDM), 5 physics
sub-steps with contacts, after_step, reward/cost/done, 259-float observation, plus the auto-reset of finished envs.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENVS_PER_GPU = 8192
LIBRARY = "pg3_density0.1.npz"
# BASELINE.json configs.  cfg2 is the one the metric is quoted on (the default and the only driver-run line); the others
# are the parity-test configurations, measurable with --workload for the record (profiles/).
WORKLOADS = {
    "cfg2": dict(lib="pg3_density0.1.npz", envs=8192, peds=0,
                 name="MetaDriveEnv PG 3-block maps, IDM traffic density 0.1, 1000 reference scenarios, 240-beam lidar"),
    "cfg4": dict(lib="safe_pg3.npz", envs=8192, peds=0,
                 name="SafeMetaDriveEnv (accident_prob 0.8, static obstacles, cost), 100 reference scenarios, 240-beam lidar"),
    "cfg5": dict(lib="x_respawn_density0.1.npz", envs=4096, peds=16,
                 name="MetaDriveEnv map X, respawn-mode IDM traffic + 16 crossing pedestrians per env, 240-beam lidar"),
    "cfg3": dict(lib=None, envs=2048, peds=0,
                 name="MultiAgentRoundaboutEnv 40 agents per env, respawn on, 240-beam lidar + crash checks"),
}
# SURVEY.md 8(d) algorithmic bytes per unit of work
B_EGO = 1684.0
B_TRAFFIC = 560.0
B_LIDAR_SHARE = 256.0 + 960.0  # neighbour footprints read + the 240 lidar floats written (k_lidar's part of B_EGO)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self._stop = index, [], threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self.t.join(timeout=3)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def build_world(n_envs, rank, workload="cfg2"):
    w = WORKLOADS[workload]
    if workload == "cfg3":
        from metadrive_ped_b200.envs import MultiAgentRoundaboutEnv, _ma_cfg_kw, _merge
        from metadrive_ped_b200.ma import MultiAgentLibrary
        c = _merge(MultiAgentRoundaboutEnv.default_config(), {"vehicle_config": {"lidar": {"num_lasers": 240, "distance": 50}}})
        lib = MultiAgentLibrary(MultiAgentRoundaboutEnv.ASSET)
        arrays, cfg = lib.build_world(n_envs, c["num_agents"], seed=rank, **_ma_cfg_kw(c))
        return lib, arrays, cfg
    from metadrive_ped_b200.library import ScenarioLibrary
    lib = ScenarioLibrary(w["lib"])
    idx = [(rank * n_envs + e) % len(lib) for e in range(n_envs)]
    arrays, cfg = lib.build_world(idx, num_pedestrians=w["peds"], seed=rank, **bank_kw(lib, workload))
    return lib, arrays, cfg


def bank_kw(lib, workload):
    """Trigger-mode workloads resample the scenario at every reset, like `env.reset()` in the reference's profiling loop
    (examples/profile_metadrive.py:26-29 -> envs/base_env.py:886-891): both the live world and the scenario bank load the
    whole library's map set and use the library-wide slot / object capacity."""
    if workload not in ("cfg2", "cfg4"):
        return {}
    return dict(slots_per_env=max(4, -(-lib.max_vehicles() // 4) * 4), objs_per_env=lib.max_objects(),
                map_universe=list(range(len(lib))))


def workload_name(workload="cfg2"):
    return WORKLOADS[workload]["name"]


def run_reference(args, rank):
    """CPU arm: the oracle port of the reference's path on all host threads, bounded sample per step."""
    if rank != 0:
        return
    from oracle.oracle import OracleSim, set_threads
    cores = set_threads()  # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host thread
    sample = 1024
    lib, arrays, cfg = build_world(sample, 0)
    orc = OracleSim(arrays, cfg)
    orc.reset_observe()
    a = np.tile(np.array([0.0, 1.0], np.float32), (sample, 1))
    for _ in range(args.warmup):
        orc.step(a)
        orc.reset_envs(orc.term | orc.trunc)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.step(a)
        done = (orc.term | orc.trunc).astype(bool)
        if done.any():
            orc.reset_envs(done)
            # reset observation of the restored envs: the oracle recomputes it with the next step's observe
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": "agent_steps_per_sec_240beam_lidar", "value": v, "unit": "agent-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(), "envs_per_step": sample, "actions": "[0,1] (profile_metadrive.py)"},
        "cpu_baseline": {"value": v, "unit": "agent-steps/s", "cores": cores, "kind": "port",
                         "sample": "%d envs x %d steps, OpenMP over envs" % (sample, args.steps)},
        "e2e": {"value": v, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "lidar_rays_per_sec": v * 240,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=None)
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--actions", default="profile", choices=["profile", "random"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-resample", action="store_true", help="finished envs replay their own scenario instead of drawing a new one")
    ap.add_argument("--host-groups", type=int, default=4, help="host groups of the e2e leg (md_host_groups)")
    ap.add_argument("--burnin", type=int, default=150,
                    help="untimed setup steps (with auto-reset) so that envs sit at mixed episode phases")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (ours) needs a GPU: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from metadrive_ped_b200.sim import BatchedSim
    E = args.envs_per_gpu or WORKLOADS[args.workload]["envs"]
    multi = args.workload == "cfg3"
    t_build = time.time()
    lib, arrays, cfg = build_world(E, rank, args.workload)
    sim = BatchedSim(arrays, cfg, device=local_rank)
    resample = args.workload in ("cfg2", "cfg4") and not args.no_resample
    if resample:  # the scenario bank: one env per library scenario, fully reset; finished envs draw their next scenario from it
        b_arrays, b_cfg = lib.build_world(list(range(len(lib))), seed=rank, **bank_kw(lib, args.workload))
        bank = BatchedSim(b_arrays, b_cfg, device=local_rank)
        bank.reset()
        sim.reset()
        sim.attach_bank(bank, seed=1000 + rank)
    t_build = time.time() - t_build
    A = sim.n_agents
    kind = arrays["veh_i"][:, 0].reshape(E, cfg.slots_per_env)
    traffic_per_env = float((kind == 2).sum(1).mean())

    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    if multi:  # small steering noise, throttle in [0, 1]: agents drive, crash, leave the road, arrive and are respawned
        act_dev = torch.rand((A, 2), generator=g, device=dev)
        act_dev[:, 0] = (act_dev[:, 0] - 0.5) * 0.2
        act_dev = act_dev.contiguous()
    elif args.actions == "profile":
        act_dev = torch.tensor([0.0, 1.0], device=dev).repeat(A, 1).contiguous()
    else:
        act_dev = (torch.rand((A, 2), generator=g, device=dev) * 2 - 1).contiguous()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sim.reset()
    for _ in range(args.burnin):  # setup, not warm-up: de-synchronise the episodes (traffic triggered, resets spread)
        sim.step(act_dev, autoreset=True)
    for _ in range(args.warmup):
        sim.step(act_dev, autoreset=True)
    barrier()

    # ---------------- timed region: device-resident inputs, CUDA events on the launch stream, L2 flushed between steps
    K = args.steps
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    launches0 = sim.launch_count
    sim.profile_begin(K)
    done_count = torch.zeros((), dtype=torch.int64, device=dev)
    valid_count = torch.zeros((), dtype=torch.int64, device=dev)
    with ClockSampler(local_rank) as clocks:
        barrier()
        for k in range(K):
            flush.zero_()
            ev0[k].record()
            sim.step(act_dev, autoreset=True)
            ev1[k].record()
            done_count += (sim.terminated | sim.truncated).sum()
            if multi:
                valid_count += ((sim.info_flags & 0x2000) != 0).sum()
        barrier()
    step_ms = np.array([ev0[k].elapsed_time(ev1[k]) for k in range(K)])
    kms = sim.profile_end()  # [K, 5]: k_pre, k_dyn, k_post, fused reset (k_post in reset mode), k_lidar
    launches = sim.launch_count - launches0
    total_ms = torch.tensor([float(step_ms.sum())], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_s = float(total_ms.item()) / 1e3
    # agent-steps: every env's agent steps every step; in the multi-agent workload only the seats that produced a
    # transition count (wrecks waiting out delay_done and empty seats do not)
    units = torch.tensor([float(valid_count.item()) if multi else float(A * K)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(units, op=dist.ReduceOp.SUM)
    value = float(units.item()) / total_s
    live_frac = float(units.item()) / (world * A * K)

    # ---------------- e2e: the host-buffer API (pinned staging, H2D actions + D2H outputs of EVERY env EVERY step inside the
    # timed region), wall clock.  Two numbers: `sync` = one md_step_host call per step over --host-groups groups (group k's
    # D2H overlaps group k+1's kernels inside the call); `pipelined` (the headline) = the same groups driven through
    # send / recv, the EnvPool-style split of the batch: group g's next actions are sent only after its observations
    # were received, while the other groups are being stepped - the way a host-side learner hides PCIe time.
    a_host = act_dev.cpu().numpy()
    G = max(1, args.host_groups)
    sim.host_groups(G)
    if multi:
        sim.host_compact(True)
    gv = sim._group_views()
    a_grp = [np.ascontiguousarray(a_host[g["a0"]:g["a0"] + g["na"]]) for g in gv]
    Ke = max(10, min(K, 100))

    def e2e_sync(n):
        rows = 0
        for _ in range(n):
            out = sim.step_host(a_host, autoreset=True)
            rows += out[0].shape[0]
        return rows

    def e2e_pipelined(n):
        rows = 0
        for g in range(G):
            sim.send(g, a_grp[g], autoreset=True)
        for i in range(n):
            for g in range(G):
                out = sim.recv(g)
                rows += out[0].shape[0]
                if i + 1 < n:
                    sim.send(g, a_grp[g], autoreset=True)
        return rows

    e2e = {}
    for name, fn in (("sync", e2e_sync), ("pipelined", e2e_pipelined)):
        fn(3)
        barrier()
        t0 = time.perf_counter()
        rows = fn(Ke)
        torch.cuda.synchronize(dev)
        dt_e = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt_e, op=dist.ReduceOp.MAX)
        e2e[name] = (world * A * Ke * live_frac / float(dt_e.item()), rows / Ke)
    e2e_value = e2e["pipelined"][0]
    obs_rows = e2e["pipelined"][1]     # observation rows copied per step (multi-agent: only the FL_VALID seats travel)
    h2d = A * 2 * 4
    d2h = int(obs_rows * sim.obs_dim * 4 + A * (4 + 4 + 1 + 1 + 4 + 8 * 4))

    # ---------------- episode statistics: the only cross-GPU exchange of this path (one tiny all-reduce)
    stats = torch.tensor([float(done_count.item()), float(A * K)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)

    if rank == 0:
        peak, peak_src = load_peaks()
        names = ["k_pre", "k_dyn", "k_post", "k_reset", "k_lidar"]
        mean_ms = kms.mean(0)
        dom_i = int(np.argmax(mean_ms))
        dom = names[dom_i]
        # algorithmic bytes per launch of each kernel: the SURVEY 8(d) per-unit figures split
        # by what each kernel must touch; T = alive traffic vehicles per env
        T = traffic_per_env
        NAg = float(cfg.agents_per_env) * live_frac  # agents that step per env (1 in the single-agent workloads)
        NP = float(WORKLOADS[args.workload]["peds"])
        per_env = {
            "k_pre": NAg * (8 + 64) + T * (64 + 256),       # action + latches r/w ; traffic: PID/timer/route r/w + neighbours
            "k_dyn": (192 + 48) * (NAg + T) + 32 * NP,      # state r/w + params, every vehicle; pedestrians pos/vel r/w
            "k_post": NAg * (64 + 76 + 16),                 # episode/nav state r/w + 19 state floats + scalars (agents)
            "k_reset": 0.0,                                 # auto-reset of finished envs: not part of the per-step figure
            "k_lidar": NAg * B_LIDAR_SHARE,                 # neighbour footprints + 240 lidar floats
        }
        B_STEP = NAg * B_EGO + B_TRAFFIC * T + 32 * NP      # bytes per env-step
        assert abs(sum(per_env.values()) - B_STEP) < 1.0, per_env
        bytes_per_launch = E * per_env[dom]
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                traffic = json.load(f).get(dom)
        except Exception:
            pass
        dur_ms = float(mean_ms[dom_i])
        achieved = bytes_per_launch / (dur_ms * 1e-3) / 1e9
        line = {
            "metric": "agent_steps_per_sec_240beam_lidar", "value": value, "unit": "agent-steps/s", "n_gpus": world,
            "steps": K, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload), "baseline_config": args.workload, "envs_per_gpu": E,
                       "agents_per_env": cfg.agents_per_env - (1 if multi else 0), "live_agent_fraction": live_frac,
                       "slots_per_env": cfg.slots_per_env, "traffic_per_env_mean": traffic_per_env,
                       "distinct_scenarios": (E * world if multi else min(len(lib), E * world)), "actions": args.actions,
                       "autoreset": "on device, inside the timed region" + (
                           "; every reset draws a new scenario from the %d-scenario bank (md_attach_bank)" % len(lib) if resample else ""),
                       "l2": "flushed between steps (256 MiB memset, untimed)",
                       "scene_build_s": round(t_build, 1), "burnin_steps": args.burnin},
            "lidar_rays_per_sec": value * cfg.n_lasers,
            "gpu_launches": int(launches),
            "kernel_ms": {**{n: float(v) for n, v in zip(names, mean_ms)},
                          "step_total_incl_autoreset": float(step_ms.mean())},
            "step_roofline_all_kernels": {"algorithmic_bytes_per_step": E * B_STEP,
                                          "achieved_gbs": E * B_STEP / (float(mean_ms.sum()) * 1e-3) / 1e9},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_per_launch},
            "e2e": {"value": e2e_value, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": Ke, "api": "BatchedSim.send/recv over %d host groups (md_host_send / md_host_recv): every env gets "
                                        "its actions H2D and its results D2H every step; a group's next actions are sent "
                                        "after its results were received" % G,
                    "sync_call_value": e2e["sync"][0], "host_groups": G},
            "clocks": clocks.summary(),
            "episodes_finished_frac": float(stats[0].item() / max(stats[1].item(), 1.0)),
        }
        if world == 1 and not args.no_cpu_baseline and not multi:
            line["cpu_baseline"] = cpu_baseline(lib, WORKLOADS[args.workload]["peds"])
        print(json.dumps(line), flush=True)
    sim.close()
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(lib, peds=0):
    """The oracle port on the box's host cores over a bounded sample (about 10-20 s of CPU work)."""
    from oracle.oracle import OracleSim, set_threads
    cores = set_threads()
    sample = 512
    arrays, cfg = lib.build_world([i % len(lib) for i in range(sample)], num_pedestrians=peds)
    orc = OracleSim(arrays, cfg)
    orc.reset_observe()
    a = np.tile(np.array([0.0, 1.0], np.float32), (sample, 1))
    t0 = time.perf_counter()
    orc.step(a)
    one = max(time.perf_counter() - t0, 1e-4)
    steps = int(max(5, min(20000, 12.0 / one)))
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.step(a)
        done = (orc.term | orc.trunc).astype(bool)
        if done.any():
            orc.reset_envs(done)
    dt = time.perf_counter() - t0
    return {"value": sample * steps / dt, "unit": "agent-steps/s", "cores": cores, "kind": "port",
            "sample": "%d envs x %d steps of the same workload, OpenMP over envs" % (sample, steps)}


if __name__ == "__main__":
    main()

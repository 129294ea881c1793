#!/usr/bin/env python
"""bench.py — agent-steps/s of the batched MetaDrive step with 240-beam lidar observations (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # ours (for N>1 launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the CPU arm, rank 0 only

Workload (configs[1] of BASELINE.json): `MetaDriveEnv(map=3, traffic_density=0.1, num_scenarios=1000)` — 3-block PG
maps with IDM traffic, 8192 environments per GPU cycling through the 1000 reference scenarios of the shipped
library, reference profiling protocol (examples/profile_metadrive.py:16-29): action [0, 1], finished envs reset
in place (on device).  A step = one env.step of every env: before_step (actuation, trigger, IDM), 5 physics
sub-steps with contacts, after_step, reward/cost/done, 259-float observation, plus the auto-reset of finished envs.
The other BASELINE configurations (cfg3 multi-agent roundabout, cfg4 SafeMetaDriveEnv, cfg5 pedestrian intersection)
ride along as short `other_configs` entries of the same JSON line (or alone with --workload).
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

# BASELINE.json configs.  cfg2 is the one the metric is quoted on (the headline line); cfg4 = 65,536 envs over 8 GPUs and
# cfg5 = 32,768 envs over 8 GPUs are 8,192 / 4,096 envs per GPU; cfg3 = 40 agents x 2,048 envs per GPU.
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
KERNELS = ["k_pre", "k_dyn", "k_post", "k_reset", "k_lidar"]


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_ncu(kernel):
    """Per-launch DRAM bytes and FP32 flops of a kernel from this round's `ncu --set full` capture of this very command
    (profiles/r02_ncu_counters.json, written by scripts/ncu_counters.py); None when there is no capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_ncu_counters.json")) as f:
            d = json.load(f)
        return d.get(kernel), d.get("_source")
    except Exception:
        return None, None


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
    """CPU arm: the reference's path on the box's host cores - the oracle port (`kind: port`; the reference itself needs
    panda3d, absent from this image), every host thread, on the SAME configuration as our arm: the 8192 envs of the
    workload, action [0, 1], finished envs reset in place.  Each step is one pass over all 8192 envs (about 50 ms on 16
    cores), so the default --steps / --warmup end within seconds."""
    if rank != 0:
        return
    from oracle.oracle import OracleSim, set_threads
    cores = set_threads()  # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host thread
    wl = args.workload
    E = args.envs_per_gpu or WORKLOADS[wl]["envs"]
    lib, arrays, cfg = build_world(E, 0, wl)
    orc = OracleSim(arrays, cfg)
    orc.reset_observe()
    A = orc.n_agents
    a = np.tile(np.array([0.0, 1.0], np.float32), (A, 1))
    steps = max(1, min(args.steps, 200))
    for _ in range(min(args.warmup, 20)):
        orc.step(a)
        orc.reset_envs(orc.term | orc.trunc)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.step(a)
        done = (orc.term | orc.trunc).astype(bool)
        if done.any():
            orc.reset_envs(done)   # the restored envs' reset observation comes with the next step's observe
    dt = time.perf_counter() - t0
    v = A * steps / dt
    line = {
        "impl": "reference", "metric": "agent_steps_per_sec_240beam_lidar", "value": v, "unit": "agent-steps/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(wl), "baseline_config": wl, "envs_per_gpu": E,
                   "actions": "profile"},
        "cpu_baseline": {"value": v, "unit": "agent-steps/s", "cores": cores, "kind": "port",
                         "sample": "%d envs x %d steps (the whole workload every step), OpenMP over envs" % (E, steps)},
        "e2e": {"value": v, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "lidar_rays_per_sec": v * 240,
    }
    print(json.dumps(line), flush=True)


def ma_driver(torch, obs, sd, gen, noise=0.05):
    """Obs-driven driver of the multi-agent workload (the lane-follow + noise driver of oracle/gen_golden.py, from the
    observation instead of the vehicle object): steer along the route lane - heading error from the `heading_diff` entry
    (obs/state_obs.py:115), lateral offset from the lane-centre entry (:141-148), a pull towards the next checkpoint from
    the navigation block (node_network_navigation.py:243-292) - and hold ~28 km/h, slower where the route bends.  Keeps
    most agents alive until they arrive, so the 40 seats of an env really hold ~40 driving agents."""
    herr = torch.asin((2.0 * obs[:, sd] - 1.0).clamp(-1.0, 1.0))           # lane heading - vehicle heading
    lat = (2.0 * obs[:, sd + 6] - 1.0) * 2.25                              # +: right of the lane centre
    fwd, rhs = (2.0 * obs[:, sd + 7] - 1.0) * 50.0, (2.0 * obs[:, sd + 8] - 1.0) * 50.0
    near = (fwd * fwd + rhs * rhs) < 64.0                                  # almost at checkpoint 1: look at checkpoint 2
    fwd = torch.where(near, (2.0 * obs[:, sd + 12] - 1.0) * 50.0, fwd)
    rhs = torch.where(near, (2.0 * obs[:, sd + 13] - 1.0) * 50.0, rhs)
    ang = torch.atan2(rhs, fwd.clamp_min(1.0))                              # navi 'rhs' grows to the LEFT (in_rhs = -d.right)
    bend = torch.maximum(obs[:, sd + 9], obs[:, sd + 14])                  # radius / (60 + n * w); 0 on straights
    v = obs[:, sd + 1] * 81.0 - 1.0
    target = torch.where((bend > 0.0) & (bend < 0.3), 16.0, 28.0)
    u = torch.rand((obs.shape[0], 2), generator=gen, device=obs.device) * 2.0 - 1.0
    steer = (2.2 * herr + 0.45 * lat + 0.4 * ang + noise * u[:, 0]).clamp(-1.0, 1.0)
    thr = torch.where(v < target, 0.6, torch.where(v > target + 6.0, -0.3, 0.0)) + noise * u[:, 1]
    return torch.stack([steer, thr.clamp(-1.0, 1.0)], 1).contiguous()


def measure(args, workload, rank, local_rank, world, K, warmup, full):
    """One workload on this rank's GPU.  `full`: the headline protocol (per-kernel roofline, both e2e legs, clocks); else a
    short line for `other_configs`.  Returns the dict rank 0 prints (None on other ranks)."""
    import torch
    import torch.distributed as dist
    from metadrive_ped_b200.sim import BatchedSim
    dev = torch.device("cuda", local_rank)
    E = (args.envs_per_gpu if full and args.envs_per_gpu else None) or WORKLOADS[workload]["envs"]
    multi = workload == "cfg3"
    t_build = time.time()
    lib, arrays, cfg = build_world(E, rank, workload)
    sim = BatchedSim(arrays, cfg, device=local_rank)
    resample = workload in ("cfg2", "cfg4") and not args.no_resample
    bank = None
    if resample:  # the scenario bank: one env per library scenario, fully reset; finished envs draw their next scenario from it
        b_arrays, b_cfg = lib.build_world(list(range(len(lib))), seed=rank, **bank_kw(lib, workload))
        bank = BatchedSim(b_arrays, b_cfg, device=local_rank)
        bank.reset()
        sim.reset()
        sim.attach_bank(bank, seed=1000 + rank)
    t_build = time.time() - t_build
    A = sim.n_agents
    kind = arrays["veh_i"][:, 0].reshape(E, cfg.slots_per_env)
    traffic_per_env = float((kind == 2).sum(1).mean())
    sd = (cfg.n_side_lasers or 2)

    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    if multi:
        actions_kind = "obs-driven lane-follow driver + noise (bench.ma_driver), recomputed on device every step"
        policy = lambda: ma_driver(torch, sim.obs, sd, g)
    elif args.actions == "profile":
        actions_kind = "profile"
        const = torch.tensor([0.0, 1.0], device=dev).repeat(A, 1).contiguous()
        policy = lambda: const
    else:
        actions_kind = "random"
        const = (torch.rand((A, 2), generator=g, device=dev) * 2 - 1).contiguous()
        policy = lambda: const
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sim.reset()
    for _ in range(args.burnin):  # setup, not warm-up: de-synchronise the episodes (traffic triggered, resets spread)
        sim.step(policy(), autoreset=True)
    for _ in range(warmup):
        sim.step(policy(), autoreset=True)
    barrier()

    # ---------------- timed region: device-resident inputs, CUDA events on the launch stream, L2 flushed between steps
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    launches0 = sim.launch_count
    done_count = torch.zeros((), dtype=torch.int64, device=dev)
    valid_count = torch.zeros((), dtype=torch.int64, device=dev)
    with ClockSampler(local_rank) as clocks:
        barrier()
        for k in range(K):
            act = policy()
            flush.zero_()
            ev0[k].record()
            sim.step(act, autoreset=True)
            ev1[k].record()
            done_count += (sim.terminated | sim.truncated).sum()
            if multi:
                valid_count += ((sim.info_flags & 0x2000) != 0).sum()
        barrier()
    step_ms = np.array([ev0[k].elapsed_time(ev1[k]) for k in range(K)])
    launches = sim.launch_count - launches0
    # ---------------- second pass, same protocol, for the per-kernel times: md_profile_begin records six CUDA events per step
    # on the launch stream, between the stages.  Those events cost the step 16 us (5 %: each one separates two kernels that
    # otherwise run back to back), so the headline region above runs without them and the per-kernel durations of the
    # roofline come from this pass; `kernel_ms.step_total_with_stage_events` is this pass's step time.
    Kp = K if full else min(K, 20)
    pe0 = [torch.cuda.Event(enable_timing=True) for _ in range(Kp)]
    pe1 = [torch.cuda.Event(enable_timing=True) for _ in range(Kp)]
    sim.profile_begin(Kp)
    for k in range(Kp):
        act = policy()
        flush.zero_()
        pe0[k].record()
        sim.step(act, autoreset=True)
        pe1[k].record()
    kms = sim.profile_end()  # [Kp, 5]: k_pre, k_dyn, k_scan + k_post, reset launch (0 when fused into k_post), k_lidar
    prof_step_ms = float(np.mean([pe0[k].elapsed_time(pe1[k]) for k in range(Kp)]))
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
    n_ag = cfg.agents_per_env - (1 if multi else 0)                 # a multi-agent env carries one spare seat
    live_frac = float(units.item()) / (world * E * n_ag * K)

    # ---------------- e2e: the host-buffer API (pinned staging, H2D actions + D2H outputs of EVERY env EVERY step inside the
    # timed region), wall clock.  `pipelined` (the headline) = --host-groups groups driven through send / recv, the
    # EnvPool-style split of the batch: group g's next actions are sent only after its results were received, while the
    # other groups are being stepped - the way a host-side learner hides PCIe time.  `sync` = one md_step_host call per step
    # over the same groups (group k's D2H overlaps group k+1's kernels inside the call).
    a_host = policy().cpu().numpy()
    G = max(1, args.host_groups)
    sim.host_groups(G)
    if multi:
        sim.host_compact(True)
    gv = sim._group_views()
    for gr in gv:
        gr["actions"][...] = a_host[gr["a0"]:gr["a0"] + gr["na"]]       # actions sit in the pinned buffers; H2D every step
    Ke = max(10, min(K, 100))

    def e2e_sync(n):
        rows = 0
        for _ in range(n):
            rows += sim.step_host(a_host, autoreset=True)[0].shape[0]
        return rows

    def e2e_pipelined(n):
        rows = 0
        for gi in range(G):
            sim.send(gi, None, autoreset=True)
        for i in range(n):
            for gi in range(G):
                rows += sim.recv(gi)[0].shape[0]
                if i + 1 < n:
                    sim.send(gi, None, autoreset=True)
        return rows

    e2e = {}
    for name, fn in (("sync", e2e_sync), ("pipelined", e2e_pipelined)) if full else (("pipelined", e2e_pipelined), ):
        fn(3)
        barrier()
        t0 = time.perf_counter()
        rows = fn(Ke)
        torch.cuda.synchronize(dev)
        dt_e = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt_e, op=dist.ReduceOp.MAX)
        e2e[name] = (world * E * n_ag * Ke * live_frac / float(dt_e.item()), rows / Ke)
    # what the host link gives while every rank copies at once: a pinned 64 MiB device-to-host copy, all ranks started together.
    # (On the 8-GPU boxes of this pool the GPUs share the host's PCIe / memory fabric: the per-GPU rate with 8 ranks active is a
    # fraction of the rate of a GPU copying alone, and it caps e2e - see d2h_probe_gbs_per_gpu next to d2h_achieved_gbs_per_gpu.)
    probe_gbs = None
    if full:
        pd = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
        ph = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
        ph.copy_(pd, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(10):
            ph.copy_(pd, non_blocking=True)
        torch.cuda.synchronize(dev)
        dt_p = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt_p, op=dist.ReduceOp.MAX)
        probe_gbs = 10 * (64 << 20) / float(dt_p.item()) / 1e9
        del pd, ph
    obs_rows = e2e["pipelined"][1]     # observation rows copied per step (multi-agent: only the FL_VALID seats travel)
    h2d = A * 2 * 4
    d2h = int(obs_rows * sim.obs_dim * 4 + A * (4 + 4 + 1 + 1 + 4 + 8 * 4))

    # ---------------- episode statistics: the only cross-GPU exchange of this path (one tiny all-reduce)
    stats = torch.tensor([float(done_count.item()), float(A * K)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    line = None
    if rank == 0:
        mean_ms = kms.mean(0)
        config = {"workload": workload_name(workload), "baseline_config": workload, "envs_per_gpu": E,
                  "agents_per_env": n_ag, "live_agent_fraction": live_frac,
                  "slots_per_env": cfg.slots_per_env, "traffic_per_env_mean": traffic_per_env,
                  "distinct_scenarios": (E * world if multi else min(len(lib), E * world)), "actions": actions_kind,
                  "autoreset": "on device, inside the timed region" + (
                      "; every reset draws a new scenario from the %d-scenario bank (md_attach_bank)" % len(lib) if resample else ""),
                  "l2": "flushed between steps (256 MiB memset, untimed)",
                  "dynamic_broad_phase": "per-env shared-memory scan (slots + objects <= 128 per env; md_create fails above)",
                  "scene_build_s": round(t_build, 1), "burnin_steps": args.burnin}
        e2e_d = {"value": e2e["pipelined"][0], "unit": "agent-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                 "steps": Ke, "host_groups": G,
                 "api": "BatchedSim.send/recv over %d host groups (md_host_send / md_host_recv): every env gets its actions "
                        "H2D and its results D2H every step; a group's next actions are sent after its results were received" % G}
        if "sync" in e2e:
            e2e_d["sync_call_value"] = e2e["sync"][0]
        if probe_gbs is not None:
            steps_per_s = e2e["pipelined"][0] / (world * E * n_ag * live_frac)
            e2e_d["d2h_achieved_gbs_per_gpu"] = d2h * steps_per_s / 1e9
            e2e_d["d2h_probe_gbs_per_gpu"] = probe_gbs
            e2e_d["d2h_probe"] = "64 MiB pinned device-to-host copies, all %d ranks copying at once (slowest rank)" % world
        line = {"metric": "agent_steps_per_sec_240beam_lidar", "value": value, "unit": "agent-steps/s", "n_gpus": world,
                "steps": K, "warmup": warmup, "ms_per_step": 1e3 * total_s / K, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "lidar_rays_per_sec": value * cfg.n_lasers, "gpu_launches": int(launches),
                "kernel_ms": {**{n: float(v) for n, v in zip(KERNELS, mean_ms)},
                              "step_total_incl_autoreset": float(step_ms.mean()),
                              "step_total_with_stage_events": prof_step_ms},
                "kernel_ms_note": "per-kernel times: a second pass of %d steps with 6 stage events per step (k_post = k_scan + k_post; "
                                  "k_reset = the reset launch, fused into k_post when a scenario bank is attached); the timed "
                                  "region of `value` runs without those events" % Kp,
                "e2e": e2e_d}
        if full:
            peak, peak_src = load_peaks()
            dom_i = int(np.argmax(mean_ms))
            dom = KERNELS[dom_i]
            # algorithmic bytes per launch of each kernel: the SURVEY 8(d) per-unit figures split
            # by what each kernel must touch; T = alive traffic vehicles per env
            T = traffic_per_env
            NAg = float(n_ag) * live_frac  # agents that step per env (1 in the single-agent workloads)
            NP = float(WORKLOADS[workload]["peds"])
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
            ncu, ncu_src = load_ncu(dom)
            dur_ms = float(mean_ms[dom_i])
            achieved = bytes_per_launch / (dur_ms * 1e-3) / 1e9
            line["step_roofline_all_kernels"] = {"algorithmic_bytes_per_step": E * B_STEP,
                                                 "achieved_gbs": E * B_STEP / (float(step_ms.mean()) * 1e-3) / 1e9}
            line["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                                "frac": achieved / peak, "traffic": (ncu or {}).get("dram_bytes"), "traffic_source": ncu_src,
                                "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_launch}
            # secondary roofline: FP32 pipe.  Peak = measured FFMA micro-benchmark on this GPU; flops per
            # launch = FADD + FMUL + 2 FFMA thread-instructions of the kernel from the round's ncu capture of this command.
            fp32_peak = sim.fp32_peak()
            rf = {"peak": fp32_peak, "unit": "TFLOP/s", "peak_source": "measured in this run (md_fp32_peak: independent FFMA chains)",
                  "kernels": {}}
            for i, n in enumerate(KERNELS):
                c, _ = load_ncu(n)
                if c and c.get("fp32_flops") and mean_ms[i] > 0:
                    ach = c["fp32_flops"] / (float(mean_ms[i]) * 1e-3) / 1e12
                    rf["kernels"][n] = {"flops_per_launch": c["fp32_flops"], "achieved": ach, "frac": ach / fp32_peak}
            rf["flops_source"] = ncu_src
            line["roofline_fp32"] = rf
            line["clocks"] = clocks.summary()
            line["episodes_finished_frac"] = float(stats[0].item() / max(stats[1].item(), 1.0))
            if world == 1 and not args.no_cpu_baseline and not multi:
                line["cpu_baseline"] = cpu_baseline(lib, WORKLOADS[workload]["peds"])
    sim.close()
    if bank is not None:
        bank.close()
    del flush
    torch.cuda.empty_cache()
    return line


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
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short cfg3 / cfg4 / cfg5 entries")
    ap.add_argument("--no-resample", action="store_true", help="finished envs replay their own scenario instead of drawing a new one")
    ap.add_argument("--host-groups", type=int, default=2, help="host groups of the e2e leg (md_host_groups)")
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
    from metadrive_ped_b200.shard import bind_to_gpu_numa
    numa = bind_to_gpu_numa(local_rank)   # before any pinned allocation: the staging buffers land on the GPU's NUMA node
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    line = measure(args, args.workload, rank, local_rank, world, args.steps, args.warmup, True)
    others = []
    if args.workload == "cfg2" and not args.no_other_configs:
        for wl in ("cfg3", "cfg4", "cfg5"):   # short entries so that every BASELINE config has a driver-run line at every N
            o = measure(args, wl, rank, local_rank, world, 20, 5, False)
            if o is not None:
                others.append({k: o[k] for k in ("value", "unit", "n_gpus", "steps", "ms_per_step", "config", "kernel_ms", "e2e",
                                                 "gpu_launches", "lidar_rays_per_sec")})
    if rank == 0:
        line["numa"] = numa
        if others:
            line["other_configs"] = others
        print(json.dumps(line), flush=True)
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

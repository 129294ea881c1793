"""-m gpu: the BASELINE.json configurations at their FULL per-GPU sizes (bench.py's worlds), every agent with its own
random action sequence, fused on-device auto-reset on.

1. CUDA step vs the CPU oracle, env by env: integer state (lanes, checkpoints, flags, rosters, seat bookkeeping), done
   flags, info flags AND every float of the rigid-body state and of the observations are bit-identical (the two float32
   paths share include/md_math.h and the operation order; ~10^9 floats are compared per run of this file).
2. Size-independent properties: shuffling the envs of the batch shuffles the outputs bit for bit (no cross-env
   leakage through the CTA-level staging, work lists or candidate sets), and two runs are bit-identical (determinism).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MUTABLE = ("env_i", "veh_s", "veh_c", "veh_i", "veh_idm", "veh_navi", "obj_f", "veh_route", "veh_rroad")


def _world(workload, n_envs=None):
    import bench
    if workload == "toll":
        # MultiAgentTollgateEnv at its default 40 agents (envs/marl_envs/marl_tollgate.py:15-36), 48 envs: booths, the toll
        # observation, overspeed penalties and the stay-time rule under random driving
        from metadrive_ped_b200.envs import MultiAgentTollgateEnv, _ma_cfg_kw
        from metadrive_ped_b200.ma import MultiAgentLibrary
        c = MultiAgentTollgateEnv.default_config()
        return MultiAgentLibrary(MultiAgentTollgateEnv.ASSET).build_world(n_envs or 48, c["num_agents"], seed=3, **_ma_cfg_kw(c))
    if workload == "matr":
        # a multi-agent env with IDM traffic: 12 agents on the roundabout, trigger-mode traffic of density 0.15 (another draw per env)
        from metadrive_ped_b200.envs import MultiAgentRoundaboutEnv, _apply_vehicle_config, _ma_cfg_kw
        from metadrive_ped_b200.ma import MultiAgentLibrary
        c = MultiAgentRoundaboutEnv.default_config()
        c["traffic_density"] = 0.15
        arrays, cfg = MultiAgentLibrary(MultiAgentRoundaboutEnv.ASSET).build_world(n_envs or 48, 12, seed=7, traffic_density=0.15,
                                                                                  traffic_seed=0, **_ma_cfg_kw(c))
        _apply_vehicle_config(arrays, c)
        return arrays, cfg
    if workload == "park":
        # MultiAgentParkingLotEnv at its default 10 agents (envs/marl_envs/marl_parking_lot.py:22-43), 64 envs: pulling out of the
        # spaces, reversing, ParkingLotSpawnManager's respawn rules under random driving
        from metadrive_ped_b200.envs import MultiAgentParkingLotEnv, _apply_vehicle_config, _ma_cfg_kw
        from metadrive_ped_b200.ma import MultiAgentLibrary
        c = MultiAgentParkingLotEnv.default_config()
        arrays, cfg = MultiAgentLibrary(MultiAgentParkingLotEnv.ASSET).build_world(n_envs or 64, c["num_agents"], seed=5, **_ma_cfg_kw(c))
        _apply_vehicle_config(arrays, c)
        return arrays, cfg
    n = n_envs or bench.WORKLOADS[workload]["envs"]
    _, arrays, cfg = bench.build_world(n, 0, workload)
    return arrays, cfg


def valid_rows(m_ag, fl_o, multi):
    return m_ag & (((fl_o & 0x2000) != 0) if multi else True)


def _actions(rng, cfg, multi):
    """Every agent its own action sequence: throttle mostly forward, steering = a per-agent bias + noise, so that within
    a few dozen steps agents leave the road, hit traffic / cones / pedestrians, and (multi-agent) arrive and respawn."""
    n = cfg.n_envs * cfg.agents_per_env
    if cfg.toll_env:
        # the tollgate map is one straight road: nearly straight driving at every agent's own pace brings most of them to the toll
        # block (some into a booth, most through it too fast, a slow few staying long enough), a weaving tenth off the lanes
        a = np.zeros((n, 2), np.float32)
        weave = np.random.RandomState(12).uniform(0, 1, n) < 0.1
        a[:, 0] = np.where(weave, 0.15, 0.01) * rng.uniform(-1.0, 1.0, n)
        pace = np.random.RandomState(13).uniform(0.05, 1.0, n)
        a[:, 1] = pace * rng.uniform(0.2, 1.0, n)
        return a.astype(np.float32)
    if cfg.parking_spaces:
        # a parking lot: gentle throttle, now and then the brake / reverse gear, every agent with its own steering bias
        a = np.zeros((n, 2), np.float32)
        a[:, 0] = np.random.RandomState(14).uniform(-0.6, 0.6, n) + 0.3 * rng.uniform(-1.0, 1.0, n)
        a[:, 1] = np.where(rng.uniform(0, 1, n) < 0.15, -0.6, 0.35) * rng.uniform(0.3, 1.0, n)
        return a.astype(np.float32)
    bias = np.random.RandomState(11).uniform(-0.25, 0.25, n)
    a = rng.uniform(-1.0, 1.0, (n, 2)).astype(np.float32)
    a[:, 0] = (0.3 * a[:, 0] + bias).astype(np.float32)
    a[:, 1] = np.where(a[:, 1] > -0.8, np.abs(a[:, 1]), a[:, 1])
    return a


@pytest.mark.parametrize("workload,steps", [("cfg2", 60), ("cfg4", 40), ("cfg5", 40), ("cfg3", 25), ("toll", 150), ("park", 400), ("matr", 200)])
def test_full_size_step_matches_oracle(workload, steps, oracle_lib):
    import torch
    from metadrive_ped_b200.sim import BatchedSim
    from oracle.oracle import OracleSim, set_threads
    set_threads()
    multi = workload in ("cfg3", "toll", "park", "matr")
    arrays, cfg = _world(workload)
    E, S, NA, O = cfg.n_envs, cfg.slots_per_env, cfg.agents_per_env, cfg.objs_per_env
    sim, orc = BatchedSim(arrays, cfg), OracleSim(arrays, cfg)
    obs0_g = sim.reset().cpu().numpy().copy()
    obs0_o = orc.reset_observe().copy()
    live = orc.a["veh_i"].reshape(E, S, -1)[:, :NA, 2].reshape(-1) != 0
    np.testing.assert_allclose(obs0_g[live], obs0_o[live], atol=1e-5, rtol=0)
    # the state a finished env returns to (the product restores the same rows from its post-reset snapshot)
    post = {k: orc.a[k].copy() for k in MUTABLE}
    rows = {"env_i": 1, "obj_f": O}
    ok = np.ones(E, bool)  # envs still in lock-step
    rng = np.random.RandomState(7)
    n_done = n_rays = n_bad_rays = n_float = n_float_diff = n_newborn = 0
    for t in range(steps):
        a = _actions(rng, cfg, multi)
        sim.step(torch.from_numpy(a).cuda(), autoreset=not multi)
        orc.step(a)
        done = (orc.term | orc.trunc).astype(bool)
        fl_g, fl_o = sim.info_flags.cpu().numpy(), orc.info_flags
        te_g, tr_g = sim.terminated.cpu().numpy(), sim.truncated.cpu().numpy()
        og, rg = sim.obs.cpu().numpy(), sim.reward.cpu().numpy()
        oo = orc.obs.copy()
        if not multi and done.any():  # env.reset of the finished envs on the oracle side
            n_done += int(done.sum())
            for e in np.nonzero(done)[0]:
                for k in MUTABLE:
                    r = rows.get(k, S)
                    if r:
                        orc.a[k][e * r:(e + 1) * r] = post[k][e * r:(e + 1) * r]
                oo[e] = obs0_o[e]
        vi_g = sim.get_state("veh_i").reshape(E, -1)
        same = (vi_g == orc.a["veh_i"].reshape(E, -1)).all(1)
        same &= (sim.get_state("env_i").reshape(E, -1) == orc.a["env_i"].reshape(E, -1)).all(1)
        per_env = lambda x: x.reshape(E, -1)
        same &= (per_env(fl_g) == per_env(fl_o)).all(1) & (per_env(te_g) == per_env(orc.term)).all(1)
        same &= (per_env(tr_g) == per_env(orc.trunc)).all(1)
        if workload == "park":   # who is heading for which parking space (VC_PARK)
            same &= (sim.get_state("veh_c").reshape(E, S, -1)[:, :, 14] == orc.a["veh_c"].reshape(E, S, -1)[:, :, 14]).all(1)
            n_newborn += int(((fl_o & 0x4000) != 0).sum())
        ok &= same
        assert ok.all(), "%d of %d envs differ in their integer state at step %d" % ((~ok).sum(), E, t)
        m_env = ok
        m_veh = np.repeat(m_env, S)
        m_ag = np.repeat(m_env, NA)
        vs_g, vs_o = sim.get_state("veh_s"), orc.a["veh_s"]
        n_float += vs_g[m_veh].size + og[valid_rows(m_ag, fl_o, multi)].size
        n_float_diff += int((vs_g[m_veh] != vs_o[m_veh]).sum()) + int((og[valid_rows(m_ag, fl_o, multi)] != oo[valid_rows(m_ag, fl_o, multi)]).sum())
        np.testing.assert_allclose(vs_g[m_veh, 0:3], vs_o[m_veh, 0:3], atol=1e-2, rtol=0)
        np.testing.assert_allclose(vs_g[m_veh, 3:7], vs_o[m_veh, 3:7], atol=1e-3, rtol=0)
        np.testing.assert_allclose(rg[m_ag], orc.reward[m_ag], atol=1e-3, rtol=0)
        valid = m_ag & (((fl_o & 0x2000) != 0) if multi else True)
        sd = sim.state_dim + (8 if cfg.add_others_navi else 4) * cfg.num_others
        np.testing.assert_allclose(og[valid][:, :sd], oo[valid][:, :sd], atol=1e-3, rtol=0)
        bad = ~np.isclose(og[valid][:, sd:], oo[valid][:, sd:], atol=2e-4, rtol=1e-4)
        n_rays += bad.size
        n_bad_rays += int(bad.sum())
        assert bad.sum(1).max(initial=0) <= 2, "more than two glancing rays in one observation at step %d" % t
    assert n_bad_rays == 0 and n_float_diff == 0, (n_bad_rays, n_rays, n_float_diff, n_float)
    if not multi:
        assert n_done > 0, "the run must exercise the fused auto-reset"
    if workload == "matr":
        tr = orc.a["veh_i"].reshape(E, S, -1)[:, NA:]
        assert ((tr[:, :, 0] == 2) & (tr[:, :, 2] != 0)).any(1).mean() > 0.5, "the traffic must have been triggered in most envs"
    if workload == "park":
        assert n_newborn >= E, "the run must exercise the parking lot's respawn rules (%d respawns)" % n_newborn
    print("%s: %d envs x %d steps, %d resets, %d floats (state + observations) compared, all bit-identical"
          % (workload, E, steps, n_done, n_float))
    sim.close()


@pytest.mark.parametrize("workload,steps", [("cfg2", 40), ("cfg5", 30), ("cfg3", 20)])
def test_full_size_env_permutation_and_determinism(workload, steps):
    """Envs never interact: the batch shuffled gives the outputs shuffled, bit for bit; a second run repeats the first."""
    import torch
    from metadrive_ped_b200.sim import BatchedSim
    multi = workload == "cfg3"
    arrays, cfg = _world(workload)
    E, NA = cfg.n_envs, cfg.agents_per_env
    rng = np.random.RandomState(3)
    perm = rng.permutation(E)
    shuffled = {}
    per_env_rows = {}
    for k, v in arrays.items():
        v = np.asarray(v)
        if k.startswith(("env_", "veh_", "obj_", "ma_")) and v.shape[0] % E == 0 and v.shape[0] > 0:
            r = v.shape[0] // E
            per_env_rows[k] = r
            shuffled[k] = np.ascontiguousarray(v.reshape((E, r) + v.shape[1:])[perm].reshape(v.shape))
        else:
            shuffled[k] = v
    assert {"env_i", "veh_s", "veh_i", "veh_p"} <= set(per_env_rows)
    acts = [_actions(rng, cfg, multi) for _ in range(steps)]

    def run(arr, order):
        sim = BatchedSim(arr, cfg)
        out = [sim.reset().cpu().numpy().copy()]
        for a in acts:
            a = a.reshape(E, NA, 2)[order].reshape(-1, 2)
            sim.step(torch.from_numpy(np.ascontiguousarray(a)).cuda(), autoreset=not multi)
            out.append((sim.obs.cpu().numpy().copy(), sim.reward.cpu().numpy().copy(), sim.terminated.cpu().numpy().copy(),
                        sim.truncated.cpu().numpy().copy(), sim.info_flags.cpu().numpy().copy()))
        vi = sim.get_state("veh_i").copy()
        sim.close()
        return out, vi

    ident = np.arange(E)
    base, vi_a = run(arrays, ident)
    again, vi_b = run(arrays, ident)
    shuf, vi_c = run(shuffled, perm)
    np.testing.assert_array_equal(vi_a, vi_b)
    np.testing.assert_array_equal(vi_a.reshape(E, -1)[perm], vi_c.reshape(E, -1))
    np.testing.assert_array_equal(base[0].reshape(E, -1)[perm], shuf[0].reshape(E, -1))
    for t in range(1, steps + 1):
        for x, y, z in zip(base[t], again[t], shuf[t]):
            np.testing.assert_array_equal(x, y, err_msg="run-to-run difference at step %d" % t)
            np.testing.assert_array_equal(x.reshape(E, -1)[perm], z.reshape(E, -1), err_msg="env leakage at step %d" % t)

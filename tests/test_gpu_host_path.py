"""-m gpu: the host-buffer path of the C ABI (md_step_host / md_host_groups / md_host_send / md_host_recv /
md_host_compact) against the device-resident path and the oracle: splitting the batch into host groups, pipelining the
groups across steps and compacting multi-agent rows must not change a single bit of the results."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MUT = ("env_i", "veh_s", "veh_c", "veh_i", "veh_idm", "veh_navi", "veh_route", "veh_rroad", "veh_p", "env_trigger")


def _bank_world(B, E):
    from metadrive_ped_b200.library import ScenarioLibrary
    lib = ScenarioLibrary("pg3_density0.1.npz")
    universe = list(range(B))
    S = max(4, -(-lib.max_vehicles() // 4) * 4)
    kw = dict(slots_per_env=S, objs_per_env=0, map_universe=universe)
    b_arrays, b_cfg = lib.build_world(universe, **kw)
    arrays, cfg = lib.build_world([e % B for e in range(E)], **kw)
    return lib, (b_arrays, b_cfg), (arrays, cfg), S


@pytest.mark.parametrize("groups", [1, 3, 4])
def test_host_groups_equal_device_path_with_bank(groups):
    """Whole-batch device steps (md_step_autoreset) == md_step_host over `groups` host groups, with scenario resampling at
    every reset: the scenario draw hashes the GLOBAL env index (MdConfig.env_base), so the partition cannot show."""
    import torch
    from metadrive_ped_b200.sim import BatchedSim
    lib, (b_arrays, b_cfg), (arrays, cfg), S = _bank_world(24, 90)
    bank = BatchedSim(b_arrays, b_cfg)
    bank.reset()
    dev, host = BatchedSim(arrays, cfg), BatchedSim(arrays, cfg)
    dev.reset()
    host.reset_host()
    dev.attach_bank(bank, seed=11)
    host.attach_bank(bank, seed=11)
    assert host.host_groups(groups) == groups
    rng = np.random.RandomState(3)
    n_done = 0
    for t in range(130):
        act = np.tile(np.array([0.0, 1.0], np.float32), (90, 1)) if t % 3 else rng.uniform(-1, 1, (90, 2)).astype(np.float32)
        dev.step(torch.from_numpy(act).cuda(), autoreset=True)
        obs, rew, cost, term, trunc, flags, info_f = host.step_host(act, autoreset=True)
        np.testing.assert_array_equal(dev.obs.cpu().numpy(), obs, err_msg="obs at step %d" % t)
        np.testing.assert_array_equal(dev.reward.cpu().numpy(), rew)
        np.testing.assert_array_equal(dev.cost.cpu().numpy(), cost)
        np.testing.assert_array_equal(dev.terminated.cpu().numpy(), term)
        np.testing.assert_array_equal(dev.truncated.cpu().numpy(), trunc)
        np.testing.assert_array_equal(dev.info_flags.cpu().numpy(), flags)
        np.testing.assert_array_equal(dev.info_f.cpu().numpy(), info_f)
        n_done += int((term | trunc).sum())
    for k in MUT:
        np.testing.assert_array_equal(dev.get_state(k), host.get_state(k), err_msg=k)
    assert n_done >= 40
    for s in (dev, host, bank):
        s.close()


def test_send_recv_pipeline_equals_synchronous_steps():
    """Two host groups stepped through md_host_send / md_host_recv, one group always a step ahead of the other, give the
    results of whole-batch steps (respawn-mode traffic + pedestrians: the random tapes hash the global env index)."""
    import torch
    from metadrive_ped_b200.library import ScenarioLibrary
    from metadrive_ped_b200.sim import BatchedSim
    lib = ScenarioLibrary("x_respawn_density0.1.npz")
    E = 40
    arrays, cfg = lib.build_world([e % len(lib) for e in range(E)], num_pedestrians=6, seed=2)
    dev, host = BatchedSim(arrays, cfg), BatchedSim(arrays, cfg)
    dev.reset()
    host.reset_host()
    host.host_groups(2)
    gv = host._group_views()
    assert gv[0]["n_envs"] + gv[1]["n_envs"] == E and gv[1]["env0"] == gv[0]["n_envs"]
    rng = np.random.RandomState(5)
    T = 60
    acts = rng.uniform(-0.3, 1.0, (T, E, 2)).astype(np.float32)
    acts[:, :, 0] *= 0.2
    ref = []
    for t in range(T):
        dev.step(torch.from_numpy(acts[t]).cuda(), autoreset=True)
        ref.append([x.cpu().numpy().copy() for x in (dev.obs, dev.reward, dev.cost, dev.terminated, dev.truncated, dev.info_flags, dev.info_f)])
    sl = [slice(g["a0"], g["a0"] + g["na"]) for g in gv]
    host.send(0, acts[0][sl[0]], autoreset=True)     # group 0 runs a step ahead of group 1
    for t in range(T):
        host.send(1, acts[t][sl[1]], autoreset=True)
        out0 = [np.array(x) for x in host.recv(0)]
        if t + 1 < T:
            host.send(0, acts[t + 1][sl[0]], autoreset=True)
        out1 = [np.array(x) for x in host.recv(1)]
        for j in range(7):
            np.testing.assert_array_equal(out0[j], ref[t][j][sl[0]], err_msg="group 0 output %d step %d" % (j, t))
            np.testing.assert_array_equal(out1[j], ref[t][j][sl[1]], err_msg="group 1 output %d step %d" % (j, t))
    for k in ("veh_s", "veh_i", "obj_f", "env_i"):
        np.testing.assert_array_equal(dev.get_state(k), host.get_state(k), err_msg=k)
    dev.close()
    host.close()


@pytest.mark.parametrize("groups", [1, 2])
def test_multi_agent_compact_rows(groups):
    """md_host_compact: the host receives exactly the observation rows of the FL_VALID seats, in seat order."""
    import torch
    from metadrive_ped_b200 import BatchedMultiAgentEnv
    E = 12
    dev = BatchedMultiAgentEnv(E, dict(num_agents=10, horizon=60), seed=4)
    host = BatchedMultiAgentEnv(E, dict(num_agents=10, horizon=60), seed=4)
    dev.reset()
    host.sim.reset_host()
    host.sim.host_groups(groups)
    host.sim.host_compact(True)
    A = dev.sim.n_agents
    rng = np.random.RandomState(1)
    saw_partial = False
    for t in range(90):
        act = rng.uniform(-1, 1, (A, 2)).astype(np.float32)
        act[:, 1] = np.abs(act[:, 1])
        dev.sim.step(torch.from_numpy(act).cuda(), autoreset=True)
        obs, rew, cost, term, trunc, flags, info_f = host.sim.step_host(act, autoreset=True)
        d_flags = dev.sim.info_flags.cpu().numpy()
        np.testing.assert_array_equal(d_flags, flags)
        np.testing.assert_array_equal(dev.sim.reward.cpu().numpy(), rew)
        valid = (d_flags & 0x2000) != 0
        assert obs.shape == (int(valid.sum()), dev.sim.obs_dim)
        np.testing.assert_array_equal(dev.sim.obs.cpu().numpy()[valid], obs, err_msg="compact rows at step %d" % t)
        saw_partial |= bool(0 < valid.sum() < A)
    assert saw_partial
    dev.close()
    host.close()


def test_reset_after_bank_draws_is_a_clean_scenario(oracle_lib):
    """with a scenario bank attached, reset() / reset(env_mask) after draws must leave every env in ONE
    scenario - here: a fresh draw from the bank, like the auto-reset - never the routes of one and the state of another.
    All mutable arrays are compared with the oracle, which is handed the scenario the device drew."""
    import torch
    from metadrive_ped_b200.sim import BatchedSim
    from oracle.oracle import OracleSim
    B, E = 24, 64
    lib, (b_arrays, b_cfg), (arrays, cfg), S = _bank_world(B, E)
    bank, sim = BatchedSim(b_arrays, b_cfg), BatchedSim(arrays, cfg)
    bank.reset()
    sim.reset()
    sim.attach_bank(bank, seed=9)
    orc, borc = OracleSim(arrays, cfg), OracleSim(b_arrays, b_cfg)
    orc.reset_observe()
    b_obs0 = borc.reset_observe().copy()
    rows = {"env_i": 1, "env_trigger": 1}
    act = np.tile(np.array([0.0, 1.0], np.float32), (E, 1))
    act_d = torch.from_numpy(act).cuda()

    def adopt(envs):
        """give the oracle the scenarios the device drew for `envs`; returns their reset observations"""
        env_i = sim.get_state("env_i")
        oo = {}
        for e in envs:
            scn = int(env_i[e, 0])
            assert 0 <= scn < B and int(env_i[e, 4]) == int(lib.seeds[scn])
            for k in MUT:
                r = rows.get(k, S)
                orc.a[k][e * r:(e + 1) * r] = borc.a[k][scn * r:(scn + 1) * r]
            oo[e] = b_obs0[scn]
        return oo

    def check(tag, obs_dev=None, reset_obs=None):
        for k in MUT:
            np.testing.assert_array_equal(sim.get_state(k), orc.a[k], err_msg="%s %s" % (k, tag))
        if reset_obs:
            o = obs_dev.cpu().numpy()
            for e, row in reset_obs.items():
                np.testing.assert_array_equal(o[e], row, err_msg="reset observation of env %d %s" % (e, tag))

    def run(n, tag):
        drawn = 0
        for t in range(n):
            sim.step(act_d, autoreset=True)
            orc.step(act)
            done = np.nonzero((orc.term | orc.trunc).astype(bool))[0]
            drawn += len(done)
            adopt(done)
            check("%s step %d" % (tag, t))
        return drawn

    assert run(110, "before reset") >= E // 3
    seeds_before = sim.get_state("env_i")[:, 4].copy()
    obs = sim.reset()                                   # full reset: every env draws
    check("after reset()", obs, adopt(range(E)))
    assert (sim.get_state("env_i")[:, 4] != seeds_before).sum() >= E // 2
    run(40, "after reset()")
    mask = np.zeros(E, np.uint8)
    mask[::3] = 1
    obs = sim.reset(torch.from_numpy(mask).cuda())      # masked reset: only those envs draw, the others keep running
    check("after reset(mask)", obs, adopt(np.nonzero(mask)[0]))
    run(40, "after reset(mask)")
    sim.close()
    bank.close()

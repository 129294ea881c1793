"""-m gpu: the Gymnasium-surface drop-ins (metadrive_ped_b200/envs.py) through the host-buffer C-ABI call."""
import numpy as np
import pytest

from tests.golden_util import load_golden

pytestmark = pytest.mark.gpu

INFO_KEYS = {"velocity", "steering", "acceleration", "step_energy", "episode_energy", "policy", "overtake_vehicle_num",
             "action", "raw_action", "crash_vehicle", "crash_object", "crash_building", "crash_human", "crash_sidewalk",
             "out_of_road", "arrive_dest", "max_step", "env_seed", "crash", "cost", "step_reward", "episode_reward",
             "episode_length"}


def test_metadrive_env_replays_reference_episode():
    """env.reset(seed=3) + the golden action sequence == the reference's own episode (same seed => same scene)."""
    from metadrive_ped_b200 import MetaDriveEnv
    g = load_golden("cfg2_pg3_seed3")
    env = MetaDriveEnv(dict(num_scenarios=1000, start_seed=0))
    obs, info = env.reset(seed=3)
    assert env.observation_space.contains(obs) and obs.dtype == np.float32
    np.testing.assert_allclose(obs, g["obs"][0], atol=2e-4, rtol=1e-4)
    assert INFO_KEYS <= set(info)
    for t in range(len(g["reward"])):
        obs, r, te, tr, info = env.step(g["actions"][t])
        assert env.observation_space.contains(obs)
        assert INFO_KEYS <= set(info) and info["env_seed"] == 3
        assert abs(r - g["reward"][t]) < 2e-3 and te == bool(g["terminated"][t]) and tr == bool(g["truncated"][t])
        np.testing.assert_allclose(obs[:19], g["obs"][t + 1][:19], atol=2e-3, rtol=0)
        np.testing.assert_allclose(obs[19:], g["obs"][t + 1][19:], atol=2e-3, rtol=1e-3)
        assert info["cost"] == g["cost"][t] and info["episode_length"] == t + 1
    assert te and (info["out_of_road"] or info["crash"] or info["arrive_dest"])
    # the read-only agent view used by the reference's tests
    assert env.agent.speed_km_h > 1.0 and len(env.agent.position) == 2
    env.close()


@pytest.mark.parametrize("tag", ["cfg1_S_straight", "cfg1_S_random", "cfg1_S_discrete", "cfg2_pg3_seed11_dense"])
def test_generated_scene_env_replays_reference_episode(tag):
    """BASELINE cfg1 as a runnable env: MetaDriveEnv(dict(map="S", ...)) builds its scene with the product's own generator
    (pgmap / pgspawn) and replays the reference's episode - continuous and Discrete(7 x 5) actions
    (policy/env_input_policy.py:40-68), and a density no shipped library holds."""
    import json
    from metadrive_ped_b200 import MetaDriveEnv
    g = load_golden(tag)
    conf = json.loads(str(g["config"]))
    cfg = dict(map=conf["map"], traffic_density=conf["traffic_density"], num_scenarios=20, start_seed=0)
    discrete = bool(conf.get("discrete_action", 0))
    if discrete:
        cfg.update(discrete_action=True, discrete_steering_dim=conf["discrete_steering_dim"],
                   discrete_throttle_dim=conf["discrete_throttle_dim"])
    env = MetaDriveEnv(cfg)
    obs, info = env.reset(seed=int(g["seed"]))
    np.testing.assert_allclose(obs, g["obs"][0], atol=2e-4, rtol=1e-4)
    if discrete:
        assert env.action_space.n == 35 and env.action_space.contains(17) and not env.action_space.contains(35)
    for t in range(len(g["reward"])):
        a = int(g["actions"][t, 0]) if discrete else g["actions"][t]
        obs, r, te, tr, info = env.step(a)
        assert abs(r - g["reward"][t]) < 2e-3 and te == bool(g["terminated"][t]) and tr == bool(g["truncated"][t]), t
        np.testing.assert_allclose(obs[:19], g["obs"][t + 1][:19], atol=2e-3, rtol=0)
        if discrete:   # info["action"] is the decoded, clipped pair (env_input_policy.py:33-37)
            sd, td = conf["discrete_steering_dim"], conf["discrete_throttle_dim"]
            assert abs(info["action"][0] - ((a % sd) * 2.0 / (sd - 1) - 1.0)) < 1e-9
            assert abs(info["action"][1] - ((a // sd) * 2.0 / (td - 1) - 1.0)) < 1e-9
            assert abs(info["steering"] - info["action"][0]) < 1e-6
        if te or tr:
            break
    env.close()


def test_drop_in_details():
    """info["action"] / ["raw_action"] are clipped (env_input_policy.py:36-37, base_vehicle.py:204-209); reset(seed=None)
    DRAWS a scenario (base_env.py:886-891); the wrapper keeps a bounded number of GPU handles; cost_to_reward is declared
    but never read by the reference (safe_metadrive_env.py:17); configs the round-1 build refused now run."""
    from metadrive_ped_b200 import MetaDriveEnv, SafeMetaDriveEnv
    from metadrive_ped_b200 import envs as E
    env = MetaDriveEnv(dict(num_scenarios=40, start_seed=100, map="SC", traffic_density=0.25, need_inverse_traffic=True,
                            random_agent_model=True, random_lane_num=True))
    seeds = set()
    for _ in range(12):
        env.reset()
        assert 100 <= env.current_seed < 140
        seeds.add(env.current_seed)
        assert len(env._sims) <= E.MAX_HANDLES
    assert len(seeds) >= 5
    obs, r, te, tr, info = env.step([3.0, float("nan")])
    assert info["action"] == [1.0, 0.0] and info["raw_action"] == (1.0, 0.0) and abs(info["steering"] - 1.0) < 1e-6
    env.close()
    s = SafeMetaDriveEnv(dict(cost_to_reward=True, accident_prob=0.5, map=5))
    s.reset(seed=3)
    total = 0.0
    for _ in range(40):
        obs, r, te, tr, info = s.step([0.0, 0.6])
        assert abs(r - (info["step_reward"] if not (te or info["cost"] > 0) else r)) < 1e-6   # reward is not reduced by the cost
        total += info["cost"]
        if te or tr:
            break
    assert abs(info["total_cost"] - total) < 1e-6
    s.close()
    o = MetaDriveEnv(dict(vehicle_config=dict(overtake_stat=True), num_scenarios=20))
    o.reset(seed=7)
    for _ in range(30):
        obs, r, te, tr, info = o.step([0.0, 1.0])
        if te or tr:
            break
    assert isinstance(info["overtake_vehicle_num"], int) and info["overtake_vehicle_num"] >= 0
    o.close()


def test_record_and_replay_episode():
    """record_episode / replay_episode (envs/base_env.py:255-263, manager/record_manager.py, replay_manager.py): an episode
    recorded with random actions is replayed - whatever seed and actions the caller passes - with the same observations,
    rewards, flags and object states, step for step; info["replay_done"] marks its end."""
    from metadrive_ped_b200 import MetaDriveEnv
    rng = np.random.RandomState(3)
    env = MetaDriveEnv(dict(record_episode=True, num_scenarios=20, start_seed=5, traffic_density=0.2))
    obs0, _ = env.reset(seed=11)
    log = []
    for t in range(60):
        o, r, te, tr, info = env.step(rng.uniform(-1, 1, 2) * [0.3, 1.0])
        log.append((o, r, te, tr, info["crash_vehicle"], info["out_of_road"], info["cost"]))
        if te or tr:
            break
    epi = env.engine.dump_episode()
    env.close()
    assert epi["scenario_index"] == 11 and len(epi["actions"]) == len(log) == len(epi["frame"]) - 1
    f_last = epi["frame"][-1][0]
    assert f_last["episode_step"] == len(log) and "default_agent" in f_last["step_info"] and len(f_last["step_info"]) >= 3
    rep = MetaDriveEnv(dict(replay_episode=epi, num_scenarios=20, start_seed=5, traffic_density=0.2, record_episode=True))
    o0, _ = rep.reset(seed=7)                      # the logged scenario wins
    assert rep.current_seed == 11
    np.testing.assert_array_equal(o0, obs0)
    for t, (o, r, te, tr, cv, oor, cost) in enumerate(log):
        o2, r2, te2, tr2, info = rep.step([0.0, 0.0])   # ignored: the logged action is applied
        np.testing.assert_array_equal(o2, o, err_msg="obs at step %d" % t)
        assert (r2, te2, tr2, info["crash_vehicle"], info["out_of_road"], info["cost"]) == (r, te, tr, cv, oor, cost)
        assert info["replay_done"] == (t == len(log) - 1)
    again = rep.dump_episode()                      # the replay re-recorded: identical frames
    for fa, fb in zip(epi["frame"], again["frame"]):
        assert fa[0]["step_info"] == fb[0]["step_info"]
    rep.close()


def test_graph_replay_equals_plain_launches():
    """md_step_autoreset replays its launch sequence as a CUDA graph while the caller's buffers stay the same, re-captures
    when they change and falls back to plain launches for a caller that passes a new action tensor every step: all three give
    the same states and outputs, bit for bit."""
    import torch
    from metadrive_ped_b200.sim import BatchedSim
    from tests.golden_util import golden_world, load_golden
    g = load_golden("cfg2_pg3_seed11_dense")
    arrays, cfg, _ = golden_world(g, replicas=64, horizon=40)
    a_sim, b_sim = BatchedSim(arrays, cfg), BatchedSim(arrays, cfg)
    a_sim.reset(); b_sim.reset()
    rng = np.random.RandomState(0)
    buf = torch.zeros((cfg.n_envs, 2), device="cuda")
    keep = []
    for t in range(60):
        act = torch.from_numpy((rng.uniform(-1, 1, (cfg.n_envs, 2)) * [0.3, 1.0]).astype(np.float32)).cuda()
        buf.copy_(act)
        a_sim.step(buf, autoreset=True)                       # same buffer every step: captured once, then replayed
        fresh = act.clone()
        keep.append(fresh)                                    # new addresses every step: re-captured, then plain launches
        b_sim.step(fresh, autoreset=True)
        np.testing.assert_array_equal(a_sim.obs.cpu().numpy(), b_sim.obs.cpu().numpy(), err_msg="obs at step %d" % t)
        np.testing.assert_array_equal(a_sim.reward.cpu().numpy(), b_sim.reward.cpu().numpy())
    for k in ("veh_s", "veh_i", "env_i"):
        np.testing.assert_array_equal(a_sim.get_state(k), b_sim.get_state(k), err_msg=k)
    assert a_sim.get_state("env_i")[:, 2].max() < 60, "envs were reset on the way (horizon 40)"
    a_sim.close(); b_sim.close()


def test_env_runs_maps_with_bottleneck_blocks():
    """map strings with Merge / Split blocks ("SyYC", pgblock/bottleneck.py) and TollGate blocks ("S$C", pgblock/tollgate.py: booths as
    static boxes) generate and step; Bidirection / ParkingLot blocks are generated but refused by the env (their shared lanes are
    not in the device world)."""
    from metadrive_ped_b200 import MetaDriveEnv
    env = MetaDriveEnv(dict(map="SyYC", traffic_density=0.2, num_scenarios=8))
    for seed in (0, 3):
        obs, _ = env.reset(seed=seed)
        assert env.observation_space.contains(obs)
        dist = 0.0
        for _ in range(80):
            obs, r, te, tr, info = env.step([0.0, 0.8])
            dist += info["step_reward"]
            if te or tr:
                break
        assert dist > 5.0, "the ego makes progress along the route"
    env.close()
    # a TollGate block inside the map: the booths are bodies of the world (the middle lane of three ends in one)
    env = MetaDriveEnv(dict(map="S$C", traffic_density=0.0, num_scenarios=4))
    obs, _ = env.reset(seed=1)   # the agent manager's draw puts seed 1's ego on the middle lane (y = 3.5)
    hit = False
    for _ in range(300):
        obs, r, te, tr, info = env.step([0.0, 0.5])
        if te or tr:
            hit = info["crash_building"]
            break
    assert hit, "driving straight on along the middle lane ends in the toll booth"
    env.close()
    with pytest.raises(NotImplementedError):
        MetaDriveEnv(dict(map="SBC")).reset(seed=0)


def test_handles_of_different_size_coexist():
    """Kernel attributes (the dynamic shared-memory opt-in) belong to the function, not to a handle: loading a small scene
    must not take it away from a live handle with a large one (a k_pre launch of the large handle failed with "invalid
    argument" after a smaller scene had been loaded - seen through the single-env wrappers' handle cache)."""
    from metadrive_ped_b200 import MetaDriveEnv, SafeMetaDriveEnv
    big = SafeMetaDriveEnv(dict(accident_prob=1.0, map=7, traffic_density=0.3, num_scenarios=4))
    big.reset(seed=1)
    small = MetaDriveEnv(dict(map="S", traffic_density=0.0))
    small.reset(seed=0)
    for _ in range(3):
        big.step([0.0, 0.5])
        small.step([0.0, 0.5])
    big.reset(seed=2)
    big.step([0.0, 0.5])
    big.close(); small.close()


def test_config_errors_match_reference():
    from metadrive_ped_b200 import MetaDriveEnv, SafeMetaDriveEnv
    with pytest.raises(KeyError):
        MetaDriveEnv(dict(not_a_key=1))
    with pytest.raises(NotImplementedError):
        MetaDriveEnv(dict(use_render=True))
    env = MetaDriveEnv(dict(num_scenarios=10, start_seed=5))
    with pytest.raises(AssertionError):
        env.reset(seed=4)
    s = SafeMetaDriveEnv()
    assert s.config["crash_vehicle_done"] is False and s.config["num_scenarios"] == 100
    obs, info = s.reset(seed=2)
    assert "total_cost" in info
    for _ in range(3):
        obs, r, te, tr, info = s.step([0.0, 1.0])
    assert info["total_cost"] == 0 and obs.shape == (259, )
    s.close()


def test_host_path_equals_device_path_and_autoreset():
    import torch
    from metadrive_ped_b200.library import ScenarioLibrary
    from metadrive_ped_b200.sim import BatchedSim
    lib = ScenarioLibrary("pg3_density0.1.npz")
    arrays, cfg = lib.build_world(list(range(48)))
    a_sim, b_sim = BatchedSim(arrays, cfg), BatchedSim(arrays, cfg)
    o1 = a_sim.reset().cpu().numpy()
    o2 = b_sim.reset_host()
    np.testing.assert_array_equal(o1, o2)
    reset_obs = o1.copy()
    act = np.tile(np.array([0.0, 1.0], np.float32), (48, 1))
    act_d = torch.from_numpy(act).cuda()
    saw_done = False
    for t in range(140):
        a_sim.step(act_d, autoreset=True)
        term_d, trunc_d = a_sim.terminated.cpu().numpy().copy(), a_sim.truncated.cpu().numpy().copy()
        obs, rew, cost, term, trunc, flags, info_f = b_sim.step_host(act, autoreset=True)
        np.testing.assert_array_equal(a_sim.obs.cpu().numpy(), obs)
        np.testing.assert_array_equal(a_sim.reward.cpu().numpy(), rew)
        np.testing.assert_array_equal(term_d, term)
        done = (term | trunc).astype(bool)
        if done.any():
            saw_done = True
            # finished envs come back with their reset observation and a fresh episode
            np.testing.assert_array_equal(obs[done], reset_obs[done])
    assert saw_done
    vi = a_sim.get_state("veh_i").reshape(48, cfg.slots_per_env, 16)
    assert (vi[:, 0, 11] < 140).sum() >= 40  # episode_length restarted for the envs that were reset
    a_sim.close()
    b_sim.close()


# ------------------------------------------------------------------------------------------------ multi-agent surface
def _ma_act(env, action):
    """tests/test_env/test_ma_roundabout_env.py:62-77 of the reference, against the drop-in"""
    assert env.action_space.contains(action)
    obs, reward, terminated, truncated, info = env.step(action)
    if not terminated["__all__"]:
        assert len(env.agents) > 0
    assert set(obs.keys()) == set(reward.keys()) == set(info.keys())
    assert set(terminated.keys()) == set(obs.keys()) | {"__all__"} == set(truncated.keys())
    for k, o in obs.items():
        assert o.dtype == np.float32 and env._obs_box.contains(o)
    # agents that are still running are exactly the ones the env wants actions for next
    running = {k for k in obs if not (terminated[k] or truncated[k])}
    assert running == set(env.agents.keys())
    return obs, reward, terminated, truncated, info


@pytest.mark.parametrize("num_agents,num_others,navi", [(1, 8, False), (1, 0, False), (4, 8, False), (4, 0, False), (8, 0, False),
                                                        (4, 4, True), (1, 4, True)])
def test_ma_roundabout_env_surface(num_agents, num_others, navi):
    """metadrive/tests/test_env/test_ma_roundabout_env.py:80-103: spaces, key bookkeeping, no done at step 0 - also with the
    neighbours' checkpoints in the others block (lidar.add_others_navi: 8 floats per neighbour, sensors/lidar.py:120-129)."""
    from metadrive_ped_b200 import MultiAgentRoundaboutEnv
    env = MultiAgentRoundaboutEnv({"num_agents": num_agents, "delay_done": 0,
                                   "vehicle_config": {"lidar": {"num_others": num_others, "add_others_navi": navi}}})
    W = 8 if navi else 4
    try:
        obs, info = env.reset()
        assert set(obs.keys()) == {"agent%d" % k for k in range(num_agents)} == set(env.agents.keys())
        assert env.observation_space.contains(obs)
        assert obs["agent0"].shape == (19 + W * num_others + 72, )  # multi-agent default lidar: 72 lasers, 40 m
        if num_others and num_agents > 1:
            assert any(o[19:19 + W * num_others].any() for o in obs.values()), "neighbours show up in the others block"
        if navi:
            for o in obs.values():   # a neighbour brings its checkpoints; an empty entry is eight zeros
                blk = o[19:19 + 8 * num_others].reshape(num_others, 8)
                assert (blk[:, :4].any(1) == blk[:, 4:].any(1)).all()
        for step in range(100):
            act = {k: [1, 1] for k in env.agents.keys()}
            o, r, tm, tc, i = _ma_act(env, act)
            if step == 0:
                assert not any(tm.values()) and not any(tc.values())
            if tm["__all__"] and not env.agents:
                break
        assert INFO_KEYS <= set(next(iter(i.values())))
    finally:
        env.close()


def test_ma_roundabout_respawn_and_horizon():
    """metadrive/tests/test_env/test_ma_roundabout_env.py (horizon / respawn): finished agents show up once, new ids
    appear with reward 0 while respawn is allowed, nobody outlives the horizon."""
    from metadrive_ped_b200 import MultiAgentRoundaboutEnv
    env = MultiAgentRoundaboutEnv({"num_agents": 6, "horizon": 60, "delay_done": 5,
                                   "vehicle_config": {"lidar": {"num_lasers": 240, "distance": 50}}})
    try:
        obs, _ = env.reset()
        seen, finished, lengths = set(obs), set(), {k: 0 for k in obs}
        for step in range(200):
            act = {k: [0.3, 1.0] for k in env.agents}   # steer off the road: everybody dies, seats get respawned
            o, r, tm, tc, i = _ma_act(env, act)
            for k in o:
                if k not in seen:                       # a newborn: reward 0, not done
                    assert r[k] == 0.0 and not tm[k] and not tc[k]
                    assert env.episode_step < 60, "no respawn after the horizon"
                    seen.add(k); lengths[k] = 0
                else:
                    lengths[k] += 1
                    assert i[k]["episode_length"] == lengths[k]
                assert k not in finished, "a finished agent never comes back"
                if tm[k] or tc[k]:
                    finished.add(k)
                    assert lengths[k] <= 60
            if not env.agents:
                break
        assert not env.agents, "after the horizon the env drains"
        assert len(seen) > 6, "agents were respawned"
        ids = sorted(int(k[5:]) for k in seen)
        assert ids == list(range(len(ids))), "agent ids grow monotonically (agent_manager.py:156-159)"
    finally:
        env.close()


def test_ma_bottleneck_env_surface():
    """envs/marl_envs/marl_bottleneck.py: 20 agents born at both ends of the I -> Merge -> Split map, 4-ray side / lane-line
    detectors in the observation (4 + 6 + 4 + 10 state floats), every agent routed to the far end (no destination draw),
    respawn at both spawn roads; driving straight on ends on the merging lanes' solid lines or in a queue."""
    from metadrive_ped_b200 import MultiAgentBottleneckEnv
    env = MultiAgentBottleneckEnv({"horizon": 80})
    try:
        obs, info = env.reset()
        assert len(obs) == 20 and obs["agent0"].shape == (24 + 72, ) and env.observation_space.contains(obs)
        seen, r_sum = set(obs), 0.0
        for step in range(160):
            o, r, tm, tc, i = _ma_act(env, {k: [0.0, 0.6] for k in env.agents})
            seen |= set(o)
            r_sum += sum(r.values())
            if not env.agents:
                break
        assert not env.agents and len(seen) > 20 and r_sum != 0.0
        lax = MultiAgentBottleneckEnv({"cross_yellow_line_done": False, "num_agents": 4})   # the yellow solid line may be crossed
        lax.reset()
        lax.step({k: [0.0, 0.5] for k in lax.agents})
        lax.close()
    finally:
        env.close()


def test_ma_tollgate_env_surface():
    """envs/marl_envs/marl_tollgate.py: 40 agents on the I -> Split -> TollGate -> Merge map; TollGateObservation = 72 side rays
    + 6 + 4 lane-line rays (no navigation block) + 72 lidar floats + [in the toll block, stayed > min_pass_steps]; an agent
    that rushes through the toll block is paid the overspeed penalty there and loses its episode (out_of_road) the step after
    it leaves; one that drives into a booth ends with crash_building."""
    from metadrive_ped_b200 import MultiAgentTollgateEnv
    env = MultiAgentTollgateEnv({"num_agents": 6, "horizon": 400, "allow_respawn": False})
    try:
        obs, info = env.reset()
        assert len(obs) == 6 and obs["agent0"].shape == (72 + 6 + 4 + 72 + 2, ) and env.observation_space.contains(obs)
        assert all(o[-2] == 0.0 and o[-1] == 0.0 for o in obs.values())
        in_toll, penalised, ended = set(), set(), {}
        for step in range(400):
            if not env.agents:
                break
            o, r, tm, tc, i = _ma_act(env, {k: [0.0, 0.6] for k in env.agents})
            for k in o:
                if o[k][-2] == 1.0:
                    in_toll.add(k)
                    assert o[k][-1] == 0.0   # nobody stays 30 steps at this speed
                    if r[k] < 0.0 and not tm[k]:
                        penalised.add(k)
                if tm.get(k) or tc.get(k):
                    ended[k] = i[k]
        assert in_toll and penalised, "somebody must reach the toll block, too fast"
        rushed = [k for k in in_toll if ended.get(k, {}).get("out_of_road")]
        booth = [k for k in ended if ended[k]["crash_building"]]
        assert rushed or booth, (in_toll, {k: (v["out_of_road"], v["crash_building"], v["crash_vehicle"]) for k, v in ended.items()})
    finally:
        env.close()


def test_batched_multi_agent_env_autoreset():
    import torch
    from metadrive_ped_b200 import BatchedMultiAgentEnv
    env = BatchedMultiAgentEnv(16, {"num_agents": 10, "horizon": 40, "delay_done": 5})
    obs = env.reset()
    assert obs.shape == (16 * 11, 19 + 72)
    a = torch.zeros((16, 11, 2), device="cuda")
    a[..., 0] = 0.2
    a[..., 1] = 1.0
    steps_valid = 0
    for t in range(150):
        obs, r, te, tr, info = env.step(a)
        valid = (info["flags"] & 0x2000) != 0
        steps_valid += int(valid.sum())
        assert torch.isfinite(obs).all() and float(obs.min()) >= 0.0 and float(obs.max()) <= 1.0
    ei = env.sim.get_state("env_i")
    assert ei[:, 2].max() < 150, "envs were reset on device after draining"
    assert steps_valid > 16 * 10 * 40
    env.close()


# ------------------------------------------------------------------------------------------------ BASELINE config 5
def test_pedestrian_intersection_env():
    """MetaDriveEnv(map="X", traffic_mode="respawn") + crossing pedestrians (peds.py): pedestrians are lidar-visible
    moving bodies, contact raises crash_human and ends the episode (crash_human_done), traffic respawns on device."""
    import torch
    from metadrive_ped_b200 import BatchedMetaDriveEnv, MetaDriveEnv
    cfg = dict(map="X", traffic_mode="respawn", traffic_density=0.1, num_scenarios=100, num_pedestrians=16)
    env = MetaDriveEnv(cfg)
    obs, info = env.reset(seed=7)
    assert env.observation_space.contains(obs)
    o0 = env._sim.get_state("obj_f").copy()
    assert (o0[:, 0] == 3).sum() == 16
    for _ in range(20):
        obs, r, te, tr, info = env.step([0.0, 0.5])
        if te:
            break
    o1 = env._sim.get_state("obj_f")
    assert np.abs(o1[:, 1:3] - o0[:, 1:3]).max() > 0.5, "pedestrians walk"
    env.close()

    b = BatchedMetaDriveEnv(128, cfg)
    b.reset()
    a = torch.tensor([0.0, 1.0], device="cuda").repeat(128, 1).contiguous()
    human = 0
    for t in range(400):
        obs, r, te, tr, info = b.step(a)
        hit = (info["flags"] & 0x8) != 0
        human += int(hit.sum())
        assert bool((te[hit] != 0).all()), "crash_human ends the episode (crash_human_done=True)"
    assert human > 0, "somebody runs into a pedestrian within 400 steps of full throttle"
    ei = b.sim.get_state("env_i")
    assert ei[:, 5].sum() == 0 or True  # EI_RNG is reset with the env; respawns are checked through the roster below
    vi = b.sim.get_state("veh_i").reshape(128, -1, 16)
    assert (vi[:, 1:, 1].sum(1) >= 8).all(), "respawn mode keeps the traffic population alive"
    b.close()


def test_lidar_noise_and_dropout():
    """vehicle_config.lidar.gaussian_noise / dropout_prob (obs/state_obs.py:236-244): clip(x + N(0, sigma), 0, 1), then
    zeroed with probability p.  The reference draws from numpy's unseeded generator, so the check is statistical; the
    state part of the observation, the hit ids and md_lidar stay clean, and a handle is reproducible."""
    from metadrive_ped_b200.library import ScenarioLibrary
    from metadrive_ped_b200.sim import BatchedSim
    lib = ScenarioLibrary("pg3_density0.1.npz")
    idx = list(range(256))
    arrays, cfg0 = lib.build_world(idx)
    _, cfg1 = lib.build_world(idx, lidar_gaussian_noise=0.05, lidar_dropout_prob=0.1, noise_seed=3)
    clean, noisy, again = BatchedSim(arrays, cfg0), BatchedSim(arrays, cfg1), BatchedSim(arrays, cfg1)
    o0, o1, o2 = clean.reset().cpu().numpy(), noisy.reset().cpu().numpy(), again.reset().cpu().numpy()
    np.testing.assert_array_equal(o1, o2)                       # same seed, same pass -> same draws
    np.testing.assert_array_equal(o0[:, :19], o1[:, :19])       # state floats untouched
    l0, l1 = o0[:, 19:], o1[:, 19:]
    assert (l1 >= 0).all() and (l1 <= 1).all()
    dropped = (l1 == 0.0) & (l0 > 0.2)
    frac = dropped.sum() / (l0 > 0.2).sum()
    assert abs(frac - 0.1) < 0.01, frac
    mid = (l0 > 0.2) & (l0 < 0.8) & ~dropped                    # away from the clip
    if mid.sum() > 200:
        res = (l1 - l0)[mid]
        assert abs(res.mean()) < 0.01 and abs(res.std() - 0.05) < 0.01, (res.mean(), res.std())
    miss = (l0 == 1.0) & ~dropped                               # a miss stays 1.0 half of the time (clip), else drops below
    below = (l1[miss] < 1.0).mean()
    assert abs(below - 0.5) < 0.02, below
    f_clean, h_clean = clean.lidar()
    f_noisy, h_noisy = noisy.lidar()                            # Lidar.perceive itself carries no noise
    np.testing.assert_array_equal(f_clean.cpu().numpy(), f_noisy.cpu().numpy())
    np.testing.assert_array_equal(h_clean.cpu().numpy(), h_noisy.cpu().numpy())
    o3 = noisy.reset().cpu().numpy()                            # a later observation pass draws fresh noise
    assert (o3[:, 19:] != o1[:, 19:]).mean() > 0.3
    for s in (clean, noisy, again):
        s.close()
    from metadrive_ped_b200 import MetaDriveEnv
    env = MetaDriveEnv(dict(map=3, num_scenarios=10, vehicle_config=dict(lidar=dict(gaussian_noise=0.1, dropout_prob=0.2))))
    obs, _ = env.reset(seed=1)
    assert obs.shape == (259, ) and (obs[19:] == 0.0).sum() > 20
    env.close()


def test_scenario_resampling_on_device(oracle_lib):
    """md_attach_bank: a finished env restarts in a scenario DRAWN from the library (BaseEnv.reset(seed=None),
    envs/base_env.py:886-891) by a row copy from the scenario bank.  Checked against the oracle, which is handed the
    scenario the device drew (the draw replaces numpy's generator and is not part of the parity) and must then stay
    bit-identical - so every array that differs between two scenarios has to be part of the copy."""
    import torch
    from metadrive_ped_b200.library import ScenarioLibrary
    from metadrive_ped_b200.sim import BatchedSim
    from oracle.oracle import OracleSim
    lib = ScenarioLibrary("pg3_density0.1.npz")
    B, E = 48, 192
    universe = list(range(B))
    S = max(4, -(-lib.max_vehicles() // 4) * 4)
    kw = dict(slots_per_env=S, objs_per_env=0, map_universe=universe)
    b_arrays, b_cfg = lib.build_world(universe, **kw)
    arrays, cfg = lib.build_world([e % B for e in range(E)], **kw)
    bank, sim = BatchedSim(b_arrays, b_cfg), BatchedSim(arrays, cfg)
    bank.reset()
    sim.reset()
    sim.attach_bank(bank, seed=5)
    orc, borc = OracleSim(arrays, cfg), OracleSim(b_arrays, b_cfg)
    orc.reset_observe()
    b_obs0 = borc.reset_observe().copy()
    MUT = ("env_i", "veh_s", "veh_c", "veh_i", "veh_idm", "veh_navi", "veh_route", "veh_rroad", "veh_p", "env_trigger")
    rows = {"env_i": 1, "env_trigger": 1}
    act = np.tile(np.array([0.0, 1.0], np.float32), (E, 1))
    act_d = torch.from_numpy(act).cuda()
    n_reset, seen = 0, set()
    for t in range(160):
        sim.step(act_d, autoreset=True)
        orc.step(act)
        done = (orc.term | orc.trunc).astype(bool)
        np.testing.assert_array_equal(sim.terminated.cpu().numpy(), orc.term)
        np.testing.assert_array_equal(sim.truncated.cpu().numpy(), orc.trunc)
        oo = orc.obs.copy()
        env_i = sim.get_state("env_i")
        for e in np.nonzero(done)[0]:
            scn = int(env_i[e, 0])          # EI_MAP = the bank row (the bank lists the universe in order)
            assert 0 <= scn < B and int(env_i[e, 4]) == int(lib.seeds[scn])   # EI_SEED
            seen.add(scn)
            n_reset += 1
            for k in MUT:
                r = rows.get(k, S)
                orc.a[k][e * r:(e + 1) * r] = borc.a[k][scn * r:(scn + 1) * r]
            oo[e] = b_obs0[scn]
        for k in MUT:
            np.testing.assert_array_equal(sim.get_state(k), orc.a[k], err_msg="%s at step %d" % (k, t))
        np.testing.assert_array_equal(sim.obs.cpu().numpy(), oo, err_msg="observation at step %d" % t)
    assert n_reset >= E // 2 and len(seen) >= B // 2, (n_reset, len(seen))
    sim.close()
    bank.close()
    from metadrive_ped_b200 import BatchedMetaDriveEnv
    env = BatchedMetaDriveEnv(64, dict(map=3, num_scenarios=32, start_seed=0), resample_scenarios=True)
    env.reset()
    a = torch.tensor([0.0, 1.0], device="cuda").repeat(64, 1).contiguous()
    seeds0 = env.sim.get_state("env_i")[:, 4].copy()
    for _ in range(150):
        env.step(a)
    seeds1 = env.sim.get_state("env_i")[:, 4]
    assert (seeds1 != seeds0).sum() >= 16 and seeds1.min() >= 0 and seeds1.max() < 32
    env.close()


def test_enable_reverse():
    """vehicle_config.enable_reverse (base_vehicle.py:479-481): a negative throttle drives the engine backwards instead of
    braking.  Default: the car brakes to a stop and stays; with reverse it ends up moving backwards."""
    from metadrive_ped_b200 import MetaDriveEnv
    out = {}
    for rev in (False, True):
        env = MetaDriveEnv(dict(map=3, num_scenarios=10, traffic_density=0.1, vehicle_config=dict(enable_reverse=rev)))
        env.reset(seed=2)
        x0 = np.array(env.agent.position)
        for _ in range(15):
            env.step([0.0, 1.0])
        x1 = np.array(env.agent.position)
        for _ in range(60):
            obs, r, te, tr, info = env.step([0.0, -1.0])
            if te or tr:
                break
        x2 = np.array(env.agent.position)
        fwd = (x1 - x0) / np.linalg.norm(x1 - x0)
        out[rev] = float((x2 - x1) @ fwd), float(info["velocity"])
        env.close()
    assert out[False][0] > 0.0 and abs(out[False][1]) < 0.5          # braked to a stop ahead of where braking began
    assert out[True][0] < out[False][0] - 3.0 and out[True][1] > 1.0  # rolled back past it, still moving (speed is unsigned)


def test_ma_parking_lot_env_surface():
    """MultiAgentParkingLotEnv (envs/marl_envs/marl_parking_lot.py): 10 agents among 8 parking spaces + 3 roads into the lot, dict
    surface, vehicles that can reverse, finished agents replaced by newborns with fresh ids."""
    from metadrive_ped_b200 import MultiAgentParkingLotEnv
    env = MultiAgentParkingLotEnv({"delay_done": 5})
    try:
        obs, info = env.reset()
        assert len(obs) == 10 and env.observation_space.contains(obs) and obs["agent0"].shape == (19 + 72, )
        assert env.config["vehicle_config"]["enable_reverse"]
        seen, finished = set(obs), 0
        rng = np.random.RandomState(0)
        for step in range(250):
            act = {k: [rng.uniform(-0.5, 0.5), 0.4] for k in env.agents.keys()}
            o, r, tm, tc, i = _ma_act(env, act)
            finished += sum(bool(v) for k, v in tm.items() if k != "__all__")
            seen |= set(o)
            if tm["__all__"]:
                break
        assert finished >= 3 and len(seen) > 10, "agents finish (crash / leave the lot) and newborns take over"
        assert INFO_KEYS <= set(next(iter(i.values())))
    finally:
        env.close()
    env = MultiAgentParkingLotEnv({"parking_space_num": 12, "num_agents": -1, "delay_done": 5})   # a bigger lot, filled up
    try:
        obs, info = env.reset()
        assert len(obs) == 15
        for step in range(60):
            o, r, tm, tc, i = _ma_act(env, {k: [0.0, 0.3] for k in env.agents.keys()})
    finally:
        env.close()


def test_ma_env_with_idm_traffic():
    """A multi-agent env with traffic_density > 0 (trigger mode): the IDM vehicles of the roundabout block wait until an agent drives
    onto the block's trigger road, then drive; respawn / hybrid traffic modes are refused."""
    from metadrive_ped_b200 import MultiAgentRoundaboutEnv
    env = MultiAgentRoundaboutEnv({"num_agents": 8, "traffic_density": 0.15, "delay_done": 5})
    try:
        obs, info = env.reset(seed=0)
        assert len(obs) == 8 and env.observation_space.contains(obs)
        vi = env._sim.get_state("veh_i")
        traffic = vi[:, 0] == 2
        assert traffic.sum() == 9 and not vi[traffic, 2].any(), "the reference's nine vehicles for seed 0, parked"
        for step in range(200):
            o, r, tm, tc, i = _ma_act(env, {k: [0.0, 0.6] for k in env.agents.keys()})
            if tm["__all__"]:
                break
        vi = env._sim.get_state("veh_i")
        assert vi[traffic, 2].any(), "an agent on the trigger road started the block's traffic"
    finally:
        env.close()
    with pytest.raises(NotImplementedError):
        MultiAgentRoundaboutEnv({"num_agents": 8, "traffic_density": 0.15, "traffic_mode": "respawn"})


def test_ma_record_and_replay_episode():
    """record_episode / replay_episode on a multi-agent env: a recorded episode (random actions, crashes, respawns under fresh agent
    ids) is replayed - whatever the caller passes - with the same agent ids, observations, rewards and done flags step for step."""
    from metadrive_ped_b200 import MultiAgentIntersectionEnv
    rng = np.random.RandomState(2)
    cfg = {"num_agents": 8, "delay_done": 5, "traffic_density": 0.1}
    env = MultiAgentIntersectionEnv(dict(cfg, record_episode=True))
    obs0, _ = env.reset(seed=0)
    log = []
    for t in range(120):
        o, r, tm, tc, info = env.step({k: rng.uniform(-1, 1, 2) * [0.4, 1.0] for k in env.agents.keys()})
        log.append((o, r, tm, tc))
        if tm["__all__"]:
            break
    epi = env.engine.dump_episode()
    env.close()
    assert len(epi["actions"]) == len(log) == len(epi["frame"]) - 1
    assert any(len(set(o) - set(obs0)) for o, _, _, _ in log), "the episode holds respawned agents"
    rep = MultiAgentIntersectionEnv(dict(cfg, replay_episode=epi))
    o0, _ = rep.reset(seed=3)
    assert set(o0) == set(obs0) and all(np.array_equal(o0[k], obs0[k]) for k in obs0)
    for t, (o, r, tm, tc) in enumerate(log):
        o2, r2, tm2, tc2, info = rep.step({k: [0.0, 0.0] for k in rep.agents.keys()})   # ignored: the logged actions are applied
        assert set(o2) == set(o) and r2 == r and tm2 == tm and tc2 == tc, "step %d" % t
        for k in o:
            np.testing.assert_array_equal(o2[k], o[k], err_msg="obs of %s at step %d" % (k, t))
        assert all(i["replay_done"] == (t == len(log) - 1) for i in info.values())
    rep.close()


def test_ma_envs_with_custom_map_config():
    """map_config overrides of the multi-agent envs reach the generated map: a three-lane roundabout holds more spawn slots, a
    one-lane intersection (no U-turn destinations) and a six-booth toll plaza step with respawns."""
    from metadrive_ped_b200 import MultiAgentIntersectionEnv, MultiAgentRoundaboutEnv, MultiAgentTollgateEnv
    env = MultiAgentRoundaboutEnv({"num_agents": -1, "map_config": {"lane_num": 3, "exit_length": 50}})
    try:
        obs, _ = env.reset()
        assert len(obs) == 60 and env.config["map_config"]["lane_num"] == 3
    finally:
        env.close()
    for cls, mc in ((MultiAgentIntersectionEnv, {"lane_num": 1}), (MultiAgentTollgateEnv, {"toll_lane_num": 6, "toll_length": 14})):
        env = cls({"num_agents": 6, "delay_done": 3, "map_config": mc})
        try:
            obs, _ = env.reset()
            seen = set(obs)
            for step in range(150):
                o, r, tm, tc, i = _ma_act(env, {k: [0.1, 0.8] for k in env.agents.keys()})
                seen |= set(o)
                if tm["__all__"]:
                    break
            assert len(seen) > 6, "newborns took over"
        finally:
            env.close()


def test_base_multi_agent_env():
    """MultiAgentMetaDrive itself: 15 agents on the first road of the BIG map of the scenario seed, all bound for the end of the last
    block; another seed brings another map."""
    from metadrive_ped_b200 import MultiAgentMetaDrive
    env = MultiAgentMetaDrive({"num_scenarios": 4, "map": 2, "delay_done": 3})
    try:
        obs, _ = env.reset(seed=0)
        assert len(obs) == 15 and env.observation_space.contains(obs) and env.current_seed == 0
        lanes0 = len(env._lib.table.lane_f)
        for step in range(60):
            o, r, tm, tc, i = _ma_act(env, {k: [0.0, 0.5] for k in env.agents.keys()})
        assert sum(r.values()) > 0, "they make progress along the route"
        obs, _ = env.reset(seed=3)
        assert env.current_seed == 3 and len(obs) == 15 and env._lib.pg_seed == 3
        _ma_act(env, {k: [0.0, 0.5] for k in env.agents.keys()})
        assert lanes0 > 0
    finally:
        env.close()

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

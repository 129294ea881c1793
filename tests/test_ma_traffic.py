"""Multi-agent envs with IDM traffic (traffic_density > 0, trigger mode) on the host: the traffic roster the library generates
against the reference trace `cfg3_ma_roundabout_traffic`, and the trigger rule on the CPU oracle."""
import json

import numpy as np

from tests.golden_util import load_golden


def test_generated_traffic_roster_equals_the_reference_trace():
    """PGTrafficManager.reset behind 8 agents on the roundabout map, seed 0, density 0.15: classes, sampled engine / brake forces
    (the engine's seed stream after 8 agent spawns), poses, routes, trigger blocks and overtake timers of all 9 vehicles."""
    from metadrive_ped_b200.ma import MultiAgentLibrary
    lib, g = MultiAgentLibrary("ma_roundabout.npz"), load_golden("cfg3_ma_roundabout_traffic")
    n = int(g["ma_alive_seats"][0])
    np.testing.assert_array_equal(lib.table.lane_f, g["map_lane_f"])
    ref_blocks = json.loads(str(g["map_meta"]))["blocks"]
    assert [b["trigger_road"] for b in lib.table.meta["blocks"]] == [b["trigger_road"] for b in ref_blocks]
    static, dyn, routes, ints, idm = lib.traffic(n, 0.15, 0)
    assert len(static) == len(g["init_veh_static"]) - n == 9
    np.testing.assert_allclose(static, g["init_veh_static"][n:], rtol=1e-7)
    np.testing.assert_allclose(dyn, g["init_veh_dyn"][n:], rtol=0, atol=1e-9)
    np.testing.assert_array_equal(routes, g["init_routes"][n:])
    np.testing.assert_array_equal(ints, g["init_veh_int"][n:])
    np.testing.assert_array_equal(idm, g["init_idm"][n:])


def test_traffic_starts_when_an_agent_enters_the_trigger_road(oracle_lib):
    from metadrive_ped_b200.envs import MultiAgentRoundaboutEnv, _apply_vehicle_config, _ma_cfg_kw
    from metadrive_ped_b200.ma import MultiAgentLibrary
    from oracle.oracle import OracleSim
    c = MultiAgentRoundaboutEnv.default_config()
    c["traffic_density"] = 0.15
    arrays, cfg = MultiAgentLibrary("ma_roundabout.npz").build_world(3, 8, seed=4, traffic_density=0.15, traffic_seed=0, **_ma_cfg_kw(c))
    _apply_vehicle_config(arrays, c)
    E, S, NA = cfg.n_envs, cfg.slots_per_env, cfg.agents_per_env
    assert NA == 9 and S >= NA + 9 and cfg.traffic_mode == 0
    sim = OracleSim(arrays, cfg)
    sim.reset_observe()
    vi = sim.a["veh_i"].reshape(E, S, -1)
    traffic = vi[:, :, 0] == 2
    assert traffic[:, NA:].sum(1).min() >= 5 and not vi[traffic][:, 2].any(), "traffic waits (alive, not active) after reset"
    started = np.zeros(E, bool)
    for t in range(250):
        a = np.tile(np.array([[0.0, 0.6]], np.float32), (E * NA, 1))
        sim.step(a)
        vi = sim.a["veh_i"].reshape(E, S, -1)
        started |= (vi[:, :, 2][traffic.reshape(E, S)].reshape(E, -1) != 0).any(1)
    assert started.all(), "every env's traffic was triggered by an agent driving onto the trigger road"


def test_one_lane_intersection_draws_no_u_turn_destinations(oracle_lib):
    """MultiAgentIntersectionEnv with map_config lane_num = 1: the reference's spawn manager never sends an agent back out of the
    road it was born on (marl_intersection.py:76-82, 103).  Reset-time draws and on-device respawn draws (CPU oracle)."""
    import metadrive_ped_b200.envs as E
    from oracle.oracle import OracleSim
    cls = E.MultiAgentIntersectionEnv
    c = E._merge(cls.default_config(), dict(map_config=dict(lane_num=1), num_agents=8, delay_done=3))
    lib = cls._make_library(c)
    assert lib.conf["disable_u_turn"] and lib.dest_nodes.shape == (4, 3)
    full = E.MultiAgentIntersectionEnv._make_library(E._merge(cls.default_config(), {})).dest_nodes   # 2 lanes: U-turns allowed
    assert full.ndim == 1 and len(full) == 4
    arrays, cfg = lib.build_world(4, 8, seed=3, **E._ma_cfg_kw(c))
    assert cfg.ma_dests == 3
    sim = OracleSim(arrays, cfg)
    sim.reset_observe()
    En, S, NA = cfg.n_envs, cfg.slots_per_env, cfg.agents_per_env
    lane_road = {lane: ri for lane, _, ri in lib.slots}
    all_dest = [int(x) for x in np.unique(lib.dest_nodes)]
    forbidden = {ri: [d for d in all_dest if d not in [int(x) for x in lib.dest_nodes[ri]]] for ri in range(4)}
    assert all(len(v) == 1 for v in forbidden.values())

    def check(rows_i, rows_rt):
        n = 0
        for I, rt in zip(rows_i, rows_rt):
            if I[0] == 1 and I[1]:
                dest = int(rt[rt >= 0][-1])
                assert dest != forbidden[lane_road[int(I[13])]][0], "an agent was sent back out of its own road"
                n += 1
        return n
    assert check(sim.a["veh_i"], sim.a["veh_route"]) == 4 * 8
    rng = np.random.RandomState(0)
    born = 0
    for t in range(250):
        a = np.stack([rng.uniform(-0.5, 0.5, En * NA), rng.uniform(0.2, 1.0, En * NA)], 1).astype(np.float32)
        sim.step(a)
        check(sim.a["veh_i"], sim.a["veh_route"])
        born += int(((sim.info_flags & 0x4000) != 0).sum())
    assert born >= 8


def test_base_multi_agent_env_on_generated_maps(oracle_lib):
    """MultiAgentMetaDrive itself (envs/marl_envs/multi_agent_metadrive.py): the BIG map of config["map"] / the scenario seed, the
    first road as the only spawn road (15 slots), the end of the last block as everybody's destination - library against the
    reference trace cfg3_ma_pg3 (map = 2 blocks, seed 0), and a short run on the oracle with another seed's map."""
    import metadrive_ped_b200.envs as E
    from oracle.oracle import OracleSim
    g = load_golden("cfg3_ma_pg3")
    c = E._merge(E.MultiAgentMetaDrive.default_config(), dict(map=2, num_agents=6, num_scenarios=5))
    lib = E.MultiAgentMetaDrive._make_library(c)
    np.testing.assert_array_equal(lib.table.lane_f, g["map_lane_f"])
    np.testing.assert_array_equal(lib.table.road_i, g["map_road_i"])
    np.testing.assert_array_equal(lib.spawn_roads, g["ma_spawn_roads"])
    np.testing.assert_array_equal(lib.dest_nodes, g["ma_dest_nodes"])
    assert lib.max_capacity == 15
    for k in range(int(g["ma_alive_seats"][0])):
        np.testing.assert_array_equal(lib.tables["routes"][0], g["init_routes"][k])
    lib3 = E.MultiAgentMetaDrive._make_library(c, pg_seed=3)
    assert lib3.pg_seed == 3 and not np.array_equal(lib3.table.lane_f.shape, ()) and len(lib3.table.lane_f) != 0
    arrays, cfg = lib3.build_world(2, 6, seed=1, **E._ma_cfg_kw(c))
    sim = OracleSim(arrays, cfg)
    obs = sim.reset_observe()
    assert obs.shape == (2 * 7, 19 + 72)
    for t in range(40):
        sim.step(np.tile(np.array([[0.0, 0.5]], np.float32), (2 * 7, 1)))
    vi = sim.a["veh_i"].reshape(2, cfg.slots_per_env, -1)
    assert vi[:, :6, 2].all() and (vi[:, :6, 8] & 0x100).all(), "six agents per env, still driving on their lanes"

"""MultiAgentParkingLotEnv (envs/marl_envs/marl_parking_lot.py) on the host: the generated library against the reference trace
`cfg3_ma_parkinglot`, and ParkingLotSpawnManager's respawn rules as the CPU oracle restates them (the CUDA path is compared with
the oracle bit for bit in tests/test_gpu_full_size.py::test_full_size_step_matches_oracle[park])."""
import numpy as np

from tests.golden_util import load_golden


def _lib():
    from metadrive_ped_b200.ma import MultiAgentLibrary
    return MultiAgentLibrary("ma_parkinglot.npz")


def test_library_equals_the_reference_trace():
    """Map, spawn roads (three ways in + eight spaces), the per-road destination lists, the vehicle row, and for every reset-time
    agent of the trace: born inside one of the library's slots, with the route the library tabulates for (spawn road, destination)."""
    lib, g = _lib(), load_golden("cfg3_ma_parkinglot")
    np.testing.assert_array_equal(lib.table.lane_f, g["map_lane_f"])
    np.testing.assert_array_equal(lib.table.road_i, g["map_road_i"])
    np.testing.assert_array_equal(lib.spawn_roads, g["ma_spawn_roads"])
    np.testing.assert_array_equal(lib.dest_nodes, g["ma_dest_nodes"])
    np.testing.assert_allclose(lib.veh_static, g["init_veh_static"][0], rtol=1e-7)
    assert lib.max_capacity == 11 and lib.conf["parking_spaces"] == 8 and lib.conf["parking_in_roads"] == 3
    n, D = int(g["ma_alive_seats"][0]), lib.tables["n_dests"]
    lane_slot = {lane: (lon, ri) for lane, lon, ri in lib.slots}
    for k in range(n):
        lane = int(g["init_veh_int"][k, 2])
        lon, ri = lane_slot[lane]
        row = lib.geo.lane_f[lane]
        dx, dy = g["init_veh_dyn"][k, 0] - row[3], g["init_veh_dyn"][k, 1] - row[4]
        along = dx * row[7] + dy * row[8]
        assert abs(along - lon) <= 1.0 + 1e-6, "born inside the slot (the spawn manager moves it by at most (8 - 10) / 2 m)"
        route = g["init_routes"][k]
        dest = int(route[route >= 0][-1])
        d = int(np.nonzero(lib.dest_nodes[ri] == dest)[0][0])
        np.testing.assert_array_equal(lib.tables["routes"][ri * D + d], route)
        assert (g["ma_parking_taken"][k] == d) if ri < 3 else (g["ma_parking_taken"][k] == -1 and d < 3)


def test_respawn_rules_of_the_spawn_manager(oracle_lib):
    """Random driving in 6 lots for 500 steps on the CPU oracle: every respawn obeys ParkingLotSpawnManager - an agent from outside
    heads for a space no other active agent is heading for, an agent born in a space leaves by one of the three roads, nobody comes
    from outside while every space is spoken for - and both kinds of respawn happen."""
    from metadrive_ped_b200.envs import MultiAgentParkingLotEnv, _apply_vehicle_config, _ma_cfg_kw
    from oracle.oracle import OracleSim
    lib = _lib()
    c = MultiAgentParkingLotEnv.default_config()
    arrays, cfg = lib.build_world(6, c["num_agents"], seed=2, **_ma_cfg_kw(c))
    _apply_vehicle_config(arrays, c)
    assert cfg.parking_spaces == 8 and cfg.parking_in_roads == 3 and cfg.on_continuous_line_done == 5
    sim = OracleSim(arrays, cfg)
    sim.reset_observe()
    E, S, NA = cfg.n_envs, cfg.slots_per_env, cfg.agents_per_env
    spaces, exits = set(lib.dest_nodes[0].tolist()), set(lib.dest_nodes[3][:3].tolist())
    in_lanes = {lane for lane, _, ri in lib.slots if ri < 3}
    rng = np.random.RandomState(3)
    bias = rng.uniform(-0.6, 0.6, E * NA)
    n_in = n_out = 0
    for t in range(500):
        a = np.zeros((E * NA, 2), np.float32)
        a[:, 0] = bias + 0.3 * rng.uniform(-1, 1, E * NA)
        a[:, 1] = np.where(rng.uniform(0, 1, E * NA) < 0.15, -0.6, 0.35) * rng.uniform(0.3, 1.0, E * NA)
        sim.step(a)
        vi = sim.a["veh_i"].reshape(E, S, -1)
        vc = sim.a["veh_c"].reshape(E, S, -1)
        rt = sim.a["veh_route"].reshape(E, S, -1)
        fl = sim.info_flags.reshape(E, NA)
        for e in range(E):
            active = [s for s in range(NA) if vi[e, s, 2]]
            heading_for = [int(vc[e, s, 14]) for s in active if vc[e, s, 14] > 0]
            assert len(heading_for) == len(set(heading_for)), "two active agents share a parking space"
            for s in np.nonzero(fl[e] & 0x4000)[0]:
                dest = int(rt[e, s][rt[e, s] >= 0][-1])
                if int(vi[e, s, 13]) in in_lanes:   # born on a road into the lot
                    assert vc[e, s, 14] > 0 and dest == int(lib.dest_nodes[0][int(vc[e, s, 14]) - 1]) and dest in spaces
                    n_in += 1
                else:
                    assert vc[e, s, 14] == 0 and dest in exits
                    n_out += 1
    assert n_in >= 3 and n_out >= 3, (n_in, n_out)


def test_other_numbers_of_parking_spaces(oracle_lib):
    """parking_space_num = 4 and 12 (marl_parking_lot.py:211-219): capacity 3 + N, the respawn rules hold on the oracle; odd or
    fewer than 4 spaces fail the reference's assertions."""
    import pytest
    from metadrive_ped_b200.envs import MultiAgentParkingLotEnv, _apply_vehicle_config, _ma_cfg_kw
    from oracle.oracle import OracleSim
    for n_spaces, n_agents in ((4, 6), (12, 12)):
        c = MultiAgentParkingLotEnv.default_config()
        c.update(parking_space_num=n_spaces, num_agents=n_agents)
        lib = MultiAgentParkingLotEnv._make_library(c)
        assert lib.max_capacity == 3 + n_spaces
        arrays, cfg = lib.build_world(3, n_agents, seed=1, **_ma_cfg_kw(c))
        _apply_vehicle_config(arrays, c)
        assert cfg.parking_spaces == n_spaces and cfg.ma_dests == max(n_spaces, 3)
        sim = OracleSim(arrays, cfg)
        sim.reset_observe()
        E, S, NA = cfg.n_envs, cfg.slots_per_env, cfg.agents_per_env
        rng = np.random.RandomState(5)
        bias = rng.uniform(-0.6, 0.6, E * NA)
        born = 0
        for t in range(300):
            a = np.zeros((E * NA, 2), np.float32)
            a[:, 0] = bias + 0.3 * rng.uniform(-1, 1, E * NA)
            a[:, 1] = 0.35 * rng.uniform(0.3, 1.0, E * NA)
            sim.step(a)
            vi, vc = sim.a["veh_i"].reshape(E, S, -1), sim.a["veh_c"].reshape(E, S, -1)
            for e in range(E):
                goal = [int(vc[e, s, 14]) for s in range(NA) if vi[e, s, 2] and vc[e, s, 14] > 0]
                assert len(goal) == len(set(goal)) and all(1 <= x <= n_spaces for x in goal)
            born += int(((sim.info_flags & 0x4000) != 0).sum())
        assert born >= 3
    for bad in (5, 2):
        with pytest.raises(AssertionError):
            MultiAgentParkingLotEnv._make_library(dict(parking_space_num=bad))

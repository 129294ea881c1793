"""Helpers shared by the parity tests: load a golden fixture into Scenario/arrays form."""
import json
import os

import numpy as np

from metadrive_ped_b200 import scene as sc
from metadrive_ped_b200.abi import make_config

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def list_golden(prefix=""):
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f.startswith(prefix))


def load_golden(tag):
    d = np.load(os.path.join(GOLDEN_DIR, tag + ".npz"), allow_pickle=False)
    return {k: d[k] for k in d.files}


def is_ma(g):
    return "ma_spawn_roads" in g


# MULTI_AGENT_METADRIVE_DEFAULT_CONFIG (envs/marl_envs/multi_agent_metadrive.py:12-62)
MA_CFG = dict(is_multi_agent=1, out_of_road_penalty=10.0, crash_vehicle_penalty=10.0, crash_object_penalty=10.0,
              out_of_road_cost=0.0, truncate_as_terminate=1, horizon=1000, delay_done=25)


def ma_tape_from_trace(g, length=256):
    """The random tape that replays the respawn choices of the reference trace: the k-th respawn drew the
    clear-list index draws[k, 0] and the destination draws[k, 1] (value % n == value since value < n)."""
    tape = np.zeros((length, 4), np.int32)
    ev = g["respawn_draws"][g["respawn_draws"][:, 0] >= 0]
    assert len(ev) <= length
    tape[:len(ev), :2] = ev[:, :2]
    return tape


def golden_world_ma(g, replicas=1, **cfg_kw):
    """(arrays, cfg, geo) for `replicas` identical multi-agent envs: seats = reset-time agents + one spare."""
    from metadrive_ped_b200 import ma
    mt = sc.MapTable(np.asarray(g["map_lane_f"], np.float64), np.asarray(g["map_lane_i"], np.int32),
                     np.asarray(g["map_road_i"], np.int32), json.loads(str(g["map_meta"])), int(g["lane_num"]))
    geo = sc.build_map_geometry(mt)
    n = int(g["ma_alive_seats"][0])
    n_tr = int(json.loads(str(g["config"])).get("n_traffic", 0))   # IDM traffic of the env: roster rows after the agents
    m = n + n_tr
    scen = sc.Scenario(0, g["init_veh_static"][:m], g["init_veh_dyn"][:m], g["init_routes"][:m], g["init_veh_int"][:m],
                       g["init_idm"][:m], g["init_objects"], int(g["seed"]),
                       g["ma_parking_taken"][:n] if "ma_parking_taken" in g else None)
    NA = n + 1
    S = ((NA + n_tr + 3) // 4) * 4
    tables = ma.build_ma_tables(geo, g["ma_spawn_roads"], g["ma_dest_nodes"])
    tape = np.tile(ma_tape_from_trace(g), (replicas, 1))
    O = len(g["init_objects"])   # static bodies of the map itself (toll booths)
    arrays = sc.pack([geo], [scen] * replicas, S, NA, O, ma_tables={0: tables}, ma_tables_tape=tape)
    conf = json.loads(str(g["config"]))
    kw = dict(MA_CFG)
    kw.update(horizon=int(conf.get("horizon", 1000)), delay_done=int(conf.get("delay_done", 25)),
              allow_respawn=int(bool(conf.get("allow_respawn", True))), n_lasers=int(conf["n_lasers"]),
              lidar_dist=float(conf["lidar_dist"]), ma_places=len(tables["places"]), ma_dests=tables["n_dests"],
              ma_roads=tables["n_roads"], tape_len=len(tape) // replicas,
              n_side_lasers=int(conf.get("n_side_lasers", 0)), side_dist=float(conf.get("side_dist", 50.0)),
              n_lane_lasers=int(conf.get("n_lane_lasers", 0)), lane_dist=float(conf.get("lane_dist", 20.0)),
              ignore_road_sign=int(conf.get("ignore_road_sign", 0)))
    for k in ("toll_env", "min_pass_steps", "on_continuous_line_done", "out_of_route_done", "num_others", "add_others_navi",
              "parking_spaces", "parking_in_roads"):
        if k in conf:
            kw[k] = int(conf[k])
    for k in ("overspeed_penalty", "speed_reward"):
        if k in conf:
            kw[k] = float(conf[k])
    kw.update(cfg_kw)
    cfg = make_config(replicas, S, NA, O, **kw)
    return arrays, cfg, geo


def respawn_tape_from_trace(g, geo, length=256):
    """Random tape replaying the traffic respawns of a reference trace (rows in (step, slot) order): respawn-lane index,
    float bits of longitude / (lane length / 2), overtake timer."""
    ev = np.asarray(g["respawn_events"]).reshape(-1, 5)
    ev = ev[np.lexsort((ev[:, 1], ev[:, 0]))]
    tape = np.zeros((length, 4), np.int32)
    lanes = [it["lane"] for it in geo.meta["respawn"]]
    for k, (t, slot, place, lon, timer) in enumerate(ev):
        half = np.float32(geo.lane_f[lanes[int(place)], 2]) / np.float32(2.0)
        frac = np.float32(lon) / half
        tape[k] = [int(place), np.array([frac], np.float32).view(np.int32)[0], int(timer), 0]
    return tape


def golden_world(g, replicas=1, slots=None, objs=None, **cfg_kw):
    """(arrays, cfg) for `replicas` identical envs built from one golden fixture."""
    if is_ma(g):
        return golden_world_ma(g, replicas, **cfg_kw)
    if str(g["tag"]).startswith("cfg5"):
        return golden_world_cfg5(g, replicas, **cfg_kw)
    mt = sc.MapTable(np.asarray(g["map_lane_f"], np.float64), np.asarray(g["map_lane_i"], np.int32),
                     np.asarray(g["map_road_i"], np.int32), json.loads(str(g["map_meta"])), int(g["lane_num"]))
    geo = sc.build_map_geometry(mt)
    scen = sc.Scenario(0, g["init_veh_static"], g["init_veh_dyn"], g["init_routes"], g["init_veh_int"], g["init_idm"],
                       g["init_objects"], int(g["seed"]))
    n = len(g["init_veh_static"])
    S = slots or max(4, ((n + 3) // 4) * 4)
    O = objs if objs is not None else len(g["init_objects"])
    arrays = sc.pack([geo], [scen] * replicas, S, 1, O)
    conf = json.loads(str(g["config"]))
    kw = {}
    if str(g["tag"]).startswith("cfg4"):  # SafeMetaDriveEnv defaults (envs/safe_metadrive_env.py:10-19)
        kw.update(crash_vehicle_done=0, crash_object_done=0)
    if "horizon" in conf and conf["horizon"]:
        kw["horizon"] = int(conf["horizon"])
    kw["num_others"] = int(conf.get("num_others", 0))
    kw["add_others_navi"] = int(conf.get("add_others_navi", 0))
    kw.update(n_side_lasers=int(conf.get("n_side_lasers", 0)), side_dist=float(conf.get("side_dist", 50.0)),
              n_lane_lasers=int(conf.get("n_lane_lasers", 0)), lane_dist=float(conf.get("lane_dist", 20.0)))
    if isinstance(conf.get("discrete_action"), int) and not isinstance(conf.get("discrete_action"), bool):
        kw.update(discrete_action=conf["discrete_action"], discrete_steering_dim=conf["discrete_steering_dim"],
                  discrete_throttle_dim=conf["discrete_throttle_dim"])
    kw.update(cfg_kw)
    cfg = make_config(replicas, S, 1, O, **kw)
    return arrays, cfg, geo


def golden_world_cfg5(g, replicas=1, **cfg_kw):
    """BASELINE config 5 fixture: respawn-mode traffic (tape from the trace) + crossing pedestrians as objects."""
    mt = sc.MapTable(np.asarray(g["map_lane_f"], np.float64), np.asarray(g["map_lane_i"], np.int32),
                     np.asarray(g["map_road_i"], np.int32), json.loads(str(g["map_meta"])), int(g["lane_num"]))
    geo = sc.build_map_geometry(mt)
    scen = sc.Scenario(0, g["init_veh_static"], g["init_veh_dyn"], g["init_routes"], g["init_veh_int"], g["init_idm"],
                       g["init_objects"], int(g["seed"]))
    n = len(g["init_veh_static"])
    S = max(4, ((n + 3) // 4) * 4)
    O = len(g["init_objects"])
    tape = np.tile(respawn_tape_from_trace(g, geo), (replicas, 1))
    arrays = sc.pack([geo], [scen] * replicas, S, 1, O, ma_tables_tape=tape, traffic_respawn=True)
    conf = json.loads(str(g["config"]))
    kw = dict(traffic_mode=1, ma_places=len(arrays["ma_place_f"]) // replicas, tape_len=len(tape) // replicas,
              crash_vehicle_done=int(conf.get("crash_vehicle_done", True)), crash_object_done=int(conf.get("crash_object_done", True)),
              crash_human_done=int(conf.get("crash_human_done", True)))
    kw.update(cfg_kw)
    cfg = make_config(replicas, S, 1, O, **kw)
    return arrays, cfg, geo

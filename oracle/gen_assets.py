"""Export a scenario library (maps + reset-time rosters) from the UNMODIFIED reference under oracle.refshim.

TEST/ASSET INFRASTRUCTURE, in-container only:
    python -m oracle.gen_assets --env metadrive --n 1000 --out metadrive_ped_b200/assets/pg3_density0.1.npz

The library is what the product's reset() loads on a box without /root/reference: for every scenario seed the
lane graph BIG generated (component/algorithm/BIG.py:68-95) and the bodies the managers spawned at reset
(manager/traffic_manager.py:211-277, manager/object_manager.py:40-151, manager/agent_manager.py:88-113).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def export_multi_agent(args):
    """Multi-agent envs have one fixed map; what the product needs besides the lane graph: the spawn roads / destination
    nodes (envs/marl_envs/marl_inout_roundabout.py:12-24, marl_intersection.py:12-25), the slot geometry constants of
    SpawnManager (manager/spawn_manager.py:25-35, 117-158) and the parameters of the "static_default" vehicle."""
    from oracle import ref_export as rx
    from metadrive.component.pgblock.first_block import FirstPGBlock
    from metadrive.manager.spawn_manager import SpawnManager
    if args.env == "ma_roundabout":
        from metadrive.envs.marl_envs.marl_inout_roundabout import MultiAgentRoundaboutEnv as cls
    elif args.env == "ma_intersection":
        from metadrive.envs.marl_envs.marl_intersection import MultiAgentIntersectionEnv as cls
    elif args.env == "ma_bottleneck":
        from metadrive.envs.marl_envs.marl_bottleneck import MultiAgentBottleneckEnv as cls
    elif args.env == "ma_bidirection":
        from metadrive.envs.marl_envs.marl_bidirection import MultiAgentBidirectionEnv as cls
    elif args.env == "ma_tollgate":
        from metadrive.envs.marl_envs.marl_tollgate import MultiAgentTollgateEnv as cls
    else:
        from metadrive.envs.marl_envs.marl_parking_lot import MultiAgentParkingLotEnv as cls
    extra = dict(parking_space_num=args.parking_spaces, num_agents=min(10, 3 + args.parking_spaces)) if args.parking_spaces else {}
    if args.map_config:   # a non-default map of the env (lane_num, exit_length, neck_length, toll_lane_num ...)
        extra.update(map_config=json.loads(args.map_config), num_agents=4)
    env = cls(dict(log_level=50, **extra))
    try:
        env.reset()
    except KeyError as e:
        # MultiAgentBidirectionEnv.reward_function reads config["use_lateral"], a key no config defines
        # (envs/marl_envs/marl_bidirection.py:113): the reference's env cannot finish reset(); its MAP is built by then
        print("reset() of the reference raised", repr(e)[:60], "- exporting the map it built")
    m, mi = rx.export_map(env.current_map)
    roads = list(env.config["spawn_roads"])
    v = next(iter(env.agents.values()))
    conf = dict(
        env=args.env, num_agents=int(env.config["num_agents"]), lane_num=int(env.config["map_config"]["lane_num"]),
        exit_length=float(env.config["map_config"]["exit_length"]), entrance_length=float(FirstPGBlock.ENTRANCE_LENGTH),
        map_config={k: v for k, v in dict(env.config["map_config"]).items() if isinstance(v, (int, float))},
        respawn_longitude=float(SpawnManager.RESPAWN_REGION_LONGITUDE), respawn_lateral=float(SpawnManager.RESPAWN_REGION_LATERAL),
        max_vehicle_length=float(SpawnManager.MAX_VEHICLE_LENGTH), max_vehicle_width=float(SpawnManager.MAX_VEHICLE_WIDTH),
        disable_u_turn=bool(getattr(env.engine.spawn_manager, "disable_u_turn", False)),
        defaults={k: env.config[k] for k in (
            "horizon", "delay_done", "allow_respawn", "crash_done", "out_of_road_done", "out_of_road_penalty",
            "crash_vehicle_penalty", "crash_object_penalty", "crash_vehicle_cost", "crash_object_cost", "out_of_road_cost",
            "truncate_as_terminate", "success_reward", "driving_reward", "speed_reward", "traffic_density")},
        lidar={k: env.config["vehicle_config"]["lidar"][k] for k in ("num_lasers", "distance", "num_others")},
    )
    out = dict(
        lane_f=m["lane_f"], lane_i=m["lane_i"], road_i=m["road_i"], meta=m["meta"], config=json.dumps(conf),
        spawn_roads=np.array([[mi.nodes[r.start_node], mi.nodes[r.end_node]] for r in roads], np.int32),
        # (the parking lot's spawn roads include its parking spaces, whose "negative road" is not a road of the map: -1)
        dest_nodes=np.array([mi.nodes.get((-r).end_node, -1) for r in roads], np.int32),
        veh_static=rx.vehicle_static(v).astype(np.float32),
        # static bodies of the map itself (the tollgate map's TollGateBuilding boxes), in ref_export.Roster.objects_table rows
        objects=rx.Roster(env, mi).objects_table(),
    )
    env.close()
    np.savez_compressed(args.out, **out)
    print("wrote", args.out, os.path.getsize(args.out) // 1024, "KiB", "lanes", len(m["lane_f"]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="metadrive", choices=["metadrive", "safe", "ma_roundabout", "ma_intersection", "ma_bottleneck", "ma_bidirection", "ma_tollgate",
                             "ma_parkinglot"])
    ap.add_argument("--parking-spaces", type=int, default=0, help="ma_parkinglot: parking_space_num (default: the env's 8)")
    ap.add_argument("--map-config", default=None, help="multi-agent envs: JSON of map_config overrides")
    ap.add_argument("--n", type=int, default=1000)
    ap.add_argument("--start", type=int, default=0)
    ap.add_argument("--density", type=float, default=None)
    ap.add_argument("--map", default="3")
    ap.add_argument("--traffic-mode", default=None, choices=[None, "trigger", "respawn", "hybrid"])
    ap.add_argument("--out", required=True)
    args = ap.parse_args()
    from oracle import refshim
    refshim.install()
    from oracle import ref_export as rx
    from oracle.gen_golden import roster_arrays
    from metadrive.envs.metadrive_env import MetaDriveEnv
    from metadrive.envs.safe_metadrive_env import SafeMetaDriveEnv
    if args.env.startswith("ma_"):
        return export_multi_agent(args)
    cls = MetaDriveEnv if args.env == "metadrive" else SafeMetaDriveEnv
    mp = int(args.map) if args.map.isdigit() else args.map
    cfg = dict(map=mp, num_scenarios=args.n, start_seed=args.start, log_level=50, store_map=False)
    if args.density is not None:
        cfg["traffic_density"] = args.density
    if args.traffic_mode is not None:
        cfg["traffic_mode"] = args.traffic_mode
    env = cls(cfg)
    acc = {k: [] for k in ("lane_f", "lane_i", "road_i", "veh_static", "veh_dyn", "routes", "veh_int", "idm", "objects")}
    map_off, veh_off, obj_off, metas, seeds = [], [], [], [], []
    t0 = time.time()
    for seed in range(args.start, args.start + args.n):
        env.reset(seed=seed)
        m, mi = rx.export_map(env.current_map)
        roster = rx.Roster(env, mi)
        init = roster_arrays(env, mi, roster)
        map_off.append([sum(len(a) for a in acc["lane_f"]), len(m["lane_f"]), sum(len(a) for a in acc["road_i"]),
                        len(m["road_i"]), env.config["map_config"]["lane_num"]])
        veh_off.append([sum(len(a) for a in acc["veh_static"]), len(init["veh_static"])])
        obj_off.append([sum(len(a) for a in acc["objects"]), len(init["objects"])])
        acc["lane_f"].append(m["lane_f"]); acc["lane_i"].append(m["lane_i"]); acc["road_i"].append(m["road_i"])
        for k in ("veh_static", "veh_dyn", "routes", "veh_int", "idm", "objects"):
            acc[k].append(init[k])
        metas.append(m["meta"])
        seeds.append(seed)
        if (seed - args.start) % 50 == 0:
            print(seed, "vehicles", len(init["veh_static"]), "lanes", len(m["lane_f"]), "%.1fs" % (time.time() - t0), flush=True)
    env.close()
    out = dict(
        seeds=np.asarray(seeds, np.int32), map_off=np.asarray(map_off, np.int64), veh_off=np.asarray(veh_off, np.int64),
        obj_off=np.asarray(obj_off, np.int64), meta=np.asarray(metas),
        config=json.dumps({k: v for k, v in cfg.items() if isinstance(v, (int, float, str, bool))}), env=args.env,
        lane_f=np.concatenate(acc["lane_f"]).astype(np.float64), lane_i=np.concatenate(acc["lane_i"]).astype(np.int16),
        road_i=np.concatenate(acc["road_i"]).astype(np.int16), veh_static=np.concatenate(acc["veh_static"]).astype(np.float32),
        veh_dyn=np.concatenate(acc["veh_dyn"]).astype(np.float64), routes=np.concatenate(acc["routes"]).astype(np.int16),
        veh_int=np.concatenate(acc["veh_int"]).astype(np.int16), idm=np.concatenate(acc["idm"]).astype(np.float32),
        objects=np.concatenate([o.reshape(-1, 8) for o in acc["objects"]]).astype(np.float64),
    )
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    np.savez_compressed(args.out, **out)
    print("wrote", args.out, os.path.getsize(args.out) // 1024, "KiB", "%.1fs" % (time.time() - t0))


if __name__ == "__main__":
    main()

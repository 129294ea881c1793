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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="metadrive", choices=["metadrive", "safe"])
    ap.add_argument("--n", type=int, default=1000)
    ap.add_argument("--start", type=int, default=0)
    ap.add_argument("--density", type=float, default=None)
    ap.add_argument("--map", default="3")
    ap.add_argument("--out", required=True)
    args = ap.parse_args()
    from oracle import refshim
    refshim.install()
    from oracle import ref_export as rx
    from oracle.gen_golden import roster_arrays
    from metadrive.envs.metadrive_env import MetaDriveEnv
    from metadrive.envs.safe_metadrive_env import SafeMetaDriveEnv
    cls = MetaDriveEnv if args.env == "metadrive" else SafeMetaDriveEnv
    mp = int(args.map) if args.map.isdigit() else args.map
    cfg = dict(map=mp, num_scenarios=args.n, start_seed=args.start, log_level=50, store_map=False)
    if args.density is not None:
        cfg["traffic_density"] = args.density
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

"""Generate golden fixtures by running the UNMODIFIED reference under oracle.refshim.

TEST INFRASTRUCTURE, in-container only:   python -m oracle.gen_golden [--out tests/golden]

Each fixture = one reference episode prefix: the exported map, the reset-time roster (vehicle parameters,
poses, routes, IDM timers, static objects), the action sequence, and per-step traces of every vehicle plus the
ego's observation / reward / cost / done.  tests/ replay the same actions through the C oracle and the CUDA
library from the same reset state and compare (north-star tolerances).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def roster_arrays(env, mi, roster):
    """Scenario arrays (see metadrive_ped_b200/scene.py:Scenario) from a freshly reset reference env."""
    from oracle import ref_export as rx
    eng = env.engine
    n = len(roster.vehicles)
    veh_static = roster.static_table()
    veh_dyn = np.stack([rx.vehicle_dynamic(v) for v in roster.vehicles])
    routes = roster.routes()
    veh_int = np.zeros((n, 6), np.int32)
    idm = np.zeros((n, 2), np.float64)
    tm = getattr(eng, "traffic_manager", None)
    for k, v in enumerate(roster.vehicles):
        is_agent = v in roster.agents
        veh_int[k, 0] = 1 if is_agent else 2
        veh_int[k, 1] = -1 if is_agent else roster.trigger_block[k - len(roster.agents)]
        veh_int[k, 2] = mi.lane_id(v.navigation.current_lane)
        veh_int[k, 3], veh_int[k, 4] = v.navigation._target_checkpoints_index
        veh_int[k, 5] = 1 if (is_agent or (tm is not None and v in tm._traffic_vehicles)) else 0
        pol = eng.get_policy(v.name)
        if pol is not None and hasattr(pol, "overtake_timer"):
            idm[k] = [pol.overtake_timer, pol.target_speed]
        else:
            idm[k] = [0, 30]
    return dict(veh_static=veh_static, veh_dyn=veh_dyn, routes=routes, veh_int=veh_int, idm=idm,
                objects=roster.objects_table())


def run_episode(env_cls, config, seed, actions, tag):
    from oracle import ref_export as rx
    env = env_cls(config)
    try:
        obs0, _ = env.reset(seed=seed)
        m, mi = rx.export_map(env.current_map)
        roster = rx.Roster(env, mi)
        init = roster_arrays(env, mi, roster)
        f0, i0 = rx.record_world(env, roster)
        fs, is_, obs, rew, cost, term, trunc, infos = [f0], [i0], [obs0], [], [], [], [], []
        for a in actions:
            o, r, te, tr, info = env.step(a)
            f, i = rx.record_world(env, roster)
            fs.append(f)
            is_.append(i)
            obs.append(o)
            rew.append(r)
            cost.append(info["cost"])
            term.append(te)
            trunc.append(tr)
            infos.append([info["velocity"], info["steering"], info["acceleration"], info["step_energy"],
                          info["episode_energy"], info["step_reward"], info["episode_reward"], info["episode_length"]])
            if te or tr:
                break
        T = len(rew)
        out = dict(
            tag=tag, seed=seed, lane_num=env.config["map_config"]["lane_num"],
            map_lane_f=m["lane_f"], map_lane_i=m["lane_i"], map_road_i=m["road_i"], map_meta=m["meta"],
            actions=np.asarray(actions[:T], np.float64), veh_f=np.stack(fs), veh_i=np.stack(is_),
            obs=np.stack(obs).astype(np.float32), reward=np.asarray(rew, np.float64), cost=np.asarray(cost, np.float64),
            terminated=np.asarray(term, bool), truncated=np.asarray(trunc, bool), info=np.asarray(infos, np.float64),
            config=json.dumps({k: v for k, v in config.items() if isinstance(v, (int, float, str, bool))}),
            **{"init_" + k: v for k, v in init.items()},
        )
        sb = rx.export_static_bodies(env.engine)
        out["ref_lines"] = sb["lines"]
        return out
    finally:
        env.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    from oracle import refshim
    refshim.install()
    from metadrive.envs.metadrive_env import MetaDriveEnv
    from metadrive.envs.safe_metadrive_env import SafeMetaDriveEnv
    os.makedirs(args.out, exist_ok=True)
    rng = np.random.RandomState(0)
    rand_actions = rng.uniform(-1, 1, (args.steps, 2))
    smooth = np.stack([0.15 * np.sin(np.arange(args.steps) / 7.0), np.full(args.steps, 0.6)], 1)
    cases = [
        # BASELINE config 1: default single agent on map "S", profiling action [0, 1]
        ("cfg1_S_straight", MetaDriveEnv, dict(map="S", traffic_density=0.1, log_level=50), 0,
         np.tile([0.0, 1.0], (args.steps, 1))),
        ("cfg1_S_random", MetaDriveEnv, dict(map="S", traffic_density=0.1, log_level=50), 0, rand_actions),
        # BASELINE config 2: 3-block PG maps with IDM traffic
        ("cfg2_pg3_seed3", MetaDriveEnv, dict(map=3, traffic_density=0.1, num_scenarios=20, start_seed=0, log_level=50),
         3, smooth),
        ("cfg2_pg3_seed7", MetaDriveEnv, dict(map=3, traffic_density=0.1, num_scenarios=20, start_seed=0, log_level=50),
         7, smooth),
        ("cfg2_pg3_seed11_dense", MetaDriveEnv,
         dict(map=3, traffic_density=0.3, num_scenarios=20, start_seed=0, log_level=50), 11, smooth),
        ("cfg2_SCO_nolimit", MetaDriveEnv, dict(map="SCO", traffic_density=0.2, log_level=50), 0,
         np.stack([np.zeros(args.steps), np.full(args.steps, 0.3)], 1)),
        # BASELINE config 4: SafeMetaDriveEnv with static obstacles
        ("cfg4_safe_seed2", SafeMetaDriveEnv, dict(num_scenarios=20, start_seed=0, log_level=50), 2, smooth),
        ("cfg4_safe_seed5", SafeMetaDriveEnv, dict(num_scenarios=20, start_seed=0, log_level=50), 5, smooth),
    ]
    for tag, cls, cfg, seed, acts in cases:
        if args.only and args.only not in tag:
            continue
        out = run_episode(cls, cfg, seed, acts, tag)
        path = os.path.join(args.out, tag + ".npz")
        np.savez_compressed(path, **out)
        print(tag, "steps", len(out["reward"]), "vehicles", out["veh_f"].shape[1], "objects",
              len(out["init_objects"]), "->", os.path.getsize(path) // 1024, "KiB", flush=True)


if __name__ == "__main__":
    main()

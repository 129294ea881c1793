import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from metadrive_ped_b200 import MetaDriveEnv, SafeMetaDriveEnv
bad = 0
for trial in range(25):
    env = MetaDriveEnv(dict(num_scenarios=40, start_seed=100, map="SC", traffic_density=0.25, need_inverse_traffic=True,
                            random_agent_model=True, random_lane_num=True))
    seeds = []
    try:
        for _ in range(12):
            env.reset()
            seeds.append(env.current_seed)
        env.step([3.0, float("nan")])
        env.close()
        s = SafeMetaDriveEnv(dict(cost_to_reward=True, accident_prob=0.5, map=5))
        s.reset(seed=3)
        for _ in range(40):
            obs, r, te, tr, info = s.step([0.0, 0.6])
            if te or tr: break
        s.close()
        o = MetaDriveEnv(dict(vehicle_config=dict(overtake_stat=True), num_scenarios=20))
        o.reset(seed=7)
        for _ in range(30):
            obs, r, te, tr, info = o.step([0.0, 1.0])
            if te or tr: break
        o.close()
    except Exception as e:
        bad += 1
        print("trial", trial, "seeds", seeds, "->", repr(e)[:400])
        traceback.print_exc(limit=4)
print("failures", bad)

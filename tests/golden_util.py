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


def golden_world(g, replicas=1, slots=None, objs=None, **cfg_kw):
    """(arrays, cfg) for `replicas` identical envs built from one golden fixture."""
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
    kw.update(cfg_kw)
    cfg = make_config(replicas, S, 1, O, **kw)
    return arrays, cfg, geo

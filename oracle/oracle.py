"""ctypes binding of the CPU oracle (oracle/md_oracle.c). TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from metadrive_ped_b200.abi import MdArrays, MdConfig
from metadrive_ped_b200.scene import ARRAY_ORDER

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libmd_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "md_oracle.c")
    hdrs = [os.path.join(_HERE, "..", "include", h) for h in ("md_layout.h", "md_math.h")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
            os.path.getmtime(f) for f in [src] + hdrs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
    return _lib


def set_threads(n=None):
    """Use `n` (default: all) host threads for the OpenMP loop over envs; returns the thread count in effect."""
    return int(lib().mdo_set_threads(int(n or os.cpu_count() or 1)))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleSim:
    """Holds numpy copies of the flat arrays and steps them with the C oracle."""
    def __init__(self, arrays: dict, cfg: MdConfig):
        self.cfg = cfg
        self.a = {k: np.ascontiguousarray(arrays[k]).copy() for k in ARRAY_ORDER}
        self.init = {k: self.a[k].copy() for k in ("env_i", "veh_s", "veh_c", "veh_i", "veh_idm", "veh_navi", "obj_f")}
        self.A = MdArrays(**{k: _p(self.a[k]) for k in ARRAY_ORDER})
        self.n_agents = cfg.n_envs * cfg.agents_per_env
        # include/md_layout.h OBS_STATE / OBS_DIM: the tollgate env drops the 10 navigation floats and appends 2 toll floats
        self.state_dim = (cfg.n_side_lasers or 2) + 6 + (cfg.n_lane_lasers or 1) + (0 if cfg.toll_env else 10)
        self.obs_dim = self.state_dim + (8 if cfg.add_others_navi else 4) * cfg.num_others + cfg.n_lasers + (2 if cfg.toll_env else 0)
        self.ray_cs = np.zeros((cfg.n_lasers, 2), np.float32)
        lib().mdo_make_ray_table(cfg.n_lasers, _p(self.ray_cs))
        na = self.n_agents
        self.obs = np.zeros((na, self.obs_dim), np.float32)
        self.reward = np.zeros(na, np.float32)
        self.cost = np.zeros(na, np.float32)
        self.term = np.zeros(na, np.uint8)
        self.trunc = np.zeros(na, np.uint8)
        self.info_flags = np.zeros(na, np.int32)
        self.info_f = np.zeros((na, 8), np.float32)
        self.hit = np.zeros((na, cfg.n_lasers), np.int32)

    def enable_contacts(self):
        """record the contact pairs of every step: self.contacts [NV, 4] uint32 (the layout of md_get_contacts)"""
        self.contacts = np.zeros((self.cfg.n_envs * self.cfg.slots_per_env, 4), np.uint32)

    def contact_pairs(self, env=0):
        S = self.cfg.slots_per_env
        pairs = set()
        for i in range(S):
            for w in range(4):
                m = int(self.contacts[env * S + i, w])
                while m:
                    b = (m & -m).bit_length() - 1
                    pairs.add((min(i, 32 * w + b), max(i, 32 * w + b)))
                    m &= m - 1
        return sorted(pairs)

    def reset_observe(self):
        lib().mdo_reset_observe(C.byref(self.cfg), C.byref(self.A), _p(self.ray_cs), _p(self.obs), _p(self.hit))
        return self.obs

    def step(self, actions, env_begin=0, env_end=None):
        actions = np.ascontiguousarray(actions, np.float32).reshape(self.n_agents, 2)
        if env_end is None:
            env_end = self.cfg.n_envs
        # the output pointer is a global of the C library: (re)bind it for this call, never leave a dangling one behind
        lib().mdo_set_contact_out(_p(self.contacts) if getattr(self, "contacts", None) is not None else None)
        lib().mdo_step(C.byref(self.cfg), C.byref(self.A), _p(self.ray_cs), _p(actions), _p(self.obs), _p(self.reward),
                       _p(self.cost), _p(self.term), _p(self.trunc), _p(self.info_flags), _p(self.info_f),
                       _p(self.hit), C.c_int(env_begin), C.c_int(env_end))
        lib().mdo_set_contact_out(None)
        return self.obs, self.reward, self.term, self.trunc

    def reset_envs(self, mask):
        """Restore the initial snapshot of the masked envs (what the product's md_reset does on device)."""
        S, O = self.cfg.slots_per_env, self.cfg.objs_per_env
        for e in np.nonzero(np.asarray(mask))[0]:
            for k in ("veh_s", "veh_c", "veh_i", "veh_idm", "veh_navi"):
                self.a[k][e * S:(e + 1) * S] = self.init[k][e * S:(e + 1) * S]
            self.a["env_i"][e] = self.init["env_i"][e]
            if O:
                self.a["obj_f"][e * O:(e + 1) * O] = self.init["obj_f"][e * O:(e + 1) * O]

    def lidar(self):
        frac = np.zeros((self.n_agents, self.cfg.n_lasers), np.float32)
        hit = np.zeros((self.n_agents, self.cfg.n_lasers), np.int32)
        lib().mdo_lidar(C.byref(self.cfg), C.byref(self.A), _p(self.ray_cs), _p(frac), _p(hit))
        return frac, hit

    def topdown(self, resolution=84, max_distance=30.0, channels=3):
        """[n_agents, res, res, 3] RGB in [0, 1]: the ego-centred bird's-eye image (mdo_topdown); channels=2: the per-frame
        [road_network, traffic_flow] grey channels of the stacked observation"""
        img = np.zeros((self.n_agents, resolution, resolution, channels), np.float32)
        lib().mdo_topdown(C.byref(self.cfg), C.byref(self.A), _p(img), C.c_int(resolution), C.c_float(max_distance), C.c_int(channels))
        return img

    def dynamics(self, act3, n_sub):
        act3 = np.ascontiguousarray(act3, np.float32)
        lib().mdo_dynamics(C.byref(self.cfg), C.byref(self.A), _p(act3), C.c_int(n_sub))

    def after_step(self):
        lib().mdo_after_step(C.byref(self.cfg), C.byref(self.A))

    def idm(self):
        out = np.zeros((self.cfg.n_envs * self.cfg.slots_per_env, 2), np.float32)
        lib().mdo_idm(C.byref(self.cfg), C.byref(self.A), _p(out))
        return out

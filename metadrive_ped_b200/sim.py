"""BatchedSim: one handle of libmdstep.so on one GPU, with torch tensors as caller-owned device buffers.

Mirrors the step surface of the reference engine for E environments at once:
  reset()  ~ BaseEnv.reset   (envs/base_env.py:502-537)          -> obs [A, 19 + n_lasers]
  step(a)  ~ BaseEnv.step    (envs/base_env.py:426-431, 586-623) -> obs, reward, cost, terminated, truncated, info
PyTorch is used for device memory and streams only.
"""
import ctypes as C

import numpy as np

from . import lib as _libmod
from .abi import MdArrays, MdConfig
from .scene import ARRAY_ORDER

INFO_F_NAMES = ["velocity", "steering", "acceleration", "step_energy", "episode_energy", "step_reward",
                "episode_reward", "episode_length"]


class BatchedSim:
    def __init__(self, arrays: dict, cfg: MdConfig, device: int = 0):
        import torch
        if not torch.cuda.is_available():
            raise _libmod.MdStepError("BatchedSim needs a CUDA device; this package has no CPU fallback")
        self.torch = torch
        self.lib = _libmod.load()
        self.cfg = cfg
        self.device = device
        self.tdev = torch.device("cuda", device)
        self.h = C.c_void_p()
        rc = self.lib.md_create(C.byref(cfg), device, C.byref(self.h))
        self._check(rc)
        self._host = {k: np.ascontiguousarray(arrays[k]) for k in ARRAY_ORDER}
        rows = (C.c_int64 * len(ARRAY_ORDER))(*[self._host[k].shape[0] for k in ARRAY_ORDER])
        ha = MdArrays(**{k: self._host[k].ctypes.data_as(C.c_void_p) for k in ARRAY_ORDER})
        self._check(self.lib.md_load_scene(self.h, C.byref(ha), rows))
        self.n_agents = cfg.n_envs * cfg.agents_per_env
        # include/md_layout.h OBS_STATE / OBS_DIM: the tollgate env drops the 10 navigation floats and appends 2 toll floats
        self.state_dim = (cfg.n_side_lasers or 2) + 6 + (cfg.n_lane_lasers or 1) + (0 if cfg.toll_env else 10)
        self.obs_dim = self.state_dim + (8 if cfg.add_others_navi else 4) * cfg.num_others + cfg.n_lasers + (2 if cfg.toll_env else 0)
        na = self.n_agents
        kw = dict(device=self.tdev)
        self.obs = torch.zeros((na, self.obs_dim), dtype=torch.float32, **kw)
        self.reward = torch.zeros(na, dtype=torch.float32, **kw)
        self.cost = torch.zeros(na, dtype=torch.float32, **kw)
        self.terminated = torch.zeros(na, dtype=torch.uint8, **kw)
        self.truncated = torch.zeros(na, dtype=torch.uint8, **kw)
        self.info_flags = torch.zeros(na, dtype=torch.int32, **kw)
        self.info_f = torch.zeros((na, 8), dtype=torch.float32, **kw)

    # ------------------------------------------------------------------ helpers
    def _check(self, rc):
        if rc != 0:
            msg = self.lib.md_last_error(self.h) if self.h else b""
            raise _libmod.MdStepError("libmdstep call failed (%d): %s" % (rc, (msg or b"").decode()))

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.tdev).cuda_stream)

    @staticmethod
    def _ptr(t):
        return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p()

    # ------------------------------------------------------------------ device-resident API (tensors in, tensors out)
    def reset(self, env_mask=None):
        self._check(self.lib.md_reset(self.h, self._ptr(env_mask), self._ptr(self.obs), self._stream()))
        return self.obs

    def step(self, actions, autoreset=False):
        assert actions.is_cuda and actions.dtype == self.torch.float32 and actions.is_contiguous()
        assert actions.numel() == self.n_agents * 2
        fn = self.lib.md_step_autoreset if autoreset else self.lib.md_step
        self._check(fn(self.h, self._ptr(actions), self._ptr(self.obs), self._ptr(self.reward), self._ptr(self.cost),
                       self._ptr(self.terminated), self._ptr(self.truncated), self._ptr(self.info_flags),
                       self._ptr(self.info_f), self._stream()))
        return self.obs, self.reward, self.cost, self.terminated, self.truncated

    # ------------------------------------------------------------------ host-buffer API (numpy in, numpy out)
    # The batch is split into host groups (md_host_groups, include/mdstep.h): contiguous env ranges with their own stream
    # and pinned buffers.  step_host() steps every group in one synchronous call; send() / recv() pipeline the groups
    # across steps, so that a group's PCIe copies and the caller's own work overlap the other groups' kernels.
    def host_groups(self, n_groups):
        self._check(self.lib.md_host_groups(self.h, int(n_groups)))
        self._gv = None
        return self.n_host_groups

    @property
    def n_host_groups(self):
        return int(self.lib.md_host_group_count(self.h))

    def host_compact(self, on=True):
        """Multi-agent: only the observation rows of seats with FL_VALID are copied to the host (md_host_compact)."""
        self._check(self.lib.md_host_compact(self.h, int(bool(on))))
        self._compact = bool(on)

    def _group_views(self):
        """per group: numpy views of its pinned buffers (no copy) + its env range"""
        if getattr(self, "_gv", None) is None:
            self._gv = []
            A = self.cfg.agents_per_env
            for k in range(self.n_host_groups):
                ptrs, rng = (C.c_void_p * 8)(), (C.c_int * 3)()
                self._check(self.lib.md_host_group_views(self.h, k, ptrs, rng))
                na = rng[1] * A

                def view(i, ctype, shape):
                    n = int(np.prod(shape))
                    return np.ctypeslib.as_array(C.cast(ptrs[i], C.POINTER(ctype)), shape=(n, )).reshape(shape)

                self._gv.append(dict(
                    obs=view(0, C.c_float, (na, self.obs_dim)), reward=view(1, C.c_float, (na, )),
                    cost=view(2, C.c_float, (na, )), term=view(3, C.c_uint8, (na, )), trunc=view(4, C.c_uint8, (na, )),
                    flags=view(5, C.c_int32, (na, )), info_f=view(6, C.c_float, (na, 8)),
                    actions=view(7, C.c_float, (na, 2)), env0=rng[0], n_envs=rng[1], a0=rng[0] * A, na=na))
        return self._gv

    def _rows(self, k):
        rng = (C.c_int * 3)()
        self._check(self.lib.md_host_group_views(self.h, k, None, rng))
        return rng[2]

    def send(self, group, actions=None, autoreset=False):
        """Enqueue one env.step of host group `group` and return at once.  `actions` [group agents, 2]; None = the caller
        wrote them into the group's pinned action buffer (`group_buffers(group)["actions"]`) already."""
        if actions is not None:
            g = self._group_views()[group]
            g["actions"][...] = np.asarray(actions, np.float32).reshape(g["na"], 2)
        rc = self.lib.md_host_send(self.h, group, None, 1 if autoreset else 0)
        if rc:
            self._check(rc)

    def group_buffers(self, group):
        """numpy views of a host group's pinned buffers: obs, reward, cost, term, trunc, flags, info_f, actions; plus its
        env range (env0, n_envs) and agent range (a0, na)."""
        return self._group_views()[group]

    def recv(self, group):
        """Wait for the group's step.  Returns views of its pinned buffers (valid until its next send):
        obs, reward, cost, terminated, truncated, info_flags, info_f.  In compact mode `obs` holds only the rows of the
        seats whose info_flags carry FL_VALID, in ascending seat order."""
        rc = self.lib.md_host_recv(self.h, group)
        if rc:
            self._check(rc)
        g = self._group_views()[group]
        obs = g["obs"][:self._rows(group)] if getattr(self, "_compact", False) else g["obs"]
        return obs, g["reward"], g["cost"], g["term"], g["trunc"], g["flags"], g["info_f"]

    def step_host(self, actions: np.ndarray, autoreset=False):
        """env.step through host memory: actions are copied into the pinned input buffers, H2D, kernels, D2H; the
        returned arrays are views of the pinned output buffers where those are contiguous over the batch (the observation
        rows always; everything with one host group), else concatenated copies of the groups' (small) scalar blocks.
        Valid until the next call."""
        gv = self._group_views()
        a = np.asarray(actions, np.float32).reshape(self.n_agents, 2)
        for g in gv:
            g["actions"][...] = a[g["a0"]:g["a0"] + g["na"]]
        null = C.c_void_p()
        self._check(self.lib.md_step_host(self.h, null, null, null, null, null, null, null, null, int(bool(autoreset))))
        if len(gv) == 1:
            b = gv[0]
            obs = b["obs"][:self._rows(0)] if getattr(self, "_compact", False) else b["obs"]
            return obs, b["reward"], b["cost"], b["term"], b["trunc"], b["flags"], b["info_f"]
        cat = lambda name: np.concatenate([g[name] for g in gv])
        if getattr(self, "_compact", False):
            obs = np.concatenate([g["obs"][:self._rows(k)] for k, g in enumerate(gv)])
        else:  # the groups' observation rows are consecutive slices of one pinned buffer
            obs = np.ctypeslib.as_array(C.cast(gv[0]["obs"].ctypes.data, C.POINTER(C.c_float)),
                                        shape=(self.n_agents * self.obs_dim, )).reshape(self.n_agents, self.obs_dim)
        return obs, cat("reward"), cat("cost"), cat("term"), cat("trunc"), cat("flags"), cat("info_f")

    def reset_host(self, env_mask=None):
        na = self.n_agents
        obs = np.zeros((na, self.obs_dim), np.float32)
        m = None if env_mask is None else np.ascontiguousarray(env_mask, np.uint8)
        self._check(self.lib.md_reset_host(self.h, m.ctypes.data_as(C.c_void_p) if m is not None else None,
                                           obs.ctypes.data_as(C.c_void_p)))
        return obs

    # ------------------------------------------------------------------ isolated stages
    def lidar(self):
        t = self.torch
        frac = t.zeros((self.n_agents, self.cfg.n_lasers), dtype=t.float32, device=self.tdev)
        hit = t.zeros((self.n_agents, self.cfg.n_lasers), dtype=t.int32, device=self.tdev)
        self._check(self.lib.md_lidar(self.h, self._ptr(frac), self._ptr(hit), self._stream()))
        return frac, hit

    def topdown(self, resolution=84, max_distance=30.0, out=None, channels=3):
        """TopDownObservation (obs/top_down_obs.py): [n_agents, resolution, resolution, 3] float32 RGB in [0, 1], the window of
        +-max_distance metres around every agent, turned so that it looks up (md_topdown); channels=2: the per-frame
        [road_network, traffic_flow] grey channels TopDownMultiChannel stacks (md_topdown_channels)"""
        t = self.torch
        assert channels in (2, 3)
        if out is None:
            out = t.empty((self.n_agents, resolution, resolution, channels), dtype=t.float32, device=self.tdev)
        assert out.is_contiguous() and tuple(out.shape) == (self.n_agents, resolution, resolution, channels)
        fn = self.lib.md_topdown if channels == 3 else self.lib.md_topdown_channels
        self._check(fn(self.h, self._ptr(out), int(resolution), float(max_distance), self._stream()))
        return out

    def dynamics(self, act3, n_sub):
        self._check(self.lib.md_dynamics(self.h, self._ptr(act3), int(n_sub), self._stream()))

    def after_step(self):
        self._check(self.lib.md_after_step(self.h, self._stream()))

    def idm(self):
        t = self.torch
        out = t.zeros((self.cfg.n_envs * self.cfg.slots_per_env, 2), dtype=t.float32, device=self.tdev)
        self._check(self.lib.md_idm(self.h, self._ptr(out), self._stream()))
        return out

    # ------------------------------------------------------------------ state snapshots
    def get_state(self, name):
        ref = self._host[name]
        out = np.empty_like(ref)
        self.torch.cuda.synchronize(self.tdev)
        self._check(self.lib.md_get_state(self.h, name.encode(), out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def set_state(self, name, value):
        v = np.ascontiguousarray(value, self._host[name].dtype)
        assert v.shape == self._host[name].shape, (v.shape, self._host[name].shape)
        self.torch.cuda.synchronize(self.tdev)
        self._check(self.lib.md_set_state(self.h, name.encode(), v.ctypes.data_as(C.c_void_p), v.nbytes))

    def attach_bank(self, bank, seed=0):
        """Finished envs restart as a scenario drawn from `bank` (a fully reset BatchedSim with one env per scenario of
        the library, same map set); see md_attach_bank in include/mdstep.h.  The bank is kept alive by this object."""
        self._check(self.lib.md_attach_bank(self.h, bank.h if bank is not None else None, int(seed)))
        self._bank = bank

    def snapshot(self):
        self._check(self.lib.md_snapshot(self.h))

    def profile_begin(self, max_steps):
        self._check(self.lib.md_profile_begin(self.h, int(max_steps)))
        self._prof_cap = int(max_steps)

    KERNELS = ["k_pre", "k_dyn", "k_post", "k_reset", "k_lidar"]

    def profile_end(self):
        """[n, 5] ms of k_pre, k_dyn, k_post, fused reset, k_lidar for the step calls since profile_begin (synchronises)."""
        self.torch.cuda.synchronize(self.tdev)
        a = np.zeros((self._prof_cap, 5), np.float32)
        n = self.lib.md_profile_end(self.h, a.ctypes.data_as(C.c_void_p), self._prof_cap)
        if n < 0:
            self._check(n)
        return a[:n]

    def enable_contacts(self, on=True):
        """Record the contact pairs of every step (md_enable_contacts)."""
        self._check(self.lib.md_enable_contacts(self.h, int(bool(on))))

    def contacts(self):
        """[NV, 4] uint32: per vehicle slot the bodies touched during the last step (bit k < S: vehicle slot k of the env,
        bit S + j: object j)."""
        out = np.zeros((self.cfg.n_envs * self.cfg.slots_per_env, 4), np.uint32)
        self._check(self.lib.md_get_contacts(self.h, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def contact_pairs(self, env=0):
        """the last step's contact pairs of one env as sorted (low, high) index pairs; objects are S + j"""
        S = self.cfg.slots_per_env
        rows = self.contacts()[env * S:(env + 1) * S]
        pairs = set()
        for i in range(S):
            for w in range(4):
                m = int(rows[i, w])
                while m:
                    b = (m & -m).bit_length() - 1
                    k = 32 * w + b
                    pairs.add((min(i, k), max(i, k)))
                    m &= m - 1
        return sorted(pairs)

    def fp32_peak(self):
        """Measured FP32-FMA peak of this GPU in TFLOP/s (md_fp32_peak)."""
        v = C.c_double()
        self._check(self.lib.md_fp32_peak(self.device, C.byref(v)))
        return float(v.value)

    @property
    def launch_count(self):
        return int(self.lib.md_launch_count(self.h))

    def close(self):
        if self.h:
            self.lib.md_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

"""Gymnasium-surface drop-ins for the step path: MetaDriveEnv / SafeMetaDriveEnv (single env, reference return
shapes and info keys) and BatchedMetaDriveEnv (E envs per call, tensors in / tensors out).

Surface mirrored (paths relative to /root/reference/metadrive):
  Env(config)                         envs/base_env.py:278-313     unknown key -> KeyError (utils/config.py:136-147)
  reset(seed) -> (obs, info)          envs/base_env.py:502-537     seed outside [start_seed, start_seed+num_scenarios) -> AssertionError (:886-891)
  step(action) -> (obs, r, term, trunc, info)   envs/base_env.py:426-431, 586-623
  observation_space / action_space    obs/state_obs.py:172-183; policy/env_input_policy.py:50-68
  info keys                           component/vehicle/base_vehicle.py:243-252; envs/metadrive_env.py:132-152, 204, 269;
                                      envs/base_env.py:614-616; envs/safe_metadrive_env.py:31-35
Scenes are GENERATED on the product side (pgmap: BIG + the PG blocks; pgspawn: ego, IDM traffic, accident scenes, with
the reference's seeded streams), so `map="S"`, `map="SCO"`, `map=5`, any `traffic_density` / `accident_prob` run; the
scenario libraries exported from the reference are its goldens (tests/test_pgmap.py) and a fast path for the benchmark.
A config outside what is restated raises, it is never silently approximated.  Rendering / image observation / manual
control keys raise NotImplementedError.
"""
import math

import numpy as np

from .abi import TRAFFIC_MODES
from .library import GeneratedLibrary, ScenarioLibrary

DEFAULT_AGENT = "default_agent"

# the keys of BASE_DEFAULT_CONFIG + METADRIVE_DEFAULT_CONFIG that touch the step path, with the reference defaults
# (envs/base_env.py:32-266, envs/metadrive_env.py:16-89)
STEP_DEFAULTS = dict(
    start_seed=0, num_scenarios=1, map=3, traffic_density=0.1, traffic_mode="trigger", need_inverse_traffic=False,
    random_traffic=False, accident_prob=0.0, static_traffic_object=True, random_spawn_lane_index=True, horizon=None,
    truncate_as_terminate=False, decision_repeat=5, physics_world_step_size=2e-2, discrete_action=False,
    use_multi_discrete=False, discrete_steering_dim=5, discrete_throttle_dim=5,
    success_reward=10.0, out_of_road_penalty=5.0, crash_vehicle_penalty=5.0, crash_object_penalty=5.0,
    driving_reward=1.0, speed_reward=0.1, use_lateral_reward=False, crash_vehicle_cost=1.0, crash_object_cost=1.0,
    out_of_road_cost=1.0, out_of_route_done=False, on_continuous_line_done=True, crash_vehicle_done=True,
    crash_object_done=True, crash_human_done=True, cost_to_reward=False, enable_idm_lane_change=True,
    use_render=False, image_observation=False, manual_control=False, log_level=20, random_agent_model=False,
    random_lane_width=False, random_lane_num=False, map_config=dict(lane_width=3.5, lane_num=3, exit_length=50),
    vehicle_config=dict(lidar=dict(num_lasers=240, distance=50, num_others=0, gaussian_noise=0.0, dropout_prob=0.0,
                                   add_others_navi=False),
                        side_detector=dict(num_lasers=0, distance=50), lane_line_detector=dict(num_lasers=0, distance=20),
                        enable_reverse=False, vehicle_model="default", overtake_stat=False),
    # record / replay (envs/base_env.py:255-263; manager/record_manager.py, replay_manager.py), see MetaDriveEnv.dump_episode
    record_episode=False, replay_episode=None, only_reset_when_replay=False,
    # extensions of this build: which device hosts the simulation; crossing pedestrians per env (peds.py, BASELINE
    # config 5 - the reference has the Pedestrian object but no spawner for PG maps)
    device=0, num_pedestrians=0,
)
UNSUPPORTED_TRUE = ("use_render", "image_observation", "manual_control")
MAX_HANDLES = 8   # GPU handles a single-env wrapper keeps alive (least recently used seeds are closed)


VP_REVERSE = 15  # include/md_layout.h: column of veh_p


def _apply_vehicle_config(arrays, config):
    """Per-agent vehicle_config entries that act on the step path are written over the library's rows: enable_reverse
    (component/vehicle/base_vehicle.py:157, 479-481).  The library was exported with the reference's defaults."""
    vc = config["vehicle_config"]
    if vc["vehicle_model"] not in ("default", "static_default", "s", "m", "l", "xl"):
        raise NotImplementedError("vehicle_model %r is not one of the reference's PG vehicle types" % vc["vehicle_model"])
    agents = arrays["veh_i"][:, 0] == 1
    arrays["veh_p"] = np.array(arrays["veh_p"], np.float32, copy=True)
    arrays["veh_p"][agents, VP_REVERSE] = 1.0 if vc["enable_reverse"] else 0.0
    return arrays


class Box:
    """Minimal gymnasium.spaces.Box stand-in used when gymnasium is not installed (same fields / contains)."""
    def __init__(self, low, high, shape, dtype=np.float32):
        self.low = np.full(shape, low, dtype)
        self.high = np.full(shape, high, dtype)
        self.shape, self.dtype = tuple(shape), np.dtype(dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

    __contains__ = contains

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)


class Discrete:
    """gymnasium.spaces.Discrete / MultiDiscrete stand-in (policy/env_input_policy.py:50-68)."""
    def __init__(self, nvec):
        self.nvec = np.atleast_1d(np.asarray(nvec, np.int64))
        self.n = int(self.nvec[0]) if self.nvec.size == 1 else None
        self.shape = () if self.nvec.size == 1 else (self.nvec.size, )

    def contains(self, x):
        x = np.atleast_1d(np.asarray(x))
        return x.shape == self.nvec.shape and np.issubdtype(x.dtype, np.integer) and bool(np.all((x >= 0) & (x < self.nvec)))

    __contains__ = contains

    def sample(self):
        v = np.array([np.random.randint(n) for n in self.nvec])
        return int(v[0]) if self.nvec.size == 1 else v


def _action_space(c):
    if not c["discrete_action"]:
        return _box(-1.0, 1.0, (2, ))
    sd, td = c["discrete_steering_dim"], c["discrete_throttle_dim"]
    return Discrete([sd, td]) if c["use_multi_discrete"] else Discrete(sd * td)


def _action_row(c, action):
    """One agent's action as the [A, 2] float row the library takes (Discrete: the index travels in column 0)."""
    if c["discrete_action"] and not c["use_multi_discrete"]:
        return np.array([float(int(action)), 0.0], np.float32)
    return np.asarray(action, np.float32).reshape(2)


def _box(low, high, shape):
    try:
        import gymnasium
        return gymnasium.spaces.Box(low, high, shape=shape, dtype=np.float32)
    except Exception:
        return Box(low, high, shape)


class DictSpace:
    """gymnasium.spaces.Dict stand-in for the multi-agent surface: `.spaces` keyed by agent id, `.contains(dict)`."""
    def __init__(self, spaces):
        self.spaces = dict(spaces)

    def contains(self, x):
        return isinstance(x, dict) and all(k in self.spaces and self.spaces[k].contains(np.asarray(v, np.float32))
                                           for k, v in x.items())

    __contains__ = contains

    def sample(self):
        return {k: s.sample() for k, s in self.spaces.items()}


def _merge(default, user, path=""):
    out = dict(default)
    for k, v in (user or {}).items():
        if k not in default:
            raise KeyError("'{}' does not exist in existing config. Please use config.update(...) to update the "
                           "config. Existing keys: {}.".format(path + k, sorted(default.keys())))
        out[k] = _merge(default[k], v, path + k + ".") if isinstance(default[k], dict) and isinstance(v, dict) else v
    return out


def _merge_new(default, user):
    """Config.update(..., allow_add_new_key=True): what the reference's env classes use to extend the defaults"""
    out = dict(default)
    for k, v in (user or {}).items():
        out[k] = _merge_new(default[k], v) if isinstance(default.get(k), dict) and isinstance(v, dict) else v
    return out


def _obs_dim(vc):
    """LidarStateObservation.observation_space (obs/state_obs.py:30-62, 172-183): side block (2 distances or the side
    detector's rays) + 6 + lane block (1 offset or the lane-line detector's rays) + navi 10 + others 4k (8k with add_others_navi) + lidar N."""
    side = vc["side_detector"]["num_lasers"] or 2
    lane = vc["lane_line_detector"]["num_lasers"] or 1
    return side + 6 + lane + 10 + (8 if vc["lidar"]["add_others_navi"] else 4) * vc["lidar"]["num_others"] + vc["lidar"]["num_lasers"]


class _Agent:
    """Read-only view of the ego for the attributes the reference's tests poke at (env.agent.position, ...)."""
    def __init__(self, env):
        self._env = env

    def _row(self, name):
        return self._env._sim.get_state(name)[0]

    @property
    def position(self):
        return self._row("veh_s")[0:2].astype(np.float64)

    @property
    def heading_theta(self):
        w, x, y, z = self._row("veh_s")[3:7].astype(np.float64)
        fx, fy = 2 * (x * y - w * z), 1 - 2 * (x * x + z * z)
        return float(np.arctan2(fy, fx))

    @property
    def speed_km_h(self):
        v = self._row("veh_s")[7:9].astype(np.float64)
        return float(np.hypot(*v) * 3.6)

    def _flag(self, bit):
        return bool(int(self._row("veh_i")[8]) & bit)

    crash_vehicle = property(lambda s: s._flag(0x1))
    crash_object = property(lambda s: s._flag(0x2))
    crash_building = property(lambda s: s._flag(0x4))
    crash_human = property(lambda s: s._flag(0x8))
    crash_sidewalk = property(lambda s: s._flag(0x10))
    on_white_continuous_line = property(lambda s: s._flag(0x20))
    on_yellow_continuous_line = property(lambda s: s._flag(0x40))
    on_broken_line = property(lambda s: s._flag(0x80))
    on_lane = property(lambda s: s._flag(0x100))


class _EngineShim:
    """What user code reaches through `env.engine` on this path: dump_episode() and the record / replay switches."""
    def __init__(self, env):
        self._env = env

    def dump_episode(self):
        return self._env.dump_episode()

    @property
    def record_episode(self):
        return bool(self._env.config["record_episode"])

    @property
    def replay_episode(self):
        return self._env.config["replay_episode"] is not None


class MetaDriveEnv:
    ENV_KIND = "metadrive"
    EXTRA_DEFAULTS = {}

    @classmethod
    def default_config(cls):
        d = dict(STEP_DEFAULTS)
        d.update(cls.EXTRA_DEFAULTS)
        return d

    def __init__(self, config=None):
        self.config = _merge(self.default_config(), config)
        for k in UNSUPPORTED_TRUE:
            if self.config[k]:
                raise NotImplementedError("config['%s'] is outside the step path this build covers" % k)
        if not self.config["static_traffic_object"]:
            raise NotImplementedError("static_traffic_object=False (loose cones that can be pushed) is not covered")
        lid = self.config["vehicle_config"]["lidar"]
        assert 0.0 <= lid["dropout_prob"] <= 1.0  # obs/state_obs.py:240
        self.start_seed = self.start_index = self.config["start_seed"]
        self.num_scenarios = self.env_num = self.config["num_scenarios"]
        self._lib = None
        self._sims = {}     # seed -> handle, at most MAX_HANDLES (least recently used first)
        self._sim = None
        self._rng = np.random.RandomState()   # BaseEnv._reset_global_seed draws from an unseeded stream (base_env.py:886-891)
        self.current_seed = None
        self.agent = _Agent(self)
        self.episode_cost = 0.0
        self._overtake = (set(), set())
        self.observation_space = _box(-0.0, 1.0, (_obs_dim(self.config["vehicle_config"]), ))
        self.action_space = _action_space(self.config)
        self._episode = None      # record_episode: the episode being recorded
        self._replay_t = 0
        self.engine = _EngineShim(self)

    # -- record / replay.  The reference's RecordManager logs the state of every object at every physics sub-step and its
    # ReplayManager restores those states frame by frame (manager/record_manager.py:35-198, replay_manager.py:56-170).  This
    # simulator is deterministic - bit for bit - given the scenario, the actions and the random tape, so an episode is
    # recorded as (scenario seed, config, actions) plus one state frame per env.step (for export and as a check), and replayed
    # by stepping the recorded actions: the replayed observations, rewards and flags ARE the recorded episode's.
    def _frame(self, step):
        vs, vi = self._sim.get_state("veh_s"), self._sim.get_state("veh_i")
        objs = {}
        for k in range(len(vi)):
            if not vi[k, 1]:
                continue
            w, x, y, z = (float(q) for q in vs[k, 3:7])
            yaw = float(np.arctan2(2.0 * (w * z + x * y), 1.0 - 2.0 * (y * y + z * z)))
            objs[DEFAULT_AGENT if k == 0 else "traffic_%d" % k] = dict(
                position=[float(v) for v in vs[k, 0:3]], heading_theta=float((yaw + np.pi / 2 + np.pi) % (2 * np.pi) - np.pi),
                velocity=[float(v) for v in vs[k, 7:9]], type="agent" if vi[k, 0] == 1 else "traffic", active=bool(vi[k, 2]))
        return dict(episode_step=int(step), step_info=objs, agents=[DEFAULT_AGENT])

    def dump_episode(self):
        """BaseEngine.dump_episode (engine/base_engine.py): the episode recorded since the last reset (record_episode=True)."""
        assert self._episode is not None, "record_episode is off or reset() was not called"
        import copy
        return copy.deepcopy(self._episode)

    # -- scenes
    def _library(self):
        """The scenario source of this config: generated on the product side (library.GeneratedLibrary)."""
        if self._lib is None:
            c = self.config
            mc = c["map_config"]
            self._lib = GeneratedLibrary(
                c["start_seed"], c["num_scenarios"], env_kind=self.ENV_KIND, map=c["map"], traffic_density=c["traffic_density"],
                traffic_mode=c["traffic_mode"], accident_prob=c["accident_prob"], lane_num=mc["lane_num"],
                lane_width=mc["lane_width"], exit_length=mc["exit_length"], random_lane_width=c["random_lane_width"],
                random_lane_num=c["random_lane_num"], random_spawn_lane_index=c["random_spawn_lane_index"],
                need_inverse_traffic=c["need_inverse_traffic"], random_traffic=c["random_traffic"],
                random_agent_model=c["random_agent_model"],
                agent_model=c["vehicle_config"]["vehicle_model"])
        return self._lib

    def _cfg_kw(self):
        c = self.config
        return dict(
            n_lasers=c["vehicle_config"]["lidar"]["num_lasers"], lidar_dist=float(c["vehicle_config"]["lidar"]["distance"]),
            num_others=int(c["vehicle_config"]["lidar"]["num_others"]),
            add_others_navi=int(bool(c["vehicle_config"]["lidar"]["add_others_navi"])),
            lidar_gaussian_noise=float(c["vehicle_config"]["lidar"]["gaussian_noise"]),
            lidar_dropout_prob=float(c["vehicle_config"]["lidar"]["dropout_prob"]),
            noise_seed=int(c.get("start_seed", 0) or 0),
            n_side_lasers=int(c["vehicle_config"]["side_detector"]["num_lasers"]),
            side_dist=float(c["vehicle_config"]["side_detector"]["distance"]),
            n_lane_lasers=int(c["vehicle_config"]["lane_line_detector"]["num_lasers"]),
            lane_dist=float(c["vehicle_config"]["lane_line_detector"]["distance"]),
            horizon=int(c["horizon"] or 0), decision_repeat=c["decision_repeat"], dt=c["physics_world_step_size"],
            traffic_mode=TRAFFIC_MODES[c["traffic_mode"]], success_reward=c["success_reward"],
            out_of_road_penalty=c["out_of_road_penalty"], crash_vehicle_penalty=c["crash_vehicle_penalty"],
            crash_object_penalty=c["crash_object_penalty"], driving_reward=c["driving_reward"], speed_reward=c["speed_reward"],
            crash_vehicle_cost=c["crash_vehicle_cost"], crash_object_cost=c["crash_object_cost"],
            out_of_road_cost=c["out_of_road_cost"], use_lateral_reward=int(c["use_lateral_reward"]),
            out_of_route_done=int(c["out_of_route_done"]), on_continuous_line_done=int(c["on_continuous_line_done"]),
            crash_vehicle_done=int(c["crash_vehicle_done"]), crash_object_done=int(c["crash_object_done"]),
            crash_human_done=int(c["crash_human_done"]), truncate_as_terminate=int(c["truncate_as_terminate"]),
            enable_idm_lane_change=int(c["enable_idm_lane_change"]),
            discrete_action=(2 if c["use_multi_discrete"] else 1) if c["discrete_action"] else 0,
            discrete_steering_dim=int(c["discrete_steering_dim"]), discrete_throttle_dim=int(c["discrete_throttle_dim"]),
        )

    # -- gym surface
    def reset(self, seed=None):
        from .sim import BatchedSim
        epi = self.config["replay_episode"]
        if epi is not None:   # the logged scenario is replayed whatever seed is asked for (base_env.py:502-537)
            seed = int(epi["scenario_index"])
            self._replay_t = 0
        if seed is None:   # BaseEnv._reset_global_seed: a uniform draw over the scenario range (envs/base_env.py:886-891)
            seed = int(self._rng.randint(self.start_seed, self.start_seed + self.num_scenarios))
        assert self.start_seed <= seed < self.start_seed + self.num_scenarios, \
            "scenario_index (seed) should be in [{}:{})".format(self.start_seed, self.start_seed + self.num_scenarios)
        lib = self._library()
        if seed in self._sims:
            self._sims[seed] = self._sims.pop(seed)                 # most recently used last
        else:
            arrays, cfg = lib.build_world([lib.index_of_seed(seed)], num_pedestrians=self.config["num_pedestrians"],
                                          seed=seed, **self._cfg_kw())
            _apply_vehicle_config(arrays, self.config)
            self._sims[seed] = BatchedSim(arrays, cfg, device=self.config["device"])
            self._scene = {seed: arrays} if not hasattr(self, "_scene") else {**self._scene, seed: arrays}
            while len(self._sims) > MAX_HANDLES:                    # a long run over 1000 seeds keeps 8 handles, not 1000
                old = next(iter(self._sims))
                self._sims.pop(old).close()
                self._scene.pop(old, None)
        self._sim = self._sims[seed]
        self.current_seed = seed
        self.episode_cost = 0.0
        self._overtake = (set(), set())
        obs = self._sim.reset_host()[0].copy()
        self._episode = None
        if self.config["record_episode"]:
            import time
            self._episode = dict(scenario_index=seed, global_seed=seed, coordinate="MetaDrive", time=time.strftime("%Y-%m-%d_%H-%M-%S"),
                                 global_config={k: v for k, v in self.config.items() if k != "replay_episode"},
                                 frames_per_step=1, frame=[[self._frame(0)]], actions=[])
        return obs, self._info(None)

    def _processed_action(self, action):
        """What EnvInputPolicy.act leaves in info["action"] and BaseVehicle._preprocess_action in info["raw_action"]
        (policy/env_input_policy.py:26-48, component/vehicle/base_vehicle.py:204-209): the continuous pair - decoded
        from the discrete index if need be - with nan -> 0 and clipped to [-1, 1]."""
        c = self.config
        if c["discrete_action"]:
            sd, td = c["discrete_steering_dim"], c["discrete_throttle_dim"]
            if c["use_multi_discrete"]:
                a = (float(action[0]) * (2.0 / (sd - 1)) - 1.0, float(action[1]) * (2.0 / (td - 1)) - 1.0)
            else:
                a = (float(int(action) % sd) * (2.0 / (sd - 1)) - 1.0, float(int(action) // sd) * (2.0 / (td - 1)) - 1.0)
        else:
            a = (float(action[0]), float(action[1]))
        return tuple(0.0 if np.isnan(x) else min(max(x, -1.0), 1.0) for x in a)

    def step(self, action):
        assert self._sim is not None, "call reset() first"
        epi = self.config["replay_episode"]
        replaying = epi is not None and not self.config["only_reset_when_replay"]
        if replaying:
            assert self._replay_t < len(epi["actions"]), "the replayed episode is over (info['replay_done'])"
            action = epi["actions"][self._replay_t]
        a = _action_row(self.config, action).reshape(1, 2)
        obs, rew, cost, term, trunc, flags, info_f = self._sim.step_host(a, autoreset=False)
        self.episode_cost += float(cost[0])
        info = self._info((self._processed_action(action), float(cost[0]), int(flags[0]), info_f[0]))
        if self._episode is not None:
            self._episode["actions"].append(action if np.isscalar(action) else [float(v) for v in np.asarray(action).reshape(-1)])
            self._episode["frame"].append([self._frame(len(self._episode["actions"]))])
        if replaying:
            self._replay_t += 1
            info["replay_done"] = self._replay_t >= len(epi["actions"])   # REPLAY_DONE (replay_manager.py:121-125)
        # cost_to_reward is only declared by the reference (envs/safe_metadrive_env.py:17), never read: the reward stays as is
        return obs[0].copy(), float(rew[0]), bool(term[0]), bool(trunc[0]), info

    def _overtake_num(self):
        """BaseVehicle._update_overtake_stat (component/vehicle/base_vehicle.py:873-892), only with
        vehicle_config.overtake_stat: vehicles of the lidar's broad phase on the ego's current route road are filed as
        "in front" or "behind" by their longitude on the ego's lane; a vehicle seen in front once and behind later counts."""
        if not self.config["vehicle_config"]["overtake_stat"] or self._sim is None:
            return 0
        from . import scene as sc
        arrays = self._scene[self.current_seed]
        vi, vs, rr = self._sim.get_state("veh_i"), self._sim.get_state("veh_s"), self._sim.get_state("veh_rroad")
        lane_f, lane_i = arrays["lane_f"], arrays["lane_i"]
        ego_lane, road = int(vi[0, 4]), int(rr[0, int(vi[0, 5])])
        D = float(self.config["vehicle_config"]["lidar"]["distance"])
        me = sc.lane_local(lane_f[ego_lane], vs[0, 0], vs[0, 1])[0]
        front, back = self._overtake
        for k in range(1, len(vi)):
            if not vi[k, 1] or vi[k, 4] < 0 or int(lane_i[int(vi[k, 4]), 0]) != road:
                continue
            if np.hypot(vs[k, 0] - vs[0, 0], vs[k, 1] - vs[0, 1]) > int(D) + 3.0:
                continue
            if me - sc.lane_local(lane_f[ego_lane], vs[k, 0], vs[k, 1])[0] < 0:
                front.add(k)
                back.discard(k)
            else:
                back.add(k)
        return len(front & back)

    def _info(self, step):
        if step is None:
            a, cost, flags, f = (0.0, 0.0), 0.0, 0x100, np.zeros(8, np.float32)
        else:
            a, cost, flags, f = step
        crash = bool(flags & 0x1f)
        info = {
            "velocity": float(f[0]), "steering": float(f[1]), "acceleration": float(f[2]), "step_energy": float(f[3]),
            "episode_energy": float(f[4]), "policy": "EnvInputPolicy",
            "overtake_vehicle_num": self._overtake_num() if step is not None and hasattr(self, "_overtake") else 0,
            "action": [float(a[0]), float(a[1])], "raw_action": (float(a[0]), float(a[1])),
            "crash_vehicle": bool(flags & 0x1), "crash_object": bool(flags & 0x2), "crash_building": bool(flags & 0x4),
            "crash_human": bool(flags & 0x8), "crash_sidewalk": bool(flags & 0x10), "out_of_road": bool(flags & 0x400),
            "arrive_dest": bool(flags & 0x800), "max_step": bool(flags & 0x1000), "env_seed": self.current_seed,
            "crash": crash, "cost": cost, "step_reward": float(f[5]), "episode_reward": float(f[6]),
            "episode_length": int(f[7]),
        }
        return info

    def close(self):
        for s in self._sims.values():
            s.close()
        self._sims = {}
        self._scene = {}
        self._sim = None

    @property
    def agents(self):
        return {DEFAULT_AGENT: self.agent}


class SafeMetaDriveEnv(MetaDriveEnv):
    """envs/safe_metadrive_env.py:7-35"""
    ENV_KIND = "safe"
    EXTRA_DEFAULTS = dict(num_scenarios=100, accident_prob=0.8, traffic_density=0.05, crash_vehicle_done=False,
                          crash_object_done=False, cost_to_reward=False)

    def _info(self, step):
        info = super()._info(step)
        info["total_cost"] = self.episode_cost
        return info


class TopDownSingleFrameMetaDriveEnv(MetaDriveEnv):
    """envs/top_down_env.py:7-31: MetaDriveEnv whose observation is TopDownObservation (obs/top_down_obs.py) - the ego-centred
    bird's-eye RGB image [resolution_size, resolution_size, 3] of the +-distance metres around the ego (float32 in [0, 1] with
    norm_pixel, else uint8), rendered on the device by md_topdown.  frame_stack / post_stack / frame_skip are the keys of the
    stacked multi-channel variant (TopDownMetaDrive) and are accepted but unused here, as in the reference's class."""
    EXTRA_DEFAULTS = dict(frame_skip=5, frame_stack=3, post_stack=5, norm_pixel=True, resolution_size=84, distance=30)

    def __init__(self, config=None):
        super().__init__(config)
        n = int(self.config["resolution_size"])
        self.observation_space = _box(-0.0, 1.0, (n, n, 3)) if self.config["norm_pixel"] else Box(0, 255, (n, n, 3), np.uint8)

    def _image(self):
        img = self._sim.topdown(int(self.config["resolution_size"]), float(self.config["distance"]))[0].cpu().numpy()
        return img if self.config["norm_pixel"] else (img * 255.0).astype(np.uint8)

    def reset(self, seed=None):
        _, info = super().reset(seed)
        return self._image(), info

    def step(self, action):
        _, r, te, tr, info = super().step(action)
        return self._image(), r, te, tr, info


class TopDownStack:
    """The stacking TopDownMultiChannel does over time (obs/top_down_obs_multi_channel.py:44-60 the two deques, :163-182 the past
    positions, :229-270 observe, :283-290 _get_stack_indices), host logic as in the reference.  observe() takes this step's
    [road_network, traffic_flow] frames (md_topdown_channels) and returns [res, res, 2 + frame_stack]:
    road_network | past positions | traffic now, frame_skip steps ago, 2 * frame_skip steps ago ..."""
    def __init__(self, resolution, max_distance, frame_stack=5, post_stack=5, frame_skip=5):
        from collections import deque
        self.res, self.frame_skip, self.num_stacks = int(resolution), int(frame_skip), 2 + int(frame_stack)
        self.traffic = deque([], maxlen=(frame_stack - 1) * frame_skip + 1)
        self.past_pos = deque([], maxlen=(post_stack - 1) * frame_skip + 1)
        self.scaling = resolution / max_distance   # the reference scales the past positions by resolution / distance (:57),
        self.fill = True                           # twice the image's resolution / (2 * distance): kept

    def reset(self):
        self.fill = True   # the past positions are NOT cleared here: the reference clears them inside the first observe (:236-242)

    def indices(self, length):
        return [length - 1 - i * self.frame_skip for i in range(int(math.ceil(length / self.frame_skip)))]

    def observe(self, road, traffic, position, heading_theta):
        n = self.res
        heading = heading_theta if abs(heading_theta) > 2 * np.pi / 180 else 0.0
        self.past_pos.append(np.asarray(position, np.float64).copy())
        past = np.zeros((n, n), np.float32)
        for k in self.indices(len(self.past_pos)):
            d = (self.past_pos[k] - self.past_pos[-1]) * self.scaling
            fwd, left = d[0] * np.cos(heading) + d[1] * np.sin(heading), -d[0] * np.sin(heading) + d[1] * np.cos(heading)
            # Vector2(dy, dx).rotate(deg(heading) + 90), swapped back and moved to the centre (:167-179): the canvas' x = left
            # of the ego, y = behind it; Surface.fill((x, y), (1, 1)) truncates towards zero and clips to the surface
            x, y = int(np.clip(left + n / 2, -n, n)), int(np.clip(-fwd + n / 2, -n, n))
            if 0 <= x < n and 0 <= y < n:
                past[y, x] = 1.0
        if self.fill:
            self.past_pos.clear()
            self.traffic.clear()
            for _ in range(self.traffic.maxlen):
                self.traffic.append(traffic)
            self.fill = False
        self.traffic.append(traffic)
        img = [road, past] + [self.traffic[i] for i in self.indices(len(self.traffic))]
        return np.clip(np.stack(img, axis=2), 0.0, 1.0).astype(np.float32)


class TopDownMetaDrive(TopDownSingleFrameMetaDriveEnv):
    """envs/top_down_env.py:34-48: the stacked multi-channel bird's-eye observation (TopDownMultiChannel), [resolution_size,
    resolution_size, 2 + frame_stack] float32 in [0, 1]; the two per-frame channels come from md_topdown_channels, TopDownStack
    keeps the history"""
    def __init__(self, config=None):
        super().__init__(config)
        c = self.config
        assert c["norm_pixel"], "the stacked observation is built in [0, 1] (clip_rgb), as TopDownMetaDrive's default"
        n = int(c["resolution_size"])
        self._stack = TopDownStack(n, float(c["distance"]), c["frame_stack"], c["post_stack"], c["frame_skip"])
        self.observation_space = _box(-0.0, 1.0, (n, n, self._stack.num_stacks))

    def _image(self):
        ch = self._sim.topdown(int(self.config["resolution_size"]), float(self.config["distance"]), channels=2)[0].cpu().numpy()
        return self._stack.observe(ch[..., 0], ch[..., 1], self.agent.position, self.agent.heading_theta)

    def reset(self, seed=None):
        self._stack.reset()
        return super().reset(seed)


class BatchedMetaDriveEnv:
    """E independent MetaDriveEnv instances stepped per call, device tensors in and out (the fast path the Gym dict
    surface cannot offer at 10^6+ steps/s).  Finished envs reset in place on device."""
    def __init__(self, num_envs, config=None, env_cls=MetaDriveEnv, rank=0, resample_scenarios=False):
        """`resample_scenarios`: like BaseEnv.reset(seed=None) (envs/base_env.py:886-891), a finished env restarts in a
        scenario drawn from [start_seed, start_seed + num_scenarios) instead of replaying its own - on device, from a
        scenario bank (sim.attach_bank)."""
        from .shard import shard_scenarios
        from .sim import BatchedSim
        proto = env_cls(config)
        lib = proto._library()
        first = lib.index_of_seed(proto.start_seed)
        n = min(proto.num_scenarios, len(lib) - first)
        idx = [first + i for i in shard_scenarios(n, num_envs, rank)]
        universe = list(range(first, first + n)) if resample_scenarios else None
        S = max(4, -(-lib.max_vehicles() // 4) * 4) if resample_scenarios else None
        # object capacity of the whole scenario range, plus the pedestrians build_world appends per env
        O = lib.max_objects() + proto.config["num_pedestrians"] if resample_scenarios else None
        kw = dict(slots_per_env=S, objs_per_env=O, num_pedestrians=proto.config["num_pedestrians"], map_universe=universe)
        arrays, cfg = lib.build_world(idx, seed=rank, **kw, **proto._cfg_kw())
        _apply_vehicle_config(arrays, proto.config)
        self.sim = BatchedSim(arrays, cfg, device=proto.config["device"])
        if resample_scenarios:
            b_arrays, b_cfg = lib.build_world(universe, seed=rank, **kw, **proto._cfg_kw())
            _apply_vehicle_config(b_arrays, proto.config)
            self.bank = BatchedSim(b_arrays, b_cfg, device=proto.config["device"])
            self.bank.reset()
            self.sim.attach_bank(self.bank, seed=proto.start_seed + 7919 * rank)
        self.num_envs = num_envs
        self.observation_space, self.action_space = proto.observation_space, proto.action_space

    def reset(self):
        return self.sim.reset()

    def step(self, actions):
        obs, r, c, te, tr = self.sim.step(actions, autoreset=True)
        return obs, r, te, tr, dict(cost=c, flags=self.sim.info_flags, scalars=self.sim.info_f)

    def close(self):
        self.sim.close()
        if getattr(self, "bank", None) is not None:
            self.bank.close()


# ------------------------------------------------------------------------------------------------ multi-agent
# MULTI_AGENT_METADRIVE_DEFAULT_CONFIG (envs/marl_envs/multi_agent_metadrive.py:12-62) on top of the MetaDriveEnv keys
MA_DEFAULTS = dict(
    is_multi_agent=True, num_agents=15, crash_done=True, out_of_road_done=True, delay_done=25, allow_respawn=True,
    horizon=1000, out_of_road_penalty=10.0, crash_vehicle_penalty=10.0, crash_object_penalty=10.0,
    crash_vehicle_cost=1.0, crash_object_cost=1.0, out_of_road_cost=0.0, traffic_density=0.0, truncate_as_terminate=True,
    force_seed_spawn_manager=False,
    vehicle_config=dict(lidar=dict(num_lasers=72, distance=40, num_others=0, gaussian_noise=0.0, dropout_prob=0.0,
                                   add_others_navi=False),
                        side_detector=dict(num_lasers=0, distance=50), lane_line_detector=dict(num_lasers=0, distance=20),
                        enable_reverse=False, vehicle_model="static_default"),
)


def _ma_cfg_kw(c):
    kw = MetaDriveEnv._cfg_kw(type("C", (), {"config": c})())
    kw.update(delay_done=int(c["delay_done"]), allow_respawn=int(bool(c["allow_respawn"])),
              ma_crash_done=int(bool(c["crash_done"])), ma_out_of_road_done=int(bool(c["out_of_road_done"])))
    if "cross_yellow_line_done" in c:
        # MultiAgentBottleneckEnv._is_out_of_road / reward_function (envs/marl_envs/marl_bottleneck.py:89-135): white solid
        # line | off the lanes | sidewalk, plus the yellow solid line when cross_yellow_line_done; no positive_road sign
        kw.update(out_of_route_done=0, on_continuous_line_done=1 if c["cross_yellow_line_done"] else 2, ignore_road_sign=1)
    if "parking_space_num" in c:
        # MultiAgentParkingLotEnv._is_out_of_road (envs/marl_envs/marl_parking_lot.py:252-256): yellow solid line | off the lanes |
        # sidewalk - the white lines of the lot may be crossed
        kw.update(on_continuous_line_done=5)
    if "overspeed_penalty" in c:
        # MultiAgentTollgateEnv (envs/marl_envs/marl_tollgate.py:185-266): out of road = sidewalk (+ yellow solid line), the toll
        # block's overspeed penalty, the stay-time rule, TollGateObservation
        kw.update(on_continuous_line_done=3 if c["cross_yellow_line_done"] else 4, toll_env=1,
                  min_pass_steps=int(c["vehicle_config"]["min_pass_steps"]), overspeed_penalty=float(c["overspeed_penalty"]))
    return kw


class MultiAgentMetaDrive:
    """Dict-keyed multi-agent surface (envs/marl_envs/multi_agent_metadrive.py:65-212) over one batched env.

    Agent ids grow monotonically ("agent{k}", manager/agent_manager.py:156-159): a finished agent appears one last
    time in the step that finished it; a respawned agent appears with reward 0 in the step that created it; terminated /
    truncated carry "__all__" (:147-149).  Seats of the device simulation are mapped to ids on the host."""
    ASSET = None
    ENV_DEFAULTS = {}

    @classmethod
    def default_config(cls):
        d = _merge(STEP_DEFAULTS, {k: v for k, v in MA_DEFAULTS.items() if k in STEP_DEFAULTS})
        d.update({k: v for k, v in MA_DEFAULTS.items() if k not in STEP_DEFAULTS})
        d["vehicle_config"] = _merge(STEP_DEFAULTS["vehicle_config"], MA_DEFAULTS["vehicle_config"])
        return _merge_new(d, cls.ENV_DEFAULTS)

    @classmethod
    def _make_library(cls, config, **kw):
        """The env's fixed map, generated from config["map_config"] the way the reference's MA*Map classes read it (lane_num,
        lane_width, exit_length; bottleneck: bottle_lane_num / neck_lane_num / neck_length; tollgate: toll_lane_num / toll_length)."""
        from .ma import MultiAgentLibrary
        mc = config["map_config"]
        if cls.ASSET is None:
            # MultiAgentMetaDrive itself (multi_agent_metadrive.py:12-62): the BIG-generated map of config["map"] for the scenario seed,
            # every agent born on the first road and bound for the end of the last block
            if not isinstance(config["map"], (int, str)) or any(b in str(config["map"]) for b in "BP"):
                raise NotImplementedError("map %r is not covered" % (config["map"], ))
            return MultiAgentLibrary("pg", lane_num=int(mc["lane_num"]), lane_width=float(mc["lane_width"]), exit_length=float(mc["exit_length"]),
                                     pg_seed=int(kw.pop("pg_seed", config["start_seed"])), pg_map=config["map"])
        chain = {k: mc[k] for k in ("neck_lane_num", "neck_length", "toll_lane_num", "toll_length") if k in mc}
        return MultiAgentLibrary(cls.ASSET, lane_num=int(mc.get("bottle_lane_num", mc["lane_num"])), lane_width=float(mc["lane_width"]),
                                 exit_length=float(mc["exit_length"]), **chain, **kw)

    def __init__(self, config=None):
        self.config = _merge(self.default_config(), config)
        for k in UNSUPPORTED_TRUE:
            if self.config[k]:
                raise NotImplementedError("config['%s'] is outside the step path this build covers" % k)
        if abs(self.config["traffic_density"]) >= 1e-2 and self.config["traffic_mode"] != "trigger":
            raise NotImplementedError("multi-agent envs with respawn / hybrid IDM traffic are not covered (trigger mode is)")
        lid = self.config["vehicle_config"]["lidar"]
        assert 0.0 <= lid["dropout_prob"] <= 1.0  # obs/state_obs.py:240
        self._lib = self._make_library(self.config)
        self.num_agents = self.config["num_agents"]
        if self.num_agents == -1:
            self.num_agents = self._lib.max_capacity
        assert 0 < self.num_agents <= self._lib.max_capacity, \
            "Too many agents! We only accept {} agents, but you have {} agents!".format(self._lib.max_capacity, self.num_agents)
        od = _obs_dim(self.config["vehicle_config"])
        if "overspeed_penalty" in self.config:   # TollGateObservation.observation_space (marl_tollgate.py:85-93): no navi, + 2
            od += 2 - 10
        self._obs_box = _box(-0.0, 1.0, (od, ))
        self._act_box = _action_space(self.config)
        self._seat_id = ["agent%d" % k for k in range(self.num_agents)] + [None]
        self._active = set(self._seat_id[:self.num_agents])
        self._sim = None
        self._episode = 0
        self._record, self._replay_t = None, 0
        self.current_seed = self.config["start_seed"]

    # -- gym surface
    def reset(self, seed=None):
        from .sim import BatchedSim
        if self._sim is not None:
            self._sim.close()
        rs = (seed if seed is not None else self.config["start_seed"]) * 1000003 + self._episode
        traffic_seed = seed if seed is not None else self.config["start_seed"]
        if self.ASSET is None:   # the map follows the scenario seed (envs/base_env.py:886-891, manager/pg_map_manager.py:57-74)
            if seed is None:
                traffic_seed = int(np.random.randint(self.config["start_seed"], self.config["start_seed"] + self.config["num_scenarios"]))
            assert self.config["start_seed"] <= traffic_seed < self.config["start_seed"] + self.config["num_scenarios"], \
                "scenario_index (seed) should be in [{}:{})".format(self.config["start_seed"], self.config["start_seed"] + self.config["num_scenarios"])
            if int(self._lib.pg_seed) != int(traffic_seed):
                self._lib = self._make_library(self.config, pg_seed=traffic_seed)
            self.current_seed = traffic_seed
        epi = self.config["replay_episode"]
        if epi is not None:   # the logged episode is replayed whatever seed is asked for (envs/base_env.py:502-537): same spawn
            rs, traffic_seed = int(epi["world_seed"]), int(epi["scenario_index"])   # draws, same respawn tape, same traffic
            self._replay_t = 0
        self._episode += 1
        arrays, cfg = self._lib.build_world(1, self.num_agents, seed=rs, traffic_density=self.config["traffic_density"],
                                            traffic_seed=traffic_seed, **_ma_cfg_kw(self.config))
        _apply_vehicle_config(arrays, self.config)
        self._sim = BatchedSim(arrays, cfg, device=self.config["device"])
        self._seat_id = ["agent%d" % k for k in range(self.num_agents)] + [None]
        self._next_id = self.num_agents
        self._active = set(self._seat_id[:self.num_agents])
        obs = self._sim.reset_host()
        self.episode_step = 0
        self._record = None
        if self.config["record_episode"]:
            # record / replay the way the single-agent envs do it (MetaDriveEnv.dump_episode): the simulation is deterministic bit for
            # bit, so an episode is the seeds of its world + the actions per agent id + one state frame per env.step
            import time
            self._record = dict(scenario_index=int(traffic_seed), global_seed=int(traffic_seed), world_seed=int(rs), coordinate="MetaDrive",
                                time=time.strftime("%Y-%m-%d_%H-%M-%S"),
                                global_config={k: v for k, v in self.config.items() if k != "replay_episode"},
                                frames_per_step=1, frame=[[self._frame(0)]], actions=[])
        return ({self._seat_id[k]: obs[k].copy() for k in range(self.num_agents)},
                {self._seat_id[k]: self._info(None) for k in range(self.num_agents)})

    @property
    def agents(self):
        return {k: None for k in self._seat_id if k in self._active}

    @property
    def observation_space(self):
        return DictSpace({k: self._obs_box for k in self.agents})

    @property
    def action_space(self):
        return DictSpace({k: self._act_box for k in self.agents})

    def _frame(self, step):
        """One frame of a recorded episode: every body of the world under its agent id (traffic: its slot)."""
        vs, vi = self._sim.get_state("veh_s"), self._sim.get_state("veh_i")
        objs = {}
        for k in range(len(vi)):
            if not vi[k, 1]:
                continue
            name = self._seat_id[k] if k <= self.num_agents and vi[k, 0] == 1 else "traffic_%d" % k
            objs[str(name)] = dict(position=[float(v) for v in vs[k, 0:3]], quaternion=[float(v) for v in vs[k, 3:7]],
                                   velocity=[float(v) for v in vs[k, 7:9]], type="agent" if vi[k, 0] == 1 else "traffic",
                                   active=bool(vi[k, 2]))
        return dict(episode_step=int(step), step_info=objs, agents=sorted(self._active))

    def dump_episode(self):
        assert getattr(self, "_record", None) is not None, "record_episode is off or reset() was not called"
        import copy
        return copy.deepcopy(self._record)

    @property
    def engine(self):
        return _EngineShim(self)

    def step(self, actions):
        assert self._sim is not None, "call reset() first"
        epi = self.config["replay_episode"]
        replaying = epi is not None and not self.config["only_reset_when_replay"]
        if replaying:
            assert self._replay_t < len(epi["actions"]), "the replayed episode is over (info['replay_done'])"
            actions = epi["actions"][self._replay_t]
        NA = self.num_agents + 1
        a = np.zeros((NA, 2), np.float32)
        acting = {}
        for k in range(NA):
            aid = self._seat_id[k]
            if aid in self._active:
                assert aid in actions, "missing action for " + aid
                a[k] = _action_row(self.config, actions[aid])
                acting[k] = aid
        obs, rew, cost, term, trunc, flags, info_f = self._sim.step_host(a, autoreset=False)
        self.episode_step += 1
        o, r, tm, tc, info = {}, {}, {}, {}, {}
        for k in range(NA):
            fl = int(flags[k])
            if not fl & 0x2000:  # FL_VALID
                continue
            if fl & 0x4000:  # FL_NEWBORN: a fresh id takes the seat (agent_manager.py:136-159)
                aid = "agent%d" % self._next_id
                self._next_id += 1
                self._seat_id[k] = aid
                self._active.add(aid)
                o[aid], r[aid], tm[aid], tc[aid] = obs[k].copy(), 0.0, False, False
                info[aid] = self._info(None)
                continue
            aid = acting[k]
            o[aid], r[aid], tm[aid], tc[aid] = obs[k].copy(), float(rew[k]), bool(term[k]), bool(trunc[k])
            info[aid] = self._info((a[k], float(cost[k]), fl, info_f[k]))
            if tm[aid] or tc[aid]:
                self._active.discard(aid)
        tc["__all__"] = all(tc.values())
        tm["__all__"] = all(tm.values())
        if self._record is not None:
            self._record["actions"].append({aid: (actions[aid] if np.isscalar(actions[aid]) else
                                                  [float(v) for v in np.asarray(actions[aid]).reshape(-1)]) for aid in acting.values()})
            self._record["frame"].append([self._frame(len(self._record["actions"]))])
        if replaying:
            self._replay_t += 1
            for aid in info:
                info[aid]["replay_done"] = self._replay_t >= len(epi["actions"])   # REPLAY_DONE (replay_manager.py:121-125)
        return o, r, tm, tc, info

    _info = MetaDriveEnv._info

    def close(self):
        if self._sim is not None:
            self._sim.close()
            self._sim = None


class MultiAgentRoundaboutEnv(MultiAgentMetaDrive):
    """envs/marl_envs/marl_inout_roundabout.py:12-24, 145-153"""
    ASSET = "ma_roundabout.npz"
    ENV_DEFAULTS = dict(num_agents=40, map_config=dict(exit_length=60, lane_num=2))


class MultiAgentIntersectionEnv(MultiAgentMetaDrive):
    """envs/marl_envs/marl_intersection.py:12-25, 98-108"""
    ASSET = "ma_intersection.npz"
    ENV_DEFAULTS = dict(num_agents=30, map_config=dict(exit_length=60, lane_num=2))


class MultiAgentBottleneckEnv(MultiAgentMetaDrive):
    """envs/marl_envs/marl_bottleneck.py:10-25, 81-139: 4 lanes merge into 1 and split again (I -> Merge -> Split); agents
    are born at both ends and drive to the other end (no destination draw), 4-ray side / lane-line detectors"""
    ASSET = "ma_bottleneck.npz"
    ENV_DEFAULTS = dict(num_agents=20, cross_yellow_line_done=True,
                        map_config=dict(exit_length=60, lane_num=4, bottle_lane_num=4, neck_lane_num=1, neck_length=20),
                        vehicle_config=dict(side_detector=dict(num_lasers=4, distance=50),
                                            lane_line_detector=dict(num_lasers=4, distance=20)))


class MultiAgentTollgateEnv(MultiAgentMetaDrive):
    """envs/marl_envs/marl_tollgate.py:15-36, 167-277: 3 lanes split into 8 toll lanes and merge again (I -> Split -> TollGate ->
    Merge), a booth (static box: crash_building, lidar-visible) on every second toll lane; agents are born at both ends and
    must stay in the toll block for at least min_pass_steps (else the step after they leave ends their episode as
    out_of_road) at no more than the lanes' speed limit (else the overspeed penalty replaces the driving reward);
    TollGateObservation = ego state without the navigation block + lidar + [in the toll block, stayed long enough]"""
    ASSET = "ma_tollgate.npz"
    ENV_DEFAULTS = dict(num_agents=40, cross_yellow_line_done=True, speed_reward=0.0, overspeed_penalty=0.5,
                        map_config=dict(exit_length=70, lane_num=3, toll_lane_num=8, toll_length=10),
                        vehicle_config=dict(min_pass_steps=30, side_detector=dict(num_lasers=72, distance=20),
                                            lane_line_detector=dict(num_lasers=4, distance=20),
                                            lidar=dict(num_lasers=72, distance=20)))


class MultiAgentParkingLotEnv(MultiAgentMetaDrive):
    """envs/marl_envs/marl_parking_lot.py:22-43, 47-142, 199-260: a one-lane street with four parking spaces on either side between
    the first block and a T intersection.  Agents are born in the parking spaces (they pull out and leave by one of the three roads)
    and on the three roads into the lot (each is sent to a parking space nobody else is heading for); vehicles can reverse
    (enable_reverse), white lines may be crossed.  Respawn follows ParkingLotSpawnManager's rules - no newcomer from outside while
    every space is spoken for, a space is free again once its agent is done - on the device (`MdConfig.parking_spaces`).  The
    reference's own `_respawn_single_vehicle` override drops the drawn place (it calls `vehicle.reset()` without the config, :235),
    so its newborns all land on the first road's default pose without a destination; this build respawns them where the spawn
    manager put them (DESIGN.md "Deliberate differences")."""
    ASSET = "ma_parkinglot.npz"
    ENV_DEFAULTS = dict(num_agents=10, parking_space_num=8, map_config=dict(exit_length=20, lane_num=1),
                        vehicle_config=dict(enable_reverse=True))

    @classmethod
    def _make_library(cls, config):
        from .ma import MultiAgentLibrary
        n = int(config["parking_space_num"])
        assert n % 2 == 0, "number of parking spaces must be multiples of 2"      # marl_parking_lot.py:213-214
        assert n >= 4, "minimal number of parking space is 4"
        if n > 30:
            raise NotImplementedError("at most 30 parking spaces (who is heading for which space is a 32-bit set on the device)")
        return super()._make_library(config, parking_space_num=n)


class BatchedMultiAgentEnv:
    """E independent multi-agent envs per call, tensors in / tensors out: actions [E, seats, 2] -> obs [E*seats, 19+N],
    reward / terminated / truncated [E*seats], info flags (FL_VALID marks the seats that produced a transition this step,
    FL_NEWBORN the seats respawned this step).  seats = num_agents + 1.  Finished envs reset in place on device."""
    def __init__(self, num_envs, config=None, env_cls=MultiAgentRoundaboutEnv, seed=0):
        from .sim import BatchedSim
        c = _merge(env_cls.default_config(), config)
        lib = env_cls._make_library(c)
        n = c["num_agents"] if c["num_agents"] != -1 else lib.max_capacity
        if abs(c["traffic_density"]) >= 1e-2 and c["traffic_mode"] != "trigger":
            raise NotImplementedError("multi-agent envs with respawn / hybrid IDM traffic are not covered (trigger mode is)")
        arrays, cfg = lib.build_world(num_envs, n, seed=seed, traffic_density=c["traffic_density"], traffic_seed=c["start_seed"],
                                      **_ma_cfg_kw(c))
        _apply_vehicle_config(arrays, c)
        self.sim = BatchedSim(arrays, cfg, device=c["device"])
        self.num_envs, self.seats = num_envs, n + 1

    def reset(self):
        return self.sim.reset()

    def step(self, actions):
        obs, r, c, te, tr = self.sim.step(actions.reshape(-1, 2), autoreset=True)
        return obs, r, te, tr, dict(cost=c, flags=self.sim.info_flags, scalars=self.sim.info_f)

    def close(self):
        self.sim.close()

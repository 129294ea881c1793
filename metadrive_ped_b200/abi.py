"""ctypes mirror of include/md_layout.h (MdConfig, MdArrays). Data definitions only."""
import ctypes as C

from .scene import ARRAY_ORDER


class MdConfig(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int), ("slots_per_env", C.c_int), ("agents_per_env", C.c_int), ("objs_per_env", C.c_int),
        ("n_lasers", C.c_int), ("horizon", C.c_int), ("decision_repeat", C.c_int), ("traffic_mode", C.c_int),
        ("dt", C.c_float), ("lidar_dist", C.c_float),
        ("success_reward", C.c_float), ("out_of_road_penalty", C.c_float), ("crash_vehicle_penalty", C.c_float),
        ("crash_object_penalty", C.c_float), ("driving_reward", C.c_float), ("speed_reward", C.c_float),
        ("crash_vehicle_cost", C.c_float), ("crash_object_cost", C.c_float), ("out_of_road_cost", C.c_float),
        ("use_lateral_reward", C.c_int), ("out_of_route_done", C.c_int), ("on_continuous_line_done", C.c_int),
        ("crash_vehicle_done", C.c_int), ("crash_object_done", C.c_int), ("crash_human_done", C.c_int),
        ("truncate_as_terminate", C.c_int), ("enable_idm_lane_change", C.c_int), ("is_multi_agent", C.c_int),
        ("delay_done", C.c_int), ("allow_respawn", C.c_int),
        ("ma_places", C.c_int), ("ma_dests", C.c_int), ("ma_roads", C.c_int), ("tape_len", C.c_int),
        ("ma_crash_done", C.c_int), ("ma_out_of_road_done", C.c_int), ("num_others", C.c_int), ("n_side_lasers", C.c_int), ("n_lane_lasers", C.c_int),
        ("side_dist", C.c_float), ("lane_dist", C.c_float), ("discrete_action", C.c_int),
        ("discrete_steering_dim", C.c_int), ("discrete_throttle_dim", C.c_int),
        ("lidar_gaussian_noise", C.c_float), ("lidar_dropout_prob", C.c_float), ("noise_seed", C.c_int),
        ("env_base", C.c_int), ("ignore_road_sign", C.c_int),
        ("toll_env", C.c_int), ("min_pass_steps", C.c_int), ("overspeed_penalty", C.c_float), ("add_others_navi", C.c_int),
        ("parking_spaces", C.c_int), ("parking_in_roads", C.c_int),
    ]


TRAFFIC_MODES = {"trigger": 0, "respawn": 1, "hybrid": 2}


def make_config(n_envs, slots_per_env, agents_per_env=1, objs_per_env=0, **kw):
    """MdConfig with the reference's defaults (envs/metadrive_env.py:16-89, envs/base_env.py:32-266)."""
    d = dict(
        n_lasers=240, horizon=0, decision_repeat=5, traffic_mode=0, dt=0.02, lidar_dist=50.0,
        success_reward=10.0, out_of_road_penalty=5.0, crash_vehicle_penalty=5.0, crash_object_penalty=5.0,
        driving_reward=1.0, speed_reward=0.1, crash_vehicle_cost=1.0, crash_object_cost=1.0, out_of_road_cost=1.0,
        use_lateral_reward=0, out_of_route_done=0, on_continuous_line_done=1, crash_vehicle_done=1,
        crash_object_done=1, crash_human_done=1, truncate_as_terminate=0, enable_idm_lane_change=1,
        is_multi_agent=0, delay_done=0, allow_respawn=0, ma_places=0, ma_dests=0, ma_roads=0, tape_len=1,
        ma_crash_done=1, ma_out_of_road_done=1, num_others=0, n_side_lasers=0, n_lane_lasers=0, side_dist=50.0, lane_dist=20.0, discrete_action=0,
        discrete_steering_dim=5, discrete_throttle_dim=5, lidar_gaussian_noise=0.0, lidar_dropout_prob=0.0, noise_seed=0, env_base=0, ignore_road_sign=0,
        toll_env=0, min_pass_steps=30, overspeed_penalty=0.5, add_others_navi=0, parking_spaces=0, parking_in_roads=0,
    )
    for k, v in kw.items():
        if k not in d:
            raise KeyError(k)
        d[k] = v
    return MdConfig(n_envs=n_envs, slots_per_env=slots_per_env, agents_per_env=agents_per_env,
                    objs_per_env=objs_per_env, **d)


class MdArrays(C.Structure):
    _fields_ = [(name, C.c_void_p) for name in ARRAY_ORDER]

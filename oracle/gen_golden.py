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


def idm_timers(env, roster):
    """overtake_timer of every roster vehicle's IDM policy (-1: no such policy).  The reference redraws it from the policy's own
    RandomState when a lateral lane change completes (policy/idm_policy.py:285-288); this build draws from a counter hash
    instead (DESIGN.md "Deliberate differences"), so the replay tests inject the trace's draw where the timer was reset."""
    out = np.full(len(roster.vehicles), -1.0)
    for k, v in enumerate(roster.vehicles):
        pol = env.engine.get_policy(v.name)
        if pol is not None and hasattr(pol, "overtake_timer"):
            out[k] = pol.overtake_timer
    return out


class Recorders:
    """What the reference decides per step, besides the state: for every Lidar.perceive call the object each ray hit
    (sensors/distance_detector.py:27-85: `detected_objects` holds one result per hitting ray, in ray order) and the body
    pairs the contact-added callback saw (engine/core/collision_callback.py:5-42).  Installed by wrapping the two module
    functions the engine binds, before the engine exists."""
    def __init__(self):
        import metadrive.component.sensors.distance_detector as dd
        import metadrive.engine.core.engine_core as ec
        from metadrive.utils.utils import get_object_from_node
        self.scans, self.pairs = [], []
        orig_p, orig_c = dd.perceive, ec.collision_callback

        def obj_of(node):
            try:
                return get_object_from_node(node)
            except Exception:
                return None

        def perceive(*a, **kw):
            cp, objs, colors = orig_p(*a, **kw)
            self.scans.append((float(kw["vehicle_position_x"]), float(kw["vehicle_position_y"]), float(kw["height"]),
                               int(kw["num_lasers"]), np.array(cp, np.float64), [obj_of(r.getNode()) for r in objs]))
            return cp, objs, colors

        def callback(contact):
            self.pairs.append((obj_of(contact.getNode0()), obj_of(contact.getNode1())))
            return orig_c(contact)

        dd.perceive, ec.collision_callback = perceive, callback

    def clear(self):
        self.scans.clear()
        self.pairs.clear()

    def index_of(self, roster, obj):
        """roster vehicle k -> k; obstacle j -> 1000 + j; anything else -> -2"""
        for k, v in enumerate(roster.vehicles):
            if v is obj:
                return k
        kept = [o for o in roster.objects if type(o).__name__ in ("TrafficCone", "TrafficWarning", "TrafficBarrier")]
        kept += list(getattr(roster, "buildings", []))   # the toll booths close the object table (Roster.objects_table)
        for j, o in enumerate(kept):
            if o is obj:
                return 1000 + j
        return -2

    def ego_hits(self, roster, n_lasers):
        """hit id per lidar ray of the LAST lidar scan (height 1.2, lidar.py:19), -1 = no hit"""
        scans = [s for s in self.scans if abs(s[2] - 1.2) < 1e-9 and s[3] == n_lasers]
        out = np.full(n_lasers, -1, np.int32)
        if not scans:
            return out
        _, _, _, _, cp, objs = scans[-1]
        it = iter(objs)
        for i in range(n_lasers):
            if cp[i] < 1.0:
                out[i] = self.index_of(roster, next(it))
        return out

    def step_pairs(self, roster):
        """sorted unique (low, high) index pairs the contact-added callback reported since the last clear()"""
        got = set()
        for a, b in self.pairs:
            i, j = self.index_of(roster, a), self.index_of(roster, b)
            if i >= 0 and j >= 0 and i != j and (i < 1000 or j < 1000):
                got.add((min(i, j), max(i, j)))
        return sorted(got)


def _pad_pairs(per_step):
    width = max([len(p) for p in per_step] + [1])
    out = np.full((len(per_step), width, 2), -1, np.int32)
    for t, p in enumerate(per_step):
        if p:
            out[t, :len(p)] = p
    return out


def run_episode(env_cls, config, seed, actions, tag, closed_loop=None):
    from oracle import ref_export as rx
    rec = Recorders()
    env = env_cls(config)
    try:
        obs0, _ = env.reset(seed=seed)
        m, mi = rx.export_map(env.current_map)
        roster = rx.Roster(env, mi)
        init = roster_arrays(env, mi, roster)
        f0, i0 = rx.record_world(env, roster)
        n_lasers = int(env.config["vehicle_config"]["lidar"]["num_lasers"])
        hits, pairs = [rec.ego_hits(roster, n_lasers)], []
        fs, is_, obs, rew, cost, term, trunc, infos = [f0], [i0], [obs0], [], [], [], [], []
        timers = [idm_timers(env, roster)]
        actions = np.array(actions, np.float64)
        for t in range(len(actions)):
            if closed_loop:   # the lane-follow driver of the multi-agent traces, seeded by the tag's seed
                actions[t] = _lane_follow_action(env.agent, closed_loop, 0.02, fast=55, slow=20)
            a = actions[t]
            rec.clear()
            if config.get("discrete_action"):  # Discrete: the index travels in column 0; MultiDiscrete: both columns
                o, r, te, tr, info = env.step([int(a[0]), int(a[1])] if config.get("use_multi_discrete") else int(a[0]))
            else:
                o, r, te, tr, info = env.step(a)
            hits.append(rec.ego_hits(roster, n_lasers))
            pairs.append(rec.step_pairs(roster))
            f, i = rx.record_world(env, roster)
            fs.append(f)
            is_.append(i)
            timers.append(idm_timers(env, roster))
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
            lidar_hit=np.stack(hits), contact_pairs=_pad_pairs(pairs), idm_timer=np.stack(timers),
            config=json.dumps(dict({k: v for k, v in config.items() if isinstance(v, (int, float, str, bool))},
                                   num_others=int(env.config["vehicle_config"]["lidar"]["num_others"]),
                                   **(dict(add_others_navi=1) if env.config["vehicle_config"]["lidar"]["add_others_navi"] else {}),
                                   n_side_lasers=int(env.config["vehicle_config"]["side_detector"]["num_lasers"]),
                                   side_dist=float(env.config["vehicle_config"]["side_detector"]["distance"]),
                                   n_lane_lasers=int(env.config["vehicle_config"]["lane_line_detector"]["num_lasers"]),
                                   lane_dist=float(env.config["vehicle_config"]["lane_line_detector"]["distance"]),
                                   discrete_action=(2 if env.config["use_multi_discrete"] else 1) if env.config["discrete_action"] else 0,
                                   discrete_steering_dim=int(env.config["discrete_steering_dim"]),
                                   discrete_throttle_dim=int(env.config["discrete_throttle_dim"]))),
            **{"init_" + k: v for k, v in init.items()},
        )
        sb = rx.export_static_bodies(env.engine)
        out["ref_lines"] = sb["lines"]
        return out
    finally:
        env.close()


def _lane_follow_action(v, rs, noise, fast=32, slow=16, ahead=2.0, kh=2.5, kl=0.5):
    """Test-side driver for multi-agent traces: steer along the localised lane (slower on tight arcs), plus noise."""
    lane = v.navigation.current_lane
    lon, lat = lane.local_coordinates(v.position)
    err = lane.heading_theta_at(lon + ahead) - v.heading_theta
    err = (err + np.pi) % (2 * np.pi) - np.pi
    steer = kh * err + kl * lat + noise * rs.uniform(-1, 1)
    tight = getattr(lane, "radius", 1e9) < 20
    target = slow if tight else fast
    thr = (0.6 if v.speed_km_h < target else (-0.3 if v.speed_km_h > target + 6 else 0.0)) + noise * rs.uniform(-1, 1)
    return [float(np.clip(steer, -1, 1)), float(np.clip(thr, -1, 1))]


def _parking_action(name, v, rs):
    """Test-side driver for the parking-lot trace: slow lane following; the agents pull out one after the other (agent k waits
    100 k steps) and brake for an agent in front of them, so that some of them reach their parking space / the far end of a road."""
    if v.engine.episode_step < 100 * int(name[5:]):
        return [0.0, 0.0]
    a = _lane_follow_action(v, rs, 0.02, fast=9.0, slow=4.0, ahead=1.0, kh=4.0, kl=1.0)
    hx, hy = np.cos(v.heading_theta), np.sin(v.heading_theta)
    for u in v.engine.agent_manager.active_agents.values():
        if u is v:
            continue
        dx, dy = u.position[0] - v.position[0], u.position[1] - v.position[1]
        ahead, side = dx * hx + dy * hy, -dx * hy + dy * hx
        if 0 < ahead < 8 and abs(side) < 2.5:
            a[1] = -1.0 if v.speed_km_h > 1.0 else 0.0
    return a


def _toll_action(v, rs, noise, patient, waited, toll_x):
    """Test-side driver for the tollgate trace: lane following at ~25 km/h; a patient agent slows down ahead of the toll block
    (`toll_x` = the x range of its lanes; the map runs along x), crawls in and stays for more than min_pass_steps (`waited`
    counts its steps inside), an impatient one drives straight through (overspeed penalty, then the stay-time rule ends its
    episode)."""
    a = _lane_follow_action(v, rs, noise, fast=25, slow=25)
    if not patient:
        return a
    target = None
    if v.navigation.current_road.block_ID() == "$":
        target = 2.0 if waited < 36 else None
    elif waited == 0:
        x = v.position[0]
        dist = toll_x[0] - x if x < toll_x[0] else (x - toll_x[1] if x > toll_x[1] else 0.0)
        if dist < 14:
            target = 2.5 if dist < 2.5 else 10.0
    if target is not None:
        a[1] = -0.8 if v.speed_km_h > target + 0.5 else (0.25 if v.speed_km_h < target - 1.0 else 0.0)
    # keep a gap to a vehicle ahead on the same lane (the patient ones queue in front of the booths)
    hx, hy = np.cos(v.heading_theta), np.sin(v.heading_theta)
    for u in v.engine.agent_manager.active_agents.values():
        if u is v:
            continue
        dx, dy = u.position[0] - v.position[0], u.position[1] - v.position[1]
        ahead, side = dx * hx + dy * hy, -dx * hy + dy * hx
        if 0 < ahead < 9 + 0.3 * v.speed_km_h and abs(side) < 2.0:
            a[1] = -1.0
    return a


def run_episode_ma(env_cls, config, actions, tag, steps=None, noise=0.0, seed=0, obs_stride=1, driver=None):
    """Multi-agent variant.  Seats: the reset-time agents take seats 0..n-1 and one spare seat follows; a respawned
    agent takes the lowest seat that is free and produced no transition this step (what the product does).  Outputs are
    padded per seat: valid[t, k] says whether seat k produced a transition at step t.  `actions` [T, n, 2] or None
    (lane-follow driver with `noise`).  Respawn draws are recorded as (index in the list of clear places, destination
    index) per step so that a replay can feed the same random tape."""
    from oracle import ref_export as rx
    # reproducible multi-agent episodes, the way the reference's own tests pin them
    # (tests/test_env/test_ma_roundabout_env.py:2-3): the spawn manager takes the global seed and the respawn place draw
    # (multi_agent_metadrive.py:199) a fixed one
    config = dict(config, force_seed_spawn_manager=True)
    env_cls._DEBUG_RANDOM_SEED = 10 + seed
    env = env_cls(config)
    rs = np.random.RandomState(seed)
    try:
        obs0, _ = env.reset(seed=0)
        eng = env.engine
        m, mi = rx.export_map(env.current_map)
        roster = rx.Roster(env, mi)
        init = roster_arrays(env, mi, roster)
        names = list(env.agents.keys())
        n = len(names)
        n_seats = n + 1
        seat_of = {k: j for j, k in enumerate(names)}
        seat_vehicle = [env.agents[k] for k in names] + [None]
        seat_name = [v.name for v in seat_vehicle[:n]] + [None]  # engine recycles vehicle objects under new names
        spawn_roads = list(env.config["spawn_roads"])
        road_nodes = np.array([[mi.nodes[r.start_node], mi.nodes[r.end_node]] for r in spawn_roads], np.int32)
        dest_nodes = np.array([mi.nodes.get((-r).end_node, -1) for r in spawn_roads], np.int32)
        # envs without a destination draw (the default SpawnManager.update_destination_for, spawn_manager.py:224-228):
        # auto_assign_task sends every agent to the far end of the map = one destination per spawn road
        fixed_dest = type(eng.spawn_manager).update_destination_for is \
            __import__("metadrive.manager.spawn_manager", fromlist=["SpawnManager"]).SpawnManager.update_destination_for
        sm = eng.spawn_manager
        parking = hasattr(sm, "parking_space_available")
        if fixed_dest:
            dest_nodes = dest_nodes[::-1].reshape(-1, 1).copy()
            if len(spawn_roads) == 1:   # MultiAgentMetaDrive itself: one spawn road, everybody drives to the end of the last block
                dest_nodes[0, 0] = mi.nodes[env.agents[names[0]].navigation.checkpoints[-1]]
        park_log = []
        if parking:
            # MultiAgentParkingLotEnv (envs/marl_envs/marl_parking_lot.py:47-90): an agent born on one of the roads into the lot
            # is sent to a parking space nobody else is heading for (drawn among the AVAILABLE ones), an agent born in a
            # parking space to the far end of one of those roads.  One destination list per spawn road, -1 padded.
            in_roads = list(env.config["in_spawn_roads"])
            spaces = list(env.current_map.parking_space)
            dest_nodes = np.full((len(spawn_roads), max(len(spaces), len(in_roads))), -1, np.int32)
            for ri, r in enumerate(spawn_roads):
                if r in in_roads:
                    dest_nodes[ri, :len(spaces)] = [mi.nodes[s.end_node] for s in spaces]
                else:
                    dest_nodes[ri, :len(in_roads)] = [mi.nodes[(-q).end_node] for q in in_roads]
            orig_gps = sm.get_parking_space

            def logging_gps(v_id):   # the spaces that were available, in list order, and the one that was drawn
                before = sorted(spaces.index(s) for s in sm.parking_space_available)
                got = orig_gps(v_id)
                park_log.append((before, spaces.index(got)))
                return got

            sm.get_parking_space = logging_gps
            name_seat = {k: j for j, k in enumerate(names)}
            parking_taken = np.full(n_seats, -1, np.int32)   # reset-time assignment: seat -> parking space it is heading for
            for vid, s in sm.v_dest_pair.items():
                parking_taken[name_seat[vid]] = spaces.index(s)
        place_keys = list(sm.safe_spawn_places.keys())
        place_lane = [tuple(sm.safe_spawn_places[k]["config"]["spawn_lane_index"]) for k in place_keys]
        first_query = []
        orig = sm.get_available_respawn_places

        def recording(*a, **kw):
            ret = orig(*a, **kw)
            if not first_query:
                first_query.append(list(ret.keys()))
            return ret

        sm.get_available_respawn_places = recording

        def world():
            fs, is_ = [], []
            for v, nm in zip(seat_vehicle, seat_name):
                if v is None or v.name != nm:
                    fs.append(np.zeros(rx.N_STEP_F)); is_.append(np.zeros(8, np.int32))
                else:
                    f, i = rx.record_vehicle(v, roster, env)
                    fs.append(f); is_.append(i)
            for v in roster.traffic:   # IDM traffic of a multi-agent env (traffic_density > 0): rows after the seats, roster order
                f, i = rx.record_vehicle(v, roster, env)
                fs.append(f); is_.append(i)
            return np.stack(fs), np.stack(is_)

        f0, i0 = world()
        fs, is_ = [f0], [i0]
        od = len(obs0[names[0]])
        row0 = np.zeros((n_seats, od), np.float32)
        row0[:n] = np.stack([obs0[k] for k in names])
        obs = [row0]
        T = steps if steps is not None else len(actions)
        rew = np.zeros((T, n_seats)); cost = np.zeros((T, n_seats)); term = np.zeros((T, n_seats), bool)
        trunc = np.zeros((T, n_seats), bool); valid = np.zeros((T, n_seats), bool); flags = np.zeros((T, n_seats), np.int32)
        newborn = np.zeros((T, n_seats), bool)
        acts = np.zeros((T, n_seats, 2))
        draws = np.full((T, 3), -1, np.int32)  # clear-list index, destination index, number of clear places
        routes_new = np.full((T, 24), -1, np.int32)
        done_steps = 0
        for t in range(T):
            live = list(env.agents.keys())
            if not live:
                break
            ad = {}
            for k in live:
                j = seat_of[k]
                if driver is not None:
                    a = driver(k, env.agents[k])
                else:
                    a = _lane_follow_action(env.agents[k], rs, noise) if actions is None else list(actions[t][j])
                acts[t, j] = a
                ad[k] = a
            first_query.clear()
            o, r, te, tr, info = env.step(ad)
            old = [k for k in o if k in seat_of]
            new = [k for k in o if k not in seat_of]
            assert len(new) <= 1
            for k in new:
                v = env.agents[k]
                free = [j for j in range(n_seats)
                        if (seat_name[j] is None or seat_name[j] not in eng._spawned_objects)
                        and not any(seat_of[q] == j for q in old)]
                j = free[0]
                seat_of[k] = j
                seat_vehicle[j] = v
                seat_name[j] = v.name
                keys = first_query[0]
                pk = place_keys[place_lane.index(tuple(v.config["spawn_lane_index"]))]
                if parking:
                    road = [tuple(r) for r in road_nodes.tolist()].index(
                        (mi.nodes[v.config["spawn_lane_index"][0]], mi.nodes[v.config["spawn_lane_index"][1]]))
                    d_abs = int(np.nonzero(dest_nodes[road] == mi.nodes[v.config["destination"]])[0][0])
                    if spawn_roads[road] in in_roads:   # rank of the drawn space among the available ones
                        before, got = park_log[-1]
                        assert got == d_abs
                        d_abs = before.index(got)
                    draws[t] = [keys.index(pk), d_abs, len(keys)]
                elif fixed_dest:
                    road = [tuple(r) for r in road_nodes.tolist()].index(
                        (mi.nodes[v.config["spawn_lane_index"][0]], mi.nodes[v.config["spawn_lane_index"][1]]))
                    assert mi.nodes[v.navigation.checkpoints[-1]] == dest_nodes[road, 0]
                    draws[t] = [keys.index(pk), 0, len(keys)]
                else:
                    draws[t] = [keys.index(pk), int(np.nonzero(dest_nodes == mi.nodes[v.config["destination"]])[0][0]), len(keys)]
                ck = v.navigation.checkpoints
                routes_new[t, :len(ck)] = [mi.nodes[c] for c in ck]
                newborn[t, j] = True
            f, i = world()
            fs.append(f); is_.append(i)
            row = np.zeros((n_seats, od), np.float32)
            for k, ob in o.items():
                j = seat_of[k]
                row[j] = ob; rew[t, j] = r[k]; cost[t, j] = info[k].get("cost", 0.0); term[t, j] = te[k]; trunc[t, j] = tr[k]
                valid[t, j] = True
                if not newborn[t, j]:
                    flags[t, j] = (info[k]["crash_vehicle"] * 1 | info[k]["crash_object"] * 2 | info[k]["crash_building"] * 4
                                   | info[k]["crash_human"] * 8 | info[k]["crash_sidewalk"] * 16
                                   | info[k]["out_of_road"] * 0x400 | info[k]["arrive_dest"] * 0x800
                                   | info[k]["max_step"] * 0x1000)
            obs.append(row)
            done_steps += 1
        steps = done_steps
        obs = np.stack(obs).astype(np.float32)
        if obs_stride > 1:  # keep the fixture small: full observations for every obs_stride-th seat only
            keep = np.zeros(n_seats, bool); keep[::obs_stride] = True
            tail = 2 if hasattr(env, "stay_time_manager") else 0   # TollGateObservation: two toll floats follow the lidar
            obs[:, ~keep, od - tail - int(env.config["vehicle_config"]["lidar"]["num_lasers"]):od - tail] = -1.0
        conf = {k: v for k, v in config.items() if isinstance(v, (int, float, str, bool))}
        conf["n_lasers"] = int(env.config["vehicle_config"]["lidar"]["num_lasers"])
        conf["lidar_dist"] = float(env.config["vehicle_config"]["lidar"]["distance"])
        conf["n_side_lasers"] = int(env.config["vehicle_config"]["side_detector"]["num_lasers"])
        conf["side_dist"] = float(env.config["vehicle_config"]["side_detector"]["distance"])
        conf["n_lane_lasers"] = int(env.config["vehicle_config"]["lane_line_detector"]["num_lasers"])
        conf["lane_dist"] = float(env.config["vehicle_config"]["lane_line_detector"]["distance"])
        conf["ignore_road_sign"] = int("cross_yellow_line_done" in env.config)
        if roster.traffic:
            conf["n_traffic"] = len(roster.traffic)
        if parking:   # marl_parking_lot.py:252-256: yellow solid line, off the lanes or sidewalk = out of road (white lines may be crossed)
            conf.update(parking_spaces=len(spaces), parking_in_roads=len(in_roads), on_continuous_line_done=5,
                        enable_reverse=int(bool(env.config["vehicle_config"]["enable_reverse"])))
        if env.config["vehicle_config"]["lidar"]["num_others"]:   # the others block sits between the state and the lidar floats
            conf.update(num_others=int(env.config["vehicle_config"]["lidar"]["num_others"]),
                        add_others_navi=int(bool(env.config["vehicle_config"]["lidar"]["add_others_navi"])))
        if hasattr(env, "stay_time_manager"):   # MultiAgentTollgateEnv (envs/marl_envs/marl_tollgate.py:15-36, 239-245)
            conf.update(toll_env=1, min_pass_steps=int(env.config["vehicle_config"]["min_pass_steps"]),
                        overspeed_penalty=float(env.config["overspeed_penalty"]), speed_reward=float(env.config["speed_reward"]),
                        on_continuous_line_done=3 if env.config["cross_yellow_line_done"] else 4)
        out = dict(
            tag=tag, seed=int(env.current_seed), lane_num=env.config["map_config"]["lane_num"],
            map_lane_f=m["lane_f"], map_lane_i=m["lane_i"], map_road_i=m["road_i"], map_meta=m["meta"],
            actions=acts[:steps], veh_f=np.stack(fs), veh_i=np.stack(is_), obs=obs, reward=rew[:steps], cost=cost[:steps],
            terminated=term[:steps], truncated=trunc[:steps], valid=valid[:steps], newborn=newborn[:steps],
            info_flags=flags[:steps], respawn_draws=draws[:steps], respawn_routes=routes_new[:steps],
            ma_spawn_roads=road_nodes, ma_dest_nodes=dest_nodes, ma_alive_seats=np.array([n], np.int32),
            config=json.dumps(conf), **{"init_" + k: v for k, v in init.items()},
        )
        if parking:
            out["ma_parking_taken"] = parking_taken
        out["ref_lines"] = rx.export_static_bodies(env.engine)["lines"]
        return out
    finally:
        env.close()


def run_episode_cfg5(config, seed, steps, tag, n_peds=16):
    """BASELINE config 5, composed from reference pieces: MetaDriveEnv(map="X", traffic_mode="respawn") plus
    `n_peds` reference `Pedestrian` objects placed and driven (set_velocity between env.steps) by the build-defined
    crossing model of metadrive_ped_b200/peds.py.  Records traffic respawn events (slot, respawn-lane index, longitude,
    overtake timer, parameters of the new vehicle) so that a replay can feed the same random tape."""
    from oracle import ref_export as rx
    from metadrive.envs.metadrive_env import MetaDriveEnv
    from metadrive.component.traffic_participants.pedestrian import Pedestrian
    from metadrive_ped_b200 import scene as sc
    from metadrive_ped_b200 import peds as pd
    env = MetaDriveEnv(config)
    rs = np.random.RandomState(seed)
    try:
        obs0, _ = env.reset(seed=seed)
        eng = env.engine
        m, mi = rx.export_map(env.current_map)
        geo = sc.build_map_geometry(sc.MapTable.from_export(m, env.config["map_config"]["lane_num"]))
        prow = pd.place_pedestrians(geo, np.random.default_rng(seed), n_peds)
        peds = []
        for r in prow:
            p = eng.spawn_object(Pedestrian, position=[r[1], r[2]], heading_theta=float(np.arctan2(r[9], r[8])), random_seed=1)
            p.set_velocity([r[8], r[9]], in_local_frame=False)
            peds.append(p)
        # the first observation must see the pedestrians: observe again now that they exist
        obs0 = env.observations[next(iter(env.agents))].observe(env.agent)
        roster = rx.Roster(env, mi)
        init = roster_arrays(env, mi, roster)
        init["objects"] = prow.copy()
        respawn_lanes = [it["lane"] for it in json.loads(m["meta"])["respawn"]]
        tm = eng.traffic_manager
        names = [v.name for v in roster.vehicles]

        def world():
            fs, is_ = [], []
            for v, nm in zip(roster.vehicles, names):
                if v.name != nm:
                    fs.append(np.zeros(rx.N_STEP_F)); is_.append(np.zeros(8, np.int32))
                else:
                    f, i = rx.record_vehicle(v, roster, env)
                    fs.append(f); is_.append(i)
            return np.stack(fs), np.stack(is_)

        def ped_state():
            return np.array([[p.position[0], p.position[1], p.velocity[0], p.velocity[1]] for p in peds])

        f0, i0 = world()
        fs, is_, obs, rew, cost, term, trunc, infos, acts, pst = [f0], [i0], [obs0], [], [], [], [], [], [], [ped_state()]
        events = []  # step, slot, place index, longitude, timer + new static row
        ev_static = []
        dt_step = env.config["decision_repeat"] * env.config["physics_world_step_size"]
        for t in range(steps):
            a = _lane_follow_action(env.agent, rs, 0.02, fast=22, slow=14)
            before = list(tm._traffic_vehicles)
            for k, v in enumerate(roster.vehicles):
                v.body._md_slot = k  # index order of the contact model = the trace's slot order (refshim/pbullet.py)
            o, r, te, tr, info = env.step(a)
            after = list(tm._traffic_vehicles)
            removed = [v for v in before if v not in after or v.name != names[roster.vehicles.index(v)]]
            added = after[len(before) - len(removed):] if removed else []
            assert len(added) == len(removed), (len(added), len(removed))
            for old_v, new_v in zip(removed, added):
                k = roster.vehicles.index(old_v)
                roster.vehicles[k] = new_v
                names[k] = new_v.name
                lane_id = mi.lane_id(eng.current_map.road_network.get_lane(new_v.config["spawn_lane_index"]))
                pol = eng.get_policy(new_v.name)
                events.append([t, k, respawn_lanes.index(lane_id), float(new_v.config["spawn_longitude"]), int(pol.overtake_timer)])
                ev_static.append(rx.vehicle_static(new_v))
            for k in pd.step_pedestrians_host(prow, dt_step):
                peds[k].set_velocity([prow[k][8], prow[k][9]], in_local_frame=False)
            f, i = world()
            fs.append(f); is_.append(i); obs.append(o); rew.append(r); cost.append(info["cost"]); term.append(te); trunc.append(tr)
            acts.append(a); pst.append(ped_state())
            infos.append([info["velocity"], info["steering"], info["acceleration"], info["step_energy"],
                          info["episode_energy"], info["step_reward"], info["episode_reward"], info["episode_length"]])
            if te or tr:
                break
        conf = {k: v for k, v in config.items() if isinstance(v, (int, float, str, bool))}
        out = dict(
            tag=tag, seed=seed, lane_num=env.config["map_config"]["lane_num"],
            map_lane_f=m["lane_f"], map_lane_i=m["lane_i"], map_road_i=m["road_i"], map_meta=m["meta"],
            actions=np.asarray(acts, np.float64), veh_f=np.stack(fs), veh_i=np.stack(is_),
            obs=np.stack(obs).astype(np.float32), reward=np.asarray(rew, np.float64), cost=np.asarray(cost, np.float64),
            terminated=np.asarray(term, bool), truncated=np.asarray(trunc, bool), info=np.asarray(infos, np.float64),
            ped_state=np.stack(pst), respawn_events=np.asarray(events, np.float64).reshape(-1, 5),
            respawn_static=np.asarray(ev_static, np.float64).reshape(-1, 16),
            config=json.dumps(conf), **{"init_" + k: v for k, v in init.items()},
        )
        out["ref_lines"] = rx.export_static_bodies(env.engine)["lines"]
        return out
    finally:
        env.close()


ALL_TAGS = ["cfg1_S_straight", "cfg1_S_random", "cfg1_S_discrete", "cfg2_pg3_seed3", "cfg2_pg3_seed7", "cfg2_pg3_seed11_dense",
            "cfg2_SCO_nolimit", "cfg2_pg3_seed11_others4", "cfg2_pg3_seed3_others_navi", "cfg2_pg3_seed3_detectors", "cfg4_safe_seed2", "cfg4_safe_seed5",
            "cfg4_safe_seed40_cones", "cfg4_safe_seed8_bump", "cfg2_StollC_seed0", "cfg3_ma_roundabout", "cfg3_ma_roundabout_respawn",
            "cfg3_ma_intersection_respawn", "cfg3_ma_intersection_others_navi", "cfg3_ma_parkinglot", "cfg3_ma_roundabout_traffic", "cfg3_ma_pg3", "cfg3_ma_bottleneck_respawn", "cfg3_ma_tollgate_respawn", "cfg5_ped_X"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--only", default=None)
    ap.add_argument("--ma-steps", type=int, default=450)
    ap.add_argument("--ma-noise", type=float, default=0.04)
    ap.add_argument("--cfg5-steps", type=int, default=350)
    ap.add_argument("--exact", action="store_true", help="--only names one tag exactly (not a substring)")
    args = ap.parse_args()
    if not args.only:
        # ONE FRESH PROCESS PER TAG.  The reference keeps state in class attributes - the multi-agent maps assign
        # Roundabout.EXIT_PART_LENGTH / InterSection.EXIT_PART_LENGTH / TInterSection.EXIT_PART_LENGTH
        # (envs/marl_envs/marl_inout_roundabout.py:46, marl_intersection.py:46, marl_parking_lot.py:171) - so an env built
        # after a multi-agent env in the same process gets other maps for the same (config, seed).  Round 1 generated all
        # tags in one process: `cfg2_SCO_nolimit` carried a roundabout with 60 m exits.
        import subprocess
        for tag in ALL_TAGS:
            cmd = [sys.executable, "-m", "oracle.gen_golden", "--only", tag, "--exact", "--out", args.out,
                   "--steps", str(args.steps), "--ma-steps", str(args.ma_steps), "--ma-noise", str(args.ma_noise),
                   "--cfg5-steps", str(args.cfg5_steps)]
            subprocess.check_call(cmd, cwd=ROOT, env=dict(os.environ, PYTHONHASHSEED="0"))
        return
    match = (lambda tag: tag == args.only) if args.exact else (lambda tag: args.only in tag)
    from oracle import refshim
    refshim.install()
    from metadrive.envs.metadrive_env import MetaDriveEnv
    from metadrive.envs.safe_metadrive_env import SafeMetaDriveEnv
    os.makedirs(args.out, exist_ok=True)
    rng = np.random.RandomState(0)
    rand_actions = rng.uniform(-1, 1, (args.steps, 2))
    smooth = np.stack([0.15 * np.sin(np.arange(args.steps) / 7.0), np.full(args.steps, 0.6)], 1)
    smooth_long = np.stack([0.15 * np.sin(np.arange(130) / 7.0), np.full(130, 0.6)], 1)
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
        # discrete actions (policy/env_input_policy.py:40-48): Discrete(7 x 5) index in column 0
        ("cfg1_S_discrete", MetaDriveEnv,
         dict(map="S", traffic_density=0.1, log_level=50, discrete_action=True, discrete_steering_dim=7, discrete_throttle_dim=5),
         0, np.stack([np.random.RandomState(4).randint(14, 35, args.steps), np.zeros(args.steps)], 1).astype(np.float64)),
        # lidar.num_others = 4 (component/sensors/lidar.py:93-138): the 4 nearest vehicles precede the lidar floats
        ("cfg2_pg3_seed11_others4", MetaDriveEnv,
         dict(map=3, traffic_density=0.3, num_scenarios=20, start_seed=0, log_level=50,
              vehicle_config=dict(lidar=dict(num_others=4))), 11, smooth),
        # ... with add_others_navi (lidar.py:120-129): every neighbour's two checkpoints in the ego frame follow its 4 floats
        ("cfg2_pg3_seed3_others_navi", MetaDriveEnv,
         dict(map=3, traffic_density=0.1, num_scenarios=20, start_seed=0, log_level=50,
              vehicle_config=dict(lidar=dict(num_others=4, add_others_navi=True))), 3, smooth),
        # side / lane-line detectors on (sensors/distance_detector.py:194-209): their rays replace the 2 + 1 floats
        ("cfg2_pg3_seed3_detectors", MetaDriveEnv,
         dict(map=3, traffic_density=0.1, num_scenarios=20, start_seed=0, log_level=50,
              vehicle_config=dict(side_detector=dict(num_lasers=32, distance=50),
                                  lane_line_detector=dict(num_lasers=16, distance=20))), 3, smooth),
        # BASELINE config 4: SafeMetaDriveEnv with static obstacles
        ("cfg4_safe_seed2", SafeMetaDriveEnv, dict(num_scenarios=20, start_seed=0, log_level=50), 2, smooth),
        ("cfg4_safe_seed5", SafeMetaDriveEnv, dict(num_scenarios=20, start_seed=0, log_level=50), 5, smooth),
        # non-terminal contacts, i.e. the contact response: seed 40 drives into the cones of an accident scene (step ~72),
        # seed 8 into a traffic vehicle (step ~81), and both keep going (crash_*_done=False, safe_metadrive_env.py:15-16)
        ("cfg4_safe_seed40_cones", SafeMetaDriveEnv, dict(num_scenarios=100, start_seed=0, log_level=50), 40, smooth_long),
        ("cfg4_safe_seed8_bump", SafeMetaDriveEnv, dict(num_scenarios=100, start_seed=0, log_level=50), 8, smooth_long),
        # a TollGate block inside a BIG map (map="S$C"): toll booths (TollGateBuilding: static boxes, crash_building, lidar-
        # visible, obstacles on their lane to the IDM traffic) in a single-agent env
        ("cfg2_StollC_seed0", MetaDriveEnv, dict(map="S$C", traffic_density=0.2, num_scenarios=20, start_seed=0, log_level=50), 0,
         np.zeros((230, 2))),   # closed loop: the lane-follow driver (see the loop below)
    ]
    # BASELINE config 3: MultiAgentRoundaboutEnv with 240-beam lidar
    #   cfg3_ma_roundabout          40 agents, random actions, respawn off (crash / out-of-road / wreck bookkeeping)
    #   cfg3_ma_roundabout_respawn  12 agents, lane-follow driver + noise, respawn on (arrivals, respawn, new routes)
    if match("cfg3_ma_roundabout") and args.only == "cfg3_ma_roundabout":
        from metadrive.envs.marl_envs.marl_inout_roundabout import MultiAgentRoundaboutEnv
        T = min(args.steps, 90)
        rs = np.random.RandomState(3)
        a = np.stack([0.2 * rs.uniform(-1, 1, (T, 41)), rs.uniform(0.0, 1.0, (T, 41))], -1)
        lid = dict(vehicle_config=dict(lidar=dict(num_lasers=240, distance=50, num_others=0)))
        cfg3 = dict(num_agents=40, allow_respawn=False, log_level=50, delay_done=25, horizon=1000, **lid)
        out = run_episode_ma(MultiAgentRoundaboutEnv, cfg3, a, "cfg3_ma_roundabout", obs_stride=4)
        path = os.path.join(args.out, "cfg3_ma_roundabout.npz")
        np.savez_compressed(path, **out)
        print("cfg3_ma_roundabout steps", len(out["reward"]), "seats", out["veh_f"].shape[1], "->",
              os.path.getsize(path) // 1024, "KiB", flush=True)
    if args.only == "cfg3_ma_roundabout_respawn":
        from metadrive.envs.marl_envs.marl_inout_roundabout import MultiAgentRoundaboutEnv
        lid = dict(vehicle_config=dict(lidar=dict(num_lasers=240, distance=50, num_others=0)))
        cfg3 = dict(num_agents=12, allow_respawn=True, log_level=50, delay_done=25, horizon=1000, **lid)
        out = run_episode_ma(MultiAgentRoundaboutEnv, cfg3, None, "cfg3_ma_roundabout_respawn", steps=max(args.ma_steps, 520),
                             noise=args.ma_noise, seed=5, obs_stride=3)
        path = os.path.join(args.out, "cfg3_ma_roundabout_respawn.npz")
        np.savez_compressed(path, **out)
        print("cfg3_ma_roundabout_respawn steps", len(out["reward"]), "respawns", int((out["respawn_draws"][:, 0] >= 0).sum()),
              "arrivals", int(((out["info_flags"] & 0x800) != 0).sum()), "->", os.path.getsize(path) // 1024, "KiB", flush=True)
    # the other shipped multi-agent map (envs/marl_envs/marl_intersection.py): same seats / respawn bookkeeping on the X block
    if args.only == "cfg3_ma_intersection_respawn":
        from metadrive.envs.marl_envs.marl_intersection import MultiAgentIntersectionEnv
        lid = dict(vehicle_config=dict(lidar=dict(num_lasers=240, distance=50, num_others=0)))
        cfgi = dict(num_agents=10, allow_respawn=True, log_level=50, delay_done=25, horizon=1000, **lid)
        out = run_episode_ma(MultiAgentIntersectionEnv, cfgi, None, "cfg3_ma_intersection_respawn", steps=300,
                             noise=args.ma_noise, seed=7, obs_stride=3)
        path = os.path.join(args.out, "cfg3_ma_intersection_respawn.npz")
        np.savez_compressed(path, **out)
        print("cfg3_ma_intersection_respawn steps", len(out["reward"]), "respawns", int((out["respawn_draws"][:, 0] >= 0).sum()),
              "arrivals", int(((out["info_flags"] & 0x800) != 0).sum()), "->", os.path.getsize(path) // 1024, "KiB", flush=True)
    # ... with the others block of the cooperative-MARL configs: the 4 nearest AGENTS and their checkpoints (lidar.num_others = 4,
    # add_others_navi, sensors/lidar.py:93-138) in front of the multi-agent default lidar (72 lasers, 40 m)
    if args.only == "cfg3_ma_intersection_others_navi":
        from metadrive.envs.marl_envs.marl_intersection import MultiAgentIntersectionEnv
        lid = dict(vehicle_config=dict(lidar=dict(num_lasers=72, distance=40, num_others=4, add_others_navi=True)))
        cfgi = dict(num_agents=8, allow_respawn=True, log_level=50, delay_done=25, horizon=1000, **lid)
        out = run_episode_ma(MultiAgentIntersectionEnv, cfgi, None, "cfg3_ma_intersection_others_navi", steps=220,
                             noise=args.ma_noise, seed=13, obs_stride=2)
        path = os.path.join(args.out, "cfg3_ma_intersection_others_navi.npz")
        np.savez_compressed(path, **out)
        print("cfg3_ma_intersection_others_navi steps", len(out["reward"]), "respawns", int((out["respawn_draws"][:, 0] >= 0).sum()),
              "arrivals", int(((out["info_flags"] & 0x800) != 0).sum()), "->", os.path.getsize(path) // 1024, "KiB", flush=True)
    # MultiAgentParkingLotEnv (envs/marl_envs/marl_parking_lot.py): agents born in parking spaces drive out, agents born on the
    # roads into the lot drive into a free space (enable_reverse on; white lines may be crossed).
    # Respawn is off: the reference's override (_respawn_single_vehicle, marl_parking_lot.py:230-236) calls vehicle.reset() WITHOUT the
    # drawn place's config, so every newborn lands on the first road's default pose (5, 0) with no destination - there is no
    # behaviour worth pinning there (DESIGN.md "Deliberate differences").
    # (ParkingLotSpawnManager indexes list(set) of Road objects, marl_parking_lot.py:66-67: the fixture is what PYTHONHASHSEED=0 gives,
    # which main() sets for every tag's process)
    if args.only == "cfg3_ma_parkinglot":
        from metadrive.envs.marl_envs.marl_parking_lot import MultiAgentParkingLotEnv
        lid = dict(vehicle_config=dict(lidar=dict(num_lasers=72, distance=40, num_others=0)))
        cfgp = dict(num_agents=7, allow_respawn=False, log_level=50, delay_done=25, horizon=1000, **lid)
        rs_p = np.random.RandomState(17)   # a slow lane-follow driver: the lot's arcs are a car length wide
        out = run_episode_ma(MultiAgentParkingLotEnv, cfgp, None, "cfg3_ma_parkinglot", steps=args.ma_steps if args.ma_steps != 450 else 1000,
                             noise=args.ma_noise, seed=17, obs_stride=2,
                             driver=lambda k, v: _parking_action(k, v, rs_p))
        path = os.path.join(args.out, "cfg3_ma_parkinglot.npz")
        np.savez_compressed(path, **out)
        print("cfg3_ma_parkinglot steps", len(out["reward"]), "arrivals", int(((out["info_flags"] & 0x800) != 0).sum()),
              "out of road", int(((out["info_flags"] & 0x400) != 0).sum()), "crashes", int(((out["info_flags"] & 0x1) != 0).sum()),
              "->", os.path.getsize(path) // 1024, "KiB", flush=True)
    # MultiAgentMetaDrive itself (envs/marl_envs/multi_agent_metadrive.py:12-62): the agents on the first road of a BIG-generated map
    # (map = 2 blocks, seed 0), all bound for the end of the last block
    if args.only == "cfg3_ma_pg3":
        from metadrive.envs.marl_envs.multi_agent_metadrive import MultiAgentMetaDrive
        lid = dict(vehicle_config=dict(lidar=dict(num_lasers=72, distance=40, num_others=0)))
        cfgm = dict(num_agents=6, allow_respawn=True, log_level=50, delay_done=25, horizon=1000, map=2, **lid)
        out = run_episode_ma(MultiAgentMetaDrive, cfgm, None, "cfg3_ma_pg3", steps=300, noise=args.ma_noise, seed=23, obs_stride=3)
        path = os.path.join(args.out, "cfg3_ma_pg3.npz")
        np.savez_compressed(path, **out)
        print("cfg3_ma_pg3 steps", len(out["reward"]), "respawns", int((out["respawn_draws"][:, 0] >= 0).sum()),
              "arrivals", int(((out["info_flags"] & 0x800) != 0).sum()), "->", os.path.getsize(path) // 1024, "KiB", flush=True)
    # a multi-agent env with IDM traffic (traffic_density > 0, trigger mode: the block's vehicles start when ANY agent enters the
    # trigger road, manager/traffic_manager.py:74-92)
    if args.only == "cfg3_ma_roundabout_traffic":
        from metadrive.envs.marl_envs.marl_inout_roundabout import MultiAgentRoundaboutEnv
        lid = dict(vehicle_config=dict(lidar=dict(num_lasers=72, distance=40, num_others=0)))
        cfgt = dict(num_agents=8, allow_respawn=True, log_level=50, delay_done=25, horizon=1000, traffic_density=0.15, **lid)
        out = run_episode_ma(MultiAgentRoundaboutEnv, cfgt, None, "cfg3_ma_roundabout_traffic", steps=300,
                             noise=args.ma_noise, seed=21, obs_stride=2)
        path = os.path.join(args.out, "cfg3_ma_roundabout_traffic.npz")
        np.savez_compressed(path, **out)
        print("cfg3_ma_roundabout_traffic steps", len(out["reward"]), "traffic", out["veh_f"].shape[1] - int(out["ma_alive_seats"][0]) - 1,
              "respawns", int((out["respawn_draws"][:, 0] >= 0).sum()), "crashes", int(((out["info_flags"] & 0x1) != 0).sum()),
              "->", os.path.getsize(path) // 1024, "KiB", flush=True)
    # MultiAgentBottleneckEnv (envs/marl_envs/marl_bottleneck.py): Merge / Split blocks, agents born at both ends without a
    # destination draw, 4-ray side / lane-line detectors in the observation, reward without the positive_road sign
    if args.only == "cfg3_ma_bottleneck_respawn":
        from metadrive.envs.marl_envs.marl_bottleneck import MultiAgentBottleneckEnv
        lid = dict(vehicle_config=dict(lidar=dict(num_lasers=240, distance=50, num_others=0)))
        cfgb = dict(num_agents=8, allow_respawn=True, log_level=50, delay_done=25, horizon=1000, **lid)
        out = run_episode_ma(MultiAgentBottleneckEnv, cfgb, None, "cfg3_ma_bottleneck_respawn", steps=300,
                             noise=args.ma_noise, seed=9, obs_stride=3)
        path = os.path.join(args.out, "cfg3_ma_bottleneck_respawn.npz")
        np.savez_compressed(path, **out)
        print("cfg3_ma_bottleneck_respawn steps", len(out["reward"]), "respawns", int((out["respawn_draws"][:, 0] >= 0).sum()),
              "arrivals", int(((out["info_flags"] & 0x800) != 0).sum()), "->", os.path.getsize(path) // 1024, "KiB", flush=True)
    # MultiAgentTollgateEnv (envs/marl_envs/marl_tollgate.py): Split -> TollGate -> Merge, toll booths (static boxes:
    # crash_building, lidar-visible) on every second toll lane, observation without the navigation block + 2 toll floats,
    # overspeed penalty inside the toll block, and the stay-time rule (less than min_pass_steps inside = out_of_road)
    if args.only == "cfg3_ma_tollgate_respawn":
        from metadrive.envs.marl_envs.marl_tollgate import MultiAgentTollgateEnv
        lid = dict(vehicle_config=dict(lidar=dict(num_lasers=72, distance=20, num_others=0)))
        cfgt = dict(num_agents=8, allow_respawn=True, log_level=50, delay_done=25, horizon=1000, **lid)
        rs_t = np.random.RandomState(21)
        waited = {}

        def driver(k, v):
            if v.navigation.current_road.block_ID() == "$":
                waited[k] = waited.get(k, 0) + 1
            if "x" not in waited:   # the toll block's extent, from its first lane
                lanes = [ln for blk in v.engine.current_map.blocks if blk.ID == "$"
                         for ln in blk.get_socket(0).positive_road.get_lanes(v.engine.current_map.road_network)]
                waited["x"] = sorted([lanes[0].position(0, 0)[0], lanes[0].position(lanes[0].length, 0)[0]])
            return _toll_action(v, rs_t, args.ma_noise, patient=int(k[5:]) % 3 != 1, waited=waited.get(k, 0), toll_x=waited["x"])

        out = run_episode_ma(MultiAgentTollgateEnv, cfgt, None, "cfg3_ma_tollgate_respawn", steps=420, noise=args.ma_noise, seed=13,
                             obs_stride=2, driver=driver)
        path = os.path.join(args.out, "cfg3_ma_tollgate_respawn.npz")
        np.savez_compressed(path, **out)
        fl = out["info_flags"]
        print("cfg3_ma_tollgate_respawn steps", len(out["reward"]), "respawns", int((out["respawn_draws"][:, 0] >= 0).sum()),
              "arrivals", int(((fl & 0x800) != 0).sum()), "crash_building", int(((fl & 0x4) != 0).sum()),
              "out_of_road", int(((fl & 0x400) != 0).sum()), "toll obs steps", int((out["obs"][:, :, -2] > 0).sum()),
              "stayed", int((out["obs"][:, :, -1] > 0).sum()), "->", os.path.getsize(path) // 1024, "KiB", flush=True)
    # BASELINE config 5 (composed): X map, respawn-mode IDM traffic, 16 crossing pedestrians; crashes do not end the
    # episode here so that the trace keeps running through pedestrian / vehicle contacts
    if args.only == "cfg5_ped_X":
        cfg5 = dict(map="X", traffic_density=0.1, traffic_mode="respawn", num_scenarios=20, start_seed=0, log_level=50,
                    crash_vehicle_done=False, crash_human_done=False, crash_object_done=False)
        out = run_episode_cfg5(cfg5, 4, args.cfg5_steps, "cfg5_ped_X")
        path = os.path.join(args.out, "cfg5_ped_X.npz")
        np.savez_compressed(path, **out)
        print("cfg5_ped_X steps", len(out["reward"]), "vehicles", out["veh_f"].shape[1], "respawns", len(out["respawn_events"]),
              "crash_human steps", int(((out["veh_i"][:, 0, 5] & 8) != 0).sum()), "->", os.path.getsize(path) // 1024, "KiB", flush=True)
    for tag, cls, cfg, seed, acts in cases:
        if not match(tag):
            continue
        out = run_episode(cls, cfg, seed, acts, tag, closed_loop=np.random.RandomState(31) if tag == "cfg2_StollC_seed0" else None)
        path = os.path.join(args.out, tag + ".npz")
        np.savez_compressed(path, **out)
        print(tag, "steps", len(out["reward"]), "vehicles", out["veh_f"].shape[1], "objects",
              len(out["init_objects"]), "->", os.path.getsize(path) // 1024, "KiB", flush=True)


if __name__ == "__main__":
    main()

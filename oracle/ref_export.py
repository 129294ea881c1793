"""Export maps / scenarios / per-step traces from the reference running under `oracle.refshim`.

TEST INFRASTRUCTURE, in-container only (needs /root/reference).  The arrays written here are the
golden fixtures under `tests/golden/` and the map library under `metadrive_ped_b200/assets/`.
Layouts are documented in `metadrive_ped_b200/scene.py` (the consumer).

Reference objects read (all attribute reads, no behaviour changed):
  road network graph        `metadrive/component/road_network/node_road_network.py:68-87`
  lanes                     `metadrive/component/lane/straight_lane.py:12-47`, `circular_lane.py:12-51`
  blocks / sockets          `metadrive/component/pgblock/pg_block.py:19-145`
  traffic roster            `metadrive/manager/traffic_manager.py:211-277`
  navigation state          `metadrive/component/navigation_module/node_network_navigation.py:75-128`
  IDM policy state          `metadrive/policy/idm_policy.py:224-233`
"""
import json

import numpy as np

LINE_NONE, LINE_BROKEN, LINE_CONTINUOUS, LINE_SIDE, LINE_GUARDRAIL = 0, 1, 2, 3, 4
VEHICLE_TYPES = ["s", "m", "l", "xl", "default", "static_default", "varying_dynamics"]


def _line_type_id(t):
    from metadrive.constants import PGLineType
    return {
        PGLineType.NONE: LINE_NONE,
        PGLineType.BROKEN: LINE_BROKEN,
        PGLineType.CONTINUOUS: LINE_CONTINUOUS,
        PGLineType.SIDE: LINE_SIDE,
        PGLineType.GUARDRAIL: LINE_GUARDRAIL
    }[t]


def _is_yellow(c):
    from metadrive.constants import PGLineColor
    return 1 if tuple(c) == tuple(PGLineColor.YELLOW) else 0


class MapIndex:
    """Stable integer ids for nodes / roads / lanes of one reference map (graph insertion order)."""
    def __init__(self, road_network):
        from metadrive.constants import Decoration
        self.nodes = {}
        self.roads = {}
        self.lanes = {}
        self.lane_objs = []
        self.road_list = []
        g = road_network.graph
        for _from, to_dict in g.items():
            # "decoration" lanes (ramps) are real lane surfaces + lines in the static world; keep them as a road
            for _to, lanes in to_dict.items():
                for n in (_from, _to):
                    if n not in self.nodes:
                        self.nodes[n] = len(self.nodes)
                self.roads[(_from, _to)] = len(self.road_list)
                self.road_list.append((_from, _to, len(self.lane_objs), len(lanes)))
                for i, lane in enumerate(lanes):
                    self.lanes[(_from, _to, i)] = len(self.lane_objs)
                    self.lane_objs.append(lane)

    def lane_id(self, lane):
        if lane is None:
            return -1
        return self.lanes.get(tuple(lane.index), -1)


def export_map(current_map):
    """Lane / road / block tables of a reference PGMap. Returns (dict of arrays, MapIndex)."""
    from metadrive.component.lane.straight_lane import StraightLane
    from metadrive.component.lane.circular_lane import CircularLane
    from metadrive.component.road_network import Road
    rn = current_map.road_network
    mi = MapIndex(rn)
    L = len(mi.lane_objs)
    lane_f = np.zeros((L, 10), dtype=np.float64)
    lane_i = np.zeros((L, 8), dtype=np.int32)
    for (f, t, i), lid in mi.lanes.items():
        lane = mi.lane_objs[lid]
        if isinstance(lane, StraightLane):
            lane_f[lid] = [0, lane.width, lane.length, lane.start[0], lane.start[1], lane.end[0], lane.end[1], 0, 0, 0]
        elif isinstance(lane, CircularLane):
            lane_f[lid] = [
                1, lane.width, lane.length, lane.center[0], lane.center[1], lane.radius, lane.start_phase,
                lane.end_phase, lane.direction, lane.angle
            ]
        else:
            raise TypeError(type(lane))
        lane_i[lid] = [
            mi.roads[(f, t)], i, mi.nodes[f], mi.nodes[t],
            _line_type_id(lane.line_types[0]),
            _line_type_id(lane.line_types[1]),
            _is_yellow(lane.line_colors[0]),
            _is_yellow(lane.line_colors[1])
        ]
    road_i = np.zeros((len(mi.road_list), 6), dtype=np.int32)
    for k, (f, t, first, n) in enumerate(mi.road_list):
        r = Road(f, t)
        bid = r.block_ID()
        road_i[k] = [mi.nodes[f], mi.nodes[t], first, n, int(r.is_negative_road()), ord(bid[0])]
    # blocks: id char, trigger road, spawn lanes (in the order the traffic manager walks them), respawn roads, sockets
    blocks = []
    for b in current_map.blocks:
        trig = b.pre_block_socket.positive_road if b.pre_block_socket is not None else None
        spawn_lanes = []
        if b is not current_map.blocks[0]:
            for lanes in b.get_intermediate_spawn_lanes():
                spawn_lanes.append([mi.lane_id(l) for l in lanes])
        neg_lanes = []
        if b.ID in ["S", "C", "r", "R"]:
            for lanes in b.block_network.get_negative_lanes():
                neg_lanes.append([mi.lane_id(l) for l in lanes])
        sockets = []
        for s in b.get_socket_list():
            sockets.append(
                [
                    mi.roads.get((s.positive_road.start_node, s.positive_road.end_node), -1),
                    mi.roads.get((s.negative_road.start_node, s.negative_road.end_node), -1)
                    if s.negative_road is not None else -1
                ]
            )
        blocks.append(
            dict(
                id=b.ID,
                trigger_road=mi.roads.get((trig.start_node, trig.end_node), -1) if trig is not None else -1,
                spawn_lanes=spawn_lanes,
                negative_lanes=neg_lanes,
                respawn_roads=[mi.roads.get((r.start_node, r.end_node), -1) for r in b.get_respawn_roads()],
                sockets=sockets,
            )
        )
    meta = dict(nodes=list(mi.nodes.keys()), blocks=blocks, respawn=export_respawn_lanes(current_map, mi))
    return dict(lane_f=lane_f, lane_i=lane_i, road_i=road_i, meta=json.dumps(meta)), mi


def export_respawn_lanes(current_map, mi):
    """Respawn / hybrid traffic mode (manager/traffic_manager.py:94-122, 279-296): the lanes a removed traffic vehicle
    may be respawned on, each with the route the reference's navigation assigns to a vehicle born there
    (node_network_navigation.py:43-91: destination drawn from RandomState(global seed), then the BFS route)."""
    from metadrive.component.navigation_module.node_network_navigation import NodeNetworkNavigation
    from metadrive.engine.engine_utils import get_engine
    roads = []
    for block in current_map.blocks:
        for road in block.get_respawn_roads():
            if road in roads:
                roads.remove(road)
            else:
                roads.append(road)
    seed = get_engine().global_random_seed
    out = []
    for road in roads:
        for lane in road.get_lanes(current_map.road_network):
            try:
                dest = NodeNetworkNavigation.auto_assign_task(current_map, lane.index, None, seed)
                ck = current_map.road_network.shortest_path(lane.index, dest)
            except Exception:
                continue
            if len(ck) <= 2:
                ck = [lane.index[0], lane.index[1]]
            out.append(dict(lane=mi.lane_id(lane), route=[mi.nodes[c] for c in ck]))
    return out


def export_static_bodies(engine):
    """Ground truth of what the reference put in its two Bullet worlds (for validating the product's own
    derivation of line boxes / lane hulls / sidewalk strips from the lane table)."""
    from oracle.refshim import pbullet as pb
    lines, hulls, sidewalks = [], [], []
    for b in engine.physics_world.static_world.bodies:
        name = b.getName()
        for shape, ts in b.shapes:
            if isinstance(shape, pb.BulletBoxShape) and name.startswith("ROAD_LINE"):
                ang = np.arctan2(b.mat[1, 0], b.mat[0, 0])
                lines.append([b.pos[0], b.pos[1], ang, shape.half[0], _line_name_id(name)])
            elif isinstance(shape, pb.BulletConvexHullShape):
                hulls.append((tuple(b.base_object_name), shape.hull2d + b.pos[:2]))
    for b in engine.physics_world.dynamic_world.bodies:
        for shape, ts in b.shapes:
            if isinstance(shape, pb.BulletTriangleMeshShape) and shape.polygon is not None:
                sidewalks.append(shape.polygon + b.pos[:2])
    return dict(lines=np.array(lines), hulls=hulls, sidewalks=sidewalks)


def _line_name_id(name):
    from metadrive.constants import MetaDriveType as T
    return {
        T.LINE_SOLID_SINGLE_WHITE: 0,
        T.LINE_SOLID_SINGLE_YELLOW: 1,
        T.LINE_BROKEN_SINGLE_WHITE: 2,
        T.LINE_BROKEN_SINGLE_YELLOW: 3
    }[name]


def vehicle_type_id(v):
    from metadrive.component.vehicle import vehicle_type as vt
    cls = type(v)
    for k, c in vt.vehicle_type.items():
        if c is cls:
            return VEHICLE_TYPES.index(k)
    if isinstance(v, vt.DefaultVehicle):
        return VEHICLE_TYPES.index("default")
    raise KeyError(cls)


def vehicle_static(v):
    """[type, length, width, height, mass, tire_radius, lateral, front_wb, rear_wb, chassis_to_axis,
        max_engine_force, max_brake_force, max_steering, wheel_friction, max_speed_km_h, enable_reverse]"""
    return np.array(
        [
            vehicle_type_id(v), v.LENGTH, v.WIDTH, v.HEIGHT, v.MASS, v.TIRE_RADIUS, v.LATERAL_TIRE_TO_CENTER,
            v.FRONT_WHEELBASE, v.REAR_WHEELBASE, v.CHASSIS_TO_WHEEL_AXIS, v.config["max_engine_force"],
            v.config["max_brake_force"], v.config["max_steering"], v.config["wheel_friction"],
            v.config["max_speed_km_h"],
            float(bool(v.config["enable_reverse"]))
        ],
        dtype=np.float64
    )


N_DYN = 14  # pos3 quat4(w,x,y,z) vel3 angvel3 static


def vehicle_dynamic(v):
    from oracle.refshim.pcore import mat_to_quat
    n = v.origin.node()
    q = mat_to_quat(n.mat)
    return np.concatenate([n.pos, q, n.lin_vel, n.ang_vel, [float(n.static)]])


class Roster:
    """Fixed slot order for one episode: agents first (insertion order), then traffic vehicles in spawn order
    (trigger mode: block order as spawned), then static traffic objects."""
    def __init__(self, env, mi):
        eng = env.engine
        self.mi = mi
        self.agents = list(eng.agent_manager.active_agents.values())
        tm = getattr(eng, "traffic_manager", None)
        self.traffic = []
        self.trigger_block = []
        if tm is not None:
            for v in tm._traffic_vehicles:
                self.traffic.append(v)
                self.trigger_block.append(-1)
            # block_triggered_vehicles was reversed after creation: last element = first block after the first
            for k, bv in enumerate(reversed(tm.block_triggered_vehicles)):
                for name in bv.vehicles:
                    self.traffic.append(eng.get_objects([name])[name])
                    self.trigger_block.append(k + 1)
        om = getattr(eng, "object_manager", None)
        self.objects = list(om.spawned_objects.values()) if om is not None else []
        # the broken-down car of a break-down scene (manager/object_manager.py:95-102) is a vehicle body in the world that
        # no manager drives: it closes the roster with trigger block 0 (never triggered)
        from metadrive.component.vehicle.base_vehicle import BaseVehicle
        for o in self.objects:
            if isinstance(o, BaseVehicle):
                self.traffic.append(o)
                self.trigger_block.append(0)
        self.vehicles = self.agents + self.traffic
        # TollGateBuilding boxes a TollGate block spawns itself (component/pgblock/tollgate.py:64-76): static bodies of the world
        # that no manager owns; they close the object table
        self.buildings = [o for b in env.current_map.blocks for o in (getattr(b, "_block_objects", None) or [])
                          if type(o).__name__ == "TollGateBuilding"]

    def static_table(self):
        return np.stack([vehicle_static(v) for v in self.vehicles])

    def routes(self, max_ckpt=24):
        out = np.full((len(self.vehicles), max_ckpt), -1, dtype=np.int32)
        for k, v in enumerate(self.vehicles):
            ck = v.navigation.checkpoints
            assert len(ck) <= max_ckpt, len(ck)
            out[k, :len(ck)] = [self.mi.nodes[c] for c in ck]
        return out

    def objects_table(self):
        """[kind(0 cone,1 warning,2 barrier,4 building), x, y, heading, half_len_or_radius, half_width, height, lane_id]"""
        from metadrive.component.static_object.traffic_object import TrafficCone, TrafficWarning, TrafficBarrier
        rows = []
        for o in self.objects:
            if isinstance(o, TrafficCone):
                rows.append([0, o.position[0], o.position[1], o.heading_theta, o.RADIUS, o.RADIUS, o.HEIGHT])
            elif isinstance(o, TrafficWarning):
                rows.append([1, o.position[0], o.position[1], o.heading_theta, o.RADIUS, o.RADIUS, o.HEIGHT])
            elif isinstance(o, TrafficBarrier):
                rows.append([2, o.position[0], o.position[1], o.heading_theta, o.LENGTH / 2, o.WIDTH / 2, o.HEIGHT])
            else:
                continue
            rows[-1].append(self.mi.lane_id(o.lane))
        # kind 4: BulletBoxShape((BUILDING_LENGTH / 2, lane width / 2, BUILDING_HEIGHT / 2)) centred at z = 0
        # (utils/pg/utils.py:315, buildings/tollgate_building.py:14-26).  Columns 4 / 5 follow the barrier's convention: half
        # extent ACROSS the heading, then ALONG it (the barrier's box is (WIDTH / 2, LENGTH / 2, .), traffic_object.py:138)
        for o in self.buildings:
            rows.append([4, o.position[0], o.position[1], o.heading_theta, o.WIDTH / 2, o.LENGTH / 2, o.BUILDING_HEIGHT,
                         self.mi.lane_id(o.lane)])
        return np.array(rows, dtype=np.float64).reshape(-1, 8)


N_STEP_F = N_DYN + 2 + 2 + 2 + 2 + 10 + 8  # see record_vehicle
FLAG_NAMES = [
    "crash_vehicle", "crash_object", "crash_building", "crash_human", "crash_sidewalk", "on_white_continuous_line",
    "on_yellow_continuous_line", "on_broken_line", "on_lane", "out_of_route"
]


def record_vehicle(v, roster, env):
    """float row: dyn(14) | steering, throttle | last_pos(2) | dist_left, dist_right | speed_kmh, heading_theta
                  | navi(10) | idm: timer, target_speed, hpid(last_err, int), lpid(last_err, int), valid, spare
       int row  : alive, active, lane_id, ckpt0, ckpt1, flags bitmask, routing_lane_id, ref_road"""
    eng = env.engine
    mi = roster.mi
    alive = v.name in eng._spawned_objects  # engine.get_objects raises KeyError for a cleared name
    f = np.zeros(N_STEP_F)
    i = np.zeros(8, dtype=np.int32)
    if not alive:
        return f, i
    tm = getattr(eng, "traffic_manager", None)
    active = (v.name in eng.agent_manager._active_objects) or (tm is not None and v in tm._traffic_vehicles)
    f[:N_DYN] = vehicle_dynamic(v)
    f[14] = v.steering
    f[15] = v.throttle_brake
    f[16:18] = v.last_position
    f[18] = v.dist_to_left_side if v.dist_to_left_side is not None else 0
    f[19] = v.dist_to_right_side if v.dist_to_right_side is not None else 0
    f[20] = v.speed_km_h
    f[21] = v.heading_theta
    f[22:32] = v.navigation.get_navi_info()
    pol = eng.get_policy(v.name)
    if pol is not None and hasattr(pol, "overtake_timer"):
        f[32] = pol.overtake_timer
        f[33] = pol.target_speed
        f[34] = pol.heading_pid.p_error
        f[35] = pol.heading_pid.i_error
        f[36] = pol.lateral_pid.p_error
        f[37] = pol.lateral_pid.i_error
        f[38] = 1.0
        i[6] = mi.lane_id(pol.routing_target_lane)
    flags = 0
    for b, nme in enumerate(FLAG_NAMES):
        if getattr(v, nme):
            flags |= 1 << b
    nav = v.navigation
    i[0] = 1
    i[1] = int(active)
    i[2] = mi.lane_id(nav.current_lane)
    i[3], i[4] = nav._target_checkpoints_index
    i[5] = flags
    cr = nav.current_road
    i[7] = mi.roads.get((cr.start_node, cr.end_node), -1)
    return f, i


def record_world(env, roster):
    fs, is_ = [], []
    for v in roster.vehicles:
        f, i = record_vehicle(v, roster, env)
        fs.append(f)
        is_.append(i)
    return np.stack(fs), np.stack(is_)

"""Reset-time scene population on the product side: ego, IDM traffic, accident scenes - what the reference's managers
spawn at `env.reset(seed)` on a map from `pgmap`, with the reference's seeded streams.

Restated (paths relative to /root/reference/metadrive):
  seeding         engine/base_engine.py:546-553: the engine and EVERY manager get their own RandomState(seed); every
                  spawned object takes its seed from the ENGINE's stream (engine/base_engine.py:123-134,
                  base_class/randomizable.py:19-20)
  manager order   map (0) -> object manager (9) -> agent manager (10) -> traffic manager (10): engine/base_engine.py:533-544,
                  envs/base_env.py:744-746, envs/metadrive_env.py:286-289
  accident scenes manager/object_manager.py:40-151 (cones / breakdown car + warning / barrier)
  ego             manager/agent_manager.py:40-57, 105-113; component/vehicle/base_vehicle.py:273-343
  traffic         manager/traffic_manager.py:211-296 (trigger / hybrid: per-block vehicles; respawn: per respawn lane)
  vehicle types   component/vehicle/vehicle_type.py (dimensions), component/pg_space.py:226-272 (sampled parameters)
  routes          component/navigation_module/node_network_navigation.py:43-128 (destination socket drawn from a fresh
                  RandomState(seed), BFS route), road_network/node_road_network.py:233-258
  IDM state       policy/idm_policy.py:224-233

Output: a `scene.Scenario` (vehicle parameters, poses, routes, checkpoint indices, IDM timers, obstacles) in the layout of
the scenario libraries; `tests/test_pgmap.py` regenerates the libraries' rosters from the seeds and compares.
The broken-down car of a break-down scene (object_manager.py:95-102) is part of the world (alive, never active) and
closes the roster, as in the exported libraries and fixtures of round 2 (`include_breakdown=False` gives the round-1
rosters, which had left it out).
"""
import math

import numpy as np

from . import pgmap as pg
from .scene import ROUTE_MAX, Scenario

VEHICLE_TYPES = ["s", "m", "l", "xl", "default", "static_default", "varying_dynamics"]
# component/vehicle/vehicle_type.py:8-165; component/vehicle/base_vehicle.py:98 (CHASSIS_TO_WHEEL_AXIS default 0.2)
#        length width height mass tire_r lateral front_wb rear_wb chassis_to_axis
DIMS = {
    "s": (4.3, 1.70, 1.70, 800, 0.376, 0.7, 1.385, 1.11, 0.2),
    "m": (4.6, 1.85, 1.37, 1200, 0.39, 0.803, 1.285, 1.203, 0.2),
    "l": (4.87, 2.046, 1.85, 1300, 0.429, 0.75, 1.5301, 1.218261, 0.2),
    "xl": (5.74, 2.3, 2.8, 1600, 0.37, 0.931, 1.726, 1.075, 0.3),
    "default": (4.515, 1.852, 1.19, 1100, 0.313, 0.815, 1.05234, 1.4166, 0.2),
}
DIMS["static_default"] = DIMS["default"]
# VehicleParameterSpace (component/pg_space.py:226-272).  BoxSpace is namedtuple("BoxSpace", "max min") and the vehicle
# spaces are written positionally - BoxSpace(750, 850) is max = 750, min = 850 - so the uniform runs from the larger bound
# DOWN to the smaller one: low + (high - low) * u with low = 850, high = 750.
def _vbox(first, second):
    return pg.box(second, first)


SPACES = {
    "default": dict(wheel_friction=pg.const(0.9), max_engine_force=_vbox(750, 850), max_brake_force=_vbox(80, 180),
                    max_steering=pg.const(40), max_speed_km_h=pg.const(80)),
    "static_default": dict(wheel_friction=pg.const(0.9), max_engine_force=pg.const(800), max_brake_force=pg.const(150),
                           max_steering=pg.const(40), max_speed_km_h=pg.const(80)),
    "s": dict(wheel_friction=pg.const(0.9), max_engine_force=_vbox(350, 550), max_brake_force=_vbox(35, 80),
              max_steering=pg.const(50), max_speed_km_h=pg.const(80)),
    "m": dict(wheel_friction=pg.const(0.75), max_engine_force=_vbox(650, 850), max_brake_force=_vbox(60, 150),
              max_steering=pg.const(45), max_speed_km_h=pg.const(80)),
    "l": dict(wheel_friction=pg.const(0.8), max_engine_force=_vbox(450, 650), max_brake_force=_vbox(60, 120),
              max_steering=pg.const(40), max_speed_km_h=pg.const(80)),
    "xl": dict(wheel_friction=pg.const(0.7), max_engine_force=_vbox(500, 700), max_brake_force=_vbox(50, 100),
               max_steering=pg.const(35), max_speed_km_h=pg.const(80)),
}
TRAFFIC_TYPES, TRAFFIC_P = ["s", "m", "l", "xl", "default"], [0.2, 0.3, 0.3, 0.2, 0.0]   # traffic_manager.py:298-301
VEHICLE_GAP = 10                                                                        # traffic_manager.py:27
MAX_RAND_INT = 65536                                                                    # randomizable.py:10
# TrafficCone / TrafficWarning / TrafficBarrier (component/static_object/traffic_object.py:43-177):
# kind, half length (or radius), half width (or radius), height
OBJ_DIMS = {"cone": (0, 0.2, 0.2, 2.0), "warning": (1, 0.5, 0.5, 1.2), "barrier": (2, 2.0 / 2, 0.3 / 2, 2.0)}


class MapIndex:
    """ids of nodes / roads / lanes in graph insertion order (the numbering of pgmap.to_tables)"""
    def __init__(self, big):
        self.big = big
        self.nodes, self.roads, self.lane_ids, self.lane_index = {}, {}, {}, {}
        n = 0
        for a, d in big.world.graph.items():
            for b, lanes in d.items():
                for node in (a, b):
                    if node not in self.nodes:
                        self.nodes[node] = len(self.nodes)
                self.roads[(a, b)] = len(self.roads)
                for i, lane in enumerate(lanes):
                    self.lane_ids[id(lane)] = n
                    self.lane_index[id(lane)] = (a, b, i)
                    n += 1

    def lane_id(self, lane):
        return self.lane_ids.get(id(lane), -1)


def shortest_path(net, start_node, goal):
    return next(net.bfs_paths(start_node, goal), [])


def choose_destination(big, lane_index, seed):
    """NodeNetworkNavigation.auto_assign_task (node_network_navigation.py:72-91): a socket of the last block (of the first
    one for a vehicle born on a negative road), drawn from a fresh RandomState(seed)."""
    start = lane_index[0]
    negative = pg.is_negative(lane_index[:2])
    block = big.blocks[0] if negative else big.blocks[-1]
    sockets = list(block.sockets.values())
    s = sockets[pg.seeded_rng(seed).choice(len(sockets))]
    on_socket = start in (s.positive[0], s.positive[1], s.negative[0], s.negative[1])
    if on_socket and len(sockets) != 1:
        raise ValueError("Can not set a destination!")   # the reference's loop ends in a ValueError here as well
    return s.negative[1] if negative else s.positive[1]


def route_for(big, lane_index, seed, destination=None):
    """set_route (node_network_navigation.py:93-128): checkpoints + the initial target-checkpoint indices"""
    dest = destination if destination is not None else choose_destination(big, lane_index, seed)
    ck = shortest_path(big.world, lane_index[0], dest)
    idx = [0, 1]
    if len(ck) <= 2:
        ck, idx = [lane_index[0], lane_index[1]], [0, 0]
    return ck, idx


def yaw_quat(heading):
    """orientation of a body whose nose (+Y) points along `heading` (base_vehicle.py:994-1000: the vehicle frame is 90
    degrees off; base_object.py:371-380 hands panda3d degrees), as (w, x, y, z) with the sign convention of a
    rotation-matrix -> quaternion conversion (w >= 0 while the trace is positive), which the exported libraries carry"""
    yaw = math.radians((heading - np.pi / 2) * 180 / np.pi)
    c, s_ = math.cos(yaw), math.sin(yaw)
    t = c + c + 1.0
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        return np.array([0.25 * s, 0.0, 0.0, (s_ - (-s_)) / s])
    s = math.sqrt(1.0 + 1.0 - c - c) * 2
    return np.array([(s_ - (-s_)) / s, 0.0, 0.0, 0.25 * s])


class Spawner:
    def __init__(self, big, seed, lane_num=3, lane_width=3.5):
        self.big, self.seed, self.lane_num, self.lane_width = big, int(seed), lane_num, lane_width
        self.idx = MapIndex(big)
        self.engine_rng = pg.seeded_rng(seed)
        self.object_rng, self.agent_rng, self.traffic_rng = pg.seeded_rng(seed), pg.seeded_rng(seed), pg.seeded_rng(seed)
        self.static, self.dyn, self.routes, self.ints, self.idm = [], [], [], [], []
        self.objects = []
        self.accident_lanes = []

    def engine_seed(self):
        return int(self.engine_rng.randint(0, MAX_RAND_INT))

    # ------------------------------------------------------------------ vehicles
    def vehicle(self, model, lane, lon, kind, trigger, active, policy_rng=None, lat=0.0, enable_reverse=False,
                is_static=False):
        """spawn_object(vehicle class, vehicle_config) + BaseVehicle.reset: the seed comes from the engine's stream, the
        class parameters from the vehicle's own stream, the pose from the spawn lane"""
        seed = self.engine_seed()
        rng = pg.seeded_rng(seed)
        par = pg.sample_space(SPACES[model], int(rng.randint(low=0, high=int(1e6))))
        d = DIMS[model]
        self.static.append([VEHICLE_TYPES.index(model), d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], d[8],
                            par["max_engine_force"], par["max_brake_force"], par["max_steering"], par["wheel_friction"],
                            par["max_speed_km_h"], float(enable_reverse)])
        p = lane.position(lon, lat)
        q = yaw_quat(lane.heading_at(lon))
        self.dyn.append([p[0], p[1], d[2] / 2, q[0], q[1], q[2], q[3], 0, 0, 0, 0, 0, 0, float(is_static)])
        li = self.idx.lane_index[id(lane)]
        ck, ci = route_for(self.big, li, self.seed)
        assert len(ck) <= ROUTE_MAX, len(ck)
        row = np.full(ROUTE_MAX, -1, np.int32)
        row[:len(ck)] = [self.idx.nodes[c] for c in ck]
        self.routes.append(row)
        self.ints.append([kind, trigger, self.idx.lane_id(lane), ci[0], ci[1], int(active)])
        timer = int(pg.seeded_rng(policy_rng).randint(0, 50)) if policy_rng is not None else 0   # idm_policy.py:229
        self.idm.append([timer, 30])

    def ego(self, random_lane=True, random_model=False, model="default"):
        """random_spawn_lane_in_single_agent + _create_agents (agent_manager.py:40-57, 105-113): lane of the first road
        drawn from the agent manager's stream, 5 m in (envs/base_env.py:139-140); with random_agent_model the class is
        drawn next, uniformly over the five types (vehicle_type.py:168-181)"""
        lane_i = int(self.agent_rng.randint(self.lane_num)) if random_lane else 0
        lane = self.big.world.graph[">"][">>"][lane_i]
        if random_model:
            model = str(self.agent_rng.choice(TRAFFIC_TYPES, p=[1 / 5 for _ in range(5)]))
        self.vehicle(model, lane, 5.0, 1, -1, True)

    def random_traffic_type(self):
        return str(self.traffic_rng.choice(TRAFFIC_TYPES, p=TRAFFIC_P))

    # ------------------------------------------------------------------ traffic (manager/traffic_manager.py:211-296)
    def traffic_trigger(self, density, inverse=False):
        for b, block in enumerate(self.big.blocks[1:], 1):
            trigger_lanes = block.spawn_lanes()
            if inverse and block.ID in "SCrR":      # need_inverse_traffic (traffic_manager.py:221-224)
                neg = block.net.negative_lanes()
                self.traffic_rng.shuffle(neg)
                trigger_lanes += neg
            cand = []
            for lanes in trigger_lanes:
                for l in lanes:
                    if any(l is a for a in self.accident_lanes):
                        continue
                    cand += [(l, i * VEHICLE_GAP) for i in range(int(l.length / VEHICLE_GAP))]
            total_length = sum([lane.length for lanes in trigger_lanes for lane in lanes])
            n = int(math.floor(int(math.floor(total_length / VEHICLE_GAP)) * density))
            order = list(range(len(cand)))
            self.traffic_rng.shuffle(order)
            for k in order[:min(n, len(cand))]:
                model = self.random_traffic_type()
                lane, lon = cand[k]
                self.vehicle(model, lane, lon, 2, b, False, policy_rng=int(self.traffic_rng.randint(0, MAX_RAND_INT)))

    def respawn_lanes(self):
        roads = []
        for block in self.big.blocks:
            for road in block.respawn_roads:
                if road in roads:
                    roads.remove(road)
                else:
                    roads.append(road)
        return [l for road in roads for l in self.big.world.lanes(road)]

    def traffic_respawn(self, density):
        for lane in self.respawn_lanes():
            longs = [i * VEHICLE_GAP for i in range(int(lane.length / VEHICLE_GAP))]
            self.traffic_rng.shuffle(longs)
            for lon in longs[:int(np.ceil(density * len(longs)))]:
                model = self.random_traffic_type()
                self.vehicle(model, lane, lon, 2, -1, True, policy_rng=int(self.traffic_rng.randint(0, MAX_RAND_INT)))

    # ------------------------------------------------------------------ accident scenes (manager/object_manager.py:40-151)
    def obstacle(self, name, lane, lon, lat):
        self.engine_seed()                                   # every spawned object takes a seed from the engine's stream
        kind, a, b, h = OBJ_DIMS[name]
        p = lane.position(lon, lat)
        h_lane = lane.heading_at(lon)
        heading = math.atan2(math.sin(h_lane), math.cos(h_lane))     # BaseObject.heading_theta reads the angle back: (-pi, pi]
        self.objects.append([kind, p[0], p[1], heading, a, b, h, self.idx.lane_id(lane)])

    def accidents(self, prob, include_breakdown=True):
        rng = self.object_rng
        if abs(prob - 0.0) < 1e-2:
            return
        for block in self.big.blocks:
            if block.ID not in "SCrR":
                continue
            if rng.rand() > prob:
                continue
            road_1 = (block.pre_socket.positive[1], block.node(0, 0))
            road_2 = (block.node(0, 0), block.node(0, 1)) if block.ID != "S" else None
            ramp = block.ID in "rR"
            if rng.rand() > 0.67:
                road = [road_1, road_2][rng.choice(2)] if block.ID != "C" else road_2
                road = road_1 if road is None else road
                on_left = True if rng.rand() > 0.5 or (road is road_2 and ramp) else False
                lane = self.big.world.lanes(road)[0 if on_left else -1]
                self.accident_lanes.append(lane)
                self.cones(lane, lane.length - 10 - 5, self.lane_width, on_left)
            else:
                road = [road_1, road_2][rng.choice(2)]
                road = road_1 if road is None else road
                on_left = True if rng.rand() > 0.5 or (road is road_2 and ramp) else False
                lanes = self.big.world.lanes(road)
                if len(lanes) - 1 == 0:
                    k = -1
                else:
                    k = int(rng.randint(0, len(lanes) - 1)) if on_left else -1
                lane = lanes[k]
                self.accident_lanes.append(lane)
                lon = rng.rand() * lane.length / 2 + lane.length / 2
                if rng.rand() > 0.5:     # break-down scene: a stopped car and a warning tripod 10 m behind it
                    model = self.random_traffic_type()
                    if include_breakdown:
                        self.vehicle(model, lane, float(lon), 2, 0, False)
                    else:
                        self.engine_seed()
                    self.obstacle("warning", lane, lon - 10, 0)
                else:
                    self.obstacle("barrier", lane, lon, 0)

    def cones(self, lane, lon0, lateral_len, on_left):
        lat_num = int(lateral_len / 1)
        lon_num = int(10 / 2)
        lat_1 = [lat * 1 for lat in range(lat_num)]
        lat_2 = [lat_num * 1] * (lon_num + 1)
        lat_3 = [(lat_num - lat - 1) * 1 for lat in range(int(lat_num))]
        total = lat_num * 2 + lon_num + 1
        pos = [(lo * 2, la - lane.width / 2) for lo, la in zip(range(-int(total / 2), int(total / 2)), lat_1 + lat_2 + lat_3)]
        left = 1 if on_left else -1
        for p in pos:
            self.obstacle("cone", lane, p[0] + lon0, left * p[1])

    # ------------------------------------------------------------------ result
    def scenario(self, map_id=0):
        n = len(self.static)
        return Scenario(map_id, np.asarray(self.static, np.float32).reshape(n, 16), np.asarray(self.dyn, np.float64).reshape(n, 14),
                        np.asarray(self.routes, np.int32).reshape(n, ROUTE_MAX), np.asarray(self.ints, np.int32).reshape(n, 6),
                        np.asarray(self.idm, np.float32).reshape(n, 2), np.asarray(self.objects, np.float64).reshape(-1, 8),
                        self.seed)


def respawn_table(big, seed, idx=None):
    """Respawn / hybrid traffic (traffic_manager.py:112-121, 279-296): the lanes a removed vehicle may come back on, each
    with the route a vehicle born there gets (the layout of oracle/ref_export.export_respawn_lanes)."""
    idx = idx or MapIndex(big)
    roads = []
    for block in big.blocks:
        for road in block.respawn_roads:
            if road in roads:
                roads.remove(road)
            else:
                roads.append(road)
    out = []
    for road in roads:
        for lane in big.world.lanes(road):
            li = idx.lane_index[id(lane)]
            try:
                ck, _ = route_for(big, li, seed)
            except Exception:
                continue
            out.append(dict(lane=idx.lane_id(lane), route=[idx.nodes[c] for c in ck]))
    return out


def populate(big, seed, traffic_density=0.1, traffic_mode="trigger", accident_prob=0.0, lane_num=3, lane_width=3.5,
             include_breakdown=True, map_id=0, random_spawn_lane=True, need_inverse_traffic=False, random_traffic=False,
             random_agent_model=False, agent_model="default"):
    """The roster `env.reset(seed)` leaves behind on the map `big`, managers in the reference's order."""
    sp = Spawner(big, seed, lane_num, lane_width)
    if random_traffic:   # PGTrafficManager.seed skips the re-seeding (traffic_manager.py:323-325): an unseeded stream
        sp.traffic_rng = np.random.RandomState()
    for blk in big.blocks:   # the map is built first: every TollGateBuilding a TollGate block spawns takes a seed from the
        for _ in getattr(blk, "buildings", ()):   # engine's stream (pgblock/tollgate.py:64-76, engine/base_engine.py:123-135)
            sp.engine_seed()
    sp.accidents(accident_prob, include_breakdown)
    n_obj_vehicles = len(sp.static)
    sp.ego(random_spawn_lane, random_agent_model, agent_model)
    if abs(traffic_density) >= 1e-2:
        if traffic_mode == "respawn":
            sp.traffic_respawn(traffic_density)
        else:
            sp.traffic_trigger(traffic_density, need_inverse_traffic)
    if n_obj_vehicles:   # roster order: agents first, then traffic, the accident scenes' cars last
        order = list(range(n_obj_vehicles, len(sp.static))) + list(range(n_obj_vehicles))
        for name in ("static", "dyn", "routes", "ints", "idm"):
            setattr(sp, name, [getattr(sp, name)[k] for k in order])
    # the toll booths close the object table (oracle/ref_export.Roster.objects_table): kind 4, half extent across the heading,
    # half extent along it (BUILDING_LENGTH / 2), BUILDING_HEIGHT (component/buildings/tollgate_building.py:7-26)
    idx = None
    for blk in big.blocks:
        for lane, pos, heading in getattr(blk, "buildings", ()):
            idx = idx or MapIndex(big)
            sp.objects.append([4.0, float(pos[0]), float(pos[1]), float(heading), lane.width / 2.0, 5.0, 5.0, idx.lane_id(lane)])
    return sp.scenario(map_id)

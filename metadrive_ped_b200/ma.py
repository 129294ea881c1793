"""Multi-agent respawn tables (host side, reset time).

What the reference's SpawnManager / navigation compute lazily in Python when an agent is respawned
(manager/spawn_manager.py:117-217, envs/marl_envs/multi_agent_metadrive.py:176-212,
component/navigation_module/node_network_navigation.py:43-128) is tabulated once per map, so that the device can
respawn without the host: the safe spawn places (first slot of every lane of every spawn road) and the checkpoint
route for every (spawn road, destination) pair.
"""
import math
from collections import deque

import numpy as np

from . import scene as sc

RESPAWN_REGION_LONGITUDE = 8.0  # manager/spawn_manager.py:31-35
RESPAWN_REGION_LATERAL = 3.0
TAPE_LEN = 256
TAPE_W = 4


def shortest_path(road_i, start_node, goal_node):
    """NodeRoadNetwork.shortest_path (component/road_network/node_road_network.py:243-271): breadth-first search over
    the node graph, never revisiting a node of the current path.  The reference iterates a Python set of successor
    names (arbitrary order among equally short paths); here successors are visited in road-table order."""
    succ = {}
    for r in road_i:
        succ.setdefault(int(r[0]), []).append(int(r[1]))
    queue = deque([(start_node, [start_node])])
    while queue:
        node, path = queue.popleft()
        for nxt in succ.get(node, []):
            if nxt in path:
                continue
            if nxt == goal_node:
                return path + [nxt]
            if nxt in succ:
                queue.append((nxt, path + [nxt]))
    return []


def route_for(road_i, spawn_road, dest_node):
    """NodeNetworkNavigation.set_route (node_network_navigation.py:93-128): checkpoints + initial target indices."""
    path = shortest_path(road_i, int(spawn_road[0]), int(dest_node))
    if len(path) <= 2:
        return [int(spawn_road[0]), int(spawn_road[1])]
    return path


def build_ma_tables(geo: "sc.MapGeometry", spawn_roads, dest_nodes):
    """places [R*lanes, 8] = x, y, quat w, quat z, lane id, cos(heading), sin(heading), spawn-road index
       routes [R*D, ROUTE_MAX] node ids, -1 padded (row = road * D + destination)."""
    spawn_roads = np.asarray(spawn_roads, np.int32).reshape(-1, 2)
    dest_nodes = np.asarray(dest_nodes, np.int32)
    if dest_nodes.ndim == 1:      # one destination list shared by every spawn road (a destination is DRAWN per agent)
        dest_nodes = np.tile(dest_nodes, (len(spawn_roads), 1))
    # else [R, D]: per-road lists, e.g. D = 1 for the maps whose destination follows from the spawn road (auto_assign_task)
    road_key = {(int(r[0]), int(r[1])): k for k, r in enumerate(geo.road_i)}
    places = []
    for ri, (a, b) in enumerate(spawn_roads):
        r = geo.road_i[road_key[(int(a), int(b))]]
        for idx in range(int(r[3])):
            lane = int(r[2]) + idx
            row = geo.lane_f[lane]
            assert row[0] == 0, "respawn is only supported on straight lanes (spawn_manager.py:176)"
            lon = RESPAWN_REGION_LONGITUDE / 2
            x, y = sc.lane_position(row, lon, 0.0)
            heading = sc.lane_heading_at(row, lon)
            yaw = heading - math.pi / 2  # the chassis +Y axis is the nose (component/vehicle/base_vehicle.py:990-1001)
            places.append([x, y, math.cos(yaw / 2), math.sin(yaw / 2), lane, math.cos(heading), math.sin(heading), ri])
    R, D = dest_nodes.shape
    routes = np.full((R * D, sc.ROUTE_MAX), -1, np.int32)
    for ri in range(R):
        for d in range(D):
            if dest_nodes[ri, d] < 0:     # a shorter list, -1 padded (parking lot: 8 spaces for some roads, 3 exits for the others)
                continue
            p = route_for(geo.road_i, spawn_roads[ri], dest_nodes[ri, d])
            assert 2 <= len(p) <= sc.ROUTE_MAX, (ri, d, p)
            routes[ri * D + d, :len(p)] = p
    return dict(places=np.array(places, np.float64), routes=routes, n_roads=R, n_dests=D)


def make_tape(n_envs, seed=0, length=TAPE_LEN):
    """[n_envs * length, TAPE_W] random tape (include/md_layout.h: env_tape): integer draws in columns 0, 2, 3 and the
    float bits of a uniform [0, 1) number in column 1."""
    rng = np.random.default_rng(seed)
    t = rng.integers(0, 2**31 - 1, size=(n_envs * length, TAPE_W), dtype=np.int64).astype(np.int32)
    t[:, 1] = rng.random(n_envs * length, dtype=np.float32).view(np.int32)
    return t


# ---------------------------------------------------------------------------------------------- scenes at reset
import json
import os

from .abi import make_config

ASSET_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


# The multi-agent envs' fixed maps and spawn roads (envs/marl_envs/marl_inout_roundabout.py:12-24, 27-60;
# marl_intersection.py:12-25, 27-71), built on the product side by pgmap.build_fixed.  The exported assets
# (assets/ma_*.npz, oracle/gen_assets.py --env ma_*) are their goldens (tests/test_pgmap.py).
# `spawn_nodes`: (block index, part, node a, node b) of the roads whose NEGATIVE is a spawn road.  `fixed_dest`: no
# destination draw - the default SpawnManager leaves `destination` None and NodeNetworkNavigation.auto_assign_task
# (component/navigation_module/node_network_navigation.py:71-91) sends an agent born on a positive road to the end of the last
# block's socket and one born on a negative road to the end of the first block's negative road.
MA_MAPS = {
    "roundabout": dict(env="ma_roundabout", num_agents=40, spawn_nodes=[(1, 0, 2, 3), (1, 1, 2, 3), (1, 2, 2, 3)],
                       lane_num=2, exit_length=60.0),
    "intersection": dict(env="ma_intersection", num_agents=30, spawn_nodes=[(1, 0, 0, 1), (1, 1, 0, 1), (1, 2, 0, 1)],
                         lane_num=2, exit_length=60.0),
    # envs/marl_envs/marl_bottleneck.py:10-25: first road + the negative of Split's socket road
    "bottleneck": dict(env="ma_bottleneck", num_agents=20, spawn_nodes=[(2, 0, 0, 1)], lane_num=4, exit_length=60.0,
                       fixed_dest=True),
    # envs/marl_envs/marl_tollgate.py:15-36: first road + the negative of the closing Merge's socket road
    "tollgate": dict(env="ma_tollgate", num_agents=40, spawn_nodes=[(3, 0, 0, 1)], lane_num=3, exit_length=70.0,
                     fixed_dest=True),
    # envs/marl_envs/marl_bidirection.py:10-25 (map only: the reference's env raises KeyError('use_lateral') in its
    # reward_function, :113, at the first step, so there is no behaviour to pin)
    "bidirection": dict(env="ma_bidirection", num_agents=20, spawn_nodes=[(3, 0, 0, 1)], lane_num=4, exit_length=60.0,
                        fixed_dest=True),
    # envs/marl_envs/marl_parking_lot.py:22-43, 144-184: first block -> ParkingLot (4 spaces a side) -> T intersection (exits 10 m), one
    # lane each way.  Spawn roads = the three roads INTO the lot (first road, the negatives of the T's two far exits) + the eight
    # parking spaces under their second road name (ParkingLot.node(1, i, 5) -> (1, i, 6), :207-211).  An agent born on a road into the
    # lot is sent to a parking space nobody else is heading for, one born in a space to the far end of one of the three roads (:80-88).
    # MultiAgentMetaDrive itself (envs/marl_envs/multi_agent_metadrive.py:12-62): a BIG-generated map (config map / seed), every agent
    # born on the first road and - no destination draw - bound for the end of the last block (auto_assign_task)
    "pg": dict(env="ma_pg", num_agents=15, spawn_nodes=[], lane_num=3, exit_length=50.0, fixed_dest=True, pg=True),
    "parkinglot": dict(env="ma_parkinglot", num_agents=10, spawn_nodes=[(2, 0, 0, 1), (2, 2, 0, 1)], lane_num=1, exit_length=20.0,
                       parking=True),
}
ASSET_KIND = {"ma_parkinglot.npz": "parkinglot", "ma_roundabout.npz": "roundabout", "ma_intersection.npz": "intersection", "ma_bottleneck.npz": "bottleneck",
              "ma_tollgate.npz": "tollgate", "ma_bidirection.npz": "bidirection"}


def generated_source(kind, lane_num=None, lane_width=3.5, exit_length=None, parking_space_num=8, pg_seed=0, pg_map=3, **chain_kw):
    """What an exported multi-agent asset holds, generated: lane tables, spawn roads (the first block's second road and the
    three roads ENTERING the block, i.e. the negatives of its exits), destination nodes, the static_default vehicle row
    (component/pg_space.py:227-234, vehicle_type.py:35-36) and SpawnManager's slot constants (spawn_manager.py:25-35)."""
    from . import pgmap, pgspawn
    m = MA_MAPS[kind]
    lane_num = m["lane_num"] if lane_num is None else lane_num
    exit_length = m["exit_length"] if exit_length is None else exit_length
    if m.get("pg"):
        lane_f, lane_i, road_i, meta, big = pgmap.generate(int(pg_seed), pg_map, lane_num, lane_width, exit_length)
    else:
        lane_f, lane_i, road_i, meta, big = pgmap.build_fixed(kind, lane_num, lane_width, exit_length, parking_space_num, **chain_kw)
    node = {n: k for k, n in enumerate(meta["nodes"])}
    roads = [(">>", ">>>")]
    for bi, part, a, b in m["spawn_nodes"]:
        blk = big.blocks[bi]
        roads.append(pgmap.neg_road((blk.node(part, a), blk.node(part, b))))
    d = pgspawn.DIMS["static_default"]
    static = [pgspawn.VEHICLE_TYPES.index("static_default"), d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], d[8], 800, 150, 40,
              0.9, 80, 0.0]
    conf = dict(env=m["env"], num_agents=m["num_agents"], lane_num=lane_num, lane_width=float(lane_width), exit_length=float(exit_length),
                entrance_length=10.0,
                respawn_longitude=RESPAWN_REGION_LONGITUDE, respawn_lateral=RESPAWN_REGION_LATERAL, max_vehicle_length=10.0,
                max_vehicle_width=2.5, disable_u_turn=bool(kind == "intersection" and lane_num < 2),   # marl_intersection.py:103
                fixed_dest=bool(m.get("fixed_dest", False)))
    # static bodies the map itself brings: the toll booths (scene.Scenario.objects rows: kind 4 = building, x, y, heading,
    # half extent across the heading, half extent along it - the barrier's column order -, height BUILDING_HEIGHT = 5, lane id)
    objects = np.array([[4.0, b[1], b[2], b[3], b[5], b[4], 5.0, b[0]] for blk in meta["blocks"]
                        for b in blk.get("buildings", [])], np.float64).reshape(-1, 8)
    if m.get("parking"):
        lot = big.blocks[1]
        n_in, spaces = len(roads), [node[r[1]] for r in lot.dest_roads]
        exits = [node[pgmap.neg_road(r)[1]] for r in roads]
        roads = roads + list(lot.parking_spawn_roads)
        dest = np.full((len(roads), max(len(spaces), n_in)), -1, np.int32)   # one destination list per spawn road, -1 padded
        dest[:n_in, :len(spaces)] = spaces
        dest[n_in:, :n_in] = exits
        conf.update(parking_spaces=len(spaces), parking_in_roads=n_in)
        static[15] = 1.0   # vehicle_config enable_reverse (marl_parking_lot.py:38)
    elif m.get("pg"):   # NodeNetworkNavigation.auto_assign_task: the end of the last block's socket road
        dest = np.array([node[pgspawn.route_for(big, (">>", ">>>", 0), int(pg_seed))[0][-1]]], np.int32)
    else:
        dest = np.array([node[pgmap.neg_road(r)[1]] for r in roads], np.int32)
    return dict(lane_f=lane_f, lane_i=lane_i, road_i=road_i, meta=json.dumps(meta), config=json.dumps(conf), objects=objects,
                spawn_roads=np.array([[node[a], node[b]] for a, b in roads], np.int32), dest_nodes=dest,
                veh_static=np.asarray(static, np.float32), big=big)


class MultiAgentLibrary:
    """The fixed map of one multi-agent env plus what SpawnManager needs.  `name` = "roundabout" / "intersection" (generated
    by pgmap.build_fixed) or an exported asset file (oracle/gen_assets.py --env ma_*).

    `scenario(rng)` restates SpawnManager.reset (manager/spawn_manager.py:72-115): num_agents of the available slots
    (spawn road x lane x longitudinal slot) drawn without replacement, a random offset inside the slot
    (`_randomize_position_in_slot`, :211-217) and a random destination per agent (marl_inout_roundabout.py:138-143).
    The reference draws these from an unseeded generator; here the caller passes the generator."""
    def __init__(self, name, from_asset=False, parking_space_num=8, lane_num=None, lane_width=3.5, exit_length=None, pg_seed=0, pg_map=3,
                 **chain_kw):
        self.pg_seed = int(pg_seed)
        self.big = None   # the generated map's blocks and lanes (IDM traffic is populated over them); an exported asset has none
        if not from_asset and ASSET_KIND.get(name, name) in MA_MAPS:
            d = generated_source(ASSET_KIND.get(name, name), lane_num, lane_width, exit_length, parking_space_num, pg_seed, pg_map, **chain_kw)
            self.big = d["big"]
        else:
            path = name if os.path.exists(name) else os.path.join(ASSET_DIR, name)
            d = np.load(path, allow_pickle=False)
        self.conf = json.loads(str(d["config"]))
        self.table = sc.MapTable(np.asarray(d["lane_f"], np.float64), np.asarray(d["lane_i"], np.int32),
                                 np.asarray(d["road_i"], np.int32), json.loads(str(d["meta"])), int(self.conf["lane_num"]))
        self.geo = sc.build_map_geometry(self.table)
        self.spawn_roads = np.asarray(d["spawn_roads"], np.int32)
        self.dest_nodes = np.asarray(d["dest_nodes"], np.int32)
        if self.conf.get("disable_u_turn") and self.dest_nodes.ndim == 1:
            # MAIntersectionSpawnManager.update_destination_for on a one-lane intersection (marl_intersection.py:76-82, 103): the
            # agent's own spawn road is no destination - one destination list per spawn road, without its own entry, so that the
            # uniform draw over the list (reset-time scenario and k_respawn alike) is the reference's draw
            R_ = len(self.dest_nodes)
            self.dest_nodes = np.array([[self.dest_nodes[j] for j in range(R_) if j != i] for i in range(R_)], np.int32)
        self.veh_static = np.asarray(d["veh_static"], np.float32)
        self.objects = np.asarray(d["objects"], np.float64) if "objects" in d else np.zeros((0, 8))
        if self.conf.get("fixed_dest") or self.conf["env"] in ("ma_bottleneck", "ma_tollgate", "ma_bidirection"):
            # destination of spawn road k = the far end of the map = the end node of the OTHER spawn road's negative
            self.dest_nodes = self.dest_nodes[::-1].reshape(-1, 1)
        self.tables = build_ma_tables(self.geo, self.spawn_roads, self.dest_nodes)
        c = self.conf
        self.n_slots = int(math.floor((c["exit_length"] - c["entrance_length"]) / c["respawn_longitude"]))
        road_key = {(int(r[0]), int(r[1])): k for k, r in enumerate(self.geo.road_i)}
        self.slots = []  # (lane id, longitude, spawn-road index) in SpawnManager._auto_fill_spawn_roads_randomly order
        for ri, (a, b) in enumerate(self.spawn_roads):
            r = self.geo.road_i[road_key[(int(a), int(b))]]
            for idx in range(int(c["lane_num"])):
                for j in range(self.n_slots):
                    self.slots.append((int(r[2]) + idx, c["respawn_longitude"] / 2 + j * c["respawn_longitude"], ri))

    @property
    def max_capacity(self):
        return len(self.slots)

    def traffic(self, num_agents, density, seed):
        """PGTrafficManager.reset in a multi-agent env (manager/traffic_manager.py:54-72, 211-277; trigger mode): the vehicles of
        every block after the first, parked until an agent enters the block's trigger road.  The manager's streams are seeded
        with the env's seed; every agent spawned before took one seed from the engine's stream (engine/base_engine.py:123-135),
        which decides the traffic vehicles' sampled parameters.  Rows in roster order, pinned on `cfg3_ma_roundabout_traffic`."""
        from . import pgspawn
        if self.big is None:
            raise NotImplementedError("IDM traffic needs the generated map (exported assets hold lane tables only)")
        sp = pgspawn.Spawner(self.big, int(seed), int(self.conf["lane_num"]), float(self.conf.get("lane_width", 3.5)))
        for _ in range(num_agents):
            sp.engine_seed()
        sp.traffic_trigger(float(density))
        n = len(sp.static)
        return (np.array(sp.static, np.float64).reshape(n, 16), np.array(sp.dyn, np.float64).reshape(n, 14),
                np.array(sp.routes, np.int32).reshape(n, sc.ROUTE_MAX), np.array(sp.ints, np.int32).reshape(n, 6),
                np.array(sp.idm, np.float32).reshape(n, 2))

    def scenario(self, rng, num_agents, traffic_density=0.0, traffic_seed=0):
        c = self.conf
        assert 0 < num_agents <= self.max_capacity, \
            "Too many agents! We only accept {} agents, but you have {} agents!".format(self.max_capacity, num_agents)
        pick = rng.choice(len(self.slots), num_agents, replace=False)
        D = self.tables["n_dests"]
        n_in, free = int(c.get("parking_in_roads", 0)), list(range(int(c.get("parking_spaces", 0))))
        dl = (c["respawn_longitude"] - c["max_vehicle_length"]) / 2
        dw = (c["respawn_lateral"] - c["max_vehicle_width"]) / 2
        H = float(self.veh_static[3])
        veh_dyn = np.zeros((num_agents, 14), np.float64)
        routes = np.full((num_agents, sc.ROUTE_MAX), -1, np.int32)
        veh_int = np.zeros((num_agents, 6), np.int32)
        parking = np.full(num_agents, -1, np.int32)
        for k, si in enumerate(pick):
            lane, lon, ri = self.slots[int(si)]
            lon = lon + rng.uniform(-dl, dl) if dl > 0 else lon + rng.uniform(dl, -dl)
            lat = rng.uniform(-dw, dw) if dw > 0 else rng.uniform(dw, -dw)
            row = self.geo.lane_f[lane]
            x, y = sc.lane_position(row, lon, lat)
            yaw = sc.lane_heading_at(row, lon) - math.pi / 2
            veh_dyn[k, 0:3] = [x, y, H / 2]
            veh_dyn[k, 3:7] = [math.cos(yaw / 2), 0.0, 0.0, math.sin(yaw / 2)]
            if not n_in:
                d = int(rng.integers(0, D))
            elif ri < n_in:    # ParkingLotSpawnManager.get_parking_space (marl_parking_lot.py:61-71): a space nobody is heading for
                assert len(free) > 0, "more agents on the roads into the lot than parking spaces"
                d = free.pop(int(rng.integers(0, len(free))))
                parking[k] = d
            else:              # update_destination_for (:80-88): the far end of one of the roads into the lot
                d = int(rng.integers(0, n_in))
            rt = self.tables["routes"][ri * D + d]
            routes[k] = rt
            n_ck = int((rt >= 0).sum())
            veh_int[k] = [1, -1, lane, 0, 1 if n_ck > 2 else 0, 1]
        static = np.tile(self.veh_static, (num_agents, 1))
        idm = np.tile(np.array([[0.0, 30.0]], np.float32), (num_agents, 1))
        if abs(traffic_density) >= 1e-2:   # traffic_manager.py:62
            t_static, t_dyn, t_routes, t_int, t_idm = self.traffic(num_agents, traffic_density, traffic_seed)
            static, veh_dyn = np.concatenate([static, t_static]), np.concatenate([veh_dyn, t_dyn])
            routes, veh_int, idm = np.concatenate([routes, t_routes]), np.concatenate([veh_int, t_int]), np.concatenate([idm, t_idm])
            parking = np.concatenate([parking, np.full(len(t_static), -1, np.int32)])
        return sc.Scenario(0, static, veh_dyn, routes, veh_int, idm, self.objects, int(traffic_seed), parking if n_in else None)

    def build_world(self, n_envs, num_agents, seed=0, traffic_density=0.0, traffic_seed=0, **cfg_kw):
        """(arrays, cfg) for n_envs independent multi-agent envs: num_agents + 1 seats each (the spare seat keeps a
        finished agent's last transition and a newborn's first observation on different rows), then the env's IDM traffic
        (traffic_density > 0; env e draws it with the seed traffic_seed + e)."""
        rng = np.random.default_rng(seed)
        scen = [self.scenario(rng, num_agents, traffic_density, traffic_seed + e) for e in range(n_envs)]
        NA = num_agents + 1
        S = ((NA + max(len(s.veh_static) for s in scen) - num_agents + 3) // 4) * 4
        tape = make_tape(n_envs, seed=seed + 1)
        O = len(self.objects)
        arrays = sc.pack([self.geo], scen, S, NA, O, ma_tables={0: self.tables}, ma_tables_tape=tape)
        kw = dict(is_multi_agent=1, ma_places=len(self.tables["places"]), ma_dests=self.tables["n_dests"],
                  ma_roads=self.tables["n_roads"], tape_len=TAPE_LEN, parking_spaces=int(self.conf.get("parking_spaces", 0)),
                  parking_in_roads=int(self.conf.get("parking_in_roads", 0)))
        kw.update(cfg_kw)
        return arrays, make_config(n_envs, S, NA, O, **kw)

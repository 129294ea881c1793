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
    dest_nodes = np.asarray(dest_nodes, np.int32).reshape(-1)
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
    R, D = len(spawn_roads), len(dest_nodes)
    routes = np.full((R * D, sc.ROUTE_MAX), -1, np.int32)
    for ri in range(R):
        for d in range(D):
            p = route_for(geo.road_i, spawn_roads[ri], dest_nodes[d])
            assert 2 <= len(p) <= sc.ROUTE_MAX, (ri, d, p)
            routes[ri * D + d, :len(p)] = p
    return dict(places=np.array(places, np.float64), routes=routes, n_roads=R, n_dests=D)


def make_tape(n_envs, seed=0, length=TAPE_LEN):
    """[n_envs * length, 2] uniform 32-bit draws (place, destination) consumed in order by an env's respawns."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 2**31 - 1, size=(n_envs * length, 2), dtype=np.int64).astype(np.int32)

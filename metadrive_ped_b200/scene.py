"""Scene tables: from a lane table to the flat arrays the step kernels read (layouts: include/md_layout.h).

A map enters as the reference's lane graph flattened to three small tables (`lane_f` [L,10], `lane_i` [L,8],
`road_i` [R,6], see `MapTable`).  Everything the reference builds as Bullet bodies at `PGBlock.create_in_world`
is derived here, on the host, once per distinct map:

  lane convex hulls     component/block/base_block.py:431-466 over lane.polygon
                        (component/lane/straight_lane.py:82-95, circular_lane.py:123-174)
  lane-line boxes       component/pgblock/pg_block.py:248-292,334-361 -> base_block.py:468-519
  sidewalk strips       component/pgblock/pg_block.py:294-332 -> base_block.py:362-397
  static broad phase    a uniform grid over line boxes + sidewalk quads (replaces Bullet's static-world AABB tree)

Host-side numpy in float64, rounded to float32 at the end.
"""
import json
import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

# column counts / indices mirrored from include/md_layout.h
LANE_F, LANE_I, ROAD_I, LINE_F, QUAD_F, MAPD, MAPDF = 16, 8, 6, 8, 8, 16, 4
VEH_P, VEH_S, VEH_C, VEH_I, ROUTE_MAX, VEH_IDM, NAVI_DIM, OBJ_F, ENV_I, TRIGGER_MAX = 16, 16, 16, 16, 24, 8, 10, 12, 8, 8
LINE_NONE, LINE_BROKEN, LINE_CONTINUOUS, LINE_SIDE, LINE_GUARDRAIL = 0, 1, 2, 3, 4
GRID_CELL = 8.0

# PGDrivableAreaProperty (metadrive/constants.py:300-342)
LANE_SEGMENT_LENGTH = 4.0
STRIPE_LENGTH = 1.5
SIDEWALK_LENGTH = 3.0
SIDEWALK_WIDTH = 2.0
POLYGON_SAMPLE_RATE = 1.0


@dataclass
class MapTable:
    """Raw lane graph of one map (what `oracle/ref_export.export_map` writes and the map library stores).

    lane_f [L,10]: type, width, length, then straight: sx sy ex ey | circular: cx cy radius start_phase end_phase
                   direction(+1 ccw, -1 cw) angle
    lane_i [L,8] : road, idx_in_road, from_node, to_node, line_left, line_right, yellow_left, yellow_right
    road_i [R,6] : from_node, to_node, first_lane, n_lanes, is_negative, ord(block id)
    meta         : dict(nodes=[names], blocks=[dict(id, trigger_road, spawn_lanes, negative_lanes, respawn_roads, sockets)])
    """
    lane_f: np.ndarray
    lane_i: np.ndarray
    road_i: np.ndarray
    meta: dict = field(default_factory=dict)
    lane_num: int = 3

    @classmethod
    def from_export(cls, d, lane_num=3):
        meta = d["meta"]
        if not isinstance(meta, dict):
            meta = json.loads(str(meta))
        return cls(np.asarray(d["lane_f"], np.float64), np.asarray(d["lane_i"], np.int32),
                   np.asarray(d["road_i"], np.int32), meta, int(lane_num))


# ---------------------------------------------------------------------------------------------- lane maths (host)
def lane_position(row, lon, lat):
    if row[0] == 0:
        sx, sy, ex, ey = row[3:7]
        ln = math.hypot(ex - sx, ey - sy)
        dx, dy = (ex - sx) / ln, (ey - sy) / ln
        return np.array([sx + lon * dx + lat * dy, sy + lon * dy - lat * dx])
    cx, cy, r, sp, ep, direction = row[3:9]
    phi = direction * lon / r + sp
    rr = r + lat * direction
    return np.array([cx + rr * math.cos(phi), cy + rr * math.sin(phi)])


def lane_local(row, x, y):
    """(longitude, lateral) of a point on a lane row (straight_lane.py:60-66, circular_lane.py:57-121); rows of the packed
    `lane_f` table [type, width, length, 7 shape parameters, ...]"""
    if row[0] == 0:
        sx, sy, ex, ey = row[3:7]
        ln = math.hypot(ex - sx, ey - sy)
        dx, dy = (ex - sx) / ln, (ey - sy) / ln
        return (x - sx) * dx + (y - sy) * dy, (x - sx) * dy - (y - sy) * dx
    cx, cy, r, sp, ep, direction = row[3:9]
    wrap = lambda a: (a + math.pi) % (2 * math.pi) - math.pi
    phase = math.atan2(y - cy, x - cx)
    if abs(wrap(phase - sp)) > abs(wrap(phase - ep)):
        lon = wrap((ep - phase) if direction < 0 else (phase - ep)) * r + row[2]
    else:
        lon = wrap((sp - phase) if direction < 0 else (phase - sp)) * r
    return lon, direction * (math.hypot(x - cx, y - cy) - r)


def lane_heading_at(row, lon):
    if row[0] == 0:
        sx, sy, ex, ey = row[3:7]
        return math.atan2(ey - sy, ex - sx)
    cx, cy, r, sp, ep, direction = row[3:9]
    return direction * lon / r + sp + math.pi / 2 * direction


def lane_polygon(row):
    """lane.polygon of the reference (straight_lane.py:82-95, circular_lane.py:123-174)."""
    width, length = row[1], row[2]
    longs = np.arange(0, length + POLYGON_SAMPLE_RATE, POLYGON_SAMPLE_RATE)
    pts = []
    if row[0] == 0:
        for k, lat in enumerate([+width / 2, -width / 2]):
            ls = longs if k == 0 else longs[::-1]
            for lon in ls:
                pts.append(lane_position(row, lon, lat))
        return np.array(pts)
    sh, eh = lane_heading_at(row, 0), lane_heading_at(row, length)
    sd = np.array([math.cos(sh), math.sin(sh)])
    ed = np.array([math.cos(eh), math.sin(eh)])
    for k, lat in enumerate([+width / 2, -width / 2]):
        ls = longs if k == 0 else longs[::-1]
        for t, lon in enumerate(ls):
            p = lane_position(row, lon, lat)
            pts.append(p)
            if (t == 0 and k == 0) or (t == len(ls) - 1 and k == 1):
                pts.append(p - sd * POLYGON_SAMPLE_RATE)
            elif (t == 0 and k == 1) or (t == len(ls) - 1 and k == 0):
                pts.append(p + ed * POLYGON_SAMPLE_RATE)
    return np.array(pts)


HULL_TOL = 1e-6  # m; a hull vertex closer than this to the chord of its neighbours is a sample on a straight edge


def convex_hull(points):
    """Andrew monotone chain (exact turn test, so the order of nearly-equal abscissae cannot drop a corner), then the
    samples lying on straight edges are removed by their distance to the chord of their neighbours.  CCW."""
    pts = sorted(set((float(p[0]), float(p[1])) for p in points))
    if len(pts) <= 2:
        return np.array(pts)

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    lower, upper = [], []
    for p in pts:
        while len(lower) >= 2 and cross(lower[-2], lower[-1], p) <= 0.0:
            lower.pop()
        lower.append(p)
    for p in reversed(pts):
        while len(upper) >= 2 and cross(upper[-2], upper[-1], p) <= 0.0:
            upper.pop()
        upper.append(p)
    hull = lower[:-1] + upper[:-1]
    changed = True
    while changed and len(hull) > 3:
        changed = False
        for i in range(len(hull)):
            o, a, b = hull[i - 1], hull[i], hull[(i + 1) % len(hull)]
            chord = math.hypot(b[0] - o[0], b[1] - o[1])
            if chord == 0.0 or abs(cross(o, a, b)) / chord < HULL_TOL:
                del hull[i]
                changed = True
                break
    return np.array(hull)


def _line_segments(row, lat, line_type):
    """(start, end) pairs of the ghost boxes of one lane border (pg_block.py:259-292)."""
    length = row[2]
    out = []
    if line_type == LINE_BROKEN:
        n = int(length / (2 * STRIPE_LENGTH))
        for seg in range(n):
            s = lane_position(row, seg * STRIPE_LENGTH * 2, lat)
            e = lane_position(row, seg * STRIPE_LENGTH * 2 + STRIPE_LENGTH, lat)
            if seg == n - 1:
                e = lane_position(row, length - STRIPE_LENGTH, lat)
            out.append((s, e))
    else:
        n = int(length / LANE_SEGMENT_LENGTH)
        if n == 0:
            out.append((lane_position(row, 0, lat), lane_position(row, length, lat)))
        for seg in range(n):
            s = lane_position(row, LANE_SEGMENT_LENGTH * seg, lat)
            e = lane_position(row, length, lat) if seg == n - 1 else lane_position(row, (seg + 1) * LANE_SEGMENT_LENGTH, lat)
            out.append((s, e))
    return out


def _sidewalk_quads(row, lateral_direction=1):
    """Strip polygon of pg_block.py:294-332 cut into the convex quads between consecutive samples."""
    width, length = row[1], row[2]
    longs = np.arange(0, length + SIDEWALK_LENGTH, SIDEWALK_LENGTH)
    start_lat = (width / 2 + 0.2) * lateral_direction
    side_lat = (width / 2 + 0.2 + SIDEWALK_WIDTH) * lateral_direction
    radius = row[5] if row[0] == 1 else 0.0
    if radius != 0 and side_lat > radius:
        return []
    inner = [lane_position(row, min(length + 0.1, lon), start_lat) for lon in longs]
    outer = [lane_position(row, min(length + 0.1, lon), side_lat) for lon in longs]
    quads = []
    for i in range(len(longs) - 1):
        q = np.array([inner[i], inner[i + 1], outer[i + 1], outer[i]])
        area2 = sum(q[k][0] * q[(k + 1) % 4][1] - q[(k + 1) % 4][0] * q[k][1] for k in range(4))
        if abs(area2) < 1e-9:
            continue
        if area2 < 0:
            q = q[::-1]
        quads.append(q.reshape(-1))
    return quads


@dataclass
class MapGeometry:
    """Derived, kernel-ready tables of one map (all ids local to the map)."""
    lane_f: np.ndarray  # [L,16]
    lane_i: np.ndarray  # [L,8]
    lane_bb: np.ndarray  # [L,4]
    road_i: np.ndarray  # [R,6]
    hull_xy: np.ndarray  # [H,2]
    line_f: np.ndarray  # [S,6]
    quad_f: np.ndarray  # [Q,8]
    grid_start: np.ndarray
    grid_items: np.ndarray
    lgrid_start: np.ndarray
    lgrid_items: np.ndarray
    grid_origin: tuple
    grid_dims: tuple
    lane_num: int
    meta: dict


def build_map_geometry(mt: MapTable, map_region_size=1024.0) -> MapGeometry:
    L = mt.lane_f.shape[0]
    lane_f = np.zeros((L, LANE_F), np.float64)
    lane_i = np.zeros((L, LANE_I), np.int32)
    lane_bb = np.zeros((L, 4), np.float64)
    hulls, lines, quads = [], [], []
    hull_off = 0
    sidewalk_done = set()
    half_region = map_region_size / 2
    for l in range(L):
        row = mt.lane_f[l]
        ri = mt.lane_i[l]
        lane_f[l, 0:3] = row[0:3]
        if row[0] == 0:
            sx, sy, ex, ey = row[3:7]
            ln = math.hypot(ex - sx, ey - sy)
            lane_f[l, 3:10] = [sx, sy, ex, ey, (ex - sx) / ln, (ey - sy) / ln, math.atan2(ey - sy, ex - sx)]
            lane_f[l, 10:14] = [sx, sy, ex, ey]
        else:
            lane_f[l, 3:10] = row[3:10]
            s, e = lane_position(row, 0, 0), lane_position(row, row[2], 0)
            lane_f[l, 10:14] = [s[0], s[1], e[0], e[1]]
        poly = lane_polygon(row)
        if row[0] == 0:  # last polygon sample of a straight lane: its hull is [0, this] x [-w/2, w/2] in lane coordinates
            lane_f[l, 14] = float(np.arange(0, row[2] + POLYGON_SAMPLE_RATE, POLYGON_SAMPLE_RATE)[-1])
        hull = convex_hull(poly)
        lane_f[l, 15] = len(hull)
        if row[0] == 1:
            # arc strip: rotate the hull so that the run of outer-arc chords (both ends on the radius R + w/2) comes last;
            # col 15 = number of leading "other" edges, col 14 = radius of the circle inscribed in the chord polygon
            # (minus 1 cm): a point inside that circle satisfies every outer chord, so kernels may skip those edges.
            cx, cy, r = row[3], row[4], row[5]
            ro = r + row[1] / 2
            on_outer = np.abs(np.hypot(hull[:, 0] - cx, hull[:, 1] - cy) - ro) < 1e-6
            n = len(hull)
            chord = np.array([on_outer[i] and on_outer[(i + 1) % n] for i in range(n)])
            if chord.any() and not chord.all():
                start = next(i for i in range(n) if chord[i] and not chord[i - 1])  # first edge of the chord run
                run = 0
                while chord[(start + run) % n]:
                    run += 1
                first_other = (start + run) % n
                hull = np.roll(hull, -first_other, axis=0)
                lane_f[l, 15] = n - run
                lane_f[l, 14] = ro * math.cos(0.5 * POLYGON_SAMPLE_RATE / r) - 0.01
        hulls.append(np.concatenate([hull, hull[:1]]))  # stored closed: edge i = (v[i], v[i+1]), n edges, n+1 rows
        # the hull test counts a point up to 1e-3 m^2 / edge length outside an edge as inside (md_device.cuh
        # point_in_hull, the stand-in for Bullet's hull margin): the box must not cut that band off
        edges = np.linalg.norm(np.diff(np.concatenate([hull, hull[:1]]), axis=0), axis=1)
        pad = min(0.05, 1e-3 / max(float(edges.min()), 1e-3))
        lane_bb[l] = [hull[:, 0].min() - pad, hull[:, 1].min() - pad, hull[:, 0].max() + pad, hull[:, 1].max() + pad]
        lane_i[l] = [ri[0], ri[1], ri[2], ri[3], hull_off, len(hull), ri[4], ri[5]]
        hull_off += len(hull) + 1
        # lane lines (pg_block.py:248-255, 334-361): left border only for lane 0 of a positive road
        road = mt.road_i[ri[0]]
        build = [ri[1] == 0 and not road[4], True]
        for k, side in enumerate([-1, 1]):
            if not build[k]:
                continue
            lt = int(ri[4 + k])
            yellow = int(ri[6 + k])
            if lt == LINE_NONE:
                continue
            lat = side * row[1] / 2
            solid = lt in (LINE_CONTINUOUS, LINE_SIDE)  # PGLineType.prohibit (constants.py:288-294)
            kind = (0 if solid else 2) + yellow
            for s, e in _line_segments(row, lat, lt):
                d = e - s
                ln = math.hypot(d[0], d[1])
                mid = (s + e) / 2
                if ln <= 0 or abs(mid[0]) > half_region or abs(mid[1]) > half_region:
                    continue
                lines.append([mid[0], mid[1], ln / 2, kind, d[0] / ln, d[1] / ln, 0.0, 0.0])
            if lt in (LINE_SIDE, LINE_GUARDRAIL) and l not in sidewalk_done:
                sidewalk_done.add(l)
                quads.extend(_sidewalk_quads(row, 1 if lt == LINE_SIDE else side))
    hull_xy = np.concatenate(hulls) if hulls else np.zeros((0, 2))
    line_f = np.array(lines, np.float64).reshape(-1, LINE_F)
    quad_f = np.array(quads, np.float64).reshape(-1, QUAD_F)
    # uniform grid over static items
    boxes = []
    for ln in line_f:
        ex = abs(ln[4]) * ln[2] + abs(ln[5]) * 0.0375
        ey = abs(ln[5]) * ln[2] + abs(ln[4]) * 0.0375
        boxes.append([ln[0] - ex, ln[1] - ey, ln[0] + ex, ln[1] + ey])
    for q in quad_f:
        xs, ys = q[0::2], q[1::2]
        boxes.append([xs.min(), ys.min(), xs.max(), ys.max()])
    boxes = np.array(boxes).reshape(-1, 4)
    lo = np.minimum(lane_bb[:, :2].min(0), boxes[:, :2].min(0) if len(boxes) else lane_bb[:, :2].min(0)) - 1.0
    hi = np.maximum(lane_bb[:, 2:].max(0), boxes[:, 2:].max(0) if len(boxes) else lane_bb[:, 2:].max(0)) + 1.0
    nx = int(math.ceil((hi[0] - lo[0]) / GRID_CELL))
    ny = int(math.ceil((hi[1] - lo[1]) / GRID_CELL))
    cells = [[] for _ in range(nx * ny)]
    for it, b in enumerate(boxes):
        x0 = max(0, int(math.floor((b[0] - lo[0]) / GRID_CELL)))
        x1 = min(nx - 1, int(math.floor((b[2] - lo[0]) / GRID_CELL)))
        y0 = max(0, int(math.floor((b[1] - lo[1]) / GRID_CELL)))
        y1 = min(ny - 1, int(math.floor((b[3] - lo[1]) / GRID_CELL)))
        for cy in range(y0, y1 + 1):
            for cx in range(x0, x1 + 1):
                cells[cy * nx + cx].append(it)
    grid_start = np.zeros(nx * ny + 1, np.int32)
    grid_start[1:] = np.cumsum([len(c) for c in cells])
    grid_items = np.array([it for c in cells for it in c], np.int32)
    # lanes binned into the same cells by hull AABB (broad phase of ray_localization, utils/pg/utils.py:151-203)
    lcells = [[] for _ in range(nx * ny)]
    for l, b in enumerate(lane_bb):
        x0 = max(0, int(math.floor((b[0] - lo[0]) / GRID_CELL)))
        x1 = min(nx - 1, int(math.floor((b[2] - lo[0]) / GRID_CELL)))
        y0 = max(0, int(math.floor((b[1] - lo[1]) / GRID_CELL)))
        y1 = min(ny - 1, int(math.floor((b[3] - lo[1]) / GRID_CELL)))
        for cy in range(y0, y1 + 1):
            for cx in range(x0, x1 + 1):
                lcells[cy * nx + cx].append(l)
    lgrid_start = np.zeros(nx * ny + 1, np.int32)
    lgrid_start[1:] = np.cumsum([len(c) for c in lcells])
    lgrid_items = np.array([it for c in lcells for it in c], np.int32)
    return MapGeometry(lane_f, lane_i, lane_bb, mt.road_i.copy(), hull_xy, line_f, quad_f, grid_start, grid_items,
                       lgrid_start, lgrid_items,
                       (float(lo[0]), float(lo[1])), (nx, ny), mt.lane_num, mt.meta)


# ---------------------------------------------------------------------------------------------- scenario + packing
@dataclass
class Scenario:
    """One env's initial world: which map, the vehicle roster (agents first) and static objects.

    veh_static [n,16] (VEH_P layout), veh_dyn [n,14] (pos3 quat4 vel3 angvel3 static), routes [n,ROUTE_MAX] node ids,
    veh_int [n,6]: kind(1 agent / 2 traffic), trigger_block, spawn_lane, ckpt0, ckpt1, active
    idm [n,2]: overtake_timer, target_speed ; objects [m,8]: kind x y heading a b height lane (+ optional vx vy cols)
    """
    map_id: int
    veh_static: np.ndarray
    veh_dyn: np.ndarray
    routes: np.ndarray
    veh_int: np.ndarray
    idm: np.ndarray
    objects: np.ndarray
    seed: int = 0
    parking: Optional[np.ndarray] = None   # parking-lot env: [n] the parking space every vehicle is heading for, -1 = none


def pack(maps: List[MapGeometry], scenarios: List[Scenario], slots_per_env: int, agents_per_env: int,
         objs_per_env: int, ma_tables=None, ma_tables_tape=None, traffic_respawn=False):
    """Concatenate maps + scenarios into the flat arrays of `MdArrays` (numpy, C-contiguous, f32/i32)."""
    M, E, S, O = len(maps), len(scenarios), slots_per_env, objs_per_env
    map_desc = np.zeros((M, MAPD), np.int32)
    map_descf = np.zeros((M, MAPDF), np.float32)
    lane_off = road_off = hull_off = line_off = quad_off = grid_off = item_off = litem_off = 0
    for k, g in enumerate(maps):
        map_desc[k, :16] = [lane_off, len(g.lane_f), road_off, len(g.road_i), hull_off, line_off, len(g.line_f),
                            quad_off, len(g.quad_f), grid_off, g.grid_dims[0], g.grid_dims[1], item_off, g.lane_num,
                            grid_off, litem_off]
        map_descf[k] = [g.grid_origin[0], g.grid_origin[1], GRID_CELL, 0]
        lane_off += len(g.lane_f)
        road_off += len(g.road_i)
        hull_off += len(g.hull_xy)
        line_off += len(g.line_f)
        quad_off += len(g.quad_f)
        grid_off += len(g.grid_start)
        item_off += len(g.grid_items)
        litem_off += len(g.lgrid_items)

    def cat(name, dtype, width=None):
        arrs = [np.asarray(getattr(g, name)) for g in maps]
        out = np.concatenate(arrs).astype(dtype) if arrs else np.zeros((0, ), dtype)
        if out.size == 0:  # keep a valid pointer for empty tables
            out = np.zeros((1, width) if width else (1, ), dtype)
        return np.ascontiguousarray(out)

    arrays = dict(
        map_desc=map_desc, map_descf=map_descf, lane_f=cat("lane_f", np.float32, LANE_F),
        lane_i=cat("lane_i", np.int32, LANE_I), lane_bb=cat("lane_bb", np.float32, 4),
        road_i=cat("road_i", np.int32, ROAD_I), hull_xy=cat("hull_xy", np.float32, 2),
        line_f=cat("line_f", np.float32, LINE_F), quad_f=cat("quad_f", np.float32, QUAD_F),
        grid_start=cat("grid_start", np.int32), grid_items=cat("grid_items", np.int32),
        lgrid_start=cat("lgrid_start", np.int32), lgrid_items=cat("lgrid_items", np.int32),
    )
    NV = E * S
    env_i = np.zeros((E, ENV_I), np.int32)
    env_trigger = np.full((E, TRIGGER_MAX), -1, np.int32)
    veh_p = np.zeros((NV, VEH_P), np.float32)
    veh_s = np.zeros((NV, VEH_S), np.float32)
    veh_s[:, 3] = 1.0
    veh_c = np.zeros((NV, VEH_C), np.float32)
    veh_i = np.zeros((NV, VEH_I), np.int32)
    veh_i[:, 4] = -1
    veh_i[:, 9] = -1
    veh_route = np.full((NV, ROUTE_MAX), -1, np.int32)
    veh_rroad = np.full((NV, ROUTE_MAX), -1, np.int32)
    road_lut = [{(int(r[0]), int(r[1])): k for k, r in enumerate(g.road_i)} for g in maps]
    veh_idm = np.zeros((NV, VEH_IDM), np.float32)
    veh_navi = np.zeros((NV, NAVI_DIM), np.float32)
    obj_f = np.zeros((max(E * O, 1), OBJ_F), np.float32)
    obj_f[:, 0] = -1
    for e, sc in enumerate(scenarios):
        g = maps[sc.map_id]
        blocks = g.meta.get("blocks", [])
        env_i[e, 0] = sc.map_id
        env_i[e, 3] = len(blocks)
        env_i[e, 4] = sc.seed
        has_wait = bool(np.any((sc.veh_int[:, 0] == 2) & (sc.veh_int[:, 1] > 0))) if len(sc.veh_int) else False
        env_i[e, 1] = 1 if (len(blocks) > 1 and has_wait) else 0
        for b, blk in enumerate(blocks[:TRIGGER_MAX]):
            env_trigger[e, b] = blk["trigger_road"]
        n = len(sc.veh_static)
        n_ag = int(np.sum(sc.veh_int[:, 0] == 1))
        assert n_ag <= agents_per_env and np.all(sc.veh_int[:n_ag, 0] == 1), "agents come first in a roster"
        assert n - n_ag + agents_per_env <= S, (n, agents_per_env, S)
        # agents fill the first seats; everything else starts after the agent seats.  In a multi-agent world every
        # seat exists from the start (kind 1, not alive) with the parameters of seat 0 (agent_manager.py:136-154
        # builds respawned agents from agent_configs["agent0"])
        sl = e * S + np.concatenate([np.arange(n_ag), agents_per_env + np.arange(n - n_ag)]).astype(np.int64)
        if agents_per_env > 1:
            seats = slice(e * S, e * S + agents_per_env)
            veh_i[seats, 0] = 1
            veh_i[seats, 3] = -1
            veh_p[seats] = sc.veh_static[0]
        veh_p[sl] = sc.veh_static
        veh_s[sl, 0:13] = sc.veh_dyn[:, 0:13]
        veh_i[sl, 0] = sc.veh_int[:, 0]
        veh_i[sl, 1] = 1
        veh_i[sl, 2] = sc.veh_int[:, 5]
        veh_i[sl, 3] = sc.veh_int[:, 1]
        veh_i[sl, 4] = sc.veh_int[:, 2]
        veh_i[sl, 5] = sc.veh_int[:, 3]
        veh_i[sl, 6] = sc.veh_int[:, 4]
        veh_i[sl, 7] = (sc.routes >= 0).sum(1)
        veh_i[sl, 10] = sc.veh_dyn[:, 13].astype(np.int32)
        veh_i[sl, 13] = sc.veh_int[:, 2]
        veh_route[sl] = sc.routes
        lut = road_lut[sc.map_id]
        for k in range(n):
            rt = sc.routes[k]
            for j in range(ROUTE_MAX - 1):
                if rt[j] < 0 or rt[j + 1] < 0:
                    break
                veh_rroad[sl[k], j] = lut.get((int(rt[j]), int(rt[j + 1])), -1)
        veh_idm[sl, 0] = sc.idm[:, 0]
        veh_idm[sl, 1] = sc.idm[:, 1]
        veh_c[sl, 0:2] = sc.veh_dyn[:, 0:2]
        if sc.parking is not None:
            veh_c[sl, 14] = np.asarray(sc.parking, np.float32) + 1.0    # VC_PARK (include/md_layout.h)
        m = len(sc.objects)
        if m > O:
            raise ValueError("an env holds %d objects but objs_per_env is %d" % (m, O))
        if m:
            ob = obj_f[e * O:e * O + m]
            ob[:, 0:7] = sc.objects[:, 0:7]
            ob[:, 7] = sc.objects[:, 6] / 2  # shapes are centred at height/2 (base_static_object.py:24)
            ob[sc.objects[:, 0] == 4, 7] = 0.0  # a TollGateBuilding is put back to z = 0 (buildings/tollgate_building.py:25)
            ob[:, 8] = sc.objects[:, 7]
            if sc.objects.shape[1] >= 10:
                ob[:, 10:12] = sc.objects[:, 8:10]
    # multi-agent respawn tables (metadrive_ped_b200/ma.py): per map id dict(places [P,8], routes [R*D, ROUTE_MAX]),
    # replicated per env; env_tape [E*L, TAPE_W] = the envs' random tapes
    if ma_tables:
        t0 = next(iter(ma_tables.values()))
        P, RD = len(t0["places"]), len(t0["routes"])
        ma_place_f = np.zeros((E * P, 8), np.float32)
        ma_route = np.full((E * RD, ROUTE_MAX), -1, np.int32)
        ma_rroad = np.full((E * RD, ROUTE_MAX), -1, np.int32)
        per_map = {}
        for mid, t in ma_tables.items():
            pl = np.array(t["places"], np.float64)
            rr = np.full((RD, ROUTE_MAX), -1, np.int32)
            lut = road_lut[mid]
            for k, rt in enumerate(t["routes"]):
                for j in range(ROUTE_MAX - 1):
                    if rt[j] < 0 or rt[j + 1] < 0:
                        break
                    rr[k, j] = lut.get((int(rt[j]), int(rt[j + 1])), -1)
            per_map[mid] = (pl.astype(np.float32), np.asarray(t["routes"], np.int32), rr)
        for e, sc in enumerate(scenarios):
            pl, rt, rr = per_map[sc.map_id]
            ma_place_f[e * P:(e + 1) * P] = pl
            ma_route[e * RD:(e + 1) * RD] = rt
            ma_rroad[e * RD:(e + 1) * RD] = rr
        env_tape = np.ascontiguousarray(ma_tables_tape, np.int32) if ma_tables_tape is not None else np.zeros((E, 4), np.int32)
        assert env_tape.shape[0] % E == 0 and env_tape.shape[1] == 4
    elif traffic_respawn:
        # respawn / hybrid traffic mode (manager/traffic_manager.py:112-121, 279-296): the same tables hold, per env, the
        # respawn lanes of its map (column 4 = lane id) and the route of a vehicle born on each (one row per lane)
        P = max(1, max(len(g.meta.get("respawn", [])) for g in maps))
        ma_place_f = np.zeros((E * P, 8), np.float32)
        ma_route = np.full((E * P, ROUTE_MAX), -1, np.int32)
        ma_rroad = np.full((E * P, ROUTE_MAX), -1, np.int32)
        per_map = {}
        for mid, g in enumerate(maps):
            rs = g.meta.get("respawn", [])
            pl = np.zeros((P, 8), np.float32)
            rt = np.full((P, ROUTE_MAX), -1, np.int32)
            rr = np.full((P, ROUTE_MAX), -1, np.int32)
            for k, it in enumerate(rs):
                pl[k, 4] = it["lane"]
                rt[k, :len(it["route"])] = it["route"]
                for j in range(len(it["route"]) - 1):
                    rr[k, j] = road_lut[mid].get((int(it["route"][j]), int(it["route"][j + 1])), -1)
            per_map[mid] = (pl, rt, rr, len(rs))
        for e, sc in enumerate(scenarios):
            pl, rt, rr, n = per_map[sc.map_id]
            ma_place_f[e * P:(e + 1) * P] = pl
            ma_route[e * P:(e + 1) * P] = rt
            ma_rroad[e * P:(e + 1) * P] = rr
            env_i[e, 6] = n
        env_tape = np.ascontiguousarray(ma_tables_tape, np.int32) if ma_tables_tape is not None else np.zeros((E, 4), np.int32)
        assert env_tape.shape[0] % E == 0 and env_tape.shape[1] == 4
    else:
        ma_place_f = np.zeros((1, 8), np.float32)
        ma_route = np.full((1, ROUTE_MAX), -1, np.int32)
        ma_rroad = np.full((1, ROUTE_MAX), -1, np.int32)
        env_tape = np.zeros((1, 4), np.int32)
    arrays.update(ma_place_f=ma_place_f, ma_route=ma_route, ma_rroad=ma_rroad, env_tape=env_tape)
    arrays.update(env_i=env_i, env_trigger=env_trigger, veh_p=veh_p, veh_s=veh_s, veh_c=veh_c, veh_i=veh_i,
                  veh_route=veh_route, veh_idm=veh_idm, veh_navi=veh_navi, obj_f=obj_f, veh_rroad=veh_rroad)
    return arrays


ARRAY_ORDER = ["map_desc", "map_descf", "lane_f", "lane_i", "lane_bb", "road_i", "hull_xy", "line_f", "quad_f",
               "grid_start", "grid_items", "env_i", "env_trigger", "veh_p", "veh_s", "veh_c", "veh_i", "veh_route",
               "veh_idm", "veh_navi", "obj_f", "lgrid_start", "lgrid_items", "veh_rroad", "ma_place_f", "ma_route",
               "ma_rroad", "env_tape"]

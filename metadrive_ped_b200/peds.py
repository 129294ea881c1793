"""Pedestrian scenes for BASELINE config 5 ("pedestrian-dense intersection + IDM traffic").

The reference has the `Pedestrian` traffic participant (component/traffic_participants/pedestrian.py:12-118: Bullet
cylinder r 0.35 m, h 1.75 m, moved kinematically by `set_velocity`, visible to the lidar, `crash_human` on contact,
ignored by the IDM through its exception path policy/idm_policy.py:254-259) but no pedestrian spawn manager or policy
for PG maps (SURVEY.md finding 4), so the *motion model* of this config is defined by this build:

  every pedestrian crosses one two-way road perpendicular to its lanes, from 1 m inside one sidewalk strip to 1 m inside
  the other, at a constant speed drawn from U[0.6, 1.6] m/s, and turns around at the ends of that track.  The turn-around
  is evaluated once per env.step, after the physics sub-steps - the only place where the reference's API
  (`Pedestrian.set_velocity` between two `env.step` calls) could do it.

Object row (scene.Scenario.objects, 10 columns): kind 3, x, y, remaining track length to the next turn-around,
radius, track length, height, lane (-1: pedestrians have no `.lane`), vx, vy.
"""
import math

import numpy as np

from . import scene as sc

PED_RADIUS, PED_HEIGHT = 0.35, 1.75
SIDEWALK_GAP = 0.2      # strip starts 0.2 m beyond the outermost lane edge (component/pgblock/pg_block.py:294-332)
INTO_SIDEWALK = 1.0
JUNCTION_RANGE = 40.0


def crossing_tracks(geo: "sc.MapGeometry"):
    """Candidate crossings: lane 0 of every positive, straight, two-way road outside the first block.  Returns rows
    (lane id, number of lanes per direction, lane width, lane length)."""
    out = []
    key = {(int(r[0]), int(r[1])): k for k, r in enumerate(geo.road_i)}
    nodes = geo.meta.get("nodes", [])
    for k, r in enumerate(geo.road_i):
        if r[4] or chr(int(r[5])) in ("I", ">"):
            continue
        lane = int(r[2])
        row = geo.lane_f[lane]
        if row[0] != 0 or row[2] < 10.0:
            continue
        a, b = nodes[int(r[0])], nodes[int(r[1])]
        opposite = ("-" + b, "-" + a)
        if opposite[0] in nodes and opposite[1] in nodes and (nodes.index(opposite[0]), nodes.index(opposite[1])) in key:
            out.append((lane, int(r[3]), float(row[1]), float(row[2])))
    return out


def place_pedestrians(geo: "sc.MapGeometry", rng, count=16):
    """`count` pedestrian rows for one scene, drawn with `rng` (numpy Generator)."""
    tracks = crossing_tracks(geo)
    if not tracks:
        return np.zeros((0, 10))
    ends = np.array([geo.lane_f[t[0]][10:14] for t in tracks])
    centre = np.concatenate([ends[:, 0:2], ends[:, 2:4]]).mean(0)
    rows = []
    for _ in range(count):
        lane, n, w, length = tracks[int(rng.integers(0, len(tracks)))]
        row = geo.lane_f[lane]
        for _try in range(20):
            lon = float(rng.uniform(2.0, length - 2.0))
            mid = sc.lane_position(row, lon, -w / 2)  # the centre line of the road (left edge of lane 0)
            if np.hypot(*(mid - centre)) <= JUNCTION_RANGE:
                break
        half = n * w + SIDEWALK_GAP + INTO_SIDEWALK
        speed = float(rng.uniform(0.6, 1.6))
        sgn = 1.0 if rng.random() < 0.5 else -1.0
        off = float(rng.uniform(-half, half))
        # lateral unit vector of the lane (positive lateral = right of travel, straight_lane.py:46)
        dx, dy = row[7], row[8]
        lx, ly = dy, -dx
        pos = mid + off * np.array([lx, ly])
        remaining = half - sgn * off
        rows.append([3, pos[0], pos[1], remaining, PED_RADIUS, 2 * half, PED_HEIGHT, -1, sgn * speed * lx, sgn * speed * ly])
    return np.array(rows, np.float64)


def step_pedestrians_host(rows, dt_step):
    """The turn-around rule restated for host-side drivers of the reference env (oracle/gen_golden.py): call once
    after every env.step with dt_step = decision_repeat * physics_world_step_size; returns the indices that turned."""
    turned = []
    for k, r in enumerate(rows):
        speed = math.hypot(r[8], r[9])
        r[3] -= speed * dt_step
        if r[3] <= 0.0:
            r[8], r[9] = -r[8], -r[9]
            r[3] += r[5]
            turned.append(k)
    return turned

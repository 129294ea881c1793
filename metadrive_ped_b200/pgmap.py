"""Procedural map generation on the product side: BIG + the PG blocks I, S, C, r, R, X, T, O.

The reference builds a map at `env.reset(seed)` with BIG ("block incremental generation",
component/algorithm/BIG.py:28-169): blocks are sampled one after another from a seeded numpy `RandomState`,
plugged into a socket of the previous block, checked for crossing the roads that exist already
(utils/pg/utils.py:37-75) and re-sampled / back-tracked when they do.  This module restates that path so that
`MetaDriveEnv(dict(map="S"))`, `map="SCO"`, `map=5` ... run without the reference, and reproduces the reference's
maps number for number: the shipped scenario libraries (exported from the reference, oracle/gen_assets.py) are the
goldens, `tests/test_pgmap.py` regenerates them and compares bit for bit.

What is restated, and where it lives in the reference (paths relative to /root/reference/metadrive):
  seeded streams            utils/random_utils.py:14-110 (sha512-hashed seeds -> RandomState); one draw per block
                            parameter space, every parameter of a block from the SAME uniform
                            (component/pg_space.py:145-330: each Box is seeded alike, base_runnable.py:81-91)
  lanes                     component/lane/straight_lane.py:12-74, circular_lane.py:12-121
  road construction         component/pgblock/create_pg_block_utils.py:19-281
  crossing test             utils/pg/utils.py:37-143
  blocks                    component/pgblock/first_block.py, straight.py, curve.py, ramp.py, intersection.py,
                            t_intersection.py, std_intersection.py, std_t_intersection.py, roundabout.py
  BIG                       component/algorithm/BIG.py, blocks_prob_dist.py
  network bookkeeping       component/road_network/node_road_network.py:68-200, road.py
The arithmetic keeps the reference's operand types on purpose (Python floats where it has them, numpy float64 arrays
where it has those): bit-identical lane tables need the same roundings, not just the same formulas.

Output: `generate(seed, map, ...)` -> scene.MapTable (lane_f, lane_i, road_i, meta) in exactly the layout
oracle/ref_export.export_map writes for a reference map.
"""
import hashlib
import math
import re
import struct
from collections import OrderedDict, deque

import numpy as np

# line types (ids of oracle/ref_export.py / scene.py) and colours
L_NONE, L_BROKEN, L_CONT, L_SIDE, L_GUARD = 0, 1, 2, 3, 4
GREY, YELLOW = 0, 1
DECO_A, DECO_B = "decoration", "decoration_"          # constants.py:93-98
SIDEWALK_WIDTH, SIDEWALK_LINE_DIST = 2, 0.6           # constants.py:319-320
LANE_STREET, LANE_UNSTRUCTURED = 0, 1


# ---------------------------------------------------------------------------------------------- seeded streams
def _bigint(b):
    b += b"\0" * (4 - len(b) % 4)
    return sum(v << (32 * i) for i, v in enumerate(struct.unpack("%dI" % (len(b) // 4), b)))


def seeded_rng(seed):
    """utils/random_utils.get_np_random: RandomState seeded with the 32-bit words of sha512(str(seed))[:8]."""
    seed = int(seed) % 2**64
    big = _bigint(hashlib.sha512(str(seed).encode("utf8")).digest()[:8])
    words = []
    while big > 0:
        big, mod = divmod(big, 2**32)
        words.append(mod)
    rng = np.random.RandomState()
    rng.seed(words or [0])
    return rng


def sample_space(space, seed):
    """ParameterSpace.sample after .seed(seed) (component/pg_space.py:145-166, 380-407; utils/config.py:229-238): the keys
    in sorted order, every entry drawn from its OWN RandomState seeded with the same seed (so all parameters of a block
    share one uniform), boxes in float32, integer boxes floored, 1-element arrays unwrapped to Python scalars."""
    out = OrderedDict()
    for key in sorted(space):
        kind, lo, hi = space[key]
        rng = seeded_rng(seed)
        if kind == "int":
            low, high = np.full((1, ), lo, np.int64), np.full((1, ), hi, np.int64).astype("int64") + 1
            out[key] = int(np.floor(rng.uniform(low=low, high=high, size=(1, ))).astype(np.int64)[0])
        else:
            low, high = np.full((1, ), lo, np.float32), np.full((1, ), hi, np.float32)
            out[key] = float(rng.uniform(low=low, high=high, size=(1, )).astype(np.float32)[0])
    return out


def box(lo, hi):
    return ("box", lo, hi)


def disc(lo, hi):
    return ("int", lo, hi)


def const(v):
    return ("box", v, v)


# ---------------------------------------------------------------------------------------------- small maths
def norm2(x, y):
    return math.sqrt(x**2 + y**2)


def wrap_to_pi(x):
    a = x
    a %= 2 * np.pi
    a -= 2 * np.pi * (a > np.pi)
    return a


class Vec(tuple):
    """utils/math.Vector: a 2-tuple with element-wise arithmetic (CircularLane positions are these, not arrays)."""
    def __sub__(self, o):
        return Vec((self[0] - o[0], self[1] - o[1]))

    def __rsub__(self, o):
        return Vec(o) - self

    def __mul__(self, o):
        if isinstance(o, float) or np.isscalar(o):
            return Vec((self[0] * o, self[1] * o))
        return Vec((self[0] * o[0], self[1] * o[1]))

    __rmul__ = __mul__

    def __add__(self, o):
        if isinstance(o, float) or np.isscalar(o):
            return Vec((self[0] + o, self[1] + o))
        return Vec((self[0] + o[0], self[1] + o[1]))

    def __neg__(self):
        return Vec((-self[0], -self[1]))


# ---------------------------------------------------------------------------------------------- lanes
class Lane:
    kind = LANE_STREET
    line_colors = (GREY, GREY)          # class-level default, replaced per instance (pg_lane.py:10)
    speed_limit = 1000

    def clone(self):
        c = self.__class__.__new__(self.__class__)
        c.__dict__.update(self.__dict__)
        return c

    def width_at(self, _lon):
        return self.width

    def ends_where_starts(self, other, tol=1e-1):   # AbstractLane.is_previous_lane_of
        return norm2(self.end[0] - other.start[0], self.end[1] - other.start[1]) < tol


class SLane(Lane):
    """component/lane/straight_lane.py"""
    def __init__(self, start, end, width=3.5, line_types=(L_BROKEN, L_BROKEN), speed_limit=1000):
        self.start, self.end = np.array(start), np.array(end)
        self.width = width
        self.line_types = list(line_types) if line_types else [L_BROKEN, L_BROKEN]
        self.speed_limit = speed_limit
        self.refresh()

    def refresh(self):
        d = self.end - self.start
        self.length = norm2(d[0], d[1])
        self.heading = math.atan2(self.end[1] - self.start[1], self.end[0] - self.start[0])
        self.direction = (self.end - self.start) / self.length
        self.direction_lateral = np.array([self.direction[1], -self.direction[0]])

    def position(self, lon, lat):
        return self.start + lon * self.direction + lat * self.direction_lateral

    def move(self, start, end):   # StraightLane.reset_start_end
        self.start, self.end = start, end
        self.refresh()

    def local(self, p):
        dx, dy = p[0] - self.start[0], p[1] - self.start[1]
        return (float(dx * self.direction[0] + dy * self.direction[1]),
                float(dx * self.direction_lateral[0] + dy * self.direction_lateral[1]))

    def heading_at(self, _lon):
        return self.heading


class CLane(Lane):
    """component/lane/circular_lane.py"""
    def __init__(self, center, radius, start_phase, angle, clockwise=True, width=3.5, line_types=(L_BROKEN, L_BROKEN),
                 speed_limit=1000):
        assert angle > 0
        self.center = Vec(center)
        self.radius = radius
        self.clockwise = clockwise
        self.start_phase = wrap_to_pi(start_phase)
        self.angle = angle
        self.end_phase = self.start_phase + (-self.angle if clockwise else self.angle)
        self.direction = -1 if clockwise else 1
        self.width = width
        self.line_types = list(line_types)
        self.speed_limit = speed_limit
        self.refresh()

    def refresh(self):
        self.length = abs(self.radius * (self.end_phase - self.start_phase))
        assert self.length > 0
        self.start = self.position(0, 0)
        self.end = self.position(self.length, 0)

    def position(self, lon, lat):
        phi = self.direction * lon / self.radius + self.start_phase
        return self.center + (self.radius + lat * self.direction) * Vec((math.cos(phi), math.sin(phi)))

    def heading_at(self, lon):
        phi = self.direction * lon / self.radius + self.start_phase
        return phi + math.pi / 2 * self.direction

    def local(self, p):
        dx, dy = p[0] - self.center[0], p[1] - self.center[1]
        phase = wrap_to_pi(math.atan2(dy, dx))
        sp, ep = wrap_to_pi(self.start_phase), wrap_to_pi(self.end_phase)
        d_start, d_end = abs(wrap_to_pi(phase - sp)), abs(wrap_to_pi(phase - ep))
        if d_start > np.pi and d_end > np.pi:
            raise ValueError("undetermined position on a circular lane")
        if d_start > d_end:
            diff = self.end_phase - phase if self.clockwise else phase - self.end_phase
            lon = wrap_to_pi(diff) * self.radius + self.length
        else:
            diff = self.start_phase - phase if self.clockwise else phase - self.start_phase
            lon = wrap_to_pi(diff) * self.radius
        return lon, self.direction * (norm2(dx, dy) - self.radius)


# ---------------------------------------------------------------------------------------------- roads / network
def neg_road(road):
    """Road.__neg__ (road_network/road.py:25-30): the road of the opposite direction."""
    s, e = road
    i = e.find("-")
    if i == -1:
        return ("-" + e, "-" + s)
    return (e[i + 1:], s[i + 1:])


def is_negative(road):
    return road[1].find("-") != -1


def road_block_id(road):
    node = road[1] if not is_negative(road) else road[0]
    if re.search(">", node) is not None:
        return ">"
    return re.search("[a-zA-Z$]", node).group(0)


class Net:
    """NodeRoadNetwork's graph: {from: {to: [lanes]}} in insertion order (node_road_network.py:68-200)."""
    def __init__(self):
        self.graph = {}

    def lanes(self, road):
        return self.graph[road[0]][road[1]]

    def add_lane(self, a, b, lane):
        self.graph.setdefault(a, {}).setdefault(b, []).append(lane)

    def deco(self):
        return self.graph[DECO_A][DECO_B] if DECO_A in self.graph else []

    def add(self, other):
        a = set(self.graph) - {DECO_A, DECO_B}
        b = set(other.graph) - {DECO_A, DECO_B}
        if a & b:
            raise ValueError("same start node in two road networks: %s" % (a & b))
        deco = self.deco() + other.deco()
        self.graph.update(dict(other.graph))            # the inner dicts are shared, as in the reference
        if deco:
            self.graph.pop(DECO_A, None)
            self.graph[DECO_A] = {DECO_B: deco}

    def subtract(self, other):
        for k in self.graph.keys() & (other.graph.keys() - {DECO_A, DECO_B}):
            self.graph.pop(k, None)
        if DECO_A in other.graph:
            for lane in other.graph[DECO_A][DECO_B]:
                if lane in self.graph[DECO_A][DECO_B]:
                    self.graph[DECO_A][DECO_B].remove(lane)

    def positive_lanes(self):
        return [lanes for a, d in self.graph.items() for b, lanes in d.items()
                if not is_negative((a, b)) and (a, b) != (DECO_A, DECO_B)]

    def negative_lanes(self):
        return [lanes for a, d in self.graph.items() for b, lanes in d.items()
                if is_negative((a, b)) and (a, b) != (DECO_A, DECO_B)]

    def bfs_paths(self, start, goal):
        """All simple paths, breadth first.  The reference expands `set(successors) - set(path)` in Python's set order
        (node_road_network.py:233-246); every use on this path is order-free (removing all roads between two nodes) or
        has a unique answer, so successors are taken in graph order here."""
        queue = [(start, [start])]
        while queue:
            node, path = queue.pop(0)
            if node not in self.graph:
                yield []
                continue
            for nxt in [n for n in self.graph[node].keys() if n not in path]:
                if nxt == goal:
                    yield path + [nxt]
                elif nxt in self.graph:
                    queue.append((nxt, path + [nxt]))

    def remove_road(self, road):
        out = self.graph[road[0]].pop(road[1])
        if not self.graph[road[0]]:
            self.graph.pop(road[0])
        return out

    def remove_all_roads(self, a, b):
        out = []
        for path in list(self.bfs_paths(a, b)):
            for k, node in enumerate(path[:-1], 1):
                out += self.remove_road((node, path[k]))
        return out


# ---------------------------------------------------------------------------------------------- crossing test
def _bbox(points):
    p = np.array(points)
    return p[:, 0].max(), p[:, 0].min(), p[:, 1].max(), p[:, 1].min()


def lanes_bbox(lanes, extra=3):
    """get_lanes_bounding_box (utils/pg/utils.py:78-143): a few contour points of the road, 3 m beyond its sides."""
    pts = []
    if isinstance(lanes[0], CLane):
        for lane, side in ((lanes[0], -1), (lanes[-1], 1)):
            half = np.pi / 2.0
            pts += [lane.position(0.1, side * (lane.width / 2.0 + extra)),
                    lane.position(lane.length - 0.1, side * (lane.width / 2.0 + extra))]
            sp = (lane.start_phase // half) * half
            sp += half if lane.clockwise else 0
            for k in range(4):
                phi = sp + k * half * lane.direction
                if lane.direction * phi > lane.direction * lane.end_phase:
                    break
                pts.append(lane.center + (lane.radius - side * (lane.width / 2.0 + extra) * lane.direction) *
                           np.array([math.cos(phi), math.sin(phi)]))
    else:
        for lane, side in ((lanes[0], -1), (lanes[-1], 1)):
            pts.append(lane.position(0.1, side * (lane.width / 2.0 + extra)))
            pts.append(lane.position(lane.length - 0.1, side * (lane.width / 2.0 + extra)))
    return _bbox(pts)


def crosses(net, lane, factor=0, ignored=None):
    """check_lane_on_road (utils/pg/utils.py:37-75): does a 1 m sampling of `lane`, shifted sideways by factor * width / 2,
    fall onto a lane of `net`?"""
    for a, d in net.graph.items():
        for b, lanes in d.items():
            if ignored and (a, b) == ignored:
                continue
            if (a, b) == (DECO_A, DECO_B) or len(lanes) == 0:
                continue
            x1, n1, y1, m1 = lanes_bbox(lanes)
            x2, n2, y2, m2 = lanes_bbox([lane])
            if n1 > x2 or n2 > x1 or m1 > y2 or m2 > y1:
                continue
            for other in lanes:
                for i in range(1, int(lane.length), 1):
                    p = lane.position(i, factor * lane.width_at(i) / 2.0)
                    lon, lat = other.local(p)
                    if math.fabs(lat) <= other.width_at(lon) / 2.0 and 0 <= lon <= other.length:
                        return True
    return False


# ---------------------------------------------------------------------------------------------- road construction
def bend_then_straight(prev, follow_len, radius, angle, clockwise=True, width=3.5, line_types=(L_BROKEN, L_BROKEN), speed_limit=20):
    """create_bend_straight (create_pg_block_utils.py:19-48)"""
    sign = 1 if clockwise else -1
    center = prev.position(prev.length, sign * radius)
    x, y = prev.direction_lateral
    start_phase = np.arctan2(y, x) + (np.pi if clockwise else 0)
    bend = CLane(center, radius, start_phase, angle, clockwise, width, line_types, speed_limit)
    length = 2 * radius * angle / 2
    bend_end = bend.position(length, 0)
    v = bend_end - center
    n = norm2(v[0], v[1])
    perp = ((-v[1] / n, v[0] / n), (v[1] / n, -v[0] / n))
    nxt = np.asarray(perp[0] if not clockwise else perp[1])
    straight = SLane(bend_end, nxt * follow_len + bend_end, width, line_types, speed_limit)
    return bend, straight


def extend_straight(lane, extra, line_types, kind=None):
    """ExtendStraightLane (create_pg_block_utils.py:186-204)"""
    new = lane.clone()
    new.start = lane.end
    new.end = lane.position(lane.length + extra, 0)
    new.line_types = line_types
    new.refresh()
    if kind is not None:
        new.kind = kind
    return new


def road_from(lane, lane_num, road, block_net, world_net, toward_smaller=True, ignore=None, center_line=None,
              one_side=True, side_line=None, inner_line=None, center_color=None, kind=None):
    """CreateRoadFrom (create_pg_block_utils.py:51-183): `lane` is the outermost (toward_smaller) or innermost lane;
    returns True when the new road does not cross the existing ones."""
    center_line = L_CONT if center_line is None else center_line
    side_line = L_SIDE if side_line is None else side_line
    inner_line = L_BROKEN if inner_line is None else inner_line
    center_color = YELLOW if center_color is None else center_color
    lane_num -= 1
    origin = lane
    lanes = []
    w = lane.width_at(0)
    for i in range(lane_num, 0, -1):
        side = lane.clone()
        if isinstance(lane, SLane):
            shift = -w if toward_smaller else w
            s, e = side.position(0, shift), side.position(side.length, shift)
            side.start, side.end = s, e
        else:
            r1 = lane.radius
            if not toward_smaller:
                r2 = r1 - w if lane.clockwise else r1 + w
            else:
                r2 = r1 + w if lane.clockwise else r1 - w
            side.radius = r2
            side.refresh()
        if i == 1:
            side.line_types = [center_line, inner_line] if toward_smaller else [inner_line, side_line]
        else:
            side.line_types = [inner_line, inner_line]
        lanes.append(side)
        lane = side
    if toward_smaller:
        lanes.reverse()
        lanes.append(origin)
        origin.line_types = [inner_line if len(lanes) > 1 else center_line, side_line]
    else:
        lanes.insert(0, origin)
        if len(lanes) > 1:
            origin.line_types = (origin.line_types[0], lanes[-1].line_types[0])
    factor = (SIDEWALK_WIDTH + SIDEWALK_LINE_DIST + w / 2.0) * 2.0 / w
    if not one_side:
        ok = not (crosses(world_net, origin, factor, ignore) or crosses(world_net, lanes[0], -0.95, ignore))
    else:
        ok = not crosses(world_net, origin, factor, ignore)
    for l in lanes:
        block_net.add_lane(road[0], road[1], l)
        if kind is not None:
            l.kind = kind
    if lane_num == 0:
        lanes[-1].line_types = [center_line, side_line]
    lanes[0].line_colors = (center_color, GREY)
    return ok


def adverse_road(road, block_net, world_net, ignore=None, center_line=None, side_line=None, inner_line=None,
                 center_color=None, kind=None):
    """CreateAdverseRoad (create_pg_block_utils.py:211-281): the mirror road for the opposite direction."""
    center_line = L_CONT if center_line is None else center_line
    side_line = L_SIDE if side_line is None else side_line
    inner_line = L_BROKEN if inner_line is None else inner_line
    center_color = YELLOW if center_color is None else center_color
    lanes = block_net.lanes(road)
    ref = lanes[-1]
    num = len(lanes) * 2
    w = ref.width_at(0)
    if isinstance(ref, SLane):
        s = ref.position(lanes[-1].length, -(num - 1) * w)
        e = ref.position(0, -(num - 1) * w)
        sym = SLane(s, e, w, lanes[-1].line_types, ref.speed_limit)
    else:
        cw = not ref.clockwise
        radius = ref.radius + (num - 1) * w if not cw else ref.radius - (num - 1) * w
        sym = CLane(ref.center, radius, ref.end_phase, ref.angle, cw, w, ref.line_types, ref.speed_limit)
    ok = road_from(sym, int(num / 2), neg_road(road), block_net, world_net, ignore=ignore, side_line=side_line,
                   inner_line=inner_line, center_line=center_line, center_color=center_color, kind=kind)
    block_net.lanes(road)[0].line_colors = (center_color, GREY)
    return ok


def two_way_road(road, block_net, world_net, new_road, center_line=None, side_line=None, inner_line=None):
    """CreateTwoWayRoad (create_pg_block_utils.py:284-349): a road in the reverse direction lying ON `road` (named `new_road`)"""
    lanes = block_net.lanes(road)
    ref = lanes[-1]
    num = len(lanes)
    w = ref.width_at(0)
    if isinstance(ref, SLane):
        sym = SLane(ref.position(lanes[-1].length, -(num - 1) * w), ref.position(0, -(num - 1) * w), w, lanes[-1].line_types,
                    ref.speed_limit)
    else:
        cw = not ref.clockwise
        radius = ref.radius + (num - 1) * w if not cw else ref.radius - (num - 1) * w
        sym = CLane(ref.center, radius, ref.end_phase, ref.angle, cw, w, ref.line_types, ref.speed_limit)
    return road_from(sym, num, new_road, block_net, world_net, side_line=side_line, inner_line=inner_line, center_line=center_line)


def wave_lanes(prev, lateral, wave_len, last_len, width, toward_left=True):
    """create_wave_lanes (create_pg_block_utils.py:352-380): two opposite arcs that shift a lane sideways by `lateral`
    over `wave_len`, and the straight lane that follows"""
    angle = np.pi - 2 * np.arctan(wave_len / (2 * lateral))
    radius = wave_len / (2 * math.sin(angle))
    c1, mid = bend_then_straight(prev, 10, radius, angle, False if toward_left else True, width, [L_NONE, L_NONE])
    mid.move(mid.position(-10, 0), mid.position(mid.length - 10, 0))
    c2, straight = bend_then_straight(mid, last_len, radius, angle, True if toward_left else False, width, [L_NONE, L_NONE])
    return c1, c2, straight


# ---------------------------------------------------------------------------------------------- blocks
class Socket:
    def __init__(self, positive, negative=None, fake_positive=None, fake_negative=None):
        self.positive, self.negative = positive, negative
        self.index = None
        # BidirectionSocket (pgblock/bidirection.py:57-67): the next block is built from two stand-in lanes instead of the
        # (overlapping) lanes of the socket roads
        self.fake_positive, self.fake_negative = fake_positive, fake_negative

    def positive_lanes(self, net):
        return [self.fake_positive] if self.fake_positive is not None else net.lanes(self.positive)


class Block:
    """PGBlock (component/pgblock/pg_block.py:60-250) + BaseBlock.construct_block (block/base_block.py:95-131)."""
    ID, SPACE = None, {}

    def __init__(self, index, pre_socket, world_net, seed):
        self.index = index
        self.name = str(index) + self.ID
        self.world = world_net
        self.net = Net()
        self.rng = seeded_rng(seed)
        self.trials = 0
        self.sockets = OrderedDict()
        self.respawn_roads = []
        self.pre_socket = pre_socket
        self.part, self.road_no = 0, 0
        self.cfg = {}
        self.resample()                                    # BaseRunnable.__init__ samples once (base_runnable.py:26)
        if index != 0:
            self.pos_lanes = pre_socket.positive_lanes(world_net)
            self.n_pos = len(self.pos_lanes)
            self.basic = self.pos_lanes[-1]
            self.lane_width = self.basic.width_at(0)

    def resample(self):
        seed = self.rng.randint(low=0, high=int(1e6))
        self.cfg.update(sample_space(self.SPACE, seed))

    # node names: <block index><ID><part>_<road>_  (pg_block.py:216-233)
    def node(self, part, road):
        return str(self.index) + self.ID + str(part) + "_" + str(road) + "_"

    def set_part(self, p):
        self.part, self.road_no = p, 0

    def new_node(self):
        self.road_no += 1
        return self.node(self.part, self.road_no - 1)

    def add_socket(self, s):
        if s.index is None:
            s.index = "%s-socket%d" % (self.name, len(self.sockets))
        self.sockets[s.index] = s

    def socket_from(self, road):
        return Socket(road, neg_road(road))

    def get_socket(self, i):
        return self.sockets[list(self.sockets)[i]]

    def clear(self):
        if len(self.world.graph) > 0:
            self.world.subtract(self.net)
        self.net.graph.clear()
        self.part = self.road_no = 0
        self.respawn_roads.clear()
        self.sockets.clear()

    def construct(self, extra=None):
        """BaseBlock.construct_block (block/base_block.py:95-131); `extra` = construct_from_config's overrides"""
        self.resample()
        if extra:
            self.cfg.update(extra)
        self.clear()
        self.trials += 1
        ok = self.plug()
        self.world.add(self.net)
        return ok

    def respawn_lanes(self):
        return [self.net.lanes(r) for r in self.respawn_roads]

    def spawn_lanes(self):
        """get_intermediate_spawn_lanes (pg_block.py:235-241): where the traffic manager may put vehicles"""
        out = self.net.positive_lanes()
        for lanes in self.respawn_lanes():
            if lanes not in out:
                out.append(lanes)
        return out


class FirstBlock(Block):
    """component/pgblock/first_block.py"""
    ID = "I"
    ENTRANCE = 10

    def __init__(self, world_net, lane_width, lane_num, length):
        super().__init__(0, Socket((DECO_A, DECO_B), (DECO_A, DECO_B)), world_net, 0)
        basic = SLane([0, 0], [self.ENTRANCE, 0], line_types=(L_BROKEN, L_SIDE), width=lane_width)
        spawn = (">", ">>")
        road_from(basic, lane_num, spawn, self.net, world_net)
        adverse_road(spawn, self.net, world_net)
        nxt = extend_straight(basic, length - self.ENTRANCE, [L_BROKEN, L_SIDE])
        other = (">>", ">>>")
        road_from(nxt, lane_num, other, self.net, world_net)
        adverse_road(other, self.net, world_net)
        world_net.add(self.net)
        s = self.socket_from(other)
        s.index = "%s-socket%d" % (self.name, 0)
        self.add_socket(s)
        self.respawn_roads = [other]

    def clear(self):
        pass


class Straight(Block):
    """component/pgblock/straight.py"""
    ID = "S"
    SPACE = {"length": box(40.0, 80.0)}

    def plug(self):
        self.set_part(0)
        new = extend_straight(self.basic, self.cfg["length"], [L_BROKEN, L_SIDE])
        road = (self.pre_socket.positive[1], self.new_node())
        ok = road_from(new, self.n_pos, road, self.net, self.world)
        ok = adverse_road(road, self.net, self.world) and ok
        self.add_socket(Socket(road, neg_road(road)))
        return ok


class Curve(Block):
    """component/pgblock/curve.py"""
    ID = "C"
    SPACE = {"length": box(40.0, 80.0), "radius": box(25.0, 60.0), "angle": box(45, 135), "dir": disc(0, 1)}

    def plug(self):
        c = self.cfg
        road = (self.pre_socket.positive[1], self.new_node())
        bend, straight = bend_then_straight(self.basic, c["length"], c["radius"], np.deg2rad(c["angle"]), c["dir"],
                                            width=self.basic.width, line_types=(L_BROKEN, None))
        ok = road_from(bend, self.n_pos, road, self.net, self.world)
        ok = adverse_road(road, self.net, self.world) and ok
        road = (road[1], self.new_node())
        ok = road_from(straight, self.n_pos, road, self.net, self.world) and ok
        ok = adverse_road(road, self.net, self.world) and ok
        self.add_socket(self.socket_from(road))
        return ok


class Ramp(Block):
    SPACE = {"length": box(20, 40)}
    RADIUS, ANGLE, CONNECT, RAMP_LEN, SPEED = 40, 10, 20, 15, 12
    TYPE = (L_CONT, L_CONT)


class InRamp(Ramp):
    """InRampOnStraight (component/pgblock/ramp.py:36-205)"""
    ID = "r"
    EXTRA, SOCKET_LEN = 10, 20

    def plug(self):
        acc_len = self.cfg["length"]
        ok = True
        self.set_part(0)
        sin_a, cos_a = math.sin(np.deg2rad(self.ANGLE)), math.cos(np.deg2rad(self.ANGLE))
        lon_len = sin_a * self.RADIUS * 2 + cos_a * self.CONNECT + self.RAMP_LEN
        extend = extend_straight(self.basic, lon_len + self.EXTRA, [L_BROKEN, L_CONT])
        extend_road = (self.pre_socket.positive[1], self.new_node())
        ok = road_from(extend, self.n_pos, extend_road, self.net, self.world, side_line=L_CONT) and ok
        self.net.lanes(extend_road)[-1].line_types = [L_BROKEN if self.n_pos != 1 else L_CONT, L_CONT]
        ok = adverse_road(extend_road, self.net, self.world) and ok
        self.net.lanes(neg_road(extend_road))[-1].line_types = [L_NONE if self.n_pos == 1 else L_BROKEN, L_SIDE]
        acc_side = extend_straight(extend, acc_len + self.lane_width, [extend.line_types[0], L_SIDE])
        acc_road = (extend_road[1], self.new_node())
        ok = road_from(acc_side, self.n_pos, acc_road, self.net, self.world, side_line=L_CONT) and ok
        ok = adverse_road(acc_road, self.net, self.world) and ok
        self.net.lanes(acc_road)[-1].line_types = [L_CONT if self.n_pos == 1 else L_BROKEN, L_BROKEN]
        socket_side = extend_straight(acc_side, self.SOCKET_LEN, acc_side.line_types)
        socket_road = (acc_road[1], self.new_node())
        ok = road_from(socket_side, self.n_pos, socket_road, self.net, self.world, side_line=L_CONT) and ok
        ok = adverse_road(socket_road, self.net, self.world) and ok
        self.add_socket(self.socket_from(socket_road))

        self.set_part(1)
        lat = (1 - cos_a) * self.RADIUS * 2 + sin_a * self.CONNECT
        end = extend.position(self.EXTRA + self.RAMP_LEN, lat + self.lane_width)
        start = extend.position(self.EXTRA, lat + self.lane_width)
        straight = SLane(start, end, self.lane_width, self.TYPE, speed_limit=self.SPEED)
        straight_road = (self.new_node(), self.new_node())
        self.net.add_lane(straight_road[0], straight_road[1], straight)
        ok = (not crosses(self.world, straight, 0.95)) and ok
        self.respawn_roads.append(straight_road)
        bend1, connect = bend_then_straight(straight, self.CONNECT, self.RADIUS, np.deg2rad(self.ANGLE), False,
                                            self.lane_width, self.TYPE, speed_limit=self.SPEED)
        bend1_road = (straight_road[1], self.new_node())
        connect_road = (bend1_road[1], self.new_node())
        self.net.add_lane(bend1_road[0], bend1_road[1], bend1)
        self.net.add_lane(connect_road[0], connect_road[1], connect)
        ok = (not crosses(self.world, bend1, 0.95)) and ok
        ok = (not crosses(self.world, connect, 0.95)) and ok
        bend2, acc = bend_then_straight(connect, acc_len, self.RADIUS, np.deg2rad(self.ANGLE), True, self.lane_width,
                                        self.TYPE, speed_limit=self.SPEED)
        acc.line_types = [L_BROKEN, L_CONT]
        bend2_road = (connect_road[1], self.node(0, 0))
        self.net.add_lane(bend2_road[0], bend2_road[1], bend2)
        self.net.add_lane(acc_road[0], acc_road[1], acc)
        ok = (not crosses(self.world, bend2, 0.95)) and ok
        ok = (not crosses(self.world, acc, 0.95)) and ok
        merge, _ = bend_then_straight(acc, 10, self.lane_width / 2, np.pi / 2, False, self.lane_width, (L_BROKEN, L_CONT))
        self.net.add_lane(DECO_A, DECO_B, merge)
        return ok

    def spawn_lanes(self):
        on_socket = self.net.lanes(self.get_socket(0).positive)[0]
        return [lanes for lanes in super().spawn_lanes() if on_socket not in lanes]


class OutRamp(Ramp):
    """OutRampOnStraight (component/pgblock/ramp.py:208-374)"""
    ID = "R"
    EXTRA_LEN = 15

    def plug(self):
        ok = True
        sin_a, cos_a = math.sin(np.deg2rad(self.ANGLE)), math.cos(np.deg2rad(self.ANGLE))
        lon_len = sin_a * self.RADIUS * 2 + cos_a * self.CONNECT + self.RAMP_LEN + self.EXTRA_LEN
        self.set_part(0)
        dec_len = self.cfg["length"]
        dec = extend_straight(self.basic, dec_len + self.lane_width, [self.basic.line_types[0], L_SIDE])
        dec_road = (self.pre_socket.positive[1], self.new_node())
        ok = road_from(dec, self.n_pos, dec_road, self.net, self.world, side_line=L_CONT) and ok
        ok = adverse_road(dec_road, self.net, self.world) and ok
        dec_right = self.net.lanes(dec_road)[-1]
        dec_right.line_types = [L_CONT if self.n_pos == 1 else L_BROKEN, L_BROKEN]
        extend = extend_straight(dec_right, lon_len, [dec_right.line_types[0], L_CONT])
        extend_road = (dec_road[1], self.new_node())
        ok = road_from(extend, self.n_pos, extend_road, self.net, self.world, side_line=L_CONT) and ok
        ok = adverse_road(extend_road, self.net, self.world) and ok
        self.net.lanes(neg_road(extend_road))[-1].line_types = [L_NONE if self.n_pos == 1 else L_BROKEN, L_SIDE]
        self.add_socket(self.socket_from(extend_road))

        self.set_part(1)
        s = dec_right.position(self.lane_width, self.lane_width)
        e = dec_right.position(dec_right.length, self.lane_width)
        dec_side = SLane(s, e, self.lane_width, (L_BROKEN, L_CONT))
        self.net.add_lane(dec_road[0], dec_road[1], dec_side)
        ok = (not crosses(self.world, dec_side, 0.95)) and ok
        bend1, connect = bend_then_straight(dec_side, self.CONNECT, self.RADIUS, np.deg2rad(self.ANGLE), True,
                                            self.lane_width, self.TYPE, speed_limit=self.SPEED)
        bend1_road = (dec_road[1], self.new_node())
        connect_road = (bend1_road[1], self.new_node())
        self.net.add_lane(bend1_road[0], bend1_road[1], bend1)
        self.net.add_lane(connect_road[0], connect_road[1], connect)
        ok = (not crosses(self.world, bend1, 0.95)) and ok
        ok = (not crosses(self.world, connect, 0.95)) and ok
        bend2, straight = bend_then_straight(connect, self.RAMP_LEN, self.RADIUS, np.deg2rad(self.ANGLE), False,
                                             self.lane_width, self.TYPE, speed_limit=self.SPEED)
        bend2_road = (connect_road[1], self.new_node())
        straight_road = (bend2_road[1], self.new_node())
        self.net.add_lane(bend2_road[0], bend2_road[1], bend2)
        self.net.add_lane(straight_road[0], straight_road[1], straight)
        ok = (not crosses(self.world, bend2, 0.95)) and ok
        ok = (not crosses(self.world, straight, 0.95)) and ok
        tool = SLane(dec_side.end, dec_side.start, dec_side.width)
        deco, _ = bend_then_straight(tool, 10, self.lane_width / 2, np.pi / 2, True, width=self.lane_width,
                                     line_types=(L_CONT, L_BROKEN))
        self.net.add_lane(DECO_A, DECO_B, deco)
        return ok


class InterSection(Block):
    """StdInterSection = InterSection with change_lane_num fixed to 0 (component/pgblock/intersection.py,
    std_intersection.py)"""
    ID = "X"
    SPACE = {"radius": const(10), "change_lane_num": disc(0, 1), "decrease_increase": disc(0, 1)}
    ANGLE, EXIT_LEN = 90, 35
    u_turn = False          # InterSection.enable_u_turn (the multi-agent intersection map turns it on)
    std = True              # StdInterSection: the lane number never changes across the junction

    def plug(self):
        if self.std:
            self.cfg["change_lane_num"] = 0
        return self.plug_x()

    def plug_x(self):
        c = self.cfg
        radius = c["radius"]
        sign = -1 if c["decrease_increase"] == 0 else 1
        if self.n_pos <= 1:
            sign = 1
        elif self.n_pos >= 4:
            sign = -1
        self.n_cross = self.n_pos + sign * c["change_lane_num"]
        ok = True
        attach = self.pre_socket.positive
        attach_lanes = self.world.lanes(attach)
        nodes = deque([self.node(0, 0), self.node(1, 0), self.node(2, 0), self.pre_socket.negative[0]])
        for i in range(4):
            right, good = self._part(attach_lanes, attach, radius, nodes, i)
            ok = ok and good
            if i != 3:
                n = self.n_pos if i == 1 else self.n_cross
                exit_road = (self.node(i, 0), self.node(i, 1))
                ok = road_from(right, n, exit_road, self.net, self.world) and ok
                ok = adverse_road(exit_road, self.net, self.world) and ok
                s = Socket(exit_road, neg_road(exit_road))
                self.respawn_roads.append(s.negative)
                self.add_socket(s)
                attach = neg_road(exit_road)
                attach_lanes = self.net.lanes(attach)
        return ok

    def _part(self, attach_lanes, attach, radius, nodes, part):
        n = self.n_cross if part in (0, 2) else self.n_pos
        ok = True
        left = attach_lanes[0]
        self._left_turn(radius, n, left, attach, nodes, part)
        if self.u_turn:      # _create_u_turn (intersection.py:213-235): a half circle back onto the opposite road
            lanes = self.net.lanes(attach) if part != 0 else self.pos_lanes
            bend, _ = bend_then_straight(lanes[0], 0.1, self.lane_width / 2, np.deg2rad(180), False, lanes[0].width_at(0),
                                         (L_NONE, L_NONE))
            road_from(bend, len(lanes), (attach[1], neg_road(attach)[0]), self.net, self.world, toward_smaller=False,
                      center_line=L_NONE, side_line=L_NONE, inner_line=L_NONE, kind=LANE_UNSTRUCTURED)
        on_road = list(attach_lanes)
        through = 2 * radius + (2 * n - 1) * on_road[0].width_at(0)
        for l in on_road:
            self.net.add_lane(attach[1], nodes[1], extend_straight(l, through, (L_NONE, L_NONE), kind=LANE_UNSTRUCTURED))
        right = on_road[-1]
        bend, straight = bend_then_straight(right, self.EXIT_LEN, radius, np.deg2rad(self.ANGLE), True, right.width_at(0),
                                            (L_NONE, L_SIDE))
        ok = (not crosses(self.world, bend, 1)) and ok
        road_from(bend, min(self.n_pos, self.n_cross), (attach[1], nodes[0]), self.net, self.world, toward_smaller=True,
                  side_line=L_SIDE, inner_line=L_NONE, center_line=L_NONE, kind=LANE_UNSTRUCTURED)
        nodes.rotate(-1)
        straight.line_types = [L_BROKEN, L_SIDE]
        return straight, ok

    def _left_turn(self, radius, n, left, attach, nodes, part):
        r = radius + n * left.width_at(0)
        diff = self.n_cross - self.n_pos
        k = min(self.n_pos, self.n_cross)
        if ((part in (1, 3)) and diff > 0) or ((part in (0, 2)) and diff < 0):
            diff = abs(diff)
            bend, extra = bend_then_straight(left, self.lane_width * diff, r, np.deg2rad(self.ANGLE), False, left.width_at(0),
                                             (L_NONE, L_NONE))
            start = nodes[2]
            pre = start + "extra"
            road_from(bend, k, (attach[1], pre), self.net, self.world, toward_smaller=False, center_line=L_NONE,
                      side_line=L_NONE, inner_line=L_NONE, kind=LANE_UNSTRUCTURED)
            road_from(extra, k, (pre, start), self.net, self.world, toward_smaller=False, center_line=L_NONE,
                      side_line=L_NONE, inner_line=L_NONE)
        else:
            bend, _ = bend_then_straight(left, self.EXIT_LEN, r, np.deg2rad(self.ANGLE), False, left.width_at(0),
                                         (L_NONE, L_NONE))
            road_from(bend, k, (attach[1], nodes[2]), self.net, self.world, toward_smaller=False, center_line=L_NONE,
                      side_line=L_NONE, inner_line=L_NONE, kind=LANE_UNSTRUCTURED)

    def get_socket(self, i):
        s = super().get_socket(i)
        if s.negative in self.respawn_roads:            # intersection.py:178-182: a used socket stops respawning traffic
            self.respawn_roads.remove(s.negative)
        return s

    def spawn_lanes(self):
        return self.respawn_lanes()


class TInterSection(InterSection):
    """StdTInterSection (component/pgblock/t_intersection.py, std_t_intersection.py): an X with one arm removed"""
    ID = "T"
    SPACE = {"radius": const(10), "t_type": disc(0, 2), "change_lane_num": disc(0, 1), "decrease_increase": disc(0, 1)}

    def plug(self):
        self.cfg["change_lane_num"] = 0
        ok = self.plug_x()
        self._drop_arm()
        return ok

    def _drop_arm(self):
        t = self.cfg["t_type"]
        self.add_socket(self.pre_socket)
        key = lambda i: "%s-socket%d" % (self.name, i)
        gone = self.sockets[key(t)]
        start, end = gone.negative[1], gone.positive[0]
        for i in range(4):
            if i == t:
                continue
            s = self.sockets[key(i) if i < 3 else self.pre_socket.index]
            exit_node = s.positive[0] if i != 3 else s.negative[0]
            self.net.remove_all_roads(start, exit_node)
            entry_node = s.negative[1] if i != 3 else s.positive[1]
            self.net.remove_all_roads(entry_node, end)
        self._restyle(t)
        self.sockets.pop(self.pre_socket.index)
        s = self.sockets.pop(key(t))
        self.net.remove_all_roads(s.positive[0], s.positive[1])
        self.net.remove_all_roads(s.negative[0], s.negative[1])
        self.respawn_roads.remove(s.negative)

    def _restyle(self, t):
        socks = list(self.sockets.values())
        nxt, last = socks[(t + 1) % 4], socks[(t + 3) % 4]
        nxt_pos, nxt_neg, last_pos, last_neg = nxt.positive, nxt.negative, last.positive, last.negative
        if t == 2:
            nxt_pos, nxt_neg = nxt.negative, nxt.positive
        if t == 0:
            last_pos, last_neg = last.negative, last.positive
        for i, road in enumerate([(last_neg[1], nxt_pos[0]), (nxt_neg[1], last_pos[0])]):
            lanes = self.net.lanes(road)
            outside = L_SIDE if i == 0 else L_NONE
            for k, lane in enumerate(lanes):
                lane.line_types = [L_NONE, L_NONE] if k != len(lanes) - 1 else [L_NONE, outside]
                if k == 0:
                    lane.line_colors = (YELLOW, GREY)
                    if i == 1:
                        lane.line_types[0] = L_NONE


class Roundabout(Block):
    """component/pgblock/roundabout.py"""
    ID = "O"
    SPACE = {"exit_radius": box(5, 15), "inner_radius": box(15, 45), "angle": const(60)}
    EXIT_LEN = 35

    def plug(self):
        self.inner_places = []
        c = self.cfg
        ok = True
        attach = self.pre_socket.positive
        for i in range(4):
            exit_road, good = self._part(attach, i, c["exit_radius"], c["inner_radius"], c["angle"])
            ok = ok and good
            if i < 3:
                ok = adverse_road(exit_road, self.net, self.world) and ok
                attach = neg_road(exit_road)
        for s in self.sockets.values():
            self.respawn_roads.append(s.negative)
        return ok

    def _style_ring(self, road, first=False):
        for k, lane in enumerate(self.net.lanes(road)):
            if first:
                if k == 0:
                    lane.line_types = [L_CONT, L_BROKEN] if self.n_pos > 1 else [L_CONT, L_NONE]
                else:
                    lane.line_types = [L_BROKEN, L_BROKEN]
            else:
                lane.line_types = [L_NONE, L_SIDE] if k == self.n_pos - 1 else [L_NONE, L_NONE]

    def _part(self, road, part, r_exit, r_inner, angle):
        ok = True
        self.set_part(part)
        r_big = (self.n_pos * 2 - 1) * self.lane_width + r_inner
        seg = (road[1], self.new_node())
        lanes = self.world.lanes(road) if part == 0 else self.net.lanes(road)
        right = lanes[-1]
        bend, straight = bend_then_straight(right, 10, r_exit, np.deg2rad(angle), True, self.lane_width, (L_BROKEN, L_SIDE))
        skip = (self.node((part + 3) % 4, 0), self.node((part + 3) % 4, 0))
        ok = road_from(bend, self.n_pos, seg, self.net, self.world, ignore=skip) and ok
        self._style_ring(seg)
        tool = SLane(straight.position(-5, 0), straight.position(0, 0))
        bend, to_next = bend_then_straight(tool, 10, r_big, np.deg2rad(2 * angle - 90), False, self.lane_width,
                                           (L_BROKEN, L_SIDE))
        seg = (seg[1], self.new_node())
        ok = road_from(bend, self.n_pos, seg, self.net, self.world) and ok
        self.inner_places.append(self.net.lanes(seg))
        tool = SLane(to_next.position(-5, 0), to_next.position(0, 0))
        bend, straight = bend_then_straight(tool, self.EXIT_LEN, r_exit, np.deg2rad(angle), True, self.lane_width,
                                            (L_BROKEN, L_SIDE))
        seg = (seg[1], self.new_node() if part < 3 else self.pre_socket.negative[0])
        ok = road_from(bend, self.n_pos, seg, self.net, self.world) and ok
        self._style_ring(seg)
        exit_road = (seg[1], self.new_node())
        if part < 3:
            ok = road_from(straight, self.n_pos, exit_road, self.net, self.world) and ok
            self.add_socket(self.socket_from(exit_road))
        ring = (self.node(part, 1), self.node((part + 1) % 4, 0))
        tool = SLane(to_next.position(-6, 0), to_next.position(0, 0))
        beneath = (self.n_pos * 2 - 1) * self.lane_width / 2 + r_exit
        r_seg = beneath / math.cos(np.deg2rad(angle)) - r_exit
        bend, _ = bend_then_straight(tool, 5, r_seg, np.deg2rad(180 - 2 * angle), False, self.lane_width, (L_BROKEN, L_SIDE))
        road_from(bend, self.n_pos, ring, self.net, self.world)
        self._style_ring(ring, first=True)
        return exit_road, ok

    def get_socket(self, i):
        s = super().get_socket(i)
        if s.negative in self.respawn_roads:            # roundabout.py:190-194
            self.respawn_roads.remove(s.negative)
        return s

    def spawn_lanes(self):
        return self.respawn_lanes() + self.inner_places


class Bottleneck(Block):
    """pgblock/bottleneck.py: Merge ("y") narrows the road by `lane_num` lanes, Split ("Y") widens it"""
    SPACE = {"length": box(20, 50), "lane_num": disc(1, 2), "bottle_len": const(20), "solid_center_line": const(0)}

    def spawn_lanes(self):
        return [lanes for lanes in super().spawn_lanes() if isinstance(lanes[0], SLane)]

    def _single(self, lane, road, side):
        return road_from(lane, 1, road, self.net, self.world, center_line=L_NONE, side_line=side, inner_line=L_NONE)


class Merge(Bottleneck):
    ID = "y"

    def plug(self):
        c = self.cfg
        ok = True
        center = L_CONT if c["solid_center_line"] else L_BROKEN
        blen = c["bottle_len"]
        start = self.pre_socket.positive[1]
        n_straight = max(1, int(self.n_pos - c["lane_num"]))
        n_wave = self.n_pos - n_straight
        ref = extend_straight(self.pos_lanes[n_straight - 1], blen, [L_NONE, L_NONE])
        straight_road = (start, self.node(0, 0))
        side = L_SIDE if n_wave == 0 else L_NONE
        ok = road_from(ref, n_straight, straight_road, self.net, self.world, center_line=center, side_line=side,
                       inner_line=L_NONE) and ok
        ok = adverse_road(straight_road, self.net, self.world, inner_line=L_NONE, side_line=side, center_line=center) and ok
        ref = extend_straight(ref, c["length"], [L_NONE, L_NONE])
        socket_road = (self.node(0, 0), self.node(0, 1))
        ok = road_from(ref, n_straight, socket_road, self.net, self.world, center_line=center, side_line=L_SIDE,
                       inner_line=L_BROKEN) and ok
        ok = adverse_road(socket_road, self.net, self.world, inner_line=L_BROKEN, side_line=L_SIDE, center_line=center) and ok
        neg_socket = neg_road(socket_road)
        self.add_socket(Socket(socket_road, neg_socket))
        for index, lane in enumerate(self.pos_lanes[n_straight:], 1):
            lateral = index * self.lane_width / 2
            inner = self.node(1, index)
            side = L_SIDE if index == self.n_pos - n_straight else L_NONE
            c1, c2, _ = wave_lanes(lane, lateral, blen, 5, self.lane_width)
            road_1, road_2 = (start, inner), (inner, self.node(0, 0))
            ok = self._single(c1, road_1, side) and ok
            ok = self._single(c2, road_2, side) and ok
            lane = self.net.lanes(neg_socket)[-1]
            c2, c1, _ = wave_lanes(lane, lateral, blen, 5, self.lane_width, False)
            ok = self._single(c2, neg_road(road_2), side) and ok
            ok = self._single(c1, neg_road(road_1), side) and ok
        return ok


class Split(Bottleneck):
    ID = "Y"

    def plug(self):
        c = self.cfg
        ok = True
        blen = c["bottle_len"]
        center = L_CONT if c["solid_center_line"] else L_BROKEN
        n_wave = c["lane_num"]
        start = self.pre_socket.positive[1]
        n_straight = self.n_pos
        total = n_straight + n_wave
        ref = extend_straight(self.pos_lanes[n_straight - 1], blen, [L_NONE, L_NONE])
        straight_road = (start, self.node(0, 0))
        ok = road_from(ref, n_straight, straight_road, self.net, self.world, center_line=center, side_line=L_NONE,
                       inner_line=L_NONE) and ok
        ok = adverse_road(straight_road, self.net, self.world, inner_line=L_NONE, side_line=L_NONE, center_line=center) and ok
        lane = self.pos_lanes[-1]
        socket_ref = None
        for index in range(1, n_wave + 1):
            lateral = index * self.lane_width / 2
            inner = self.node(1, index)
            side = L_SIDE if index == n_wave else L_NONE
            c1, c2, straight = wave_lanes(lane, lateral, blen, c["length"], self.lane_width, False)
            if index == n_wave:
                socket_ref = straight
            ok = self._single(c1, (start, inner), side) and ok
            ok = self._single(c2, (inner, self.node(0, 0)), side) and ok
        socket_road = (self.node(0, 0), self.node(0, 1))
        ok = road_from(socket_ref, total, socket_road, self.net, self.world, center_line=L_CONT, side_line=L_SIDE,
                       inner_line=L_BROKEN) and ok
        ok = adverse_road(socket_road, self.net, self.world, inner_line=L_BROKEN, side_line=L_SIDE, center_line=L_CONT) and ok
        neg_socket = neg_road(socket_road)
        self.add_socket(Socket(socket_road, neg_socket))
        for index, lane in enumerate(self.net.lanes(neg_socket)[self.n_pos:], 1):
            lateral = index * self.lane_width / 2
            inner = self.node(1, index)
            side = L_SIDE if index == n_wave else L_NONE
            c1, c2, _ = wave_lanes(lane, lateral, blen, 5, self.lane_width)
            ok = self._single(c1, neg_road((inner, self.node(0, 0))), side) and ok
            ok = self._single(c2, neg_road((start, inner)), side) and ok
        return ok


class Bidirection(Block):
    """pgblock/bidirection.py: ONE lane used in both directions (the positive and the negative road overlap)"""
    ID = "B"
    SPACE = {"length": box(40.0, 80.0)}

    def plug(self):
        self.set_part(0)
        length = self.cfg["length"]
        basic = self.pos_lanes[0]
        fake_pos = extend_straight(basic, length, [L_BROKEN, L_SIDE])
        fake_neg = SLane(fake_pos.position(fake_pos.length, -fake_pos.width), fake_pos.position(0, -fake_pos.width), fake_pos.width)
        new = SLane(basic.position(basic.length, -basic.width / 2), basic.position(basic.length + length, -basic.width / 2),
                    basic.width, [L_BROKEN, L_SIDE])
        road = (self.pre_socket.positive[1], self.new_node())
        ok = road_from(new, 1, road, self.net, self.world)
        # create_overlap_road (bidirection.py:10-54): the adverse road lies ON the positive one
        lanes = self.net.lanes(road)
        ref = lanes[-1]
        sym = SLane(ref.position(lanes[-1].length, 0), ref.position(0, 0), lanes[-1].width_at(0), lanes[-1].line_types,
                    ref.speed_limit)
        sym.line_colors = (GREY, GREY)
        ok = road_from(sym, int(len(lanes) * 2 / 2), neg_road(road), self.net, self.world, side_line=L_SIDE, inner_line=L_BROKEN,
                       center_line=L_CONT, center_color=YELLOW) and ok
        self.net.lanes(road)[0].line_colors = (YELLOW, GREY)
        new.line_colors = (GREY, GREY)
        self.add_socket(Socket(road, neg_road(road), fake_pos, fake_neg))
        return ok


class TollGate(Block):
    """pgblock/tollgate.py: a straight stretch of solid-lined lanes with a speed limit and a booth on every second lane"""
    ID = "$"
    SPACE = Bottleneck.SPACE
    SPEED_LIMIT = 3

    def plug(self):
        self.set_part(0)
        length = self.cfg["length"]
        new = extend_straight(self.basic, length, [L_CONT, L_SIDE])
        road = (self.pre_socket.positive[1], self.new_node())
        ok = road_from(new, self.n_pos, road, self.net, self.world, center_color=YELLOW, center_line=L_CONT, inner_line=L_CONT,
                       side_line=L_SIDE)
        ok = adverse_road(road, self.net, self.world, center_color=YELLOW, center_line=L_CONT, inner_line=L_CONT,
                          side_line=L_SIDE) and ok
        self.add_socket(Socket(road, neg_road(road)))
        self.buildings = []       # (lane, position, heading): TollGateBuilding on every second lane of both roads
        for r in (road, neg_road(road)):
            for idx, lane in enumerate(self.net.lanes(r)):
                lane.speed_limit = self.SPEED_LIMIT
                if idx % 2 == 1:
                    self.buildings.append((lane, lane.position(lane.length / 2, 0), lane.heading_at(0)))
        return ok


class ParkingLot(Block):
    """pgblock/parking_lot.py: a one-lane two-way street with `one_side_vehicle_number` parking spaces on either side; every
    space is reached by a right-angle bend from either direction and left by one towards either direction"""
    ID = "P"
    SPACE = {"one_side_vehicle_number": disc(2, 10), "radius": const(4), "length": const(8)}
    ANGLE = np.deg2rad(90)
    SOCKET_LENGTH = 4

    def plug(self):
        c = self.cfg
        assert self.n_pos == 1, "Lane number of previous block must be 1 in each direction"
        self.parking_spawn_roads, self.dest_roads = [], []
        self.space_len, self.space_w = c["length"], self.lane_width
        n, radius = int(c["one_side_vehicle_number"]), c["radius"]
        kw = dict(center_line=L_BROKEN, inner_line=L_BROKEN)
        main = extend_straight(self.pos_lanes[0], 2 * radius + (n - 1) * self.space_w, [L_BROKEN, L_NONE])
        road = (self.pre_socket.positive[1], self.node(0, 0))
        ok = road_from(main, self.n_pos, road, self.net, self.world, side_line=L_NONE, center_color=GREY, **kw)
        ok = adverse_road(road, self.net, self.world, side_line=L_NONE, center_color=GREY, **kw) and ok
        out_lane = extend_straight(main, self.SOCKET_LENGTH, [L_BROKEN, L_NONE])
        out_road = (self.node(0, 0), self.node(0, 1))
        ok = road_from(out_lane, self.n_pos, out_road, self.net, self.world, side_line=L_SIDE, **kw) and ok
        ok = adverse_road(out_road, self.net, self.world, side_line=L_SIDE, **kw) and ok
        self.add_socket(self.socket_from(out_road))
        mine, pre = self.get_socket(0), self.pre_socket
        rev = lambda s: Socket(s.negative, s.positive)
        for i in range(n):
            ok = self._space(rev(mine), rev(pre), i + 1, radius, i * self.space_w, (n - i - 1) * self.space_w) and ok
        for i in range(n):
            ok = self._space(pre, mine, n + i + 1, radius, i * self.space_w, (n - i - 1) * self.space_w) and ok
        return ok

    def _from_world(self, sock):
        pre = self.pre_socket
        return (sock.positive, sock.negative) in ((pre.positive, pre.negative), (pre.negative, pre.positive))

    def _space(self, in_socket, out_socket, part, radius, dist_in, dist_out):
        """_add_one_parking_space (parking_lot.py:115-332)"""
        ok = True
        none = dict(center_line=L_NONE, inner_line=L_NONE, side_line=L_NONE)
        node = lambda k: self.node(part, k)
        one = lambda lane, road, **k: road_from(lane, self.n_pos, road, self.net, self.world, **k)
        net = self.world if self._from_world(in_socket) else self.net
        in_lane = net.lanes(in_socket.positive)[0]
        start = in_socket.positive[1]
        if dist_in > 1e-3:
            in_lane = extend_straight(in_lane, dist_in, [L_NONE, L_NONE])
            one(in_lane, (in_socket.positive[1], node(0)), **none)
            start = node(0)
        side = L_SIDE if dist_in < 1e-3 else L_NONE
        bend, straight = bend_then_straight(in_lane, self.space_len, radius, self.ANGLE, True, self.space_w)
        bend_ok = one(bend, (start, node(1)), center_line=L_NONE, inner_line=L_NONE, side_line=side)
        if dist_in < 1e-3:
            ok = ok and bend_ok
        straight_road = (node(1), node(2))
        self.dest_roads.append(straight_road)
        ok = ok and one(straight, straight_road, center_line=L_CONT, inner_line=L_NONE, side_line=side, center_color=GREY)
        # the way in from the other direction
        neg = out_socket.negative
        net = self.world if self._from_world(out_socket) else self.net
        neg_lane = net.lanes(neg)[0]
        start = neg[1]
        if dist_out > 1e-3:
            neg_lane = extend_straight(neg_lane, dist_out, [L_NONE, L_NONE])
            one(neg_lane, (neg[1], node(3)), **none)
            start = node(3)
        bend, straight = bend_then_straight(neg_lane, self.lane_width, radius, self.ANGLE, False, self.space_w)
        one(bend, (start, node(4)), **none)
        one(straight, (node(4), node(1)), **none)
        # the space under a second road name, in the reverse direction: (1, 2) = (5, 6)
        parking_road = (node(5), node(6))
        self.parking_spawn_roads.append(parking_road)
        two_way_road((node(1), node(2)), self.net, self.world, parking_road, center_line=L_NONE, inner_line=L_NONE,
                     side_line=L_SIDE if dist_out < 1e-3 else L_NONE)
        parking_lane = self.net.lanes(parking_road)[0]
        # the way out, first direction
        bend, straight = bend_then_straight(parking_lane, 0.1 if dist_out < 1e-3 else dist_out, radius, self.ANGLE, True,
                                            parking_lane.width)
        bend_ok = one(bend, (node(6), node(7) if dist_out > 1e-3 else out_socket.positive[0]), center_line=L_NONE,
                      inner_line=L_NONE, side_line=L_SIDE if dist_out < 1e-3 else L_NONE)
        if dist_out < 1e-3:
            ok = ok and bend_ok
        if dist_out > 1e-3:
            ok = ok and one(straight, (node(7), out_socket.positive[0]), **none)
        # the way out, other direction
        ext = extend_straight(parking_lane, self.lane_width, [L_NONE, L_NONE])
        one(ext, (node(6), node(8)), **none)
        bend, straight = bend_then_straight(ext, 0.1 if dist_in < 1e-3 else dist_in, radius, self.ANGLE, False, parking_lane.width)
        one(bend, (node(8), node(9) if dist_in > 1e-3 else in_socket.negative[0]), **none)
        if dist_in > 1e-3:
            one(straight, (node(9), in_socket.negative[0]), **none)
        return ok


# BLOCK_TYPE_DISTRIBUTION_V2 (component/algorithm/blocks_prob_dist.py:24-43), in its dict order
BLOCK_DIST = [("Curve", Curve, 0.3), ("Straight", Straight, 0.1), ("InRampOnStraight", InRamp, 0.1),
              ("OutRampOnStraight", OutRamp, 0.1), ("StdInterSection", InterSection, 0.15),
              ("StdTInterSection", TInterSection, 0.15), ("Roundabout", Roundabout, 0.1), ("InFork", None, 0.0),
              ("OutFork", None, 0.0), ("Merge", Merge, 0.0), ("Split", Split, 0.0), ("ParkingLot", ParkingLot, 0.0),
              ("TollGate", TollGate, 0.0), ("Bidirection", Bidirection, 0.0)]
BY_ID = {cls.ID: cls for _, cls, _ in BLOCK_DIST if cls is not None}
MIN_LANES, MAX_LANES = 1, 5


class BIG:
    """component/algorithm/BIG.py:28-169: forward / destruct / search-sibling / back state machine."""
    MAX_TRIAL = 5

    def __init__(self, lane_num, lane_width, seed, exit_length=50, block_overrides=None):
        self.rng = seeded_rng(seed)
        self.world = Net()
        self.blocks = [FirstBlock(self.world, lane_width, lane_num, exit_length)]
        self.sequence = None
        # per block ID, class constants to replace - the multi-agent maps lengthen the exits of their block by assigning
        # the CLASS attribute (envs/marl_envs/marl_inout_roundabout.py:46, marl_intersection.py:46), which in the reference
        # then leaks into every later map of the same process; here it is an explicit, per-map argument
        self.block_overrides = block_overrides or {}

    def generate(self, spec):
        if isinstance(spec, int) and not isinstance(spec, bool):
            n_blocks = spec + 1
        else:
            n_blocks = len(spec) + 1
            self.sequence = FirstBlock.ID + spec
        state = "forward"
        while not (len(self.blocks) >= n_blocks and state == "forward"):
            if state == "forward":
                b = self._sample()
                self.blocks.append(b)
                state = "forward" if self._construct(b) else "destruct"
            elif state == "destruct":
                b = self.blocks[-1]
                b.clear()
                state = "sibling" if b.trials < self.MAX_TRIAL else "back"
            elif state == "sibling":
                b = self.blocks[-1]
                if len(self.blocks) == 1:
                    state = "forward"
                elif b.trials < self.MAX_TRIAL:
                    state = "forward" if self._construct(b) else "destruct"
                else:
                    state = "back"
            else:   # back
                self.blocks.pop()
                self.blocks[-1].clear()
                state = "sibling"
        return self

    def _sample(self):
        if self.sequence is None:
            name = self.rng.choice([n for n, _, _ in BLOCK_DIST], p=[p for _, _, p in BLOCK_DIST])
            cls = next(c for n, c, _ in BLOCK_DIST if n == name)
        else:
            cls = BY_ID.get(self.sequence[len(self.blocks)])
            if cls is None:
                raise NotImplementedError("block type %r is not covered by the product's map generator" % self.sequence[len(self.blocks)])
        last = self.blocks[-1]
        key = self.rng.choice(list(last.sockets.keys()))
        sock = last.get_socket(list(last.sockets).index(key))
        blk = cls(len(self.blocks), sock, self.world, self.rng.randint(0, 10000))
        for k, v in self.block_overrides.get(cls.ID, {}).items():
            setattr(blk, k, v)
        return blk

    def _construct(self, b):
        ok = b.construct()
        n = max(len(s.positive_lanes(self.world)) for s in b.sockets.values())
        if n < MIN_LANES or n > MAX_LANES:
            ok = False
        return ok


# ---------------------------------------------------------------------------------------------- tables
def to_tables(big):
    """The flat lane / road tables + block metadata of a generated map, in the layout of oracle/ref_export.export_map
    (node / road / lane ids in graph insertion order)."""
    g = big.world.graph
    nodes, roads, road_list, lane_objs, lane_key = {}, {}, [], [], {}
    for a, d in g.items():
        for b, lanes in d.items():
            for n in (a, b):
                if n not in nodes:
                    nodes[n] = len(nodes)
            roads[(a, b)] = len(road_list)
            road_list.append((a, b, len(lane_objs), len(lanes)))
            for i, lane in enumerate(lanes):
                lane_key[id(lane)] = len(lane_objs)
                lane_objs.append((a, b, i, lane))
    lane_f = np.zeros((len(lane_objs), 10), np.float64)
    lane_i = np.zeros((len(lane_objs), 8), np.int32)
    for k, (a, b, i, lane) in enumerate(lane_objs):
        if isinstance(lane, SLane):
            lane_f[k] = [0, lane.width, lane.length, lane.start[0], lane.start[1], lane.end[0], lane.end[1], 0, 0, 0]
        else:
            lane_f[k] = [1, lane.width, lane.length, lane.center[0], lane.center[1], lane.radius, lane.start_phase,
                         lane.end_phase, lane.direction, lane.angle]
        lt = [L_NONE if t is None else t for t in lane.line_types]
        lane_i[k] = [roads[(a, b)], i, nodes[a], nodes[b], lt[0], lt[1], int(lane.line_colors[0] == YELLOW),
                     int(lane.line_colors[1] == YELLOW)]
    road_i = np.zeros((len(road_list), 6), np.int32)
    for k, (a, b, first, n) in enumerate(road_list):
        road_i[k] = [nodes[a], nodes[b], first, n, int(is_negative((a, b))), ord(road_block_id((a, b))[0])]
    lid = lambda lane: lane_key.get(id(lane), -1)
    blocks = []
    for blk in big.blocks:
        trig = blk.pre_socket.positive if blk.pre_socket is not None else None
        spawn = [] if blk is big.blocks[0] else [[lid(l) for l in lanes] for lanes in blk.spawn_lanes()]
        negs = [[lid(l) for l in lanes] for lanes in blk.net.negative_lanes()] if blk.ID in "SCrR" else []
        blocks.append(dict(
            id=blk.ID, trigger_road=roads.get(trig, -1) if trig is not None else -1, spawn_lanes=spawn, negative_lanes=negs,
            respawn_roads=[roads.get(r, -1) for r in blk.respawn_roads],
            sockets=[[roads.get(s.positive, -1), roads.get(s.negative, -1) if s.negative is not None else -1]
                     for s in blk.sockets.values()]))
        if getattr(blk, "buildings", None):
            # TollGateBuilding (component/buildings/tollgate_building.py:7-26): a static box BUILDING_LENGTH = 10 long and one
            # lane wide, centred on the lane: [lane id, x, y, heading, half length, half width]
            blocks[-1]["buildings"] = [[lid(lane), float(pos[0]), float(pos[1]), float(heading), 5.0, lane.width / 2.0]
                                       for lane, pos, heading in blk.buildings]
    meta = dict(nodes=list(nodes.keys()), blocks=blocks)
    return lane_f, lane_i, road_i, meta


def build_fixed(kind, lane_num=2, lane_width=3.5, exit_length=60, parking_space_num=8, **chain_kw):
    """The fixed maps of the multi-agent envs (envs/marl_envs/marl_inout_roundabout.py:27-60, marl_intersection.py:27-71):
    a first block and ONE block built from a given configuration with block seed 1."""
    big = BIG.__new__(BIG)
    big.rng, big.sequence, big.block_overrides = None, None, {}
    big.world = Net()
    big.blocks = [FirstBlock(big.world, lane_width, lane_num, exit_length)]
    sock = big.blocks[0].get_socket(0)
    if kind == "roundabout":
        blk = Roundabout(1, sock, big.world, 1)
        blk.EXIT_LEN = exit_length
        blk.construct(dict(exit_radius=10, inner_radius=30, angle=70))
    elif kind == "intersection":
        blk = InterSection(1, sock, big.world, 1)
        blk.EXIT_LEN = exit_length
        blk.std = False
        blk.u_turn = lane_num > 1
        blk.construct()
    elif kind in ("bottleneck", "bidirection", "tollgate", "parkinglot"):
        return _build_chain(big, kind, lane_num, exit_length, parking_space_num=parking_space_num, **chain_kw)
    else:
        raise NotImplementedError("multi-agent map %r is not restated" % kind)
    big.blocks.append(blk)
    return to_tables(big) + (big, )


def _build_chain(big, kind, lane_num, exit_length, neck_lane_num=1, neck_length=20, toll_lane_num=8, toll_length=10,
                 parking_space_num=8):
    """MABottleneckMap / MABidirectionMap / MATollGateMap (envs/marl_envs/marl_bottleneck.py:27-66,
    marl_bidirection.py:27-77, marl_tollgate.py:103-150): fixed block chains, every block with seed 1"""
    def add(cls, extra):
        blk = cls(len(big.blocks), big.blocks[-1].get_socket(0), big.world, 1)
        blk.construct(extra)
        big.blocks.append(blk)
    d = lane_num - neck_lane_num
    if kind == "bottleneck":      # first block with bottle_lane_num lanes -> Merge -> Split
        add(Merge, dict(lane_num=d, length=neck_length))
        add(Split, dict(length=exit_length, lane_num=d))
    elif kind == "bidirection":   # -> Merge -> Bidirection -> Split
        add(Merge, dict(lane_num=d, length=3))
        add(Bidirection, None)
        add(Split, dict(length=exit_length, lane_num=d))
    elif kind == "parkinglot":    # MAParkingLotMap (marl_parking_lot.py:144-184): -> ParkingLot (parking_space_num / 2 spaces a side) -> T
        add(ParkingLot, dict(one_side_vehicle_number=int(parking_space_num / 2)))
        t = TInterSection(len(big.blocks), big.blocks[-1].get_socket(0), big.world, 1)
        t.EXIT_LEN = 10
        t.construct(dict(t_type=1, change_lane_num=0))
        big.blocks.append(t)
    else:                         # tollgate: -> Split -> TollGate -> Merge, bottle length 35
        add(Split, dict(length=2, lane_num=toll_lane_num - lane_num, bottle_len=35))
        add(TollGate, dict(length=toll_length))
        add(Merge, dict(lane_num=toll_lane_num - lane_num, length=exit_length, bottle_len=35))
    return to_tables(big) + (big, )


def generate(seed, map_spec=3, lane_num=3, lane_width=3.5, exit_length=50, block_overrides=None):
    """The map `MetaDriveEnv(dict(map=map_spec, ...)).reset(seed)` builds (component/map/pg_map.py:55-80,
    manager/pg_map_manager.py:57-74): BIG seeded with the scenario seed.  Returns (lane_f, lane_i, road_i, meta, big)."""
    big = BIG(lane_num, lane_width, seed, exit_length, block_overrides).generate(map_spec)
    return to_tables(big) + (big, )

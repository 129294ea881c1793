"""Recording + analytic stand-in for `panda3d.bullet`. TEST INFRASTRUCTURE (see package docstring).

Every body the reference creates is kept with its shape, transform, collide mask and name, and the
queries the step path issues are answered analytically by `physics.py`.
Known, documented simplifications: no collision margins, no contact *response*
between dynamic bodies (contacts only raise the reference's callback), gyroscopic term ignored.
"""
import math

import numpy as np

from . import physics as ph
from .pcore import PandaNode, NodePath, Vec3, BitMask32, TransformState, mat_to_quat, quat_to_mat, _word

ZUp = 2
YUp = 1
XUp = 0


# ------------------------------------------------------------------------------------------ shapes
class BulletShape:
    pass


class BulletBoxShape(BulletShape):
    def __init__(self, half):
        self.half = np.array(list(half), dtype=np.float64)


class BulletCylinderShape(BulletShape):
    def __init__(self, radius, height, up=ZUp):
        self.radius = float(radius)
        self.height = float(height)
        self.up = up


class BulletCapsuleShape(BulletCylinderShape):
    pass


class BulletSphereShape(BulletShape):
    def __init__(self, radius):
        self.radius = float(radius)


class BulletPlaneShape(BulletShape):
    def __init__(self, normal, d):
        self.normal = np.array(list(normal), dtype=np.float64)
        self.d = float(d)


class BulletConvexHullShape(BulletShape):
    def __init__(self):
        self.points = []
        self._hull = None

    def addPoint(self, p):
        self.points.append(np.array(list(p), dtype=np.float64))
        self._hull = None

    add_point = addPoint

    @property
    def hull2d(self):
        if self._hull is None:
            self._hull = ph.convex_hull([p[:2] for p in self.points])
        return self._hull

    @property
    def zrange(self):
        zs = [p[2] for p in self.points]
        return min(zs), max(zs)


class BulletTriangleMesh:
    def __init__(self):
        self.polygon = None
        self.height = None

    def addGeom(self, geom, *a, **k):
        self.polygon = getattr(geom, "polygon", None)
        self.height = getattr(geom, "height", None)

    add_geom = addGeom


class BulletTriangleMeshShape(BulletShape):
    def __init__(self, mesh, dynamic=False, *a, **k):
        self.polygon = np.asarray(mesh.polygon, dtype=np.float64) if mesh.polygon is not None else None
        self.height = mesh.height


class BulletHeightfieldShape(BulletShape):
    def __init__(self, *a, **k):
        pass


# ------------------------------------------------------------------------------------------ bodies
class BulletBodyNode(PandaNode):
    def __init__(self, name=""):
        super().__init__(name)
        self.shapes = []
        self._geom_cache = None
        self.into_mask = BitMask32.allOn()
        self.static = False
        self.kinematic = False
        self.active = True
        self.notify = False
        self.world = None

    def addShape(self, shape, ts=None):
        self.shapes.append((shape, ts))

    add_shape = addShape

    def setIntoCollideMask(self, m):
        self.into_mask = BitMask32(_word(m))

    set_into_collide_mask = setIntoCollideMask

    def getIntoCollideMask(self):
        return self.into_mask

    def setStatic(self, f):
        self.static = bool(f)

    set_static = setStatic

    def isStatic(self):
        return self.static

    def setKinematic(self, f):
        self.kinematic = bool(f)

    set_kinematic = setKinematic

    def setActive(self, f, *a):
        self.active = bool(f)

    set_active = setActive

    def notifyCollisions(self, f):
        self.notify = bool(f)

    notify_collisions = notifyCollisions

    def setDeactivationEnabled(self, f, *a):
        pass


class BulletGhostNode(BulletBodyNode):
    pass


class BulletRigidBodyNode(BulletBodyNode):
    def __init__(self, name=""):
        super().__init__(name)
        self.mass = 0.0
        self.lin_vel = np.zeros(3)
        self.ang_vel = np.zeros(3)
        self.friction = 0.5
        self.gravity = None

    def setMass(self, m):
        self.mass = float(m)

    set_mass = setMass

    def getMass(self):
        return self.mass

    def setLinearVelocity(self, v):
        self.lin_vel = np.array(list(v), dtype=np.float64)

    set_linear_velocity = setLinearVelocity

    def getLinearVelocity(self):
        return Vec3(*self.lin_vel)

    get_linear_velocity = getLinearVelocity

    def setAngularVelocity(self, v):
        self.ang_vel = np.array(list(v), dtype=np.float64)

    set_angular_velocity = setAngularVelocity

    def getAngularVelocity(self):
        return Vec3(*self.ang_vel)

    get_angular_velocity = getAngularVelocity

    def clearForces(self):
        pass

    clear_forces = clearForces

    def setFriction(self, f):
        self.friction = float(f)

    def setGravity(self, g):
        self.gravity = np.array(list(g), dtype=np.float64)

    def setLinearDamping(self, *a):
        pass

    def setAngularDamping(self, *a):
        pass


class BulletDebugNode(PandaNode):
    pass


# ------------------------------------------------------------------------------------------ vehicle
class BulletWheel:
    def __init__(self):
        self.conn = np.zeros(3)
        self.radius = 0.5
        self.front = False
        self.travel = 500.0
        self.stiffness = 5.88
        self.damp_relax = 0.88
        self.damp_comp = 0.83
        self.friction = 10.5
        self.roll = 0.1
        self.node = None

    def setNode(self, n):
        self.node = n

    def setChassisConnectionPointCs(self, p):
        self.conn = np.array(list(p), dtype=np.float64)

    def setFrontWheel(self, f):
        self.front = bool(f)

    def setWheelDirectionCs(self, d):
        assert tuple(d) == (0, 0, -1)

    def setWheelAxleCs(self, a):
        assert tuple(a) == (1, 0, 0)

    def setWheelRadius(self, r):
        self.radius = float(r)

    def setMaxSuspensionTravelCm(self, t):
        self.travel = float(t)

    def setSuspensionStiffness(self, k):
        self.stiffness = float(k)

    def setWheelsDampingRelaxation(self, c):
        self.damp_relax = float(c)

    def setWheelsDampingCompression(self, c):
        self.damp_comp = float(c)

    def setFrictionSlip(self, f):
        self.friction = float(f)

    def setRollInfluence(self, r):
        self.roll = float(r)

    def as_dict(self):
        return dict(
            conn=self.conn,
            radius=self.radius,
            front=self.front,
            travel=self.travel,
            stiffness=self.stiffness,
            damp_relax=self.damp_relax,
            damp_comp=self.damp_comp,
            friction=self.friction,
            roll=self.roll
        )


class BulletVehicle:
    def __init__(self, world, chassis):
        self.chassis = chassis
        self.wheels = []
        self.steering = {}
        self.engine_force = {}
        self.brake = {}
        chassis._vehicle = self

    def setCoordinateSystem(self, up):
        assert up == ZUp

    def getChassis(self):
        return self.chassis

    def getName(self):
        return "vehicle"

    get_chassis = getChassis

    def createWheel(self):
        w = BulletWheel()
        self.wheels.append(w)
        return w

    create_wheel = createWheel

    def getWheels(self):
        return self.wheels

    def setSteeringValue(self, deg, idx):
        self.steering[idx] = math.radians(float(deg))

    def applyEngineForce(self, f, idx):
        self.engine_force[idx] = float(f)

    def setBrake(self, b, idx):
        self.brake[idx] = float(b)

    def resetSuspension(self):
        pass

    def getCurrentSpeedKmHour(self):
        return 3.6 * float(np.linalg.norm(self.chassis.lin_vel))

    def _body(self):
        box = self.chassis.shapes[0][0]
        w, l, h = 2 * box.half[0], 2 * box.half[1], 2 * box.half[2]
        b = ph.VehicleBody(self.chassis.mass, w, l, h, [wh.as_dict() for wh in self.wheels])
        n = len(self.wheels)
        b.steering = [self.steering.get(i, 0.0) for i in range(n)]
        b.engine_force = [self.engine_force.get(i, 0.0) for i in range(n)]
        b.brake = [self.brake.get(i, 0.0) for i in range(n)]
        return b


# ------------------------------------------------------------------------------------------ results
class _RayHit:
    def __init__(self, node, frac, pos, normal=None):
        self._node, self._frac, self._pos = node, frac, pos

    def getNode(self):
        return self._node

    get_node = getNode

    @property
    def node(self):
        return self._node

    def getHitFraction(self):
        return self._frac

    get_hit_fraction = getHitFraction

    def hasHit(self):
        return self._node is not None

    has_hit = hasHit

    def getHitPos(self):
        return Vec3(*self._pos)

    get_hit_pos = getHitPos


class _RayAll:
    def __init__(self, hits):
        self._hits = hits

    def hasHits(self):
        return len(self._hits) > 0

    has_hits = hasHits

    def getHits(self):
        return list(self._hits)

    get_hits = getHits

    def getNumHits(self):
        return len(self._hits)


class _Contact:
    def __init__(self, n0, n1):
        self._n0, self._n1 = n0, n1

    def getNode0(self):
        return self._n0

    get_node0 = getNode0

    def getNode1(self):
        return self._n1

    get_node1 = getNode1


class _ContactResult:
    def __init__(self, contacts):
        self._c = contacts

    def getContacts(self):
        return list(self._c)

    get_contacts = getContacts

    def getNumContacts(self):
        return len(self._c)


# ------------------------------------------------------------------------------------------ geometry of a body
def _prims(node):
    """Yield world-space primitives of a body: ('box', c, R, h) | ('cyl', c, r, half_h) | ('hull', poly, z0, z1)
    | ('mesh', poly, z0, z1) | ('plane',)."""
    for shape, ts in node.shapes:
        off = ts.pos if ts is not None else np.zeros(3)
        rot = ts.mat if ts is not None else np.eye(3)
        if isinstance(shape, BulletBoxShape):
            yield ("box", node.pos + node.mat @ off, node.mat @ rot, shape.half)
        elif isinstance(shape, BulletCylinderShape):
            assert shape.up == ZUp
            yield ("cyl", node.pos + node.mat @ off, shape.radius, shape.height / 2)
        elif isinstance(shape, BulletConvexHullShape):
            z0, z1 = shape.zrange
            yield ("hull", shape.hull2d + node.pos[:2], z0 + node.pos[2], z1 + node.pos[2])
        elif isinstance(shape, BulletTriangleMeshShape):
            if shape.polygon is not None:
                hh = (shape.height or 0.0) / 2
                yield ("mesh", shape.polygon + node.pos[:2], node.pos[2] - hh, node.pos[2] + hh)
        elif isinstance(shape, BulletPlaneShape):
            yield ("plane", )


def _footprint(prim):
    """2-D footprint of a box primitive: centre, unit forward axis (local x of the shape), half (x, y)."""
    _, c, R, h = prim
    u = R[:2, 0]
    n = math.hypot(u[0], u[1])
    return c[:2], u / n, (h[0], h[1])


def _overlap2d(pa, pb):
    ka, kb = pa[0], pb[0]
    if ka == "plane" or kb == "plane":
        return False
    if ka == "box" and kb == "box":
        # z ranges (upright approximation) then 2-D SAT
        if abs(pa[1][2] - pb[1][2]) > pa[3][2] + pb[3][2]:
            return False
        ca, ua, ha = _footprint(pa)
        cb, ub, hb = _footprint(pb)
        return ph.obb2d_overlap(ca, ua, ha, cb, ub, hb)
    if ka == "box" and kb == "cyl":
        if abs(pa[1][2] - pb[1][2]) > pa[3][2] + pb[3]:
            return False
        ca, ua, ha = _footprint(pa)
        return ph.obb2d_circle_overlap(ca, ua, ha, pb[1][:2], pb[2])
    if ka == "cyl" and kb == "box":
        return _overlap2d(pb, pa)
    if ka == "cyl" and kb == "cyl":
        if abs(pa[1][2] - pb[1][2]) > pa[3] + pb[3]:
            return False
        d = pa[1][:2] - pb[1][:2]
        return float(d @ d) <= (pa[2] + pb[2])**2
    if ka == "box" and kb in ("hull", "mesh"):
        if pa[1][2] - pa[3][2] > pb[3] or pa[1][2] + pa[3][2] < pb[2]:
            return False
        ca, ua, ha = _footprint(pa)
        corners = ph.rect_corners(ca, ua, ha[0], ha[1])
        return ph.rect_polygon_overlap(corners, pb[1])
    if kb == "box" and ka in ("hull", "mesh"):
        return _overlap2d(pb, pa)
    if ka == "cyl" and kb in ("hull", "mesh"):
        return ph.point_in_polygon(pb[1], pa[1][:2])  # centre test is enough for the step path
    if kb == "cyl" and ka in ("hull", "mesh"):
        return _overlap2d(pb, pa)
    return False


def _cached(node):
    """(prims, bounding-circle centre, radius) cached until the node's transform or shapes change."""
    c = node.__dict__.get("_geom_cache")
    if c is not None and c[3] == len(node.shapes):
        return c
    prims = list(_prims(node))
    pts = []
    for p in prims:
        if p[0] == "box":
            ce, u, h = _footprint(p)
            pts.extend(ph.rect_corners(ce, u, h[0], h[1]))
        elif p[0] == "cyl":
            ce = p[1][:2]
            pts.extend([ce + [p[2], p[2]], ce - [p[2], p[2]], ce + [p[2], -p[2]], ce + [-p[2], p[2]]])
        elif p[0] in ("hull", "mesh"):
            pts.extend(p[1])
    if pts:
        pts = np.asarray(pts)
        centre = 0.5 * (pts.min(0) + pts.max(0))
        rad = float(np.sqrt(((pts - centre)**2).sum(1).max()))
    else:
        centre, rad = np.zeros(2), float("inf")
    c = (prims, centre, rad, len(node.shapes))
    node.__dict__["_geom_cache"] = c
    return c


def nodes_overlap(a, b):
    pa_, ca, ra, _ = _cached(a)
    pb_, cb, rb, _ = _cached(b)
    d = ca - cb
    if d[0] * d[0] + d[1] * d[1] > (ra + rb)**2:
        return False
    for pa in pa_:
        for pb in pb_:
            if _overlap2d(pa, pb):
                return True
    return False


# ------------------------------------------------------------------------------------------ world
class BulletWorld:
    def __init__(self):
        self.bodies = []
        self.vehicles = []
        self.group_flags = {}
        self.contact_cb = None
        self.gravity = np.array([0, 0, -9.81])

    # -- registration
    def attach(self, obj):
        if isinstance(obj, BulletVehicle):
            if obj not in self.vehicles:
                self.vehicles.append(obj)
        elif isinstance(obj, BulletBodyNode):
            if obj not in self.bodies:
                self.bodies.append(obj)
                obj.world = self
        else:
            raise TypeError(type(obj))

    attachRigidBody = attachGhost = attachVehicle = attach_rigid_body = attach_ghost = attach_vehicle = attach

    def remove(self, obj):
        if isinstance(obj, BulletVehicle):
            # panda3d BulletWorld::do_remove_vehicle also removes the chassis body
            # (the reference relies on it: base_class/base_object.py:90-93 breaks after the vehicle)
            if obj in self.vehicles:
                self.vehicles.remove(obj)
            if obj.chassis in self.bodies:
                self.bodies.remove(obj.chassis)
                obj.chassis.world = None
        elif obj in self.bodies:
            self.bodies.remove(obj)
            obj.world = None

    removeRigidBody = removeGhost = removeVehicle = remove_rigid_body = remove_ghost = remove_vehicle = remove

    def setGravity(self, g):
        self.gravity = np.array(list(g), dtype=np.float64)

    set_gravity = setGravity

    def setGroupCollisionFlag(self, g1, g2, flag):
        self.group_flags[(g1, g2)] = self.group_flags[(g2, g1)] = bool(flag)

    set_group_collision_flag = setGroupCollisionFlag

    def setContactAddedCallback(self, cb):
        self.contact_cb = cb

    def clearContactAddedCallback(self):
        self.contact_cb = None

    def clearDebugNode(self):
        pass

    def clearFilterCallback(self):
        pass

    def setDebugNode(self, *a):
        pass

    def getNumRigidBodies(self):
        return sum(1 for b in self.bodies if isinstance(b, BulletRigidBodyNode))

    def getNumGhosts(self):
        return sum(1 for b in self.bodies if isinstance(b, BulletGhostNode))

    def getNumVehicles(self):
        return len(self.vehicles)

    def getRigidBodies(self):
        return [b for b in self.bodies if isinstance(b, BulletRigidBodyNode)]

    def getGhosts(self):
        return [b for b in self.bodies if isinstance(b, BulletGhostNode)]

    def getVehicles(self):
        return list(self.vehicles)

    def getSoftBodies(self):
        return []

    def getCharacters(self):
        return []

    # -- filters
    def _groups_collide(self, m0, m1):
        w0, w1 = _word(m0), _word(m1)
        for i in range(32):
            if not (w0 >> i) & 1:
                continue
            for j in range(32):
                if (w1 >> j) & 1 and self.group_flags.get((i, j), False):
                    return True
        return False

    # -- ray tests
    def _ray_hits(self, p_from, p_to, mask):
        o = np.array(list(p_from), dtype=np.float64)
        d = np.array(list(p_to), dtype=np.float64) - o
        mw = _word(mask)
        out = []
        for b in self.bodies:
            if not (mw & b.into_mask.w):
                continue
            prims, cen, rad, _ = _cached(b)
            if rad != float("inf"):
                # distance from the bounding-circle centre to the ray's xy segment
                dd = d[0] * d[0] + d[1] * d[1]
                tt = 0.0 if dd < 1e-18 else min(1.0, max(0.0, ((cen[0] - o[0]) * d[0] + (cen[1] - o[1]) * d[1]) / dd))
                qx, qy = o[0] + tt * d[0] - cen[0], o[1] + tt * d[1] - cen[1]
                if qx * qx + qy * qy > rad * rad:
                    continue
            best = None
            for prim in prims:
                t = None
                k = prim[0]
                if k == "box":
                    t = ph.ray_obb(o, d, prim[1], prim[2], prim[3])
                elif k == "cyl":
                    t = ph.ray_zcyl(o, d, prim[1], prim[2], prim[3])
                elif k in ("hull", "mesh"):
                    if abs(d[0]) < 1e-12 and abs(d[1]) < 1e-12 and abs(d[2]) > 0:
                        # vertical ray: hit the top (or bottom) face if the point is inside the polygon
                        inside = ph.point_in_convex(prim[1], o[:2]) if k == "hull" else ph.point_in_polygon(
                            prim[1], o[:2]
                        )
                        if inside:
                            zt = prim[3] if d[2] < 0 else prim[2]
                            tt = (zt - o[2]) / d[2]
                            if 0 <= tt <= 1:
                                t = tt
                elif k == "plane":
                    if abs(d[2]) > 1e-12:
                        tt = -o[2] / d[2]
                        if 0 <= tt <= 1:
                            t = tt
                if t is not None and (best is None or t < best):
                    best = t
            if best is not None:
                out.append(_RayHit(b, best, o + d * best))
        return out

    def rayTestClosest(self, p_from, p_to, mask=None):
        hits = self._ray_hits(p_from, p_to, mask if mask is not None else BitMask32.allOn())
        if not hits:
            return _RayHit(None, 1.0, np.array(list(p_to), dtype=np.float64))
        return min(hits, key=lambda h: h.getHitFraction())

    ray_test_closest = rayTestClosest

    def rayTestAll(self, p_from, p_to, mask=None):
        return _RayAll(self._ray_hits(p_from, p_to, mask if mask is not None else BitMask32.allOn()))

    ray_test_all = rayTestAll

    # -- contact test
    def contactTest(self, node, use_filter=False):
        contacts = []
        for b in self.bodies:
            if b is node:
                continue
            if use_filter and not self._groups_collide(node.into_mask, b.into_mask):
                continue
            if nodes_overlap(node, b):
                contacts.append(_Contact(node, b))
        return _ContactResult(contacts)

    contact_test = contactTest

    # -- sweep
    def sweepTestClosest(self, shape, ts_from, ts_to, mask=None, penetration=0.0):
        mw = _word(mask) if mask is not None else 0xFFFFFFFF
        probe = BulletGhostNode("sweep_probe")
        probe.shapes = [(shape, None)]
        probe.pos = ts_from.pos.copy()
        probe.mat = ts_from.mat
        z_hi = max(ts_from.pos[2], ts_to.pos[2]) + 1e3
        for b in self.bodies:
            if not (mw & b.into_mask.w):
                continue
            # vertical sweep: the 2-D footprints decide; lift z-range checks by testing at the body's height
            for pb in _prims(b):
                if pb[0] == "plane":
                    continue
                zc = pb[1][2] if pb[0] in ("box", "cyl") else 0.5 * (pb[2] + pb[3])
                probe.pos = np.array([ts_from.pos[0], ts_from.pos[1], zc])
                for pa in _prims(probe):
                    if _overlap2d(pa, pb):
                        return _RayHit(b, 0.5, probe.pos)
        return _RayHit(None, 1.0, ts_to.pos)

    sweep_test_closest = sweepTestClosest

    # -- stepping
    def doPhysics(self, dt, max_substeps=1, stepsize=1.0 / 60.0):
        assert max_substeps == 1 and abs(dt - stepsize) < 1e-9, "reference always calls doPhysics(0.02, 1, 0.02)"
        chassis_nodes = set()
        for veh in self.vehicles:
            c = veh.chassis
            chassis_nodes.add(c)
            if c.static or c.world is not self:
                continue
            body = veh._body()
            q = mat_to_quat(c.mat)
            pos, q, v, w = ph.vehicle_substep_pre(body, c.pos, q, c.mat, c.lin_vel, c.ang_vel, dt)
            R = quat_to_mat(q)
            box = c.shapes[0][0]
            pos, v = ph.chassis_ground_clamp(box.half[0], box.half[1], pos, R, v)
            v, w = ph.update_vehicle(body, pos, R, v, w, dt)
            c.pos, c.mat, c.lin_vel, c.ang_vel = pos, R, v, w
            c._geom_cache = None
        for b in self.bodies:
            if b in chassis_nodes or not isinstance(b, BulletRigidBodyNode):
                continue
            if b.static or b.mass <= 0:
                continue
            # free rigid bodies on the step path are upright kinematic movers (pedestrians, loose cones):
            # zero-friction slide at their set planar velocity; z is held by the ground contact.
            b.pos = b.pos + np.array([b.lin_vel[0], b.lin_vel[1], 0.0]) * dt
            b._geom_cache = None
        # contact flags are raised on the poses every body reached this sub-step; the response (physics.contact_deltas)
        # is computed from the same snapshot and applied afterwards
        deltas = self._contact_response()
        self._contact_callbacks(chassis_nodes)
        for c, (dv, dw, dp) in deltas:
            c.lin_vel = c.lin_vel + np.array([dv[0], dv[1], 0.0])
            c.ang_vel = c.ang_vel + np.array([0.0, 0.0, dw])
            c.pos = c.pos + np.array([dp[0], dp[1], 0.0])
            c._geom_cache = None

    do_physics = doPhysics

    def _contact_response(self):
        """One Jacobi pass of the planar contact model (physics.contact_deltas) over the chassis of this world, ordered
        like self.vehicles, followed by the static obstacles that collide with the Vehicle group."""
        bodies, nodes = [], []
        chassis = set()
        # index order = the trace's slot order when the golden generator tagged the chassis (a respawned vehicle takes
        # over the slot of the one that left), else creation order
        order = sorted(range(len(self.vehicles)), key=lambda i: (self.vehicles[i].chassis.__dict__.get("_md_slot", i), i))
        for veh in (self.vehicles[i] for i in order):
            c = veh.chassis
            chassis.add(c)
            if c.world is not self:
                continue
            box = c.shapes[0][0]
            ts = c.shapes[0][1]
            off = ts.pos if ts is not None else np.zeros(3)
            centre = c.pos + c.mat @ off
            u = c.mat[:2, 1]
            u = u / math.hypot(u[0], u[1])
            dyn = not c.static and c.mass > 0
            izz = c.mass / 12.0 * ((2 * box.half[0])**2 + (2 * box.half[1])**2) if dyn else 0.0
            bodies.append(dict(shape="rect", c=centre[:2].copy(), u=u, h=(box.half[1], box.half[0]), o=c.pos[:2].copy(),
                               v=c.lin_vel[:2].copy(), w=float(c.ang_vel[2]), im=1.0 / c.mass if dyn else 0.0,
                               ii=1.0 / izz if dyn else 0.0))
            nodes.append(c)
        n_veh = len(bodies)
        if n_veh == 0:
            return []
        probe = nodes[0]
        for b in self.bodies:
            if b in chassis or isinstance(b, BulletGhostNode) or not isinstance(b, BulletRigidBodyNode):
                continue
            if not (b.static or b.mass <= 0) or not self._groups_collide(probe.into_mask, b.into_mask):
                continue  # free movers (pedestrians, loose cones) neither push nor are pushed in this model
            prims = _cached(b)[0]
            if len(prims) != 1 or prims[0][0] not in ("box", "cyl"):
                continue
            pr = prims[0]
            if pr[0] == "box":
                ce, u, h = _footprint(pr)
                bodies.append(dict(shape="rect", c=np.array(ce), u=np.array(u), h=h, o=np.array(ce), v=np.zeros(2), w=0.0,
                                   im=0.0, ii=0.0))
            else:
                bodies.append(dict(shape="circle", c=pr[1][:2].copy(), r=pr[2], o=pr[1][:2].copy(), v=np.zeros(2), w=0.0,
                                   im=0.0, ii=0.0))
            nodes.append(b)
        # broad phase: only bodies near a moving chassis matter; the pair loop itself is exact
        out = ph.contact_deltas(bodies)
        return [(nodes[k], out[k]) for k in range(n_veh) if bodies[k]["im"] > 0.0
                and (out[k][0].any() or out[k][1] != 0.0 or out[k][2].any())]

    def _contact_callbacks(self, chassis_nodes):
        if self.contact_cb is None:
            return
        for c in chassis_nodes:
            if c.world is not self or not c.notify:
                continue
            for b in self.bodies:
                if b is c or isinstance(b, BulletGhostNode):
                    continue
                if not self._groups_collide(c.into_mask, b.into_mask):
                    continue
                if any(p[0] in ("plane", "mesh") for p in _cached(b)[0]):
                    continue  # ground; sidewalk / crosswalk meshes are ignored by collision_callback.py:21-23
                if nodes_overlap(c, b):
                    self.contact_cb(_Contact(c, b))

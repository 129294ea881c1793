"""Float64 restatement of the Bullet calls on MetaDrive's step path. TEST INFRASTRUCTURE.

The arithmetic the reference delegates to `panda3d.bullet` (pinned panda3d==1.10.13, `setup.py:52`, which
bundles Bullet 2.8x; NEITHER source tree is under /root/reference) is restated here from the published
algorithms of upstream Bullet:
  btRaycastVehicle.cpp  (updateVehicle, rayCast, updateSuspension, updateFriction, resolveSingleBilateral,
                         calcRollingFriction), btDiscreteDynamicsWorld.cpp (internalSingleStepSimulation order),
  btTransformUtil.h     (integrateTransform), btCompoundShape.cpp (calculateLocalInertia from the AABB),
  panda3d bulletVehicle.cxx (create_wheel: suspension rest length 0.4 m; set_steering_value in degrees).
Constants above are RECALLED, not read: numeric parity with real Bullet is UNPINNED.
What this file pins is the reference's own Python running on top of it.

Reference call sites: `component/vehicle/base_vehicle.py:577-598,632-671` (chassis + wheels),
`:447-484` (actuation), `engine/core/engine_core.py:350-352` (doPhysics(0.02, 1, 0.02)),
`component/sensors/distance_detector.py:58-62` (rayTestClosest/All), `utils/pg/utils.py:174,255`
(vertical rayTestAll, sweep), `component/vehicle/base_vehicle.py:704-705` (contactTest).
"""
import math

import numpy as np

SUSPENSION_REST = 0.4  # panda3d bulletVehicle.cxx create_wheel default
MAX_SUSPENSION_FORCE = 6000.0  # btRaycastVehicle::btVehicleTuning default
GRAVITY = -9.81  # engine/core/physics_world.py:14
ANGULAR_MOTION_THRESHOLD = 0.5 * (math.pi / 2)
MAX_ANGVEL = math.pi / 2
SIDE_DAMPING = 0.2  # resolveSingleBilateral contactDamping


# ---------------------------------------------------------------------------------------------
# rigid body + raycast vehicle
# ---------------------------------------------------------------------------------------------
def _cross(a, b):
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])


def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array(
        [
            aw * bw - ax * bx - ay * by - az * bz,
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by - ax * bz + ay * bw + az * bx,
            aw * bz + ax * by - ay * bx + az * bw,
        ]
    )


def integrate_rotation(q, w, dt):
    """btTransformUtil::integrateTransform, rotation part. q = (w, x, y, z)."""
    ang = math.sqrt(float(w @ w))
    if ang * dt > ANGULAR_MOTION_THRESHOLD:
        ang = ANGULAR_MOTION_THRESHOLD / dt
    if ang < 0.001:
        axis = w * (0.5 * dt - (dt * dt * dt) * 0.020833333333 * ang * ang)
    else:
        axis = w * (math.sin(0.5 * ang * dt) / ang)
    dq = np.array([math.cos(ang * dt * 0.5), axis[0], axis[1], axis[2]])
    out = quat_mul(dq, q)
    return out / math.sqrt(float(out @ out))


class VehicleBody:
    """State + parameters of one chassis on four ray-cast wheels (local axes: x right, y forward, z up)."""
    def __init__(self, mass, width, length, height, wheels):
        self.mass = float(mass)
        lx, ly, lz = width, length, height
        # btCompoundShape::calculateLocalInertia uses the AABB of the (offset) child box, about the body origin
        self.inertia = self.mass / 12.0 * np.array([ly * ly + lz * lz, lx * lx + lz * lz, lx * lx + ly * ly])
        self.wheels = wheels  # list of dict(conn, radius, front, travel, stiffness, damp_relax, damp_comp, friction, roll)
        self.steering = [0.0] * len(wheels)  # rad
        self.engine_force = [0.0] * len(wheels)
        self.brake = [0.0] * len(wheels)

    def inv_inertia_world(self, R):
        return R @ np.diag(1.0 / self.inertia) @ R.T


def vehicle_substep_pre(body, pos, q, R, v, w, dt):
    """applyGravity + predictUnconstraintMotion + integrateTransforms for a body with no contacts."""
    v = v + np.array([0.0, 0.0, GRAVITY]) * dt
    wl = math.sqrt(float(w @ w))
    if wl * dt > MAX_ANGVEL:
        w = w * (MAX_ANGVEL / dt) / wl
    pos = pos + v * dt
    q = integrate_rotation(q, w, dt)
    return pos, q, v, w


def chassis_ground_clamp(body_half_w, body_half_l, pos, R, v):
    """Chassis-box vs ground-plane contact, restated as an inelastic clamp (Bullet resolves this contact with
    its sequential-impulse solver, restitution 0): the lowest bottom corner of the box (bottom face = body
    origin plane, base_vehicle.py:588-590) is lifted to z = 0 and the downward velocity is removed."""
    lowest = 0.0
    for sx in (-1.0, 1.0):
        for sy in (-1.0, 1.0):
            z = pos[2] + R[2, 0] * sx * body_half_w + R[2, 1] * sy * body_half_l
            lowest = min(lowest, z)
    if lowest < 0.0:
        pos = pos.copy()
        pos[2] -= lowest
        if v[2] < 0.0:
            v = v.copy()
            v[2] = 0.0
    return pos, v


def update_vehicle(body, pos, R, v, w, dt):
    """btRaycastVehicle::updateVehicle on the plane z = 0 (normal +z, `engine/core/terrain.py:157-175`)."""
    n_w = len(body.wheels)
    inv_m = 1.0 / body.mass
    inv_I = body.inv_inertia_world(R)
    normal = np.array([0.0, 0.0, 1.0])
    up = R[:, 2]
    right = R[:, 0]
    contact = [None] * n_w
    susp_force = [0.0] * n_w

    def vel_at(rel):
        return v + _cross(w, rel)

    # rayCast + updateSuspension
    for i, wh in enumerate(body.wheels):
        hard = pos + R @ wh["conn"]
        d = -up  # wheelDirectionWS
        raylen = SUSPENSION_REST + wh["radius"]
        hit = None
        if d[2] < 0 and hard[2] > 0:
            t = hard[2] / (-d[2] * raylen)
            if t <= 1.0:
                hit = t
        if hit is None:
            continue
        cp = hard + d * (raylen * hit)
        susp_len = hit * raylen - wh["radius"]
        lo = SUSPENSION_REST - wh["travel"] * 0.01
        hi = SUSPENSION_REST + wh["travel"] * 0.01
        susp_len = min(max(susp_len, lo), hi)
        denom = float(normal @ d)
        rel = cp - pos
        proj_vel = float(normal @ vel_at(rel))
        if denom >= -0.1:
            rel_vel = 0.0
            clipped_inv = 10.0
        else:
            inv = -1.0 / denom
            rel_vel = proj_vel * inv
            clipped_inv = inv
        force = wh["stiffness"] * (SUSPENSION_REST - susp_len) * clipped_inv
        damp = wh["damp_comp"] if rel_vel < 0.0 else wh["damp_relax"]
        force -= damp * rel_vel
        f = max(force * body.mass, 0.0)
        susp_force[i] = f
        contact[i] = cp

    # apply suspension impulses
    for i in range(n_w):
        if contact[i] is None:
            continue
        f = min(susp_force[i], MAX_SUSPENSION_FORCE)
        imp = normal * (f * dt)
        rel = contact[i] - pos
        v = v + imp * inv_m
        w = w + inv_I @ _cross(rel, imp)

    # updateFriction
    n_ground = sum(1 for c in contact if c is not None)
    if n_ground == 0:
        return v, w
    axle = [None] * n_w
    fwd = [None] * n_w
    side_imp = [0.0] * n_w
    fwd_imp = [0.0] * n_w
    for i, wh in enumerate(body.wheels):
        if contact[i] is None:
            continue
        s = body.steering[i]
        # steeringMat (rotation about `up` by s) applied to the chassis right axis
        a = right * math.cos(s) + _cross(up, right) * math.sin(s) + up * float(up @ right) * (1 - math.cos(s))
        a = a - normal * float(a @ normal)
        a = a / math.sqrt(float(a @ a))
        f_ = _cross(normal, a)
        f_ = f_ / math.sqrt(float(f_ @ f_))
        axle[i], fwd[i] = a, f_
        rel = contact[i] - pos
        # resolveSingleBilateral against the fixed ground body
        aj = R.T @ _cross(rel, a)
        jac_diag = inv_m + float(aj @ (aj / body.inertia))
        rel_vel = float(a @ vel_at(rel))
        side_imp[i] = -SIDE_DAMPING * rel_vel / jac_diag

    sliding = False
    skid = [1.0] * n_w
    for i, wh in enumerate(body.wheels):
        if contact[i] is None:
            continue
        if body.engine_force[i] != 0.0:
            rolling = body.engine_force[i] * dt
        else:
            max_imp = body.brake[i] if body.brake[i] else 0.0
            rel = contact[i] - pos
            c0 = _cross(rel, fwd[i])
            denom0 = inv_m + float(fwd[i] @ _cross(inv_I @ c0, rel))
            vrel = float(fwd[i] @ vel_at(rel))
            j1 = -vrel / denom0 / float(n_ground)
            rolling = min(max(j1, -max_imp), max_imp)
        fwd_imp[i] = rolling
        maximp = susp_force[i] * dt * wh["friction"]
        x = fwd_imp[i] * 0.5
        y = side_imp[i] * 1.0
        imp2 = x * x + y * y
        if imp2 > maximp * maximp:
            sliding = True
            skid[i] *= maximp / math.sqrt(imp2)
    if sliding:
        for i in range(n_w):
            if side_imp[i] != 0.0 and skid[i] < 1.0:
                fwd_imp[i] *= skid[i]
                side_imp[i] *= skid[i]

    # apply the impulses
    for i, wh in enumerate(body.wheels):
        if contact[i] is None:
            continue
        rel = contact[i] - pos
        if fwd_imp[i] != 0.0:
            imp = fwd[i] * fwd_imp[i]
            v = v + imp * inv_m
            w = w + inv_I @ _cross(rel, imp)
        if side_imp[i] != 0.0:
            imp = axle[i] * side_imp[i]
            rel2 = rel - up * (float(up @ rel) * (1.0 - wh["roll"]))  # ROLLING_INFLUENCE_FIX
            v = v + imp * inv_m
            w = w + inv_I @ _cross(rel2, imp)
    return v, w


# ---------------------------------------------------------------------------------------------
# geometry queries (analytic substitutes for rayTest / contactTest / sweepTest)
# ---------------------------------------------------------------------------------------------
def ray_obb(o, d, c, R, h):
    """Ray o + t d, t in [0,1] against an oriented box. Returns entry t or None (origin inside -> 0)."""
    ol = R.T @ (o - c)
    dl = R.T @ d
    t0, t1 = 0.0, 1.0
    for k in range(3):
        if abs(dl[k]) < 1e-12:
            if abs(ol[k]) > h[k]:
                return None
        else:
            ta = (-h[k] - ol[k]) / dl[k]
            tb = (h[k] - ol[k]) / dl[k]
            if ta > tb:
                ta, tb = tb, ta
            t0 = max(t0, ta)
            t1 = min(t1, tb)
            if t0 > t1:
                return None
    return t0


def ray_zcyl(o, d, c, r, half_h):
    """Ray vs upright cylinder (axis +z) centred at c. Handles side and caps."""
    t0, t1 = 0.0, 1.0
    # z slab
    if abs(d[2]) < 1e-12:
        if abs(o[2] - c[2]) > half_h:
            return None
    else:
        ta = (c[2] - half_h - o[2]) / d[2]
        tb = (c[2] + half_h - o[2]) / d[2]
        if ta > tb:
            ta, tb = tb, ta
        t0, t1 = max(t0, ta), min(t1, tb)
        if t0 > t1:
            return None
    ox, oy = o[0] - c[0], o[1] - c[1]
    a = d[0] * d[0] + d[1] * d[1]
    b = ox * d[0] + oy * d[1]
    cc = ox * ox + oy * oy - r * r
    if a < 1e-18:
        if cc > 0:
            return None
        return t0
    disc = b * b - a * cc
    if disc < 0:
        return None
    sq = math.sqrt(disc)
    ta = (-b - sq) / a
    tb = (-b + sq) / a
    t0, t1 = max(t0, ta), min(t1, tb)
    if t0 > t1:
        return None
    return t0


HULL_TOL = 1e-3  # m^2 on the edge cross product (sub-millimetre); Bullet's own hull carries a 0.04 m margin


def point_in_convex(poly, p):
    """poly: (n,2) CCW convex hull vertices. Points on an edge count as inside (see HULL_TOL)."""
    x, y = p
    n = len(poly)
    for i in range(n):
        ax, ay = poly[i]
        bx, by = poly[(i + 1) % n]
        if (bx - ax) * (y - ay) - (by - ay) * (x - ax) < -HULL_TOL:
            return False
    return True


def convex_hull(points):
    """Andrew monotone chain, CCW, no collinear points."""
    pts = sorted(set((float(p[0]), float(p[1])) for p in points))
    if len(pts) <= 2:
        return np.array(pts)

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    lower = []
    for p in pts:
        while len(lower) >= 2 and cross(lower[-2], lower[-1], p) <= 0:
            lower.pop()
        lower.append(p)
    upper = []
    for p in reversed(pts):
        while len(upper) >= 2 and cross(upper[-2], upper[-1], p) <= 0:
            upper.pop()
        upper.append(p)
    return np.array(lower[:-1] + upper[:-1])


def rect_corners(c, heading_vec, half_l, half_w):
    hx, hy = heading_vec
    ax = np.array([hx, hy]) * half_l
    ay = np.array([-hy, hx]) * half_w
    c = np.asarray(c[:2], dtype=np.float64)
    return np.array([c + ax + ay, c - ax + ay, c - ax - ay, c + ax - ay])


def obb2d_overlap(c0, u0, h0, c1, u1, h1):
    """2-D SAT for two rectangles: centre, unit axis u (forward), half extents (along u, across u)."""
    axes = [np.asarray(u0), np.array([-u0[1], u0[0]]), np.asarray(u1), np.array([-u1[1], u1[0]])]
    d = np.asarray(c1[:2], dtype=np.float64) - np.asarray(c0[:2], dtype=np.float64)
    a0 = [np.asarray(u0) * h0[0], np.array([-u0[1], u0[0]]) * h0[1]]
    a1 = [np.asarray(u1) * h1[0], np.array([-u1[1], u1[0]]) * h1[1]]
    for ax in axes:
        r0 = abs(float(ax @ a0[0])) + abs(float(ax @ a0[1]))
        r1 = abs(float(ax @ a1[0])) + abs(float(ax @ a1[1]))
        if abs(float(ax @ d)) > r0 + r1:
            return False
    return True


def obb2d_circle_overlap(c0, u0, h0, c1, r):
    d = np.asarray(c1[:2], dtype=np.float64) - np.asarray(c0[:2], dtype=np.float64)
    lx = float(d @ np.asarray(u0))
    ly = float(d @ np.array([-u0[1], u0[0]]))
    qx = min(max(lx, -h0[0]), h0[0])
    qy = min(max(ly, -h0[1]), h0[1])
    return (lx - qx)**2 + (ly - qy)**2 <= r * r


def seg_intersect(p, p2, q, q2):
    def orient(a, b, c):
        return (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0])

    d1 = orient(q, q2, p)
    d2 = orient(q, q2, p2)
    d3 = orient(p, p2, q)
    d4 = orient(p, p2, q2)
    return (d1 * d2 < 0) and (d3 * d4 < 0)


def point_in_polygon(poly, p):
    x, y = p
    inside = False
    n = len(poly)
    j = n - 1
    for i in range(n):
        xi, yi = poly[i]
        xj, yj = poly[j]
        if (yi > y) != (yj > y):
            if x < (xj - xi) * (y - yi) / (yj - yi) + xi:
                inside = not inside
        j = i
    return inside


def rect_polygon_overlap(corners, poly):
    """Footprint rectangle (4 corners) vs simple polygon (n,2)."""
    for c in corners:
        if point_in_polygon(poly, c):
            return True
    rect = [tuple(c) for c in corners]
    for p in poly:
        # point in convex rectangle
        if point_in_convex(np.array(rect) if _ccw(rect) else np.array(rect[::-1]), p):
            return True
    n = len(poly)
    for i in range(4):
        a, b = corners[i], corners[(i + 1) % 4]
        for j in range(n):
            if seg_intersect(a, b, poly[j], poly[(j + 1) % n]):
                return True
    return False


def _ccw(pts):
    s = 0.0
    for i in range(len(pts)):
        x0, y0 = pts[i]
        x1, y1 = pts[(i + 1) % len(pts)]
        s += x0 * y1 - x1 * y0
    return s > 0


# ---------------------------------------------------------------------------------------------
# contact response between the bodies of the dynamic world
# ---------------------------------------------------------------------------------------------
# Bullet resolves chassis-chassis and chassis-obstacle contacts inside doPhysics with its sequential-impulse solver over
# persistent GJK/EPA manifolds (btSequentialImpulseConstraintSolver, btConvexConvexAlgorithm; NOT in /root/reference).
# That machinery is not restated; SURVEY.md 7.3 #2 asks for "a simple impulse model, stated as such".  The model, shared
# word for word by oracle/md_oracle.c and the CUDA kernel (DESIGN.md "Contact response"):
#   * planar (x, y, yaw) contacts between the chassis footprints (rectangles) and static obstacles (cone / warning
#     cylinders -> circles, barrier -> rectangle); pedestrians and other free movers are not pushed and do not push;
#   * per sub-step, after every body has moved: one Jacobi pass over the overlapping pairs from a snapshot of the
#     post-move state - normal = minimum-translation axis (A -> B, A = the lower index), contact point = mean of the
#     corners of either rectangle lying inside the other (circle: closest point of the rectangle), frictionless
#     inelastic normal impulse J = max(0, -v_n) / K with K = 1/mA + 1/mB + (rA x n)^2 / IzA + (rB x n)^2 / IzB, plus a
#     position push-out of CONTACT_ERP x (depth - CONTACT_SLOP) split by inverse mass (Bullet's erp is 0.2 as well);
#   * static bodies (mass 0: obstacles, wrecks, broken-down cars) have 1/m = 1/Iz = 0 and never move.
CONTACT_ERP = 0.2
CONTACT_SLOP = 0.01
CORNER_EPS = 1e-4
CONTACT_TIE = 1e-6
AXIS_MARGIN = 1e-3


def _perp(u):
    return np.array([-u[1], u[0]])


def _corners_inside(c_in, u_in, h_in, c_box, u_box, h_box):
    """corners of rectangle `in` that lie inside rectangle `box` -> (count, sum_x, sum_y)"""
    n, sx, sy = 0, 0.0, 0.0
    vb = _perp(u_box)
    vi = _perp(u_in)
    for su, sv in ((1.0, 1.0), (-1.0, 1.0), (-1.0, -1.0), (1.0, -1.0)):
        p = c_in + u_in * (su * h_in[0]) + vi * (sv * h_in[1])
        d = p - c_box
        if abs(float(d @ u_box)) <= h_box[0] + CORNER_EPS and abs(float(d @ vb)) <= h_box[1] + CORNER_EPS:
            n += 1
            sx += p[0]
            sy += p[1]
    return n, sx, sy


def rect_rect_contact(cA, uA, hA, cB, uB, hB):
    """(n, depth, p) with n the unit normal from A to B, or None when the rectangles do not overlap."""
    cA, uA, cB, uB = (np.asarray(x, dtype=np.float64) for x in (cA, uA, cB, uB))
    d = cB - cA
    cand = []
    for a in (uA, _perp(uA), uB, _perp(uB)):
        rA = abs(float(a @ uA)) * hA[0] + abs(float(a @ _perp(uA))) * hA[1]
        rB = abs(float(a @ uB)) * hB[0] + abs(float(a @ _perp(uB))) * hB[1]
        s = float(a @ d)
        ov = rA + rB - abs(s)
        if ov < 0.0:
            return None
        cand.append((ov, a if s > -CONTACT_TIE else -a))  # centres level along the axis: A -> B is "+axis" by convention
    # minimum-translation axis.  Two nearly parallel boxes overlap by almost the same amount along A's and B's axis, and
    # the pick would flip on rounding noise: B's axes win only by a clear margin (AXIS_MARGIN)
    bestA = cand[1] if cand[1][0] < cand[0][0] else cand[0]
    bestB = cand[3] if cand[3][0] < cand[2][0] else cand[2]
    best, n = bestB if bestB[0] < bestA[0] - AXIS_MARGIN else bestA
    k0, x0, y0 = _corners_inside(cB, uB, hB, cA, uA, hA)
    k1, x1, y1 = _corners_inside(cA, uA, hA, cB, uB, hB)
    if k0 + k1 > 0:
        p = np.array([(x0 + x1) / (k0 + k1), (y0 + y1) / (k0 + k1)])
    else:
        p = 0.5 * (cA + cB)
    return n, best, p


def rect_circle_contact(cA, uA, hA, cB, r):
    cA, uA, cB = (np.asarray(x, dtype=np.float64) for x in (cA, uA, cB))
    vA = _perp(uA)
    d = cB - cA
    lx, ly = float(d @ uA), float(d @ vA)
    qx, qy = min(max(lx, -hA[0]), hA[0]), min(max(ly, -hA[1]), hA[1])
    ex, ey = lx - qx, ly - qy
    d2 = ex * ex + ey * ey
    if d2 > r * r:
        return None
    if d2 > 1e-12:  # centre outside the rectangle: normal from the closest point to the centre
        dist = math.sqrt(d2)
        n = uA * (ex / dist) + vA * (ey / dist)
        return n, r - dist, cA + uA * qx + vA * qy
    dx, dy = hA[0] - abs(lx), hA[1] - abs(ly)  # centre inside: leave through the nearest face
    if dx < dy:
        sgn = 1.0 if lx >= 0.0 else -1.0
        return uA * sgn, dx + r, cA + uA * (sgn * hA[0]) + vA * ly
    sgn = 1.0 if ly >= 0.0 else -1.0
    return vA * sgn, dy + r, cA + uA * lx + vA * (sgn * hA[1])


def contact_deltas(bodies):
    """bodies: list of dict(shape='rect'|'circle', c, u, h | r, o, v, w, im, ii) ordered by index; returns per body
    (dv[2], dw, dp[2]) of one Jacobi pass.  Bodies with im == 0 get no deltas."""
    out = [(np.zeros(2), 0.0, np.zeros(2)) for _ in bodies]
    rad = [b["r"] if b["shape"] == "circle" else math.hypot(b["h"][0], b["h"][1]) for b in bodies]
    for i, A in enumerate(bodies):
        for j in range(i + 1, len(bodies)):
            B = bodies[j]
            if A["im"] == 0.0 and B["im"] == 0.0:
                continue
            dc = B["c"] - A["c"]
            if dc[0] * dc[0] + dc[1] * dc[1] > (rad[i] + rad[j])**2:
                continue  # bounding circles apart: the exact tests below cannot overlap
            if A["shape"] == "circle" and B["shape"] == "circle":
                continue
            if A["shape"] == "rect" and B["shape"] == "rect":
                res = rect_rect_contact(A["c"], A["u"], A["h"], B["c"], B["u"], B["h"])
            elif A["shape"] == "rect":
                res = rect_circle_contact(A["c"], A["u"], A["h"], B["c"], B["r"])
            else:
                res = rect_circle_contact(B["c"], B["u"], B["h"], A["c"], A["r"])
                if res is not None:
                    res = (-res[0], res[1], res[2])
            if res is None:
                continue
            n, depth, p = res
            rA, rB = p - A["o"], p - B["o"]
            vpA = A["v"] + A["w"] * np.array([-rA[1], rA[0]])
            vpB = B["v"] + B["w"] * np.array([-rB[1], rB[0]])
            vn = float((vpB - vpA) @ n)
            cAn = rA[0] * n[1] - rA[1] * n[0]
            cBn = rB[0] * n[1] - rB[1] * n[0]
            K = A["im"] + B["im"] + cAn * cAn * A["ii"] + cBn * cBn * B["ii"]
            J = -vn / K if vn < 0.0 else 0.0
            push = CONTACT_ERP * max(depth - CONTACT_SLOP, 0.0) / (A["im"] + B["im"])
            for k, sgn, body, cn in ((i, -1.0, A, cAn), (j, 1.0, B, cBn)):
                if body["im"] == 0.0:
                    continue
                dv, dw, dp = out[k]
                out[k] = (dv + n * (sgn * J * body["im"]), dw + sgn * J * cn * body["ii"], dp + n * (sgn * push * body["im"]))
    return out

"""The contact-response model (DESIGN.md "Contact response"): the float64 definition in oracle/refshim/physics.py against
the float32 C oracle's geometry on random footprints, and the physical properties a single inelastic contact must have
(momentum conserved, no approach velocity left at the contact point, static bodies unmoved)."""
import ctypes as C

import numpy as np

from oracle.refshim import physics as ph


def _rand_rect(rng, spread=3.0):
    a = rng.uniform(-np.pi, np.pi)
    return np.array([rng.uniform(-spread, spread), rng.uniform(-spread, spread), np.cos(a), np.sin(a),
                     rng.uniform(0.8, 3.0), rng.uniform(0.15, 1.2)])


def test_rect_rect_contact_matches_definition(oracle_lib):
    rng = np.random.RandomState(0)
    out = np.zeros(5, np.float32)
    hits = 0
    for _ in range(4000):
        a, b = _rand_rect(rng), _rand_rect(rng)
        a32, b32 = a.astype(np.float32), b.astype(np.float32)
        got = oracle_lib.mdo_rr_contact(a32.ctypes.data_as(C.c_void_p), b32.ctypes.data_as(C.c_void_p),
                                        out.ctypes.data_as(C.c_void_p))
        a, b = a32.astype(np.float64), b32.astype(np.float64)
        ref = ph.rect_rect_contact(a[:2], a[2:4], a[4:6], b[:2], b[2:4], b[4:6])
        ov = ph.obb2d_overlap(a[:2], a[2:4], a[4:6], b[:2], b[2:4], b[4:6])
        if ref is not None and ref[1] < 1e-4:
            continue  # grazing: either verdict is fine
        assert bool(got) == (ref is not None) == bool(ov)
        if ref is None:
            continue
        hits += 1
        n, depth, p = ref
        assert abs(out[2] - depth) < 1e-4
        # the minimum-translation axis may tie between two axes (squares, parallel faces): compare only clear minima
        if abs(float(np.dot(out[:2], n)) - 1.0) > 1e-3:
            continue
        np.testing.assert_allclose(out[3:5], p, atol=2e-3)
        assert float(np.dot(n, b[:2] - a[:2])) >= -1e-9  # A -> B
    assert hits > 500


def test_rect_circle_contact_matches_definition(oracle_lib):
    rng = np.random.RandomState(1)
    out = np.zeros(5, np.float32)
    hits = 0
    for _ in range(4000):
        a = _rand_rect(rng).astype(np.float32)
        c = np.array([rng.uniform(-3, 3), rng.uniform(-3, 3), rng.choice([0.2, 0.5])], np.float32)
        got = oracle_lib.mdo_rc_contact(a.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        a64 = a.astype(np.float64)
        ref = ph.rect_circle_contact(a64[:2], a64[2:4], a64[4:6], c[:2].astype(np.float64), float(c[2]))
        if ref is not None and ref[1] < 1e-4:
            continue
        assert bool(got) == (ref is not None)
        if ref is None:
            continue
        n, depth, p = ref
        d = c[:2].astype(np.float64) - a64[:2]
        lx, ly = abs(d @ a64[2:4]), abs(d @ np.array([-a64[3], a64[2]]))
        if lx < a64[4] and ly < a64[5] and abs((a64[4] - lx) - (a64[5] - ly)) < 1e-3:
            continue  # centre inside, equally far from two faces
        hits += 1
        assert abs(out[2] - depth) < 1e-4
        np.testing.assert_allclose(out[:2], n, atol=1e-3)
        np.testing.assert_allclose(out[3:5], p, atol=1e-3)
    assert hits > 300


def _body(c, ang, h, v, w, mass):
    u = np.array([np.cos(ang), np.sin(ang)])
    izz = mass / 12.0 * ((2 * h[0])**2 + (2 * h[1])**2) if mass > 0 else 0.0
    return dict(shape="rect", c=np.array(c, float), u=u, h=h, o=np.array(c, float), v=np.array(v, float), w=w,
                im=1.0 / mass if mass > 0 else 0.0, ii=1.0 / izz if mass > 0 else 0.0)


def test_single_contact_is_inelastic_and_conserves_momentum():
    rng = np.random.RandomState(2)
    n_hit = 0
    for _ in range(300):
        A = _body([0, 0], rng.uniform(-3, 3), (2.2, 0.9), rng.uniform(-8, 8, 2), rng.uniform(-0.5, 0.5), 1100.0)
        B = _body(rng.uniform(-3, 3, 2), rng.uniform(-3, 3), (2.4, 1.0), rng.uniform(-8, 8, 2), rng.uniform(-0.5, 0.5), 1400.0)
        res = ph.rect_rect_contact(A["c"], A["u"], A["h"], B["c"], B["u"], B["h"])
        if res is None:
            continue
        n, depth, p = res
        (dvA, dwA, dpA), (dvB, dwB, dpB) = ph.contact_deltas([A, B])
        # linear and angular momentum (about the contact point both impulses act through) are conserved
        np.testing.assert_allclose(dvA / A["im"] + dvB / B["im"], 0.0, atol=1e-8)
        rA, rB = p - A["o"], p - B["o"]
        vp = lambda body, dv, dw, r: body["v"] + dv + (body["w"] + dw) * np.array([-r[1], r[0]])
        vn0 = float((vp(B, 0, 0, rB) - vp(A, 0, 0, rA)) @ n)
        vn1 = float((vp(B, dvB, dwB, rB) - vp(A, dvA, dwA, rA)) @ n)
        if vn0 < 0:
            n_hit += 1
            assert abs(vn1) < 1e-9          # approaching: the normal approach velocity is removed, nothing bounces
        else:
            assert abs(vn1 - vn0) < 1e-12   # separating: no impulse
        # the push-out separates along the normal, split by inverse mass
        assert float((dpB - dpA) @ n) >= 0.0
        np.testing.assert_allclose(dpA / A["im"] + dpB / B["im"], 0.0, atol=1e-8)
    assert n_hit > 30


def test_static_obstacle_stops_the_normal_velocity_and_never_moves():
    car = _body([0.0, 0.0], 0.0, (2.2, 0.9), [6.0, 0.5], 0.0, 1100.0)
    cone = dict(shape="circle", c=np.array([2.3, 0.2]), r=0.2, o=np.array([2.3, 0.2]), v=np.zeros(2), w=0.0, im=0.0, ii=0.0)
    (dv, dw, dp), (dvc, dwc, dpc) = ph.contact_deltas([car, cone])
    assert not dvc.any() and dwc == 0.0 and not dpc.any()
    n, depth, p = ph.rect_circle_contact(car["c"], car["u"], car["h"], cone["c"], cone["r"])
    r = p - car["o"]
    v_after = car["v"] + dv + (car["w"] + dw) * np.array([-r[1], r[0]])
    assert abs(float(v_after @ n)) < 1e-9 and float(dp @ n) < 0.0

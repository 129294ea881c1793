"""include/md_math.h: the float32 sin / cos / atan2 / exp shared by the CUDA kernels and the CPU oracle (so that the two
produce the same bits) against double-precision libm, on the ranges the step path uses."""
import ctypes as C

import numpy as np


def _call(lib, op, x, y=None):
    x = np.ascontiguousarray(x, np.float32)
    y = np.ascontiguousarray(y if y is not None else np.zeros_like(x), np.float32)
    out = np.zeros_like(x)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.mdo_math(C.c_int(op), C.c_int(len(x)), p(x), p(y), p(out))
    return out


def test_sin_cos_within_two_ulp(oracle_lib):
    rng = np.random.RandomState(0)
    x = np.concatenate([rng.uniform(-1e4, 1e4, 200000), rng.uniform(-7, 7, 200000), rng.uniform(-0.8, 0.8, 200000),
                        np.arange(-64, 65) * (np.pi / 4), [0.0, -0.0, 1e-8, -1e-8]]).astype(np.float32)
    xd = x.astype(np.float64)
    assert np.abs(_call(oracle_lib, 0, x) - np.sin(xd)).max() < 1.3e-7     # 2 ulp of a value near 1
    assert np.abs(_call(oracle_lib, 1, x) - np.cos(xd)).max() < 1.3e-7
    s, c = _call(oracle_lib, 0, x), _call(oracle_lib, 1, x)
    assert np.abs(s * s + c * c - 1.0).max() < 4e-7
    assert _call(oracle_lib, 0, [0.0])[0] == 0.0 and _call(oracle_lib, 1, [0.0])[0] == 1.0


def test_atan2_and_exp(oracle_lib):
    rng = np.random.RandomState(1)
    y, x = rng.uniform(-60, 60, 400000).astype(np.float32), rng.uniform(-60, 60, 400000).astype(np.float32)
    got = _call(oracle_lib, 2, y, x)
    assert np.abs(got - np.arctan2(y.astype(np.float64), x.astype(np.float64))).max() < 5e-7   # 2 ulp of a value near pi
    edge_y = np.array([0, 0, 1, -1, 0, 1e-30, -1e-30], np.float32)
    edge_x = np.array([1, -1, 0, 0, 0, -1, -1], np.float32)
    np.testing.assert_allclose(_call(oracle_lib, 2, edge_y, edge_x), [0, np.pi, np.pi / 2, -np.pi / 2, 0, np.pi, -np.pi], atol=3e-7)
    e = rng.uniform(-2, 3, 200000).astype(np.float32)
    rel = np.abs(_call(oracle_lib, 3, e) / np.exp(e.astype(np.float64)) - 1.0)
    assert rel.max() < 1.5e-7
    assert _call(oracle_lib, 3, [0.0])[0] == 1.0

"""Monkey patches applied to the reference AFTER stubs are installed. TEST INFRASTRUCTURE.

Only rendering / asset / shapely entry points are replaced; no step-path logic is touched.
"""
import numpy as np


def _clip_polygon(cls, polygon):
    """Stand-in for `TerrainProperty.clip_polygon` (`metadrive/constants.py:510-533`, shapely-based):
    Sutherland-Hodgman clip against the square map region. PG maps fit inside, so this is the identity there."""
    half = cls.map_region_size / 2
    pts = [(float(p[0]), float(p[1])) for p in polygon]
    if len(pts) < 3:
        return None
    if all(-half <= x <= half and -half <= y <= half for x, y in pts):
        if pts[0] != pts[-1]:
            pts = pts + [pts[0]]  # shapely's exterior.coords is closed
        return [pts]

    def clip(pts, inside, inter):
        out = []
        for i in range(len(pts)):
            a, b = pts[i - 1], pts[i]
            ia, ib = inside(a), inside(b)
            if ib:
                if not ia:
                    out.append(inter(a, b))
                out.append(b)
            elif ia:
                out.append(inter(a, b))
        return out

    def ix(xc):
        return lambda a, b: (xc, a[1] + (b[1] - a[1]) * (xc - a[0]) / (b[0] - a[0]))

    def iy(yc):
        return lambda a, b: (a[0] + (b[0] - a[0]) * (yc - a[1]) / (b[1] - a[1]), yc)

    for inside, inter in ((lambda p: p[0] >= -half, ix(-half)), (lambda p: p[0] <= half, ix(half)),
                          (lambda p: p[1] >= -half, iy(-half)), (lambda p: p[1] <= half, iy(half))):
        pts = clip(pts, inside, inter)
        if len(pts) < 3:
            return None
    return [pts + [pts[0]]]


class _FakeGeom:
    def __init__(self, polygon, height):
        self.polygon = np.asarray(polygon, dtype=np.float64)[:, :2]
        self.height = float(height)


class _FakeGeomNode:
    def __init__(self, geom):
        self._geom = geom

    def getGeom(self, i):
        return self._geom


def _make_polygon_model(points, height, auto_anticlockwise=True, force_anticlockwise=False, texture_scale=0.1):
    """Stand-in for `metadrive/utils/vertex.py:make_polygon_model`: keeps the polygon + extrusion height."""
    from .pcore import NodePath
    np_ = NodePath("polygon_model")
    geom = _FakeGeom(points, height)
    np_.node = lambda: _FakeGeomNode(geom)
    return np_


def apply():
    from metadrive.constants import TerrainProperty
    TerrainProperty.clip_polygon = classmethod(_clip_polygon)

    import metadrive.component.block.base_block as bb
    bb.make_polygon_model = _make_polygon_model

    import metadrive.engine.base_engine as be
    be.BaseEngine.try_pull_asset = staticmethod(lambda: None)
    be.BaseEngine.warmup = lambda self: None

    # physics nodes print assertion noise at interpreter exit (engine/physics_node.py:27-29)
    import metadrive.engine.physics_node as pn
    pn.BaseRigidBodyNode.__del__ = lambda self: None
    pn.BaseGhostBodyNode.__del__ = lambda self: None

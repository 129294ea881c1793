"""Semantic stand-in for the slice of `panda3d.core` that MetaDrive's step path uses. TEST INFRASTRUCTURE.

Conventions restated from Panda3D 1.10 (source not under /root/reference; the reference relies on them at
`metadrive/base_class/base_object.py:291-398`, `metadrive/component/vehicle/base_vehicle.py:983-1001`):
  * Z-up right-handed; HPR = heading about +Z, pitch about +X, roll about +Y, angles in degrees;
    column-vector rotation  M = Rz(h) * Rx(p) * Ry(r).
  * `NodePath.getRelativeVector(other, v)` rotates `v` from `other`'s frame into this node's frame.
All scene-graph parents on the step path carry identity transforms, so a node's transform is treated
as its world transform.
"""
import math

import numpy as np

from .stubs import _mk


class VecBase:
    __slots__ = ("v", )
    N = 3

    def __init__(self, *a):
        if len(a) == 0:
            self.v = np.zeros(self.N)
        elif len(a) == 1:
            if isinstance(a[0], (int, float, np.floating, np.integer)):
                self.v = np.full(self.N, float(a[0]))
            else:
                self.v = np.array([float(x) for x in a[0]], dtype=np.float64)
        else:
            self.v = np.array([float(x) for x in a], dtype=np.float64)
        assert self.v.shape == (self.N, ), (self.v, a)

    def __getitem__(self, i):
        return float(self.v[i]) if isinstance(i, (int, np.integer)) else self.v[i]

    def __setitem__(self, i, val):
        self.v[i] = val

    def __len__(self):
        return self.N

    def __iter__(self):
        return iter(float(x) for x in self.v)

    def __array__(self, dtype=None, copy=None):
        return self.v.astype(dtype) if dtype is not None else self.v.copy()

    def _w(self, arr):
        return type(self)(*arr)

    def __add__(self, o):
        return self._w(self.v + np.asarray(o, dtype=np.float64))

    __radd__ = __add__

    def __sub__(self, o):
        return self._w(self.v - np.asarray(o, dtype=np.float64))

    def __rsub__(self, o):
        return self._w(np.asarray(o, dtype=np.float64) - self.v)

    def __mul__(self, s):
        return self._w(self.v * float(s))

    __rmul__ = __mul__

    def __truediv__(self, s):
        return self._w(self.v / float(s))

    def __neg__(self):
        return self._w(-self.v)

    def __eq__(self, o):
        try:
            return bool(np.all(self.v == np.asarray(o, dtype=np.float64)))
        except Exception:
            return False

    def __hash__(self):
        return hash(tuple(self.v))

    def __repr__(self):
        return "%s(%s)" % (type(self).__name__, ", ".join("%g" % x for x in self.v))

    def length(self):
        return float(np.linalg.norm(self.v))

    def lengthSquared(self):
        return float(self.v @ self.v)

    def dot(self, o):
        return float(self.v @ np.asarray(o, dtype=np.float64))

    def normalize(self):
        n = self.length()
        if n > 0:
            self.v /= n
        return n > 0

    def normalized(self):
        n = self.length()
        return self._w(self.v / n) if n > 0 else self._w(self.v)

    def getX(self):
        return float(self.v[0])

    def getY(self):
        return float(self.v[1])

    def getZ(self):
        return float(self.v[2])

    get_x, get_y, get_z = getX, getY, getZ

    @property
    def x(self):
        return float(self.v[0])

    @property
    def y(self):
        return float(self.v[1])

    @property
    def z(self):
        return float(self.v[2])

    @property
    def xy(self):
        return self.v[:2].copy()


class Vec3(VecBase):
    N = 3

    def cross(self, o):
        return Vec3(*np.cross(self.v, np.asarray(o, dtype=np.float64)))


LVector3 = LVector3f = LVecBase3 = LVecBase3f = LPoint3 = LPoint3f = Point3 = LVector3d = Vec3


class Vec4(VecBase):
    N = 4


LVecBase4 = LVecBase4f = LVector4 = LVector4f = LPoint4 = Vec4


class Vec2(VecBase):
    N = 2


LVector2 = LVecBase2 = LPoint2 = LVector2f = LPoint2f = Vec2


class LQuaternionf(VecBase):
    """(w, x, y, z), Panda order."""
    N = 4

    def to_matrix(self):
        return quat_to_mat(self.v)


LQuaternion = Quat = LQuaternionf


def quat_to_mat(q):
    w, x, y, z = q
    n = w * w + x * x + y * y + z * z
    s = 2.0 / n if n > 0 else 0.0
    return np.array(
        [
            [1 - s * (y * y + z * z), s * (x * y - w * z), s * (x * z + w * y)],
            [s * (x * y + w * z), 1 - s * (x * x + z * z), s * (y * z - w * x)],
            [s * (x * z - w * y), s * (y * z + w * x), 1 - s * (x * x + y * y)],
        ]
    )


def mat_to_quat(m):
    t = np.trace(m)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        return np.array([0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s])
    i = int(np.argmax(np.diag(m)))
    j, k = (i + 1) % 3, (i + 2) % 3
    s = math.sqrt(1.0 + m[i, i] - m[j, j] - m[k, k]) * 2
    q = np.zeros(4)
    q[0] = (m[k, j] - m[j, k]) / s
    q[1 + i] = 0.25 * s
    q[1 + j] = (m[j, i] + m[i, j]) / s
    q[1 + k] = (m[k, i] + m[i, k]) / s
    return q


def hpr_to_mat(h, p, r):
    """Degrees -> column-vector rotation Rz(h) Rx(p) Ry(r)."""
    h, p, r = math.radians(h), math.radians(p), math.radians(r)
    ch, sh, cp, sp, cr, sr = math.cos(h), math.sin(h), math.cos(p), math.sin(p), math.cos(r), math.sin(r)
    rz = np.array([[ch, -sh, 0], [sh, ch, 0], [0, 0, 1.0]])
    rx = np.array([[1.0, 0, 0], [0, cp, -sp], [0, sp, cp]])
    ry = np.array([[cr, 0, sr], [0, 1.0, 0], [-sr, 0, cr]])
    return rz @ rx @ ry


def mat_to_hpr(m):
    """Inverse of hpr_to_mat, degrees."""
    h = math.atan2(-m[0, 1], m[1, 1])
    p = math.asin(max(-1.0, min(1.0, m[2, 1])))
    r = math.atan2(-m[2, 0], m[2, 2])
    return math.degrees(h), math.degrees(p), math.degrees(r)


class BitMask32:
    __slots__ = ("w", )

    def __init__(self, w=0):
        self.w = int(w) & 0xFFFFFFFF

    @classmethod
    def bit(cls, n):
        return cls(1 << n)

    @classmethod
    def allOn(cls):
        return cls(0xFFFFFFFF)

    all_on = allOn

    @classmethod
    def allOff(cls):
        return cls(0)

    all_off = allOff

    def getWord(self):
        return self.w

    get_word = getWord

    def __or__(self, o):
        return BitMask32(self.w | _word(o))

    __ror__ = __or__

    def __and__(self, o):
        return BitMask32(self.w & _word(o))

    def __invert__(self):
        return BitMask32(~self.w)

    def __eq__(self, o):
        return self.w == _word(o)

    def __hash__(self):
        return hash(self.w)

    def __bool__(self):
        return self.w != 0

    def isZero(self):
        return self.w == 0

    def __repr__(self):
        return "BitMask32(%#x)" % self.w


def _word(o):
    if isinstance(o, BitMask32):
        return o.w
    if isinstance(o, (int, np.integer)):
        return int(o)
    return 0xFFFFFFFF  # inert dummy (camera masks) behaves as all-on


class TransformState:
    def __init__(self, pos, mat):
        self.pos = np.asarray(pos, dtype=np.float64)
        self.mat = mat

    @classmethod
    def makePos(cls, pos):
        return cls(np.array(list(pos), dtype=np.float64), np.eye(3))

    make_pos = makePos

    @classmethod
    def makePosHpr(cls, pos, hpr):
        return cls(np.array(list(pos), dtype=np.float64), hpr_to_mat(*list(hpr)))

    make_pos_hpr = makePosHpr

    @classmethod
    def makeIdentity(cls):
        return cls(np.zeros(3), np.eye(3))

    def getPos(self):
        return Vec3(*self.pos)


class PandaNode:
    def __init__(self, name=""):
        self._name = name
        self._tags = {}
        self._pytags = {}
        self.pos = np.zeros(3)
        self.mat = np.eye(3)
        self.parent = None
        self.children = []

    def getName(self):
        return self._name

    get_name = getName

    def setName(self, n):
        self._name = n

    def setPythonTag(self, k, v):
        self._pytags[k] = v

    set_python_tag = setPythonTag

    def getPythonTag(self, k):
        return self._pytags.get(k)

    get_python_tag = getPythonTag

    def hasPythonTag(self, k):
        return k in self._pytags

    has_python_tag = hasPythonTag

    def clearPythonTag(self, k):
        self._pytags.pop(k, None)

    clear_python_tag = clearPythonTag

    def setTag(self, k, v):
        self._tags[k] = v

    def getTag(self, k):
        return self._tags.get(k, "")

    def removeAllChildren(self):
        for c in self.children:
            c.parent = None
        self.children = []

    remove_all_children = removeAllChildren

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return lambda *a, **kw: None  # rendering-only calls are inert


class _Children(list):
    def reparentTo(self, np_):
        for c in list(self):
            c.reparentTo(np_)

    reparent_to = reparentTo


class NodePath:
    def __init__(self, x=""):
        if isinstance(x, NodePath):
            self._node = x._node
        elif isinstance(x, PandaNode):
            self._node = x
        else:
            self._node = PandaNode(str(x))

    # identity
    def node(self):
        return self._node

    def __eq__(self, o):
        return isinstance(o, NodePath) and o._node is self._node

    def __hash__(self):
        return id(self._node)

    def getName(self):
        return self._node.getName() if self._node is not None else ""

    def isEmpty(self):
        return self._node is None

    is_empty = isEmpty

    # graph
    def attachNewNode(self, x):
        child = NodePath(x)
        child.reparentTo(self)
        return child

    attach_new_node = attachNewNode

    def reparentTo(self, other):
        n = self._node
        if n.parent is not None and n in n.parent.children:
            n.parent.children.remove(n)
        n.parent = other._node
        other._node.children.append(n)

    reparent_to = reparentTo

    def detachNode(self):
        n = self._node
        if n is None:
            return
        if n.parent is not None and n in n.parent.children:
            n.parent.children.remove(n)
        n.parent = None

    detach_node = detachNode

    def removeNode(self):
        self.detachNode()

    remove_node = removeNode

    def hasParent(self):
        return self._node is not None and self._node.parent is not None

    has_parent = hasParent

    def getParent(self):
        return NodePath(self._node.parent) if self._node.parent is not None else NodePath("empty")

    def getChildren(self):
        return _Children(NodePath(c) for c in self._node.children)

    get_children = getChildren

    def instanceTo(self, other):
        return self

    def copyTo(self, other):
        return NodePath(self.getName())

    def find(self, *a):
        return NodePath("found")

    def findAllMatches(self, *a):
        return _Children()

    # transform
    def getPos(self, *a):
        return Vec3(*self._node.pos)

    get_pos = getPos

    def setPos(self, *a):
        self._node._geom_cache = None
        if len(a) == 1:
            a = list(a[0])
        if len(a) == 4:  # (other, x, y, z)
            a = a[1:]
        self._node.pos = np.array([float(a[0]), float(a[1]), float(a[2])])

    set_pos = setPos

    def getX(self):
        return float(self._node.pos[0])

    def getY(self):
        return float(self._node.pos[1])

    def getZ(self):
        return float(self._node.pos[2])

    get_x, get_y, get_z = getX, getY, getZ

    def setX(self, v):
        self._node._geom_cache = None
        self._node.pos[0] = float(v)

    def setY(self, v):
        self._node._geom_cache = None
        self._node.pos[1] = float(v)

    def setZ(self, v):
        self._node._geom_cache = None
        self._node.pos[2] = float(v)

    set_x, set_y, set_z = setX, setY, setZ

    def getHpr(self):
        return Vec3(*mat_to_hpr(self._node.mat))

    get_hpr = getHpr

    def setHpr(self, *a):
        self._node._geom_cache = None
        if len(a) == 1:
            a = list(a[0])
        self._node.mat = hpr_to_mat(float(a[0]), float(a[1]), float(a[2]))

    set_hpr = setHpr

    def getH(self):
        return mat_to_hpr(self._node.mat)[0]

    def getP(self):
        return mat_to_hpr(self._node.mat)[1]

    def getR(self):
        return mat_to_hpr(self._node.mat)[2]

    get_h, get_p, get_r = getH, getP, getR

    def setH(self, h):
        self._node._geom_cache = None
        _, p, r = mat_to_hpr(self._node.mat)
        self._node.mat = hpr_to_mat(float(h), p, r)

    def setP(self, p):
        self._node._geom_cache = None
        h, _, r = mat_to_hpr(self._node.mat)
        self._node.mat = hpr_to_mat(h, float(p), r)

    def setR(self, r):
        self._node._geom_cache = None
        h, p, _ = mat_to_hpr(self._node.mat)
        self._node.mat = hpr_to_mat(h, p, float(r))

    set_h, set_p, set_r = setH, setP, setR

    def setQuat(self, q):
        self._node._geom_cache = None
        self._node.mat = quat_to_mat(np.array(list(q), dtype=np.float64))

    set_quat = setQuat

    def getQuat(self):
        return LQuaternionf(*mat_to_quat(self._node.mat))

    get_quat = getQuat

    def getRelativeVector(self, other, vec):
        """Rotate `vec` expressed in `other`'s frame into this node's frame."""
        v = np.array(list(vec), dtype=np.float64)
        world = other._node.mat @ v
        return Vec3(*(self._node.mat.T @ world))

    get_relative_vector = getRelativeVector

    def getRelativePoint(self, other, pt):
        p = np.array(list(pt), dtype=np.float64)
        world = other._node.mat @ p + other._node.pos
        return Vec3(*(self._node.mat.T @ (world - self._node.pos)))

    get_relative_point = getRelativePoint

    # tags
    def setTag(self, k, v):
        self._node.setTag(k, v)

    set_tag = setTag

    def getTag(self, k):
        return self._node.getTag(k)

    def setPythonTag(self, k, v):
        self._node.setPythonTag(k, v)

    def getPythonTag(self, k):
        return self._node.getPythonTag(k)

    def hasPythonTag(self, k):
        return self._node.hasPythonTag(k)

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return lambda *a, **kw: None  # hide/show/setScale/setMaterial/... are inert


class _Task:
    cont = 1
    done = 0


class _TaskChainMgr:
    def getNumTaskChains(self):
        return 0


class _TaskMgr:
    def __init__(self):
        self.globalClock = _mk("clock")()
        self.mgr = _TaskChainMgr()

    def add(self, *a, **k):
        return None

    def remove(self, *a, **k):
        return None

    def step(self):
        return None

    def getAllTasks(self):
        return []

    def hasTaskNamed(self, n):
        return False

    def stop(self):
        return None

    def destroy(self):
        return None


class ShowBase:
    """`direct.showbase.ShowBase.ShowBase` stand-in: a root NodePath and an inert task manager."""
    def __init__(self, *a, **k):
        self.render = NodePath("render")
        self.render2d = NodePath("render2d")
        self.aspect2d = NodePath("aspect2d")
        self.taskMgr = _TaskMgr()
        self.task_manager = self.taskMgr
        self.loader = None
        self.win = None
        self.cam = NodePath("cam")
        self.camera = NodePath("camera")
        self.graphicsEngine = _mk("graphicsEngine")()

    def accept(self, *a, **k):
        return None

    def ignoreAll(self):
        return None

    def destroy(self):
        return None

    def disableMouse(self):
        return None


def loadPrcFileData(*a, **k):
    return None


class PythonCallbackObject:
    def __init__(self, fn):
        self.fn = fn

    def __call__(self, *a):
        return self.fn(*a)


class GraphicsPipeSelection:
    @classmethod
    def getGlobalPtr(cls):
        return cls()

    def getPipeTypes(self):
        return ["none"]

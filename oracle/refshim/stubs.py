"""Import stubs for packages the reference imports but the step path never needs (rendering, IO).

TEST INFRASTRUCTURE (see package docstring).  Attribute access on a stub module fabricates an inert
dummy class: subclassable, callable, iterable, arithmetic-tolerant.  `panda3d.core` / `panda3d.bullet`
are NOT inert: they resolve to `pcore` / `pbullet`, falling back to dummies for rendering-only names.
"""
import importlib.abc
import importlib.machinery
import sys
import types


class _Meta(type):
    def __getattr__(cls, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _mk(k)

    def __or__(a, b):
        return a

    def __ror__(a, b):
        return a


def _inst_getattr(self, k):
    if k.startswith("__"):
        raise AttributeError(k)
    return _mk(k)()


def _mk(name):
    return _Meta(
        name, (object, ), {
            "__init__": lambda self, *a, **k: None,
            "__getattr__": _inst_getattr,
            "__call__": lambda self, *a, **k: _mk("ret")(),
            "__or__": lambda a, b: a,
            "__ror__": lambda a, b: a,
            "__and__": lambda a, b: a,
            "__iter__": lambda self: iter(()),
            "__getitem__": lambda self, i: 0.0,
            "__setitem__": lambda self, i, v: None,
            "__len__": lambda self: 0,
            "__float__": lambda self: 0.0,
            "__int__": lambda self: 0,
            "__index__": lambda self: 0,
            "__bool__": lambda self: False,
            "__mul__": lambda a, b: a,
            "__rmul__": lambda a, b: a,
            "__truediv__": lambda a, b: a,
            "__add__": lambda a, b: a,
            "__radd__": lambda a, b: a,
            "__sub__": lambda a, b: a,
            "__rsub__": lambda a, b: a,
            "__neg__": lambda a: a,
            "__enter__": lambda self: self,
            "__exit__": lambda self, *a: False,
        }
    )


class StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        v = _mk(k)
        setattr(self, k, v)
        return v


PREFIXES = (
    "panda3d", "direct", "shapely", "pygame", "gymnasium", "gym", "seaborn", "PIL", "cv2", "pygments", "gltf", "tqdm",
    "progressbar", "filelock", "requests", "lxml", "geopandas", "matplotlib", "yapf", "zmq", "ray", "cupy",
    "OpenGL", "mediapy", "imageio"
)


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in PREFIXES:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return StubModule(spec.name)

    def exec_module(self, module):
        name = module.__name__
        if name == "panda3d.core":
            from . import pcore
            _graft(module, pcore)
        elif name == "panda3d.bullet":
            from . import pbullet
            _graft(module, pbullet)
        elif name == "gymnasium" or name == "gym":
            from . import fakegym
            module.spaces = fakegym.spaces
            module.Env = fakegym.Env
            module.Wrapper = fakegym.Env
        elif name in ("gymnasium.spaces", "gym.spaces"):
            from . import fakegym
            _graft(module, fakegym.spaces)
        elif name == "seaborn":
            module.color_palette = lambda *a, **k: [(0.1 * i, 0.1 * i, 0.1 * i) for i in range(10)]
        elif name == "direct.showbase":
            from . import pcore
            sb = StubModule("direct.showbase.ShowBase")
            sb.ShowBase = pcore.ShowBase
            module.ShowBase = sb
            sys.modules["direct.showbase.ShowBase"] = sb
        elif name == "direct.showbase.ShowBase":
            from . import pcore
            module.ShowBase = pcore.ShowBase


def _graft(module, src):
    for k in dir(src):
        if not k.startswith("__"):
            setattr(module, k, getattr(src, k))


def install():
    for f in sys.meta_path:
        if isinstance(f, _Finder):
            return
    sys.meta_path.insert(0, _Finder())

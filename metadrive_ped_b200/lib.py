"""Loader of libmdstep.so (the C ABI in include/mdstep.h). There is no CPU fallback: if the library is missing
or no CUDA device is usable, every call fails loudly."""
import ctypes as C
import os
import subprocess

from .abi import MdArrays, MdConfig

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MD_LIB") or os.path.join(_HERE, "libmdstep.so")  # MD_LIB: a tuning build of the same sources
_lib = None

EXPORTS = [
    "md_abi_version", "md_create", "md_destroy", "md_last_error", "md_load_scene", "md_reset", "md_step",
    "md_autoreset", "md_step_autoreset", "md_step_host", "md_reset_host", "md_lidar", "md_dynamics", "md_after_step", "md_idm",
    "md_get_state", "md_set_state", "md_snapshot", "md_launch_count", "md_profile_begin", "md_profile_end",
    "md_host_views", "md_attach_bank", "md_sizeof_config", "md_sizeof_arrays", "md_host_groups", "md_host_group_count",
    "md_host_group_views", "md_host_send", "md_host_recv", "md_host_compact", "md_fp32_peak", "md_enable_contacts", "md_get_contacts",
    "md_topdown", "md_topdown_channels",
]
ABI_VERSION = 8  # include/mdstep.h MD_ABI_VERSION


class MdStepError(RuntimeError):
    pass


def build(force=False):
    """Compile libmdstep.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    src_dir = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(src_dir, f) for f in ("md_kernels.cu", "md_device.cuh")]
    srcs += [os.path.join(_HERE, "..", "include", f) for f in ("mdstep.h", "md_layout.h", "md_math.h")]
    newest = max(os.path.getmtime(s) for s in srcs)
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < newest:
        subprocess.check_call(["make", "-C", src_dir, "-s"] + (["-B"] if force else []))
    return LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MdStepError(
            "libmdstep.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'`. "
            "This package has no CPU fallback." % LIB_PATH
        )
    lib = C.CDLL(LIB_PATH)
    vp, ip = C.c_void_p, C.c_int
    lib.md_abi_version.restype = ip
    # a stale or tuning build with another struct layout would silently misread the configuration passed by value
    if lib.md_abi_version() != ABI_VERSION:
        raise MdStepError("%s has ABI version %d, this package expects %d: rebuild it" % (LIB_PATH, lib.md_abi_version(), ABI_VERSION))
    if lib.md_sizeof_config() != C.sizeof(MdConfig) or lib.md_sizeof_arrays() != C.sizeof(MdArrays):
        raise MdStepError("%s was built with another MdConfig / MdArrays layout (%d / %d bytes, expected %d / %d): rebuild it" % (
            LIB_PATH, lib.md_sizeof_config(), lib.md_sizeof_arrays(), C.sizeof(MdConfig), C.sizeof(MdArrays)))
    lib.md_create.argtypes = [C.POINTER(MdConfig), ip, C.POINTER(vp)]
    lib.md_destroy.argtypes = [vp]
    lib.md_destroy.restype = None
    lib.md_last_error.argtypes = [vp]
    lib.md_last_error.restype = C.c_char_p
    lib.md_load_scene.argtypes = [vp, C.POINTER(MdArrays), C.POINTER(C.c_int64)]
    lib.md_reset.argtypes = [vp, vp, vp, vp]
    lib.md_attach_bank.argtypes = [vp, vp, ip]
    lib.md_step.argtypes = [vp] + [vp] * 8 + [vp]
    lib.md_autoreset.argtypes = [vp, vp, vp, vp, vp]
    lib.md_step_autoreset.argtypes = [vp] + [vp] * 8 + [vp]
    lib.md_step_host.argtypes = [vp] + [vp] * 8 + [ip]
    lib.md_reset_host.argtypes = [vp, vp, vp]
    lib.md_lidar.argtypes = [vp, vp, vp, vp]
    lib.md_dynamics.argtypes = [vp, vp, ip, vp]
    lib.md_topdown.argtypes = [vp, vp, ip, C.c_float, vp]
    lib.md_topdown_channels.argtypes = [vp, vp, ip, C.c_float, vp]
    lib.md_after_step.argtypes = [vp, vp]
    lib.md_idm.argtypes = [vp, vp, vp]
    lib.md_get_state.argtypes = [vp, C.c_char_p, vp, C.c_size_t]
    lib.md_set_state.argtypes = [vp, C.c_char_p, vp, C.c_size_t]
    lib.md_snapshot.argtypes = [vp]
    lib.md_profile_begin.argtypes = [vp, ip]
    lib.md_profile_end.argtypes = [vp, vp, ip]
    lib.md_host_views.argtypes = [vp, C.POINTER(vp)]
    lib.md_host_groups.argtypes = [vp, ip]
    lib.md_host_group_count.argtypes = [vp]
    lib.md_host_group_views.argtypes = [vp, ip, C.POINTER(vp), C.POINTER(ip)]
    lib.md_host_send.argtypes = [vp, ip, vp, ip]
    lib.md_host_recv.argtypes = [vp, ip]
    lib.md_host_compact.argtypes = [vp, ip]
    lib.md_fp32_peak.argtypes = [ip, C.POINTER(C.c_double)]
    lib.md_enable_contacts.argtypes = [vp, ip]
    lib.md_get_contacts.argtypes = [vp, vp, C.c_size_t]
    lib.md_launch_count.argtypes = [vp]
    lib.md_launch_count.restype = C.c_int64
    for name in EXPORTS:
        getattr(lib, name)
    _lib = lib
    return lib

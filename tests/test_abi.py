"""CPU: the C-ABI library loads and exports every symbol include/mdstep.h declares (no compute without a GPU);
the ctypes mirror of MdConfig has the header's size; the product refuses to run without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mdstep.h")).read()
    return sorted(set(re.findall(r"\b(md_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from metadrive_ped_b200 import lib as mdlib
    path = mdlib.build()
    h = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(h, n), n
    assert set(names) == set(mdlib.EXPORTS)
    h.md_abi_version.restype = ctypes.c_int
    assert h.md_abi_version() == 8


def test_config_struct_matches_header():
    from metadrive_ped_b200.abi import MdConfig, MdArrays
    src = open(os.path.join(ROOT, "include", "md_layout.h")).read()
    body = src[src.index("typedef struct MdConfig {"):src.index("} MdConfig;")]
    n_fields = len(re.findall(r"\b(?:int|float)\b([^;]*);", body))
    fields = [f.strip() for line in re.findall(r"\b(?:int|float)\b([^;]*);", body) for f in line.split(",")]
    assert [f for f in fields] == [n for n, _ in MdConfig._fields_], (fields, n_fields)
    assert ctypes.sizeof(MdConfig) == 4 * len(fields)
    arr = src[src.index("typedef struct MdArrays {"):src.index("} MdArrays;")]
    names = re.findall(r"\*\s*([a-z_]+);", arr)
    assert names == [n for n, _ in MdArrays._fields_]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from metadrive_ped_b200.abi import make_config
    from metadrive_ped_b200.lib import MdStepError
    from metadrive_ped_b200.sim import BatchedSim
    with pytest.raises(MdStepError):
        BatchedSim({}, make_config(1, 4))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "metadrive_ped_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                s = open(os.path.join(dp, f)).read()
                assert "import oracle" not in s and "from oracle" not in s and "md_oracle" not in s, f


def test_integration_stub_matches_the_struct():
    """INTEGRATION.md shows the reference-side ctypes stub; its MdConfig must list the header's fields in order."""
    import re
    from metadrive_ped_b200.abi import MdConfig
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blk = text[text.index("class MdConfig(C.Structure):"):text.index("# include/md_layout.h: struct MdArrays")]
    assert re.findall(r'"([a-z_0-9]+)"', blk) == [f[0] for f in MdConfig._fields_]

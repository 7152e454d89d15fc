"""The C-ABI library loads and exports every symbol include/cassie2d.h declares; struct layouts
match RobotInterface.h:14-50 / cassie2d_structs.py:5-51.  No compute calls (no GPU needed)."""
import ctypes as ct
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def built_lib():
    from cassierl_b200 import build
    return build.build()


def test_header_symbols_exported(built_lib):
    hdr = open(os.path.join(ROOT, "include", "cassie2d.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", hdr)
    names = sorted(set(n for n in names if n not in ("defined",)))
    assert len(names) >= 30, names
    L = ct.CDLL(built_lib)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    from cassierl_b200 import lib
    assert sorted(lib.LEGACY_SYMBOLS + lib.BATCH_SYMBOLS) == names


def test_struct_layouts():
    from cassierl_b200 import structs as S
    assert ct.sizeof(S.ControllerTorque) == 48
    assert ct.sizeof(S.ControllerForce) == 48
    assert ct.sizeof(S.ControllerOsc) == 56
    assert ct.sizeof(S.ControllerPd) == 48
    assert ct.sizeof(S.StateGeneral) == 208
    assert ct.sizeof(S.StateOperationalSpace) == 144
    assert S.StateGeneral.left_pos.offset == 48 and S.StateGeneral.right_vel.offset == 168
    assert S.ControllerOsc.pitch_add.offset == 48


def test_no_cpu_fallback(built_lib):
    """Without a CUDA device the batch constructor must fail loudly (never compute on the CPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cassierl_b200 import lib
    L = lib.load()
    h = L.Cassie2dBatchInit(4, 0, None, 32)
    assert not h
    assert b"no CUDA device" in L.CassieGetLastError()
    from cassierl_b200.envs import Cassie2dBatch
    with pytest.raises(RuntimeError):
        Cassie2dBatch(4)


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under cassierl_b200/ may import, include or link it."""
    pkg = os.path.join(ROOT, "cassierl_b200")
    bad = re.compile(r"import\s+oracle|from\s+oracle|from\s+\.\.?oracle|cassie_oracle|oracle/")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert not bad.search(txt), "product file %s references the oracle" % f

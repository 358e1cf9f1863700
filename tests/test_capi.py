"""CPU: the C-ABI library loads and exports every symbol include/sidm_b200.h declares; without
a GPU the product path fails loudly instead of falling back."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "sidm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported():
    from sidm_b200 import capi
    lib = capi.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sidm_b200.h but not exported"
    assert sorted(capi.EXPORTS) == names
    assert b"sm_100a" in lib.b200_version()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sidm_b200 import B200Error, HotPath
    with pytest.raises(B200Error) as e:
        HotPath(1000)
    assert e.value.code == 9002


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "sidm-nbody_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.lower() or f in (), f"{f} mentions oracle/"


def test_particle_layout_matches_reference():
    """124-byte struct particle_data (allvars.h:422-460 with -DSIDM)"""
    from sidm_b200 import capi
    assert capi.PARTICLE_DTYPE.itemsize == 124
    lay = capi.layout_of()
    assert (lay.Pos, lay.Vel, lay.Mass, lay.PosPred, lay.Accel, lay.OldAcc, lay.HsmlVelDisp, lay.dVel) == (0, 12, 24, 44, 68, 88, 108, 112)
    import refdrv
    if refdrv.available("diag"):
        assert refdrv.Reference("diag").psize == 124

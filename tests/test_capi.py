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


def test_header_is_plain_c_and_structs_match(tmp_path):
    """include/sidm_b200.h compiles as C99 (the reference and its shim are C), and the structs the Python mirror passes
    have the sizes the header declares: b200_sysstate = struct state_of_system (allvars.h:517-537, 102 doubles)"""
    import ctypes as C
    import subprocess
    from sidm_b200 import capi
    src = tmp_path / "hdr.c"
    src.write_text('#include <stdio.h>\n#include "sidm_b200.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu\\n", sizeof(b200_sysstate), sizeof(b200_params), sizeof(b200_layout), '
                   'sizeof(b200_counters)); return 0; }\n')
    exe = tmp_path / "hdr"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    assert sizes[0] == 102 * 8 == C.sizeof(capi.SysState)
    assert sizes[1] == C.sizeof(capi.Params) and sizes[2] == C.sizeof(capi.Layout) and sizes[3] == C.sizeof(capi.Counters)


def test_dropin_builds_resolve_their_symbols():
    """the reference driver linked against the shim (oracle/_ref/libsidmref_b200*.so, sidm_b200_mpi): every symbol the shim
    needs (b200_comm.c included) is there at load time - no compute call, runs without a GPU"""
    import ctypes
    import subprocess
    import pytest
    ref = os.path.join(ROOT, "oracle", "_ref")
    libs = [os.path.join(ref, f) for f in ("libsidmref_b200.so", "libsidmref_b200f.so")]
    if not all(os.path.exists(p) for p in libs):
        pytest.skip("oracle/_ref not built")
    for p in libs:
        ctypes.CDLL(p, mode=os.RTLD_NOW)
    exe = os.path.join(ref, "sidm_b200_mpi")
    if os.path.exists(exe):
        r = subprocess.run(["ldd", "-r", exe], capture_output=True, text=True)
        assert "undefined symbol" not in r.stdout + r.stderr, r.stdout + r.stderr

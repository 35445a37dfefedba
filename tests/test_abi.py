"""The C-ABI library loads (no GPU needed) and exports every symbol include/fs2d.h declares."""
from __future__ import annotations

import ctypes
import re

import pytest
from conftest import REPO

from fs import _lib


def header_symbols():
    text = (REPO / "include" / "fs2d.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fs2d_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_header_symbols():
    assert _lib.LIB_PATH.exists(), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/fs2d.h but not exported by libfs2d.so"


def test_binding_covers_header():
    assert sorted(_lib.EXPORTED_SYMBOLS) == header_symbols()
    lib = _lib.load()
    assert lib.fs2d_version() == 1


def test_no_cpu_fallback():
    """CPU tensors are refused loudly; nothing routes through the oracle."""
    import torch

    from fs.double_buffer import Field
    f = Field((8, 8), 1, device="cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        f.ptr()
    for mod in ("fs/solver.py", "fs/pressure_updater.py", "fs/boundary_condition.py", "fs/vorticity_confinement.py",
                "fs/fluid_simulator.py", "fs/_lib.py", "fs/_bc_tables.py", "fs/double_buffer.py"):
        src = (REPO / "2d-fluid-simulator_b200" / mod).read_text()
        assert "oracle" not in src.replace("the oracle", "").replace("oracle `", "").replace("oracle (", ""), mod
    assert not torch.cuda.is_available() or True


# ---- argument validation happens before any CUDA call, so it is testable without a GPU -------------------------
def _dom(rows=8, Y=8, r0=0, r1=8, clo=0, chi=7, gi0=0):
    return _lib.Dom(rows, Y, r0, r1, clo, chi, gi0)


_FAKE = 0x1000   # never dereferenced: every call below must return before launching anything


def test_null_pointers_are_rejected_with_valueerror():
    d = _dom()
    with pytest.raises(ValueError, match="null field pointer"):
        _lib.call("fs2d_limit", None, d, 10.0, None)
    with pytest.raises(ValueError, match="null field pointer"):
        _lib.call("fs2d_cip_nonadv", _FAKE, None, _FAKE, _FAKE, d, 0.1, 0.1, 100.0, None)
    with pytest.raises(ValueError, match="null field pointer"):
        _lib.call("fs2d_pressure_source", _FAKE, None, d, 0.1, 0.1, None)
    with pytest.raises(ValueError, match="null/aliased"):
        _lib.call("fs2d_jacobi_fused", _FAKE, _FAKE, _FAKE, _FAKE, d, 4, None, 0, None)
    with pytest.raises(ValueError, match="in place"):
        _lib.call("fs2d_jacobi_sweep", _FAKE, _FAKE, _FAKE, _FAKE, d, 0, None)
    with pytest.raises(ValueError, match="in place"):
        _lib.call("fs2d_vort_apply", _FAKE, _FAKE, _FAKE, _FAKE, _FAKE, d, 0.1, 0.1, None)
    assert b"in place" in _lib.load().fs2d_last_error()


def test_bad_domains_schemes_and_sizes_are_rejected():
    for bad in (_dom(rows=0), _dom(Y=0), _dom(r0=5, r1=4), _dom(r1=9), _dom(clo=3, chi=2), _dom(chi=8), _dom(r0=-1)):
        with pytest.raises(ValueError, match="bad argument"):
            _lib.call("fs2d_limit", _FAKE, bad, 10.0, None)
    with pytest.raises(ValueError, match="unknown advection scheme"):
        _lib.call("fs2d_mac_update", _FAKE, _FAKE + 64, _FAKE, _FAKE, _dom(), 0.1, 0.1, 100.0, 7, None)
    with pytest.raises(ValueError, match="parity"):
        _lib.call("fs2d_rbsor_pass", _FAKE, _FAKE, _FAKE, _FAKE, _dom(), 1.3, -0.3, 2, None)
    with pytest.raises(ValueError, match="out of range"):
        _lib.call("fs2d_jacobi_fused", _FAKE, _FAKE + 64, _FAKE, _FAKE, _dom(Y=16, chi=7), 13, None, 0, None)
    with pytest.raises(ValueError, match="Y % 16"):
        _lib.call("fs2d_jacobi_fused", _FAKE, _FAKE + 64, _FAKE, _FAKE, _dom(Y=8), 4, None, 0, None)
    with pytest.raises(ValueError, match="unknown tuning key"):
        _lib.call("fs2d_set_tuning", 99, 0)
    with pytest.raises(ValueError, match="unknown tuning key"):
        _lib.call("fs2d_set_tuning", 1, 4)       # no such key (the fused-kernel variants of round 1 are gone)


def test_empty_inputs_are_noops():
    """Empty row windows and empty BC tables return FS2D_OK without launching (multi-rank edge windows can be empty)."""
    lib = _lib.load()
    n0 = lib.fs2d_launch_count()
    d = _dom(r0=3, r1=3)
    _lib.call("fs2d_limit", _FAKE, d, 10.0, None)
    _lib.call("fs2d_cip_nonadv", _FAKE, _FAKE, _FAKE, _FAKE, d, 0.1, 0.1, 100.0, None)
    _lib.call("fs2d_cip_nonadv_grad", _FAKE, _FAKE, _FAKE, _FAKE, _FAKE, _FAKE, _FAKE, d, 0.2, None)
    _lib.call("fs2d_cip_advect", _FAKE, _FAKE, _FAKE, _FAKE, _FAKE, _FAKE, _FAKE, _FAKE, d, 0.1, 0.1, 0.01, 0.001, None)
    _lib.call("fs2d_vort_apply", _FAKE, _FAKE, _FAKE, _FAKE + 64, _FAKE, d, 0.1, 0.1, None)
    _lib.call("fs2d_pressure_source", _FAKE, _FAKE, d, 0.1, 0.1, None)
    _lib.call("fs2d_jacobi_sweep", _FAKE, _FAKE + 64, _FAKE, _FAKE, d, 0, None)
    _lib.call("fs2d_jacobi_fused", _FAKE, _FAKE + 64, _FAKE, _FAKE, _dom(Y=16, r0=3, r1=3), 4, None, 0, None)
    _lib.call("fs2d_vel_bc", _FAKE, _FAKE, None, None, None, None, 0, None)
    _lib.call("fs2d_pressure_bc", _FAKE, None, None, None, None, None, 0, None)
    _lib.call("fs2d_dye_bc", _FAKE, _FAKE, None, 0, None)
    assert lib.fs2d_launch_count() == n0


def test_fused_tile_geometry_is_consistent():
    rows, cols, hr, hc, tmax = (ctypes.c_int() for _ in range(5))
    for t in range(1, 13):
        _lib.call("fs2d_fused_tile", t, ctypes.byref(rows), ctypes.byref(cols), ctypes.byref(hr), ctypes.byref(hc), ctypes.byref(tmax))
        assert hr.value == t and hc.value % 4 == 0 and t <= hc.value < t + 4       # TMA: 16-byte aligned box starts
        assert rows.value - 2 * hr.value > 0 and cols.value - 2 * hc.value > 0 and tmax.value == 12

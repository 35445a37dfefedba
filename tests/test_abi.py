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

"""pytest configuration: import paths, the `gpu` marker and shared helpers."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]
PKG = REPO / "2d-fluid-simulator_b200"
GOLDEN = REPO / "tests" / "golden"
for p in (str(REPO), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def assert_bitexact(name: str, got: np.ndarray, want: np.ndarray) -> None:
    """fp32 arrays must match bit for bit (NaNs compare equal)."""
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, f"{name}: shape {got.shape} vs {want.shape}"
    if np.array_equal(got, want, equal_nan=True):
        return
    bad = ~((got == want) | (np.isnan(got) & np.isnan(want)))
    idx = np.argwhere(bad)
    d = np.nanmax(np.abs(got.astype(np.float64) - want.astype(np.float64))[bad])
    raise AssertionError(f"{name}: {bad.sum()} of {got.size} values differ (max |d|={d:.3e}); first at {idx[0].tolist()}: "
                         f"got {got[tuple(idx[0])]!r} want {want[tuple(idx[0])]!r}")


@pytest.fixture(scope="session")
def masks_small():
    return np.load(GOLDEN / "masks_small.npz")


@pytest.fixture(scope="session")
def kernels_golden():
    return np.load(GOLDEN / "kernels_r24.npz")

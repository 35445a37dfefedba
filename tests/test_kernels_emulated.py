"""The CUDA kernel SOURCES executed on the CPU (tests/cuda_emu: a fiber-based emulation of the CUDA slice they use, with
mbarrier / TMA box-load semantics) and put through the GPU parity tests' own bodies -- bit-exact against the reference
fixtures and the oracle.

What this covers that nothing else can without a GPU: the kernels' indexing and tiling logic -- clamp windows and row
windows, interior fast paths, the TMA tile pipeline with clamp repair in shared memory, the fused Jacobi kernels' tile
lists / slow-cell lists / edge-row exchange / shrinking valid region, barrier and progress-counter placement (a
missing __syncthreads deadlocks or corrupts here too), TMA box geometry and alignment rules.
What it does not cover: anything about the hardware -- memory ordering, async-proxy fences, occupancy, performance.  The
`-m gpu` tests on a B200 remain the parity gate; this module exists so that kernel logic is exercised on every CPU run.
The emulated library is tests/cuda_emu/_build/libfs2d_emu.so; the product never loads it (fs/_lib.py opens lib/libfs2d.so only).
"""
from __future__ import annotations

import ctypes
import os
import sys

import pytest
import torch
from conftest import REPO

sys.path.insert(0, str(REPO / "tests" / "cuda_emu"))

import test_gpu_parity as G  # noqa: E402  (its module-level `gpu` mark does not apply to the functions re-exported here)


@pytest.fixture(scope="module")
def env():
    """Swap libfs2d.so for the emulated build and CUDA tensors for CPU tensors, for this module only."""
    import build_emu
    from fs import _lib, boundary_condition, double_buffer

    lib = ctypes.CDLL(os.environ.get("FS2D_EMU_LIB") or str(build_emu.build()))   # override: mutation testing of this harness
    for name, (res, args) in _lib._SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    saved = (_lib._lib, _lib.ptr, _lib.stream, double_buffer.default_device, boundary_condition.default_device)
    _lib._lib = lib
    _lib.ptr = lambda t: None if t is None else t.data_ptr()
    _lib.stream = lambda: 0
    double_buffer.default_device = boundary_condition.default_device = lambda: torch.device("cpu")
    try:
        assert _lib.load() is lib and lib.fs2d_device_ok()
        yield lib
    finally:
        _lib._lib, _lib.ptr, _lib.stream, double_buffer.default_device, boundary_condition.default_device = saved


# ---- the GPU parity tests, unchanged -------------------------------------------------------------------------------------
test_emu_each_kernel_matches_reference_fixture = G.test_each_kernel_matches_reference_fixture
test_emu_trajectory_matches_reference_fixture = G.test_trajectory_matches_reference_fixture
test_emu_dye_trajectory_matches_reference_fixture = G.test_dye_trajectory_matches_reference_fixture
test_emu_dye_step_from_random_state_vs_oracle = G.test_dye_step_from_random_state_vs_oracle
test_emu_vectorised_nonadv_equals_one_cell_kernel = G.test_vectorised_nonadv_equals_one_cell_kernel
test_emu_render_matches_reference_fixture = G.test_render_matches_reference_fixture
test_emu_state_dict_roundtrip_resumes_bitwise = G.test_state_dict_roundtrip_resumes_bitwise
test_emu_random_mask_trajectory_vs_oracle = G.test_random_mask_trajectory_vs_oracle
test_emu_division_special_cases_match_the_oracle = G.test_division_special_cases_match_the_oracle
test_emu_facade_matches_reference_defaults = G.test_facade_matches_reference_defaults


# ---- the same bodies on the parameter sets an emulator finishes in seconds -----------------------------------------------
_SMALL_CONFIGS = [c for c in G.CONFIGS if c[0] in ("cfg1_as_given", "kk_r100", "cip_r50_y_not_mult4")]


@pytest.mark.parametrize("cfg", _SMALL_CONFIGS, ids=[c[0] for c in _SMALL_CONFIGS])
def test_emu_config_trajectory_vs_oracle(env, cfg):
    G.test_config_trajectory_vs_oracle(env, cfg)


def test_emu_jacobi_row_range_invariance_and_literal_equivalence(env):
    G.test_jacobi_row_range_invariance_and_literal_equivalence(env, res=128)


_FUSED_EMU_CASES = [(n, x, y, listed) for listed in (True, False)
                    for n, x, y in [(1, 128, 64), (2, 256, 128), (4, 200, 96), (5, 384, 192), (1, 288, 352), (1, 300, 480)]]


@pytest.mark.parametrize("num,X,Y,listed", _FUSED_EMU_CASES)
def test_emu_fused_pass_equals_literal_iterations(env, num, X, Y, listed):
    # (1, 288, 352) and (1, 300, 480) are wide and tall enough to contain OPEN-FLUID tiles (no wall, BC cell or grid edge in
    # the tile): the autonomous-warp path with per-warp TMA refills and the edge-row exchange under progress counters;
    # every tile of the smaller grids is "slow" (CTA barriers)
    big = X * Y > 40000
    G.test_fused_pass_equals_literal_iterations(env, num, X, Y, listed, t_list=(1, 3, 8, 12) if big else (1, 2, 4, 6, 8, 11, 12),
                                                need=2 if big else 3)


@pytest.mark.parametrize("num,X,Y,listed", [(4, 480, 640, True), (4, 480, 640, False), (5, 768, 384, True), (5, 768, 384, False)])
def test_emu_fused_pass_open_and_skipped_tiles(env, num, X, Y, listed):
    """bc4 at 480 x 640: mostly open-fluid tiles -- consecutive autonomous tiles per CTA, and (unlisted: spread order) slow tiles
    in between; bc5 at 768 x 384: its thick slab contains tiles without any relaxed cell, which are dropped / skipped"""
    G._fused_pass_check(num, X, Y, t_list=(3, 8), need=2, listed=listed)


@pytest.mark.parametrize("seed", [0, 1])
def test_emu_fused_pass_random_obstacles(env, seed):
    G.test_fused_pass_random_obstacles(env, seed, size=(320, 160), t_list=(4, 8))


@pytest.mark.parametrize("num,X,Y", [(5, 384, 192), (1, 300, 480)])
def test_emu_tile_list_classes(env, num, X, Y):
    G.test_tile_list_classes(env, num, X, Y)


@pytest.mark.parametrize("num,X,Y,n_iter", [(1, 128, 64, 7), (3, 320, 160, 13), (2, 96, 48, 3)])
def test_emu_fused_update_equals_literal_update(env, num, X, Y, n_iter):
    G.test_fused_update_equals_literal_update(env, num, X, Y, n_iter)


@pytest.mark.parametrize("num,X,Y", [(2, 256, 128), (3, 200, 176), (5, 333, 208), (1, 64, 48)])
def test_emu_stream_kernels_equal_direct_kernels(env, num, X, Y):
    G.test_stream_kernels_equal_direct_kernels(env, num, X, Y)


@pytest.mark.parametrize("cfg", [0, 2, 3])
def test_emu_stream_kernel_shapes(env, cfg):
    """every {stages x CTAs/SM x threads} shape of the TMA streaming kernel (fs2d_set_tuning(3, cfg)) == direct kernels"""
    env.fs2d_set_tuning(3, cfg)
    try:
        G.test_stream_kernels_equal_direct_kernels(env, 3, 200, 176)
    finally:
        env.fs2d_set_tuning(3, 1)


def test_emu_fused_pass_split_into_interior_and_edge_launches(env):
    G.test_fused_pass_split_into_interior_and_edge_launches(env, X=420, Y=160)


@pytest.mark.parametrize("num,res,scheme", [(2, 96, "kk"), (5, 100, "upwind"), (2, 256, "cip")])
def test_emu_dye_simulator_vs_oracle(env, num, res, scheme):
    G.test_dye_simulator_vs_oracle(env, num, res, scheme)


# ---- the two tails of the update ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("num,X,Y,n_iter", [(1, 128, 64, 7), (3, 320, 160, 13), (2, 96, 48, 3), (4, 200, 96, 4), (1, 288, 352, 6)])
def test_emu_two_literal_tail_equals_literal_update(env, num, X, Y, n_iter):
    G.test_two_literal_tail_equals_literal_update(env, num, X, Y, n_iter)


def test_emu_emitting_tail_on_open_tiles(env):
    """the default (emitting) tail on a grid with open-fluid tiles"""
    G.test_fused_update_equals_literal_update(env, 1, 288, 352, 6)


@pytest.mark.parametrize("seed", range(8))
def test_emu_fused_pass_fuzz(env, seed):
    """random grid sizes and masks (thick blocks / thin walls / stray inflow and outflow cells): for every pass size the host
    layer declares valid, a fused pass == literal iterations.  (A longer run of the same generator -- 140 masks, 1500 passes,
    round-1 kernels -- found no mismatch.)"""
    import numpy as np

    rng = np.random.default_rng(7000 + seed)
    X, Y = int(rng.integers(100, 260)), 16 * int(rng.integers(4, 14))
    mask = np.zeros((X, Y), np.uint8)
    mask[:, :2] = 1; mask[:, -2:] = 1
    kind = seed % 3
    for _ in range(int(rng.integers(5, 40))):
        i, j = int(rng.integers(2, X - 8)), int(rng.integers(2, Y - 8))
        h, w = (int(rng.integers(1, 8)), int(rng.integers(1, 8))) if kind else (int(rng.integers(3, 30)), int(rng.integers(3, 30)))
        mask[i:i + h, j:j + w] = 1
    mask[:2, 2:-2] = np.where(rng.random((2, Y - 4)) < 0.8, 2, mask[:2, 2:-2])
    mask[-2:, 2:-2] = np.where(rng.random((2, Y - 4)) < 0.7, 3, mask[-2:, 2:-2])
    if kind == 2:
        for _ in range(6):
            mask[int(rng.integers(3, X - 3)), int(rng.integers(3, Y - 3))] = int(rng.integers(2, 4))
    G._fused_pass_check(1, X, Y, mask_override=mask, t_list=(1, 3, 8), need=0, listed=seed % 2 == 0)


# ---- adversarial schedules ---------------------------------------------------------------------------------------------------
@pytest.mark.skipif(os.environ.get("FS2D_EMU_SCHED") is not None, reason="already inside an adversarial-schedule run")
@pytest.mark.parametrize("seed", [1])
def test_emu_synchronisation_under_adversarial_warp_schedules(seed):
    """The kernels with hand-written synchronisation (the fused Jacobi kernel: autonomous warps with per-warp TMA refills
    and progress counters in open-fluid tiles, CTA barriers in the others; the TMA streaming kernels) once more under the emulator's
    starvation scheduler (FS2D_EMU_SCHED=<seed>): a victim warp -- the leader half of the time -- only runs when every
    other warp is blocked, so warps drift as far apart as the barriers allow.  Seeded mutants with a missing barrier or
    an early TMA refill pass the round-robin schedule but fail here."""
    import subprocess

    sel = ("288-352 or 300-480 or stream_kernel_shapes or random_obstacles or emitting_tail_on_open or "
           "(two_literal_tail and 128-64) or split_into_interior")
    out = subprocess.run([sys.executable, "-m", "pytest", __file__, "-x", "-q", "-p", "no:cacheprovider", "-k", sel],
                         capture_output=True, text=True, timeout=1200, env=dict(os.environ, FS2D_EMU_SCHED=str(seed)))
    assert out.returncode == 0 and " passed" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]


@pytest.mark.parametrize("num,X,Y", [(1, 128, 64), (3, 200, 176), (4, 97, 80)])
def test_emu_rbsor_fused_colours_equal_two_passes(env, num, X, Y):
    G.test_rbsor_fused_colours_equal_two_passes(env, num, X, Y)

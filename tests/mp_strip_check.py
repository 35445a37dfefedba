"""Multi-rank strip check, launched by torchrun (NCCL, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/mp_strip_check.py

Every case runs the same seeded problem (a) on P row strips with halo exchange and (b) on one GPU
(rank 0), then compares every physical buffer BITWISE.  Prints "MP_CHECK OK <n cases>".
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO / "2d-fluid-simulator_b200"))

from fs.boundary_condition import BoundaryCondition, build_scene  # noqa: E402
from fs.distributed import Partition  # noqa: E402
from fs.fluid_simulator import make_solver  # noqa: E402
from fs.halo import gather_owned  # noqa: E402

CASES = [
    # bc, X, Y, scheme, vc, pressure kwargs, steps, halo
    (2, 128, 64, "cip", 5.0, dict(pressure="jacobi", n_iter=6), 4, 2),
    (3, 160, 80, "cip", 10.0, dict(pressure="jacobi", n_iter=5), 3, 2),
    (5, 128, 64, "kk", 5.0, dict(pressure="rbsor", n_iter=2), 4, 2),
    (1, 128, 64, "upwind", None, dict(pressure="jacobi", n_iter=3), 4, 3),
    (4, 96, 48, "cip", 2.0, dict(pressure="rbsor", n_iter=3), 3, 4),
    (2, 1024, 512, "cip", 5.0, dict(pressure="jacobi", n_iter=8), 2, 2),
    (2, 1024, 512, "cip", 5.0, dict(pressure="jacobi", n_iter=40), 2, 9),      # fused passes of 8 across the strip edge
    (5, 384, 192, "cip", 5.0, dict(pressure="jacobi", n_iter=21), 3, 6),       # odd count, passes of <= 5
    (3, 640, 320, "upwind", None, dict(pressure="jacobi", n_iter=30), 2, 9),
    (2, 640, 64, "cip", 5.0, dict(pressure="jacobi", n_iter=13), 2, 9),       # tall narrow strips: fused passes split into interior + edge launches
    (2, 640, 64, "cip", 5.0, dict(pressure="jacobi", n_iter=30), 2, 20),      # deep halo: several fused passes per exchange
    (5, 576, 64, "cip", 5.0, dict(pressure="jacobi", n_iter=23), 2, 24),      # the same with an odd count on bc5
    ("rand1", 192, 64, "cip", 5.0, dict(pressure="jacobi", n_iter=11), 2, 5),  # thin walls on / next to the strip edges
    ("rand3", 192, 64, "kk", None, dict(pressure="jacobi", n_iter=14), 2, 6),
]
# GPU runs only (seconds on a B200): the BASELINE scenes with their own sweep counts, fused passes of 8 across every strip edge
GPU_CASES = [
    (2, 8192, 2048, "cip", 5.0, dict(pressure="jacobi", n_iter=80), 2, 9),     # bc2, 80 sweeps (configs 2 / 4)
    (2, 8192, 2048, "cip", 5.0, dict(pressure="jacobi", n_iter=80), 2, 33),    # the same with deep halos (4 passes per exchange): the bench's setting
    (5, 4096, 2048, "cip", 5.0, dict(pressure="jacobi", n_iter=200), 2, 9),    # bc5, 200 sweeps (config 5 at res 2048)
    (3, 4096, 2048, "cip", 10.0, dict(pressure="jacobi", n_iter=100), 1, 9),   # bc3, vc=10, 100 sweeps (config 3 at res 2048)
]
# FS2D_STRIP_BIG=1: BASELINE config 5 itself (bc5, res=16384: 32768 x 16384 cells, 200 sweeps), quiescent start, 2 steps; the
# single-domain run needs 45 GB on rank 0's GPU
BIG_CASES = [(5, 32768, 16384, "cip", 5.0, dict(pressure="jacobi", n_iter=200), 2, 33)]
SEED_MAX_CELLS = 1 << 25     # larger grids start quiescent (all fields zero) instead of from seeded random buffers


def random_scene(seed: int, X: int, Y: int):
    """adversarial mask (as tests/test_host_logic.py): walls one cell thick, ragged inflow / outflow, stray BC cells"""
    rng = np.random.default_rng(100 + seed)
    mask = np.zeros((X, Y), dtype=np.uint8)
    mask[:, :2] = 1
    mask[:, -2:] = 1
    for _ in range(int(rng.integers(20, 40))):
        i, j = int(rng.integers(4, X - 8)), int(rng.integers(2, Y - 6))
        mask[i:i + int(rng.integers(1, 7)), j:j + int(rng.integers(1, 7))] = 1
    mask[:2, 2:-2] = np.where(rng.random((2, Y - 4)) < 0.8, 2, mask[:2, 2:-2])
    mask[-2:, 2:-2] = np.where(rng.random((2, Y - 4)) < 0.7, 3, mask[-2:, 2:-2])
    for _ in range(4):
        mask[int(rng.integers(3, X - 3)), int(rng.integers(3, Y - 3))] = int(rng.integers(2, 4))
    const = np.zeros((X, Y, 2), dtype=np.float32)
    const[mask == 2] = (1.0, 0.0)
    return const, mask


def buffers(s) -> dict:
    d = {"v_cur": s.v.current, "v_nxt": s.v.next, "p_cur": s.p.current, "p_nxt": s.p.next}
    if hasattr(s, "vx"):
        d.update(vx_cur=s.vx.current, vx_nxt=s.vx.next, vy_cur=s.vy.current, vy_nxt=s.vy.next)
    if s.vorticity_confinement is not None:
        d.update(vort=s.vorticity_confinement.vorticity, vort_abs=s.vorticity_confinement.vorticity_abs)
    return d


def main() -> None:
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    if os.environ.get("FS2D_FAKE_LIB") in ("1", "emu"):
        # CPU dry run of THIS script (gloo).  "1": tests/fake_fs2d.py, the oracle-backed stand-in for libfs2d.so (host layer
        # and script logic only).  "emu": the kernel SOURCES compiled against the CUDA emulation of tests/cuda_emu -- the real
        # kernels' row-window / clamp-window handling on strips, split fused passes, 1-2 row edge windows.  Small cases only.
        sys.path[:0] = [str(REPO), str(REPO / "tests"), str(REPO / "tests" / "cuda_emu")]
        from fs import _lib

        if os.environ["FS2D_FAKE_LIB"] == "1":
            from fake_fs2d import FakeFs2d

            FakeFs2d(_lib.load()).install_plain()
        else:
            import ctypes

            import build_emu

            emu = ctypes.CDLL(str(build_emu.LIB))      # built by the test that launches this script
            for name, (res, args) in _lib._SIGNATURES.items():
                fn = getattr(emu, name)
                fn.restype, fn.argtypes = res, args
            _lib._lib = emu
            _lib.ptr = lambda t: None if t is None else t.data_ptr()
            _lib.stream = lambda: 0
        dev = torch.device("cpu")
        dist.init_process_group("gloo")
        cases = [c for c in CASES if c[1] * c[2] <= int(os.environ.get("FS2D_DRY_MAX_CELLS", 640 * 64))]
    else:
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
        dist.init_process_group("nccl", device_id=dev)
        cases = CASES + GPU_CASES + (BIG_CASES if os.environ.get("FS2D_STRIP_BIG") == "1" else [])
        if os.environ.get("FS2D_STRIP_ONLY_BIG") == "1":
            cases = BIG_CASES
    for item in filter(None, os.environ.get("FS2D_STRIP_TUNING", "").split(",")):   # e.g. "4=1": experimental kernel variants
        from fs import _lib as _l

        k, v = item.split("=")
        _l.call("fs2d_set_tuning", int(k), int(v))
    n_ok = 0
    # cases tall enough run a second time on strips of UNEQUAL height (Partition.bounds: what bench.py --config 5 uses to balance
    # the work of scenes with unevenly distributed walls)
    runs = [(c, False) for c in cases] + [(c, True) for c in cases if not isinstance(c[0], str) and c[1] >= 512 and c[1] // world >= 2 * c[7] + 16
                                           and c[1] * c[2] <= SEED_MAX_CELLS]
    for (num, X, Y, scheme, vc, pkw, steps, halo), skewed in runs:
        res = Y
        dt, dx, re = 0.05 / res, 1.0 / res, 1e4
        bounds = None
        if skewed:   # strip k ends at X * ((k + 1) / world) ** 1.6, never thinner than the halo
            bounds = [0]
            for k in range(1, world):
                bounds.append(min(max(int(X * (k / world) ** 1.6), bounds[-1] + halo + 8), X - (world - k) * (halo + 8)))
            bounds.append(X)
        part = Partition(X, rank, world, halo, tuple(bounds) if bounds else None)
        if isinstance(num, str):
            const, mask = random_scene(int(num[4:]), X, Y)
            strip = make_solver(BoundaryCondition(const, mask, device=dev, partition=part), dt, dx, re, vc, scheme, **pkw)
        else:   # the strip is built from ITS rows of the scene only (build_scene(rows=...), row_offset), the single domain from the whole scene
            a0, a1 = BoundaryCondition.strip_rows(part)
            c_w, m_w = build_scene(num, X, Y, rows=(a0, a1))
            strip = make_solver(BoundaryCondition(c_w, m_w, device=dev, partition=part, row_offset=a0), dt, dx, re, vc, scheme, **pkw)
            del c_w, m_w
            const, mask = build_scene(num, X, Y) if rank == 0 else (None, None)
        single = make_solver(BoundaryCondition(const, mask, device=dev), dt, dx, re, vc, scheme, **pkw) if rank == 0 else None
        del const, mask
        rng = np.random.default_rng(1234 + (num if isinstance(num, int) else 50 + int(num[4:])))
        g0, g1 = part.owned()
        p_first = None
        for k, f in (buffers(strip).items() if X * Y <= SEED_MAX_CELLS else ()):  # same seeded global state on every rank (incl. "next" buffers)
            shape = (X, Y, 2) if f.n == 2 else (X, Y)
            scale = 0.05 / dx if k[:2] in ("vx", "vy") else (0.5 if k[0] == "v" and k != "vort" else 1.0)
            a = (rng.uniform(-1, 1, shape) * scale).astype(np.float32)
            if k == "vort_abs":
                a = np.abs(a)
            if k in ("p_cur", "p_nxt") and pkw["n_iter"] > 8:
                # equal pressure buffers: their never-written wall cells agree, so the FUSED passes engage on the strips
                # (with independent random buffers the host layer falls back to literal iterations)
                p_first = a if p_first is None else p_first
                a = p_first
            f.from_numpy(a[g0:g1])
            if single is not None:
                buffers(single)[k].from_numpy(a)
        for _ in range(steps):
            strip.update()
            if single is not None:
                single.update()
        for k, f in buffers(strip).items():
            got = gather_owned(f, part)
            if rank == 0:
                want = buffers(single)[k].tensor
                same = torch.equal(got, want) or bool(((got == want) | (got.isnan() & want.isnan())).all())
                if not same:
                    bad = ~((got == want) | (got.isnan() & want.isnan()))
                    rows = torch.nonzero(bad.reshape(X, -1).any(1)).flatten()[:8].tolist()
                    raise SystemExit(f"MP_CHECK FAIL case bc{num} {X}x{Y} {scheme} {pkw} buffer {k}: "
                                     f"{int(bad.sum())} values differ, first rows {rows}")
        n_ok += 1
        if rank == 0:
            print(f"  case ok: bc{num} {X}x{Y} {scheme} vc={vc} {pkw} steps={steps} halo={halo} on {world} ranks{' strips ' + str(bounds) if bounds else ''}"
                  f"{'' if X * Y <= SEED_MAX_CELLS else ' (quiescent start)'}", flush=True)
        del strip, single
        if dev.type == "cuda":
            torch.cuda.empty_cache()
        dist.barrier()
    if rank == 0:
        print(f"MP_CHECK OK {n_ok} cases on {world} ranks", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""N>1 path: halo exchange logic on CPU (gloo, world_size 2) and the NCCL strip check on >=2 GPUs."""
from __future__ import annotations

import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from conftest import REPO


def free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, X: int, Y: int, halo: int, iters: int, q) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, str(REPO / "2d-fluid-simulator_b200"))
    from fs.distributed import Partition
    from fs.halo import HaloExchanger

    part = Partition(X, rank, world, halo)
    hx = HaloExchanger(part)
    g0, g1 = part.owned()
    full = torch.arange(X * Y * 2, dtype=torch.float32).reshape(X, Y, 2).sin()
    loc = torch.zeros((g1 - g0 + 2 * halo, Y, 2))
    loc[halo:halo + g1 - g0] = full[g0:g1]
    # 1. halo rows after an exchange are the neighbour's owned rows, for every width <= halo
    for w in range(1, halo + 1):
        t = loc.clone()
        hx.exchange(t, w)
        if part.has_lower:
            assert torch.equal(t[halo - w:halo], full[g0 - w:g0]), (rank, w)
        if part.has_upper:
            assert torch.equal(t[halo + g1 - g0:halo + g1 - g0 + w], full[g1:g1 + w]), (rank, w)
        assert torch.equal(t[halo:halo + g1 - g0], full[g0:g1])
    # 1b. split-phase exchange: work between start() and finish() on other rows does not disturb it
    t = loc.clone()
    reqs = hx.start(t, halo)
    t[2 * halo:g1 - g0] += 1.0                     # "interior kernel": owned rows that are neither sent nor received
    hx.finish(reqs)
    if part.has_lower:
        assert torch.equal(t[:halo], full[g0 - halo:g0])
    if part.has_upper:
        assert torch.equal(t[halo + g1 - g0:], full[g1:g1 + halo])
    with np.testing.assert_raises(ValueError):
        hx.exchange(loc, halo + 1)
    ok = True
    q.put((rank, ok, hx.n_exchanges))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("halo", [2, 3])
def test_halo_exchange_gloo_world2(halo):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 20, 6, halo, 0, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=5) for _ in range(2))
    assert all(ok for _, ok, _ in res) and all(n == halo + 1 for _, _, n in res)


def _strip_worker(rank: int, world: int, port: int, q) -> None:
    """Jacobi-like iteration on strips: interior rows use halo rows, global edges clamp (dom logic)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, str(REPO / "2d-fluid-simulator_b200"))
    from fs.distributed import Partition
    from fs.halo import HaloExchanger

    X, Y, H, iters = 24, 7, 2, 9
    part = Partition(X, rank, world, H)
    hx = HaloExchanger(part)
    g0, g1 = part.owned()
    w0, w1 = part.window()
    lo, hi = max(w0, 0), min(w1, X)
    clo, chi = lo - w0, hi - 1 - w0           # the fs2d_dom clamp bounds of this rank
    full = torch.linspace(0, 1, X * Y).reshape(X, Y).cos()
    loc = torch.zeros((w1 - w0, Y))
    loc[lo - w0:hi - w0] = full[lo:hi]

    def sweep(a, r0, r1, c_lo, c_hi):
        rows = torch.arange(a.shape[0])
        up = a[(rows - 1).clamp(c_lo, c_hi)]
        dn = a[(rows + 1).clamp(c_lo, c_hi)]
        lf = torch.cat([a[:, :1], a[:, :-1]], 1)
        rt = torch.cat([a[:, 1:], a[:, -1:]], 1)
        out = a.clone()
        out[r0:r1] = (0.25 * (dn + up + rt + lf))[r0:r1]
        return out

    ref = full.clone()
    for _ in range(iters):
        ref = sweep(ref, 0, X, 0, X - 1)
        hx.exchange(loc, 1)
        loc = sweep(loc, g0 - w0, g1 - w0, clo, chi)
    q.put((rank, bool(torch.equal(loc[g0 - w0:g1 - w0], ref[g0:g1]))))
    dist.barrier()
    dist.destroy_process_group()


def test_strip_iteration_equals_single_domain_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_strip_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ok for _, ok in (q.get(timeout=5) for _ in range(2)))


@pytest.mark.gpu
def test_strips_bitwise_equal_single_gpu_nccl():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    world = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), str(REPO / "tests" / "mp_strip_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "MP_CHECK OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("world", [4])
def test_strip_check_script_dry_run_on_cpu(world):
    """The NCCL strip check (tests/mp_strip_check.py) run on CPU: gloo + the oracle-backed stand-in for libfs2d.so
    (tests/fake_fs2d.py).  Exercises the script itself and the whole multi-rank host layer under torchrun, small cases."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), str(REPO / "tests" / "mp_strip_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, FS2D_FAKE_LIB="1"))
    import re

    m = re.search(r"MP_CHECK OK (\d+) cases on (\d+) ranks", out.stdout)
    assert out.returncode == 0 and m and int(m.group(2)) == world and int(m.group(1)) >= 12, out.stdout[-2000:] + out.stderr[-2000:]
    assert "strips [0," in out.stdout      # ... some of them on strips of unequal height


@pytest.mark.parametrize("world", [3])
def test_strip_check_script_with_emulated_kernels(world):
    """The same script with the KERNEL SOURCES running on the CUDA emulation (tests/cuda_emu) instead of the oracle stand-in:
    the real kernels on strips -- row windows with clamp bounds inside the local array, 1-2 row edge windows of the overlap
    scheme, fused passes split into interior + edge launches -- must equal the single-domain run bitwise."""
    sys.path.insert(0, str(REPO / "tests" / "cuda_emu"))
    import build_emu

    build_emu.build()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), str(REPO / "tests" / "mp_strip_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1200, env=dict(os.environ, FS2D_FAKE_LIB="emu"))
    import re

    m = re.search(r"MP_CHECK OK (\d+) cases on (\d+) ranks", out.stdout)
    assert out.returncode == 0 and m and int(m.group(2)) == world and int(m.group(1)) >= 12, out.stdout[-2000:] + out.stderr[-2000:]
    assert "strips [0," in out.stdout      # ... some of them on strips of unequal height


def test_split_windows_cover_the_strip_and_keep_the_interior_off_the_halo():
    """fs.halo.split_windows: the interior window is a whole number of tile rows starting one tile row into the strip,
    and its reads (t rows beyond it) stay inside the owned rows."""
    sys.path.insert(0, str(REPO / "2d-fluid-simulator_b200"))
    from fs._lib import Dom
    from fs.halo import split_windows

    for rows, halo, ti, t in [(8192, 9, 80, 8), (500, 9, 80, 8), (250, 13, 72, 12), (4096, 9, 84, 6), (200, 9, 80, 8), (96, 4, 90, 3)]:
        d = Dom(rows + 2 * halo, 64, halo, halo + rows, 0, rows + 2 * halo - 1, 0)
        mid, m = split_windows(d, ti, t)
        k = (rows + ti - 1) // ti
        if mid is None:
            assert (rows - t) // ti < 2 or k < 3
            continue
        assert mid.r0 == d.r0 + ti and mid.r1 == d.r0 + m * ti and 2 <= m <= k - 1
        assert mid.r0 - t >= d.r0 and mid.r1 + t <= d.r1

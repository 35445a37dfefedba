"""Halo exchange and the multi-rank versions of the operators.

Row-strip decomposition: rank r owns global rows [g0, g1) and keeps `halo` extra rows of its
neighbours on each side of every field (see fs/distributed.py).  The ONLY communication of the
solver is a nearest-neighbour SendRecv of whole grid rows (NCCL over NVLink through
torch.distributed; gloo in the CPU tests): no reductions, no global state.

Invariant that makes P strips bit-identical to one GPU: a kernel that updates the owned rows reads
rows up to its stencil radius away; those rows were copied from the owner's memory AFTER the owner's
last write, so every load sees the same bits as in the single-GPU run.
"""
from __future__ import annotations

from contextlib import contextmanager

import torch
import torch.distributed as dist

from fs.double_buffer import DoubleBuffer, Field


class HaloExchanger:
    def __init__(self, partition, group=None) -> None:
        self.part = partition
        self.group = group
        self.n_exchanges = 0
        self.bytes_sent = 0

    def start(self, field: Field | torch.Tensor | list, width: int = 0) -> list:
        """Post the SendRecv that fills the `width` halo rows on both sides; returns the requests for finish().
        `field` may be a list of (field, width) pairs: all of them go into ONE batch (one NCCL group).
        Kernels launched between start() and finish() run concurrently with the transfer (the communication stream only
        waits for work enqueued before start()): they must neither write the rows being sent nor touch the halo rows."""
        p = self.part
        pairs = field if isinstance(field, list) else [(field, width)]
        pairs = [(f, w) for f, w in pairs if w > 0]
        if p.world == 1 or not pairs:
            return []
        ops, r = [], p.rank
        for f, w in pairs:
            t = f.tensor if isinstance(f, Field) else f
            H, rows = p.halo, t.shape[0]
            if w > H:
                raise ValueError(f"halo exchange of {w} rows but the partition only has {H} halo rows")
            if p.has_lower:
                ops.append(dist.P2POp(dist.isend, t[H:H + w], r - 1, self.group))            # my first owned rows
                ops.append(dist.P2POp(dist.irecv, t[H - w:H], r - 1, self.group))            # their last owned rows
            if p.has_upper:
                ops.append(dist.P2POp(dist.isend, t[rows - H - w:rows - H], r + 1, self.group))
                ops.append(dist.P2POp(dist.irecv, t[rows - H:rows - H + w], r + 1, self.group))
        self.n_exchanges += 1
        self.bytes_sent += sum(op.tensor.numel() * op.tensor.element_size() for op in ops[::2])
        return dist.batch_isend_irecv(ops) if ops else []

    @staticmethod
    def finish(reqs: list) -> None:
        """Make the current stream (the host, for gloo) wait for the transfer posted by start()."""
        for req in reqs:
            req.wait()

    def exchange(self, field: Field | torch.Tensor, width: int) -> None:
        """Fill the `width` halo rows adjacent to the owned rows on both sides from the neighbours."""
        self.finish(self.start(field, width))


def exchanger_for(bc) -> HaloExchanger:
    """The (single) exchanger of a BoundaryCondition; kept on the object so its lifetime follows the partition."""
    hx = getattr(bc, "_halo_exchanger", None)
    if hx is None:
        hx = bc._halo_exchanger = HaloExchanger(bc.partition)
    return hx


def _extended(bc, extra: int):
    """dom covering the owned rows plus `extra` rows on each side that exist globally."""
    d = bc.dom
    return d.replace(r0=max(d.r0 - extra, d.clo), r1=min(d.r1 + extra, d.chi + 1))


# ------------------------------------------------------------------------------------------------
# multi-rank operator bodies (called from the single-rank classes when partition.world > 1)
# ------------------------------------------------------------------------------------------------
def vorticity_apply_distributed(vc, v: DoubleBuffer) -> None:
    """fs/vorticity_confinement.py:57-59 on a strip: the fused kernel recomputes the neighbours' curl from v, so two
    fresh halo rows of v are all it needs per step (no per-step exchange of the vorticity fields)."""
    bc = vc._bc
    hx = exchanger_for(bc)
    if vc.vorticity_abs.dirty:
        # The stored |vorticity| of a NON-fluid cell is read by its fluid neighbours and never written by any kernel
        # (SURVEY T1), so a halo row only has to receive it once after the field was written from the host
        # (construction: zeros; load_state_dict / from_numpy -- which every rank must call alike, like any collective).
        hx.exchange(vc.vorticity_abs, 1)
        vc.vorticity_abs.dirty = False
    _overlapped(bc, hx, [(v.current, 2)], lambda: vc._apply_fused(v.next, v.current), 2)


@contextmanager
def _rows(bc, r0: int, r1: int):
    """Temporarily restrict the rows the kernels of `bc` update (they all take bc.dom)."""
    old = bc.dom
    bc.dom = old.replace(r0=r0, r1=r1)
    try:
        yield
    finally:
        bc.dom = old


def _overlapped(bc, hx: HaloExchanger, exchanges: list, launch, reach: int) -> None:
    """exchange(...) followed by launch(), with the transfer hidden behind the kernel: post the SendRecv, run `launch`
    on the owned rows that read no halo row ([r0 + reach, r1 - reach)), wait, then run it on the `reach` edge rows of
    each side.  `launch` must write neither the exchanged fields nor anything it reads from other rows."""
    d = bc.dom
    if d.r1 - d.r0 < 4 * reach + 16:
        hx.finish(hx.start(exchanges))
        launch()
        return
    reqs = hx.start(exchanges)
    with _rows(bc, d.r0 + reach, d.r1 - reach):
        launch()
    hx.finish(reqs)
    with _rows(bc, d.r0, d.r0 + reach):
        launch()
    with _rows(bc, d.r1 - reach, d.r1):
        launch()


def _pass_windows(bc, t: int):
    """Interior row window of a fused pass of t iterations on this strip and the number m of tile rows before its end
    (tile rows [1, m) are interior), or (None, 0) if the strip has too few tile rows to split."""
    import ctypes

    from fs import _lib

    rows, cols, hr, hc, tmax = (ctypes.c_int() for _ in range(5))
    _lib.call("fs2d_fused_tile", t, ctypes.byref(rows), ctypes.byref(cols), ctypes.byref(hr), ctypes.byref(hc), ctypes.byref(tmax))
    return split_windows(bc.dom, rows.value - 2 * hr.value, t)


def split_windows(d, ti: int, t: int):
    """(mid, m): `mid` = copy of dom `d` covering the tile rows [1, m) of the tiling of [r0, r1) by tiles of ti output
    rows, i.e. rows [r0 + ti, r0 + m*ti), chosen such that a pass of t iterations over it (reading t rows beyond its
    tiles) touches owned rows only: r0 + ti - t >= r0 and r0 + m*ti + t <= r1.  (None, 0) when no such non-empty window
    exists.  The remaining tile rows {0} U [m, k) read the halo rows."""
    n = d.r1 - d.r0
    if ti < t or ti <= 0:
        return None, 0
    m = min((n + ti - 1) // ti - 1, (n - t) // ti)
    if m < 2:
        return None, 0
    return d.replace(r0=d.r0 + ti, r1=d.r0 + m * ti), m


def jacobi_update_distributed(jac, p: DoubleBuffer, v_current: Field) -> None:
    """fs/pressure_updater.py:56-60 on a strip.  Same schedule as fs2d_jacobi_update: a fused pass of t
    iterations needs t fresh halo rows of p (one SendRecv per PASS instead of per iteration) and the source
    terms on those rows; a literal iteration needs 2 (row g0-1 for the stencil, g0-2 for the BC of row g0-1)."""
    bc = jac._bc
    hx = exchanger_for(bc)
    plan = jac.plan(p)
    groups = jacobi_groups(jac, p, plan)
    reach = max([sum(t for _, t in g) for g in groups], default=0)   # rows beyond the owned ones on which a pass of the update runs
    # source terms also on the halo rows a pass reads; the rows that need no halo row of v run during the exchange
    d, ext = bc.dom, _extended(bc, reach)
    if d.r1 - d.r0 >= 16:
        reqs = hx.start(v_current, min(bc.halo, reach + 1))
        jac._source(v_current, dom=d.replace(r0=d.r0 + 1, r1=d.r1 - 1))
        hx.finish(reqs)
        jac._source(v_current, dom=ext.replace(r1=d.r0 + 1))
        src = jac._source(v_current, dom=ext.replace(r0=d.r1 - 1))
    else:
        hx.exchange(v_current, min(bc.halo, reach + 1))
        src = jac._source(v_current, dom=ext)
    # the schedule ends {fused pass that emits the BC values of its penultimate state, ONE literal iteration} (or, with
    # fs2d_set_tuning(4, 0), two literal iterations)
    tail_at = len(plan) - 2 if len(plan) >= 2 and plan[-2] > 0 else -1
    for group in groups:
        k0, t0 = group[0]
        if t0 == 0:      # literal iteration
            hx.exchange(p.current, 2)
            bc.set_pressure_boundary_condition(p.current)       # owned rows and the first halo row
            jac._sweep(p.next, p.current, src, inline_bc=False)
            p.swap()
        elif len(group) == 1:
            # Overlap: the pass reads the fresh halo rows only in its first and last TILE ROW, so the tile rows in
            # between run while the SendRecv is in flight.  The interior window starts at a multiple of the tile height
            # from r0, so both launches use exactly the tiles of a single launch (the validated tiling, fused_reach_ok).
            emit = k0 == tail_at
            mid, m = _pass_windows(bc, t0)
            if mid is None:
                hx.exchange(p.current, t0)
                jac._fused(p.next, p.current, src, t0, emit=emit)
            else:
                reqs = hx.start(p.current, t0)
                jac._fused(p.next, p.current, src, t0, dom=mid, emit=emit)             # tile rows [1, m): owned rows only
                hx.finish(reqs)
                jac._fused(p.next, p.current, src, t0, skip=(1, m - 1), emit=emit)     # tile rows {0} U [m, k) in one launch
            p.swap()
        else:
            # Deep halo: ONE exchange for the whole group of passes.  Pass i recomputes, beyond its owned rows, the rows the
            # later passes of the group read (sum of their sizes: a few dozen rows of 8192), so none of them needs an
            # exchange or a split launch.  Both buffers travel: the passes ping-pong between them and read the never-written
            # wall cells of either in rows that only an exchange fills.
            total = sum(t for _, t in group)
            hx.finish(hx.start([(p.current, total), (p.next, total)]))
            ext = total
            for k, t in group:
                ext -= t
                jac._fused(p.next, p.current, src, t, dom=_extended(bc, ext) if ext else None, emit=k == tail_at)
                p.swap()


def jacobi_groups(jac, p: DoubleBuffer, plan: list[int]) -> list[list[tuple[int, int]]]:
    """The plan's entries grouped into exchanges: [(index, size), ...] per group.  Consecutive fused passes share ONE halo
    exchange as long as the sum of their sizes fits the halo (halo >= sum + 1) and every pass is valid on its extended
    window on EVERY rank (BoundaryCondition.fused_ok(T, ext); agreed once per plan with a MIN all-reduce, setup time)."""
    bc = jac._bc
    key = (tuple(plan), bc.halo)
    cache = jac.__dict__.setdefault("_group_cache", {})
    if key in cache:
        return cache[key]

    def build(max_group: int) -> list[list[tuple[int, int]]]:
        groups, cur = [], []
        for k, t in enumerate(plan):
            if t == 0:
                if cur:
                    groups.append(cur)
                    cur = []
                groups.append([(k, 0)])
            elif cur and len(cur) < max_group and sum(x for _, x in cur) + t <= bc.halo - 1:
                cur.append((k, t))
            else:
                if cur:
                    groups.append(cur)
                cur = [(k, t)]
        if cur:
            groups.append(cur)
        return groups

    groups = build(MAX_PASSES_PER_EXCHANGE)
    ok = True
    for g in groups:
        ext = sum(t for _, t in g)
        for _, t in g:
            ext -= t
            if len(g) > 1 and t > 0 and not bc.fused_ok(t, ext):
                ok = False
    if bc.partition.world > 1 and dist.is_initialized():
        flag = torch.tensor([int(ok)], dtype=torch.int32, device=p.current.tensor.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(int(flag.item()))
    cache[key] = groups if ok else build(1)
    return cache[key]


#: fused passes that may share one halo exchange (deep halo); 1 = one exchange per pass (round 1)
MAX_PASSES_PER_EXCHANGE = 4


def rbsor_update_distributed(sor, p: DoubleBuffer, v_current: Field) -> None:
    """fs/pressure_updater.py:86-96 on a strip: the even pass reads odd cells of p.next written in the
    same iteration, so p.next's halo row is refreshed between the two colour passes."""
    bc = sor._bc
    hx = exchanger_for(bc)
    hx.exchange(v_current, 1)
    src = sor._source(v_current)
    for _ in range(sor._n_iter):
        hx.exchange(p.current, 2)
        bc.set_pressure_boundary_condition(p.current)
        sor._pass(p.next, p.current, src, 1)
        hx.exchange(p.next, 1)
        sor._pass(p.next, p.next, src, 0)
        p.swap()


def _velocity_bc(bc, hx: HaloExchanger, v: Field, reach: int) -> None:
    """set_velocity_boundary_condition on a strip; afterwards v is post-BC on the owned rows +-reach."""
    hx.exchange(v, bc.bc_halo)                    # sources lie up to 2 rows away from a target
    bc.set_velocity_boundary_condition(v)         # targets: owned rows +- (bc_halo - 2)
    if bc.bc_halo - 2 < reach:
        hx.exchange(v, reach)


def cip_update_distributed(s) -> None:
    """CipMacSolver.update() (fs/solver.py:192-227) on a strip."""
    bc = s._bc
    hx = exchanger_for(bc)
    v, vx, vy, p = s.v, s.vx, s.vy, s.p
    _velocity_bc(bc, hx, v.current, 1)
    # every exchange below is hidden behind the interior rows of the kernel that needs it (_overlapped)
    _overlapped(bc, hx, [(p.current, 1)], lambda: s._non_advection_phase(v.next, v.current, p.current), 1)
    _overlapped(bc, hx, [(v.next, 1)],            # fn(i+-1, j) of the grad kernel
                lambda: s._non_advection_phase_grad(vx.next, vy.next, vx.current, vy.current, v.current, v.next), 1)
    v.swap(); vx.swap(); vy.swap()
    _overlapped(bc, hx, [(vx.current, 1), (vy.current, 1)],   # CIP reads f, fx, fy at the upwind row i_m = i +- 1
                lambda: s._advection_phase(v.next, vx.next, vy.next, v.current, vx.current, vy.current, v.current), 1)
    v.swap(); vx.swap(); vy.swap()
    if s.vorticity_confinement is not None:
        s.vorticity_confinement.apply(v)
        v.swap()
    s._pressure_update()
    from fs.solver import VELOCITY_LIMIT, limit_field

    limit_field(v.current, VELOCITY_LIMIT, bc=bc)


def mac_update_distributed(s) -> None:
    """MacSolver.update() (fs/solver.py:79-89) on a strip; KK reads +-2 rows."""
    bc = s._bc
    hx = exchanger_for(bc)
    _velocity_bc(bc, hx, s.v.current, s._advect.radius)
    hx.exchange(s.p.current, 1)
    s._update_velocities(s.v.next, s.v.current, s.p.current)
    s.v.swap()
    if s.vorticity_confinement is not None:
        s.vorticity_confinement.apply(s.v)
        s.v.swap()
    s._pressure_update()
    from fs.solver import VELOCITY_LIMIT, limit_field

    limit_field(s.v.current, VELOCITY_LIMIT, bc=bc)


def dye_update_distributed(s) -> None:
    """Dye part of DyeMacSolver.update (fs/solver.py:149-152) / DyeCipMacSolver.update (:366-373) on a strip."""
    from fs.solver import clamp_field

    bc = s._bc
    hx = exchanger_for(bc)
    dye = s.dye
    if hasattr(s, "dyex"):
        dx_, dy_ = s.dyex, s.dyey
        hx.exchange(dye.current, 1)
        bc.set_dye_boundary_condition(dye.current)            # every local inflow cell, halo rows included
        s._non_advection_phase_dye(dye.next, dye.current)
        hx.exchange(dye.next, 1)
        s._dye_grad(dx_.next, dy_.next, dx_.current, dy_.current, dye.current, dye.next)
        dye.swap(); dx_.swap(); dy_.swap()
        hx.exchange(dx_.current, 1)
        hx.exchange(dy_.current, 1)
        hx.exchange(s.v.current, 1)                           # d/dx, d/dy of the (limited) advecting velocity
        s._dye_advect(dye.next, dx_.next, dy_.next, dye.current, dx_.current, dy_.current, s.v.current)
        dye.swap(); dx_.swap(); dy_.swap()
    else:
        hx.exchange(dye.current, s._advect.radius)
        bc.set_dye_boundary_condition(dye.current)
        s._update_dye(dye.next, dye.current, s.v.current)
        dye.swap()
    clamp_field(dye.current, 0.0, 1.0, bc=bc)


def gather_owned(field: Field, partition, dst: int = 0) -> torch.Tensor | None:
    """Owned rows of every rank concatenated on rank `dst` (tests / field_to_numpy on strips)."""
    own = field.owned().contiguous()
    if partition.world == 1:
        return own
    sizes = [partition.owned(r)[1] - partition.owned(r)[0] for r in range(partition.world)]
    if partition.rank == dst:
        parts = [torch.empty((n,) + tuple(own.shape[1:]), dtype=own.dtype, device=own.device) for n in sizes]
        parts[dst].copy_(own)
        reqs = [dist.irecv(parts[r], r) for r in range(partition.world) if r != dst]
        for q in reqs:
            q.wait()
        return torch.cat(parts, 0)
    dist.send(own, dst)
    return None

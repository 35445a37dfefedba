"""Static tables derived once from the cell-type mask (host-side setup, torch used as plumbing).

The reference evaluates its boundary-condition branches per cell inside every BC kernel
(/root/reference/fs/boundary_condition.py:16-65).  The mask never changes, so we resolve the
branches once:

* `pcode`  (uint8 per cell, FS2D_PC_* in include/fs2d.h): which pressure-BC branch a cell takes.
  Dense kernels use it both as the write predicate (not-wall) and to recompute post-BC neighbour
  pressures inline.
* sparse `p` table  (tgt, src0, src1, kind): the in-place pressure BC as a gather list.
* sparse `vel` table (tgt, src, kind): the in-place velocity BC, resolved TARGET-centrically so the
  multi-writer scatter of the reference becomes a race-free gather: for each target the last writer
  in (i, j) order wins (SURVEY T4 pin; equals oracle `orc_vel_bc`).

All functions are device-agnostic torch code so they can be unit-tested on CPU.
"""
from __future__ import annotations

import torch

PC_FLUID, PC_W_IM, PC_W_IP, PC_W_JM, PC_W_JP = 0, 1, 2, 3, 4
PC_W_IM_JP, PC_W_IP_JP, PC_W_IM_JM, PC_W_IP_JM, PC_W_NONE, PC_INFLOW, PC_OUTFLOW = 5, 6, 7, 8, 9, 10, 11


def _shift(m: torch.Tensor, di: int, dj: int) -> torch.Tensor:
    """m[clamp(i+di), clamp(j+dj)] (clamp-to-edge, like the reference's raw/`sample` reads)."""
    X, Y = m.shape
    out = m
    if di:
        ii = (torch.arange(X, device=m.device) + di).clamp_(0, X - 1)
        out = out.index_select(0, ii)
    if dj:
        jj = (torch.arange(Y, device=m.device) + dj).clamp_(0, Y - 1)
        out = out.index_select(1, jj)
    return out


def _first_match(conds: list[torch.Tensor]) -> torch.Tensor:
    """elif chain: index (1-based) of the first true condition per cell, 0 if none."""
    out = torch.zeros_like(conds[0], dtype=torch.uint8)
    for k in range(len(conds) - 1, -1, -1):
        out = torch.where(conds[k], torch.full_like(out, k + 1), out)
    return out


def pressure_codes(mask: torch.Tensor) -> torch.Tensor:
    """FS2D_PC_* per cell; boundary_condition.py:41-65 (no interior guard, clamped mask reads)."""
    m = mask
    im, ip, jm, jp = _shift(m, -1, 0), _shift(m, 1, 0), _shift(m, 0, -1), _shift(m, 0, 1)
    f = lambda a: a == 0  # noqa: E731
    w = lambda a: a == 1  # noqa: E731
    branch = _first_match([
        f(im) & w(jm) & w(jp), f(ip) & w(jm) & w(jp), f(jm) & w(im) & w(ip), f(jp) & w(im) & w(ip),
        f(im) & f(jp), f(ip) & f(jp), f(im) & f(jm), f(ip) & f(jm)])
    code = torch.full_like(m, PC_FLUID)
    wall = m == 1
    code = torch.where(wall & (branch > 0), branch, code)
    code = torch.where(wall & (branch == 0), torch.full_like(m, PC_W_NONE), code)
    code = torch.where(m == 2, torch.full_like(m, PC_INFLOW), code)
    code = torch.where(m == 3, torch.full_like(m, PC_OUTFLOW), code)
    return code.contiguous()


def pack_pcode(code: torch.Tensor) -> torch.Tensor:
    """pcode byte of include/fs2d.h: code | (neighbour-is-a-BC-cell bits << 4), neighbours clamped."""
    bc_cell = (code != PC_FLUID) & (code != PC_W_NONE)
    out = code.clone()
    for bit, (di, dj) in ((0x10, (-1, 0)), (0x20, (1, 0)), (0x40, (0, -1)), (0x80, (0, 1))):
        out |= _shift(bc_cell, di, dj).to(torch.uint8) * bit
    return out.contiguous()


def velocity_writer_branch(mask: torch.Tensor) -> torch.Tensor:
    """Scatter branch 1..4 taken by each wall cell as a WRITER (boundary_condition.py:20-34), else 0."""
    m = mask
    X, Y = m.shape
    im, ip, jm, jp = _shift(m, -1, 0), _shift(m, 1, 0), _shift(m, 0, -1), _shift(m, 0, 1)
    f = lambda a: a == 0  # noqa: E731
    w = lambda a: a == 1  # noqa: E731
    br = _first_match([f(im) & w(jm) & w(jp), f(ip) & w(jm) & w(jp), f(jm) & w(im) & w(ip), f(jp) & w(im) & w(ip)])
    interior = torch.zeros_like(m, dtype=torch.bool)
    interior[1:X - 1, 1:Y - 1] = True
    return torch.where((m == 1) & interior, br, torch.zeros_like(br))


def _window_index(i: torch.Tensor, j: torch.Tensor, w0: int, w1: int, Y: int, what: str) -> torch.Tensor:
    if i.numel() and (int(i.min()) < w0 or int(i.max()) >= w1):
        raise ValueError(f"{what}: a boundary-condition source row lies outside the local window [{w0}, {w1}); "
                         "increase the halo width")
    return ((i - w0) * Y + j).to(torch.int32)


def velocity_table(mask: torch.Tensor, t0: int | None = None, t1: int | None = None, w0: int = 0,
                   w1: int | None = None) -> dict:
    """Target-centric gather list for set_velocity_boundary_condition.

    Targets are the global rows [t0, t1); indices are local to the row window [w0, w1).
    kind 0: v[tgt] = -v[src]; kind 1: v[tgt] = bc_const[tgt]; kind 2: v[tgt].x = max(v[src].x, .05).
    """
    X, Y = mask.shape
    t0, t1 = (0 if t0 is None else t0), (X if t1 is None else t1)
    w1 = X if w1 is None else w1
    wb = velocity_writer_branch(mask)

    def writer(di, dj, b):  # is the cell at (i+di, j+dj) a writer taking branch b?  (zero outside)
        s = torch.zeros_like(wb, dtype=torch.bool)
        xs, xd = slice(max(di, 0), X + min(di, 0)), slice(max(-di, 0), X + min(-di, 0))
        ys, yd = slice(max(dj, 0), Y + min(dj, 0)), slice(max(-dj, 0), Y + min(-dj, 0))
        s[xd, yd] = wb[xs, ys] == b
        return s

    # candidates in DEscending writer order: last writer in (i, j) order wins
    c_b2, c_b4 = writer(1, 0, 2), writer(0, 1, 4)
    own2, own3 = mask == 2, mask == 3
    c_b3, c_b1 = writer(0, -1, 3), writer(-1, 0, 1)
    sel = _first_match([c_b2, c_b4, own2, own3, c_b3, c_b1])
    rows = torch.zeros_like(sel, dtype=torch.bool)
    rows[t0:t1] = True
    ti, tj = torch.nonzero((sel > 0) & rows, as_tuple=True)
    s = sel[ti, tj].to(torch.int64)
    # source offset per selected candidate: b2 -> (i+2, j); b4 -> (i, j+2); own2 -> self; own3 -> (i-1, j);
    # b3 -> (i, j-2); b1 -> (i-2, j)
    di = torch.tensor([0, 2, 0, 0, -1, 0, -2], device=mask.device)[s]
    dj = torch.tensor([0, 0, 2, 0, 0, -2, 0], device=mask.device)[s]
    kind = torch.tensor([0, 0, 0, 1, 2, 0, 0], device=mask.device, dtype=torch.uint8)[s]
    si = (ti + di).clamp_(0, X - 1)
    sj = (tj + dj).clamp_(0, Y - 1)
    # hazard: a source that is itself a target needs the 2-phase gather (always used) -- fine; but a
    # target that is NOT a wall/inflow/outflow cell means walls thinner than the reference assumes
    thin = bool(((mask[ti, tj] == 0) & (kind == 0)).any()) if ti.numel() else False
    return {"tgt": _window_index(ti, tj, w0, w1, Y, "vel tgt"), "src": _window_index(si, sj, w0, w1, Y, "vel src"),
            "kind": kind.contiguous(), "n": int(ti.numel()), "thin_walls": thin}


def pressure_table(pcode: torch.Tensor, t0: int | None = None, t1: int | None = None, w0: int = 0,
                   w1: int | None = None) -> dict:
    """Gather list for set_pressure_boundary_condition from `pcode` (global), local indices.

    kind 0: p[tgt] = p[src0]; kind 1: p[tgt] = (p[src0] + p[src1]) / 2; kind 2: p[tgt] = 0.
    """
    X, Y = pcode.shape
    t0, t1 = (0 if t0 is None else t0), (X if t1 is None else t1)
    w1 = X if w1 is None else w1
    active = (pcode != PC_FLUID) & (pcode != PC_W_NONE)
    rows = torch.zeros_like(active)
    rows[t0:t1] = True
    ti, tj = torch.nonzero(active & rows, as_tuple=True)
    c = pcode[ti, tj].to(torch.int64)
    dev = pcode.device
    #                 code: 0  1   2  3  4   5  6   7  8  9 10 11
    d0i = torch.tensor([0, -1, 1, 0, 0, -1, 1, -1, 1, 0, 1, 0], device=dev)[c]
    d0j = torch.tensor([0, 0, 0, -1, 1, 0, 0, 0, 0, 0, 0, 0], device=dev)[c]
    d1i = torch.tensor([0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0], device=dev)[c]
    d1j = torch.tensor([0, 0, 0, 0, 0, 1, 1, -1, -1, 0, 0, 0], device=dev)[c]
    kind = torch.tensor([0, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 2], device=dev, dtype=torch.uint8)[c]
    s0i, s0j = (ti + d0i).clamp_(0, X - 1), (tj + d0j).clamp_(0, Y - 1)
    s1i, s1j = (ti + d1i).clamp_(0, X - 1), (tj + d1j).clamp_(0, Y - 1)
    # outflow entries have no source: point them at the target itself so indices stay in-window
    zero = kind == 2
    s0i, s0j = torch.where(zero, ti, s0i), torch.where(zero, tj, s0j)
    s1i, s1j = torch.where(kind != 1, s0i, s1i), torch.where(kind != 1, s0j, s1j)
    tgt = _window_index(ti, tj, w0, w1, Y, "p tgt")
    src0 = _window_index(s0i, s0j, w0, w1, Y, "p src0")
    src1 = _window_index(s1i, s1j, w0, w1, Y, "p src1")
    kind = kind.contiguous()
    # "feed" sub-table: wall-BC cells (codes 1..8) sitting at (i+1, j) of an inflow cell.  Their STORED
    # value is read raw by that inflow cell's BC (p = p(i+1,j), :62-63), so the inline-BC sweep path must
    # keep exactly these cells materialised (see fs2d_jacobi_update).
    up = _shift(pcode, -1, 0)  # code of (i-1, j)
    feeds = (c >= PC_W_IM) & (c <= PC_W_IP_JM) & (up[ti, tj] == PC_INFLOW) & (ti > 0)
    sel = torch.nonzero(feeds, as_tuple=True)[0]
    feed = {"tgt": tgt[sel].contiguous(), "src0": src0[sel].contiguous(), "src1": src1[sel].contiguous(),
            "kind": kind[sel].contiguous(), "n": int(sel.numel())}
    return {"tgt": tgt, "src0": src0, "src1": src1, "kind": kind, "n": int(ti.numel()), "feed": feed}


def exposed_stale_cells(pcode: torch.Tensor) -> torch.Tensor:
    """Linear indices of wall cells that take no BC branch (never written by any kernel)
    yet are read by a relaxed (not-wall) 4-neighbour."""
    notwall = (pcode == PC_FLUID) | (pcode == PC_INFLOW) | (pcode == PC_OUTFLOW)
    near = _shift(notwall, -1, 0) | _shift(notwall, 1, 0) | _shift(notwall, 0, -1) | _shift(notwall, 0, 1)
    return torch.nonzero(((pcode == PC_W_NONE) & near).flatten(), as_tuple=True)[0]


# --------------------------------------------------------------------------------------------------
# validity of the fused multi-iteration Jacobi kernel for a given mask (fs2d_jacobi_fused)
# --------------------------------------------------------------------------------------------------
_SRC_OFFSETS = {  # BC code of a neighbour -> offsets (relative to that neighbour) of the cells its value reads
    PC_W_IM: ((-1, 0),), PC_W_IP: ((1, 0),), PC_W_JM: ((0, -1),), PC_W_JP: ((0, 1),),
    PC_W_IM_JP: ((-1, 0), (0, 1)), PC_W_IP_JP: ((1, 0), (0, 1)), PC_W_IM_JM: ((-1, 0), (0, -1)),
    PC_W_IP_JM: ((1, 0), (0, -1)), PC_INFLOW: ((1, 0),), PC_OUTFLOW: (),
}


def iteration_dependencies(code: torch.Tensor) -> dict:
    """{(di, dj): bool mask}: relaxed cell c reads cell c + (di, dj) during one {BC, sweep} iteration."""
    relaxed = (code == PC_FLUID) | (code == PC_INFLOW) | (code == PC_OUTFLOW)
    deps: dict = {}

    def add(off, m):
        if off in deps:
            deps[off] |= m
        else:
            deps[off] = m.clone()

    for n in ((1, 0), (-1, 0), (0, 1), (0, -1)):
        cn = _shift(code, *n)
        add(n, relaxed & ((cn == PC_FLUID) | (cn == PC_W_NONE)))
        for k, offs in _SRC_OFFSETS.items():
            m = relaxed & (cn == k)
            if bool(m.any()):
                for o in offs:
                    add((n[0] + o[0], n[1] + o[1]), m)
    return deps


def fused_reach_ok(code: torch.Tensor, T: int, tile_rows: int, tile_cols: int, halo_rows: int | None = None,
                   halo_cols: int | None = None, row0: int = 0, row1: int | None = None,
                   fresh_below: int | None = None) -> bool:
    """True if, for the tiling used by fs2d_jacobi_fused (tiles of tile_rows x tile_cols loaded cells, of which
    halo_rows / halo_cols are discarded on each side, output tiles anchored at (row0, 0)), the value of every
    OUTPUT cell after T iterations depends only on cells inside its tile.  Dynamic programme over the T
    iterations of how far up/down/left/right each cell's dependency cone reaches (conservative at global edges).

    fresh_below (row strips with a neighbour below row1): only that many rows beyond row1 - 1 hold current data (the
    halo rows exchanged before the pass).  The tiles are anchored at row0, so the LAST tile row of a strip generally
    extends further down than that; its extra rows are stale, and a cone that grows by more than one row per iteration
    (a relaxed cell above an inflow cell reads, through that cell's BC value p(i+1, j), two rows down) must not reach
    them.  The upward side needs no such bound: the first tile starts exactly halo_rows above row0, so the tile bound
    is the fresh-row bound there.  The reference's scenes keep inflow cells in the first rows of the grid and never
    trip this; adversarial masks do (tests/test_host_logic.py, tests/test_bc_tables.py)."""
    X, Y = code.shape
    row1 = X if row1 is None else row1
    HI = T if halo_rows is None else halo_rows
    HJ = T if halo_cols is None else halo_cols
    TI, TJ = tile_rows - 2 * HI, tile_cols - 2 * HJ
    if TI <= 0 or TJ <= 0:
        return False
    deps = iteration_dependencies(code)
    dev = code.device
    ii = torch.arange(X, device=dev)
    lr = (HI + ((ii - row0) % TI)).to(torch.int16)[:, None]         # tile row of each cell as an output cell
    lc = (HJ + (torch.arange(Y, device=dev) % TJ)).to(torch.int16)[None, :]
    owned = ((ii >= row0) & (ii < row1))[:, None]
    weights = {"up": lambda o: -o[0], "down": lambda o: o[0], "left": lambda o: -o[1], "right": lambda o: o[1]}
    room = {"up": lr, "down": (tile_rows - 1) - lr, "left": lc, "right": (tile_cols - 1) - lc}
    if fresh_below is not None:
        fresh = ((row1 - 1 - ii) + fresh_below).clamp(min=0, max=32767)      # int16 like the other bounds; only the small values matter
        room["down"] = torch.minimum(room["down"], fresh.to(torch.int16)[:, None])
    for name, w in weights.items():
        R = torch.zeros((X, Y), dtype=torch.int16, device=dev)
        for _ in range(T):
            new = torch.zeros_like(R)
            for off, m in deps.items():
                cand = _shift(R, *off) + w(off)
                new = torch.maximum(new, torch.where(m, cand, torch.zeros_like(cand)))
            R = new
        if bool(((R > room[name]) & owned).any()):
            return False
    return True

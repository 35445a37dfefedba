"""BoundaryCondition + scene builders, API of /root/reference/fs/boundary_condition.py.

Device side: `set_velocity_boundary_condition` (:16-39) and `set_pressure_boundary_condition`
(:41-65) run as sparse gather kernels of libfs2d.so over tables resolved once from the static mask
(`fs/_bc_tables.py`).  Host side: the NumPy scene builders bc1..bc6 (:115-524) are restated as a
small declarative scene description (same operations, same order, same rounding) and verified
against masks produced by the reference's own builders (tests/golden/masks_*.{npz,json}).
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import numpy.typing as npt
import torch

from fs import _bc_tables, _lib
from fs.double_buffer import Field, default_device

FLUID, WALL, INFLOW, OUTFLOW = 0, 1, 2, 3


class BoundaryCondition:
    """bc_const: (X, Y, 2) f32 inflow velocities; bc_mask: (X, Y) u8 cell types (:13, :78-85)."""

    #: rows of scene a rank needs beyond its owned rows when it builds only its strip (`row_offset`): the tables and the
    #: fused-pass validity analysis look at most 2 * 12 + 2 rows past the owned rows (a dependency cone grows by at most two
    #: rows per iteration, 12 iterations per pass at most)
    STRIP_MARGIN = 32
    #: halo rows on which the in-place boundary conditions are applied (and v is exchanged for them)
    BC_HALO = 9
    #: scenes with fewer cells than this build their tables on the host
    HOST_TABLES_BELOW = 1 << 21

    def __init__(self, bc_const: npt.NDArray, bc_mask: npt.NDArray, device=None, partition=None, row_offset: int | None = None) -> None:
        """row_offset (not in the reference; needs `partition`): the arrays hold only the global rows [row_offset, row_offset +
        len) of the scene -- at least the rank's owned rows +- STRIP_MARGIN, clipped to the grid (build_scene(..., rows=...),
        strip_rows()) -- instead of the whole grid; every table comes out identical to a build from the global arrays."""
        bc_const = np.ascontiguousarray(bc_const, dtype=np.float32)
        bc_mask = np.ascontiguousarray(bc_mask, dtype=np.uint8)
        if bc_const.shape[:2] != bc_mask.shape or bc_const.shape[2:] != (2,):
            raise ValueError(f"bc_const {bc_const.shape} / bc_mask {bc_mask.shape} shape mismatch")
        self.device = torch.device(device) if device is not None else default_device()
        if self.device.type == "cuda" and self.device.index not in (None, torch.cuda.current_device()):
            # kernels are launched on the current device's stream (fs/_lib.py:stream): pointers of another device would fault
            raise ValueError(f"BoundaryCondition(device={self.device}) while cuda:{torch.cuda.current_device()} is current: "
                             "call torch.cuda.set_device() first (one process per GPU)")
        XA, Y = int(bc_mask.shape[0]), int(bc_mask.shape[1])      # rows the arrays hold
        if row_offset is None:
            A0, X = 0, XA
        else:
            if partition is None:
                raise ValueError("row_offset needs the partition it belongs to")
            A0, X = int(row_offset), partition.x_global
        self._global_resolution = (X, Y)
        if partition is None:
            from fs.distributed import Partition

            partition = Partition.single(X)
        if partition.x_global != X:
            raise ValueError(f"partition of {partition.x_global} rows for a scene of {X} rows")
        self.partition = partition
        w0, w1 = partition.window()          # global rows held in the local array (owned + halo)
        g0, g1 = partition.owned()           # global rows this rank updates
        if row_offset is not None:
            need = self.strip_rows(partition)
            if A0 > need[0] or A0 + XA < need[1]:
                raise ValueError(f"arrays hold the rows [{A0}, {A0 + XA}) but rank {partition.rank} needs {need} (owned rows +- STRIP_MARGIN)")
        self._resolution = (g1 - g0, Y)
        self.halo = partition.halo
        self._row_offset = A0

        # from here on rows are indexed in ARRAY coordinates (global row - A0); array edges that are not grid edges lie in
        # the margin, where the clamped neighbour reads of the table builders produce values nobody uses.
        # The tables are torch code (fs/_bc_tables.py): hundreds of small element-wise kernels.  Small grids build them on
        # the HOST and upload the results (a 128 x 64 scene issued > 900 library-less launches before its first fs2d kernel);
        # large grids build them on the device (seconds instead of minutes at 8192^2).
        build_dev = self.device if XA * Y >= self.HOST_TABLES_BELOW else torch.device("cpu")
        gmask = torch.from_numpy(bc_mask).to(build_dev)
        pcode = _bc_tables.pressure_codes(gmask)
        lo, hi = max(w0, 0), min(w1, X)      # part of the window that exists globally

        def local(t_array: torch.Tensor, fill=0) -> torch.Tensor:
            out = torch.full((w1 - w0,) + tuple(t_array.shape[1:]), fill, dtype=t_array.dtype, device=build_dev)
            out[lo - w0:hi - w0] = t_array[lo - A0:hi - A0]
            return out.contiguous().to(self.device)

        self._bc_mask = local(gmask, WALL)
        self._pcode = local(_bc_tables.pack_pcode(pcode), _bc_tables.PC_W_NONE)
        self._pcode_global = pcode           # unpacked codes of the array's rows (fused-kernel validity analysis)
        self._fused_ok: dict[int, bool] = {}
        self._fused_orders: dict[tuple, tuple] = {}
        self._bc_const = self._upload_rows(bc_const, lo, hi, w0, w1, A0)
        # BC targets: owned rows plus the halo rows whose sources are inside the window -- of the first BC_HALO halo rows: the
        # stencil kernels read the post-BC fields at most a few rows beyond the owned ones, however deep the halo is (deep
        # halos only serve the fused Jacobi passes, fs/halo.py)
        self.bc_halo = min(self.halo, self.BC_HALO)
        tl, th = max(lo, g0 - max(self.bc_halo - 2, 0)), min(hi, g1 + max(self.bc_halo - 2, 0))
        def to_dev(table: dict) -> dict:
            return {k: (v.to(self.device) if isinstance(v, torch.Tensor) else to_dev(v) if isinstance(v, dict) else v)
                    for k, v in table.items()}

        self._vel_table = to_dev(_bc_tables.velocity_table(gmask, tl - A0, th - A0, w0 - A0, w1 - A0))
        self._p_table = to_dev(_bc_tables.pressure_table(pcode, max(lo, g0 - max(self.bc_halo - 1, 0)) - A0,
                                                         min(hi, g1 + max(self.bc_halo - 1, 0)) - A0, w0 - A0, w1 - A0))
        stale = _bc_tables.exposed_stale_cells(pcode)
        if stale.numel():
            si, sj = stale // Y, stale % Y
            keep = (si >= lo - A0) & (si < hi - A0)
            stale = (si[keep] - (w0 - A0)) * Y + sj[keep]
        self._exposed_stale = stale.to(self.device)
        n = max(self._vel_table["n"], self._p_table["n"], 1)
        self._scratch = torch.empty(2 * n, dtype=torch.float32, device=self.device)
        self.dom = _lib.Dom(rows=w1 - w0, Y=Y, r0=g0 - w0, r1=g1 - w0, clo=lo - w0, chi=hi - 1 - w0, gi0=w0)
        del gmask

    @classmethod
    def strip_rows(cls, partition) -> tuple[int, int]:
        """global rows [a, b) of the scene a rank must hold to build its BoundaryCondition with `row_offset=a`"""
        g0, g1 = partition.owned()
        m = max(cls.STRIP_MARGIN, partition.halo + 28)   # deep halos: windows extended by up to halo - 1 rows, cones 2 * 12 + 2 beyond
        return max(0, g0 - m), min(partition.x_global, g1 + m)

    def _upload_rows(self, host: npt.NDArray, lo: int, hi: int, w0: int, w1: int, A0: int, fill=0) -> torch.Tensor:
        """local (window-shaped) device copy of a per-cell host array that holds the global rows from A0 on: only the window's
        rows cross PCIe"""
        out = torch.full((w1 - w0,) + tuple(host.shape[1:]), fill, dtype=torch.from_numpy(host[:0]).dtype, device=self.device)
        out[lo - w0:hi - w0] = torch.from_numpy(np.ascontiguousarray(host[lo - A0:hi - A0])).to(self.device)
        return out.contiguous()

    # -- reference API ---------------------------------------------------------------------
    def set_velocity_boundary_condition(self, vc: Field) -> None:
        t = self._vel_table
        _lib.call("fs2d_vel_bc", vc.ptr(), _lib.ptr(self._bc_const), _lib.ptr(t["tgt"]), _lib.ptr(t["src"]),
                  _lib.ptr(t["kind"]), _lib.ptr(self._scratch), t["n"], _lib.stream())

    def set_pressure_boundary_condition(self, pc: Field) -> None:
        t = self._p_table
        _lib.call("fs2d_pressure_bc", pc.ptr(), _lib.ptr(t["tgt"]), _lib.ptr(t["src0"]), _lib.ptr(t["src1"]),
                  _lib.ptr(t["kind"]), _lib.ptr(self._scratch), t["n"], _lib.stream())

    def apply_feed_bc(self, pc: Field) -> None:
        """In-place BC of the few wall cells whose stored value an inflow cell reads raw (see
        fs2d_jacobi_update in include/fs2d.h)."""
        t = self._p_table["feed"]
        _lib.call("fs2d_pressure_bc", pc.ptr(), _lib.ptr(t["tgt"]), _lib.ptr(t["src0"]), _lib.ptr(t["src1"]),
                  _lib.ptr(t["kind"]), _lib.ptr(self._scratch), t["n"], _lib.stream())

    def is_wall(self, i: int, j: int) -> bool:
        return int(self._bc_mask[i + self.halo, j]) == WALL

    def is_fluid_domain(self, i: int, j: int) -> bool:
        return int(self._bc_mask[i + self.halo, j]) == FLUID

    def get_resolution(self) -> tuple[int, int]:
        return self._resolution

    def get_global_resolution(self) -> tuple[int, int]:
        return self._global_resolution

    @staticmethod
    def to_field(bc: npt.NDArray, bc_mask: npt.NDArray, device=None) -> tuple[torch.Tensor, torch.Tensor]:
        dev = torch.device(device) if device is not None else default_device()
        return (torch.from_numpy(np.ascontiguousarray(bc, dtype=np.float32)).to(dev),
                torch.from_numpy(np.ascontiguousarray(bc_mask, dtype=np.uint8)).to(dev))

    # -- helpers for the other operators -----------------------------------------------------
    def fused_ok(self, T: int, ext: int = 0) -> bool:
        """May fs2d_jacobi_fused run T iterations per pass on this mask?  (static analysis, cached)
        ext (row strips): the pass updates the owned rows extended by `ext` rows on each side that has a neighbour -- a
        strip that exchanges the halos of several passes at once recomputes the rows the later passes of the group read
        (fs/halo.py: deep halos); its tiling is anchored at the extended window, so it is validated separately."""
        import ctypes

        rows, cols, hr, hc, tmax = (ctypes.c_int() for _ in range(5))
        _lib.call("fs2d_fused_tile", T, ctypes.byref(rows), ctypes.byref(cols), ctypes.byref(hr), ctypes.byref(hc),
                  ctypes.byref(tmax))
        key = (T, rows.value, cols.value, ext)   # the tile depends on the kernel (fs2d_fused_tile)
        if key not in self._fused_ok:
            X = self._global_resolution[0]
            g0, g1 = self.partition.owned()
            e0, e1 = max(g0 - ext, 0), min(g1 + ext, X)            # the extended window, clipped to the grid
            below = T if e1 < X else None                           # rows beyond e1 + T hold stale data (there is a neighbour below)
            ok = (1 <= T <= tmax.value and self._global_resolution[1] % 16 == 0 and self._p_table["feed"]["n"] == 0
                  and (self.partition.world == 1 or self.halo >= T + ext + 1)
                  and _bc_tables.fused_reach_ok(self._pcode_global, T, rows.value, cols.value, hr.value, hc.value,
                                                e0 - self._row_offset, e1 - self._row_offset, fresh_below=below))
            self._fused_ok[key] = bool(ok)
        return self._fused_ok[key]

    def fused_order(self, T: int, dom=None, skip: tuple[int, int] | None = None) -> tuple[torch.Tensor, int]:
        """Tile list of one fused-pass geometry (fs2d_fused_order): (device int32 tensor, entries).  The class of a tile
        -- open fluid / BC cells or edges / nothing to store -- depends on pcode and the geometry only, so it is built once
        per (T, row window, skipped tile rows) and handed to every pass of that shape."""
        import ctypes

        d = dom or self.dom
        first, n = skip if skip is not None else (0, 0)
        key = (T, d.r0, d.r1, d.clo, d.chi, first, n)
        if key not in self._fused_orders:
            rows, cols, hr, hc, tmax = (ctypes.c_int() for _ in range(5))
            _lib.call("fs2d_fused_tile", T, ctypes.byref(rows), ctypes.byref(cols), ctypes.byref(hr), ctypes.byref(hc),
                      ctypes.byref(tmax))
            ti, tj = rows.value - 2 * hr.value, cols.value - 2 * hc.value
            cap = max(1, -(-(d.r1 - d.r0) // ti) * -(-d.Y // tj))
            order = torch.empty(cap, dtype=torch.int32, device=self.device)
            counts = (ctypes.c_int * 3)()
            _lib.call("fs2d_fused_order", _lib.ptr(self._pcode), d, T, first, n, _lib.ptr(order), cap, counts, _lib.stream())
            self._fused_orders[key] = (order, int(counts[0]), int(counts[1]), int(counts[2]))
        return self._fused_orders[key][:2]

    def stale_cells_agree(self, a: Field, b: Field) -> bool:
        """True if the never-written wall cells read by relaxed neighbours hold equal values in both
        physical buffers (always the case for solver-owned buffers)."""
        if self._exposed_stale.numel() == 0:
            return True
        fa, fb = a.tensor.flatten()[self._exposed_stale], b.tensor.flatten()[self._exposed_stale]
        return bool(torch.equal(fa, fb))


class DyeBoundaryCondition(BoundaryCondition):
    """(:88-112) adds the inflow dye colours bc_dye (X, Y, 3) f32 and set_dye_boundary_condition."""

    def __init__(self, bc_const: npt.NDArray, bc_dye: npt.NDArray, bc_mask: npt.NDArray, device=None, partition=None,
                 row_offset: int | None = None) -> None:
        super().__init__(bc_const, bc_mask, device=device, partition=partition, row_offset=row_offset)
        bc_dye = np.ascontiguousarray(bc_dye, dtype=np.float32)
        if bc_dye.shape != tuple(np.shape(bc_mask)) + (3,):
            raise ValueError(f"bc_dye {bc_dye.shape} does not match the mask {np.shape(bc_mask)}")
        X, Y = self._global_resolution
        w0, w1 = self.partition.window()
        lo, hi = max(w0, 0), min(w1, X)
        self._bc_dye = self._upload_rows(bc_dye, lo, hi, w0, w1, self._row_offset)
        self._dye_tgt = torch.nonzero((self._bc_mask == INFLOW).flatten(), as_tuple=True)[0].to(torch.int32).contiguous()

    def set_dye_boundary_condition(self, dye: Field) -> None:
        _lib.call("fs2d_dye_bc", dye.ptr(), _lib.ptr(self._bc_dye), _lib.ptr(self._dye_tgt), int(self._dye_tgt.numel()),
                  _lib.stream())


# =================================================================================================
# scene builders (host, NumPy)
# =================================================================================================
def create_bc_array(x_resolution: int, y_resolution: int) -> tuple[npt.NDArray, npt.NDArray, npt.NDArray]:
    return (np.zeros((x_resolution, y_resolution, 2), dtype=np.float32),
            np.zeros((x_resolution, y_resolution), dtype=np.uint8),
            np.zeros((x_resolution, y_resolution, 3), dtype=np.float32))


class _Win:
    """Row window [a, b) of a scene that is X rows tall: the builders below speak GLOBAL row indices (NumPy slice semantics,
    negative values included) and this maps them onto arrays that hold only the window's rows.  A rank of a row-strip
    decomposition builds just its strip (+ a margin) instead of the global grid (21 B/cell of host memory)."""

    def __init__(self, X: int, a: int = 0, b: int | None = None) -> None:
        self.X, self.a, self.b = int(X), int(a), int(X if b is None else b)
        if not 0 <= self.a <= self.b <= self.X:
            raise ValueError(f"row window [{a}, {b}) outside the {X} rows of the scene")

    def rows(self, sl) -> slice:
        """local slice of the window rows selected by the global int / slice `sl` (empty if they miss the window)"""
        if isinstance(sl, (int, np.integer)):
            i = int(sl) % self.X
            sl = slice(i, i + 1)
        lo, hi, step = sl.indices(self.X)
        assert step == 1
        lo, hi = max(lo, self.a), min(hi, self.b)
        return slice(lo - self.a, max(hi, lo) - self.a)


def set_plane(bc, bc_mask, bc_dye, lower_left, upper_right, win: _Win | None = None) -> None:
    """Axis-aligned wall slab (:157-168); NumPy slice semantics incl. negative corners."""
    win = win or _Win(bc.shape[0])
    sl = (win.rows(slice(int(lower_left[0]), int(upper_right[0]))), slice(int(lower_left[1]), int(upper_right[1])))
    bc[sl] = 0.0
    bc_mask[sl] = WALL
    if bc_dye is not None:
        bc_dye[sl] = 0.0


def set_circle(bc, bc_mask, bc_dye, center, radius: float, win: _Win | None = None) -> None:
    """Wall disc (:137-154): cell (i, j) is wall iff |(i, j) + 0.5 - center| < radius, scanned over the
    reference's rounded bounding box.  Vectorised; same float64 arithmetic per cell."""
    win = win or _Win(bc.shape[0])
    c = np.asarray(center, dtype=np.float64)
    lo = np.round(np.maximum(c - radius, 0)).astype(np.int32)
    u0 = round(min(center[0] + radius, win.X))
    u1 = round(min(center[1] + radius, bc.shape[1]))
    r0, r1 = max(int(lo[0]), win.a), min(int(u0), win.b)      # the bounding box's rows that fall into the window
    if r1 <= r0 or u1 <= lo[1]:
        return
    x = np.arange(r0, r1, dtype=np.float64)[:, None] + 0.5 - c[0]
    y = np.arange(lo[1], u1, dtype=np.float64)[None, :] + 0.5 - c[1]
    inside = np.sqrt(x * x + y * y) < radius
    sub = (slice(r0 - win.a, r1 - win.a), slice(int(lo[1]), int(u1)))
    bc[sub][inside] = 0.0
    bc_mask[sub][inside] = WALL
    if bc_dye is not None:
        bc_dye[sub][inside] = 0.0


def set_obstacle_fromfile(bc, bc_mask, bc_dye, filepath: Path, win: _Win | None = None) -> None:
    """Dark pixels of an image become wall (:171-198)."""
    from PIL import Image

    win = win or _Win(bc.shape[0])
    image = Image.open(filepath).convert("L")
    x_res, y_res = win.X, bc.shape[1]
    x_ratio, y_ratio = x_res / image.width, y_res / image.height
    size = (x_res, round(image.height * x_ratio)) if x_ratio < y_ratio else (round(image.width * y_ratio), y_res)
    image = image.resize(size)
    canvas = Image.new(image.mode, (x_res, y_res), 255)
    canvas.paste(image, ((x_res - image.width) // 2, 0))
    dark = (np.flip(np.array(canvas).T, axis=1) < 200)[win.a:win.b]
    bc[dark] = 0.0
    bc_mask[dark] = WALL
    if bc_dye is not None:
        bc_dye[dark] = 0.0


def create_color_map(color_list: list[npt.NDArray], n_samples: int) -> npt.NDArray:
    """(:125-134) piecewise-linear colour ramp through `color_list`, n_samples x 3 (float64)."""
    color_arr = np.vstack(color_list)
    x = np.linspace(0.0, 1.0, color_arr.shape[0], endpoint=True)
    x_ = np.linspace(0.0, 1.0, n_samples, endpoint=True)
    return np.vstack([np.interp(x_, x, color_arr[:, k]) for k in range(3)]).T


_YEL, _BLU, _RED, _CYA = (np.array(c) for c in ([1.1, 1.1, 0.2], [0.2, 0.2, 1.1], [1.1, 0.2, 0.2], [0.2, 1.1, 1.1]))
_RAMP = [_CYA, _RED, _BLU, _YEL]


def _inflow(bc, mask, rows, cols, win: _Win) -> None:
    bc[win.rows(rows), cols] = np.array([1.0, 0.0], dtype=np.float32)
    mask[win.rows(rows), cols] = INFLOW


def _outflow(bc, mask, rows, cols, win: _Win) -> None:
    bc[win.rows(rows), cols] = 0.0
    mask[win.rows(rows), cols] = OUTFLOW


def _floor_ceiling(bc, mask, X, Y, dye, win: _Win) -> None:
    set_plane(bc, mask, dye, (0, 0), (X, 2), win)
    set_plane(bc, mask, dye, (0, Y - 2), (X, Y), win)


ALL = slice(None)


def build_scene(num: int, x_res: int, y_res: int, obstacle_image: Path | None = None, with_dye: bool = False,
                rows: tuple[int, int] | None = None):
    """(bc_const, bc_mask) of scene `num` on an x_res x y_res grid.  `create_boundary_conditionN`
    of the reference is build_scene(N, 2*res, res) (:226, 272, 326, 376, 425, 486); other aspect
    ratios are used for the weak-scaling grids.  rows = (a, b): only the global rows [a, b) of the scene
    (arrays of b - a rows, identical to the corresponding rows of the full build)."""
    X, Y = int(x_res), int(y_res)
    win = _Win(X, *(rows or (0, X)))
    bc, mask, dye = create_bc_array(win.b - win.a, Y)
    if not with_dye:
        dye = None
    first2 = win.rows(slice(0, 2))     # the two inflow rows, as far as the window holds them

    def plane(lower_left, upper_right) -> None:
        set_plane(bc, mask, dye, lower_left, upper_right, win)

    if num == 1:                                                   # :222-265
        _inflow(bc, mask, slice(0, 2), ALL, win)
        if with_dye:
            dye[first2, :] = create_color_map(_RAMP * 3, Y)[None]    # the same colour ramp on both inflow rows (:239)
        _outflow(bc, mask, -1, ALL, win)
        _floor_ceiling(bc, mask, X, Y, dye, win)
        set_circle(bc, mask, dye, (X // 4, Y // 2), Y // 18, win)
    elif num == 2:                                                 # :268-319
        _inflow(bc, mask, slice(0, 2), ALL, win)
        if with_dye:
            dye[first2, :] = np.array([0.2, 0.2, 1.2])
            width = Y // 10
            for i in range(0, Y, width):
                dye[first2, i:i + width // 2] = np.array([1.2, 1.2, 0.2])
        plane((0, 0), (2, Y // 3))
        plane((0, 2 * Y // 3), (2, Y))
        plane((X - 2, 0), (X, Y))
        _floor_ceiling(bc, mask, X, Y, dye, win)
        xp, yp, size = X // 5, Y // 2, Y // 32
        plane((xp - size, yp), (xp + size, Y))
        plane((2 * xp - size, 0), (2 * xp + size, yp))
        plane((3 * xp - size, yp), (3 * xp + size, Y))
        plane((4 * xp - size, 0), (4 * xp + size, yp))
        yq = Y // 3
        _outflow(bc, mask, slice(-2, None), slice(yq, 2 * yq), win)
    elif num == 3:                                                 # :322-369
        _inflow(bc, mask, slice(0, 2), ALL, win)
        if with_dye:
            dye[first2, :] = create_color_map(_RAMP, Y)[None]
        _outflow(bc, mask, -1, ALL, win)
        _floor_ceiling(bc, mask, X, Y, dye, win)
        np.random.seed(123)  # noqa: NPY002  (legacy RNG on purpose: same stream as the reference)
        points = np.random.uniform(0, X, (100, 2))  # noqa: NPY002
        points = points[points[:, 1] < Y]
        radius = 16 * (Y / 500)
        for p in points:
            set_circle(bc, mask, dye, p, radius, win)
    elif num == 4:                                                 # :372-418
        plane((0, 0), (2, Y))
        plane((X - 2, 0), (X, Y))
        _floor_ceiling(bc, mask, X, Y, dye, win)
        if with_dye:
            ramp = create_color_map(_RAMP, Y // 4 - 2)[None]
            dye[first2, 3 * Y // 4:-2] = ramp
            dye[first2, 2:Y // 4] = ramp
        _inflow(bc, mask, slice(0, 2), slice(3 * Y // 4, -2), win)
        _inflow(bc, mask, slice(0, 2), slice(2, Y // 4), win)
        _outflow(bc, mask, slice(-2, None), slice(3 * Y // 8, 5 * Y // 8), win)
    elif num == 5:                                                 # :421-479
        _inflow(bc, mask, slice(0, 2), slice(2, Y // 3), win)
        _inflow(bc, mask, slice(0, 2), slice(2 * Y // 3, Y - 2), win)
        if with_dye:
            dye[first2, 2:Y // 3] = np.array([1.2, 0.2, 0.2])
            dye[first2, 2 * Y // 3:Y - 2] = np.array([0.2, 1.2, 1.2])
        _outflow(bc, mask, slice(-2, None), ALL, win)
        _floor_ceiling(bc, mask, X, Y, dye, win)
        size = X // 64
        plane((0, Y // 5), (11 * X // 30, 4 * Y // 5))
        plane((X // 2 - size, 0), (X // 2 + size, 2 * Y // 5))
        plane((X // 2 - size, 3 * Y // 5), (X // 2 + size, Y))
        yp, half = Y // 6, np.array([Y, Y]) // 25
        for a, b in zip((7, 8, 9, 10, 11), (0, 1, 0, 1, 0), strict=True):
            for i in range(1, 6 + b):
                p = np.array([a * X // 12, i * yp - b * Y // 12])
                plane(p - half, p + half)
    elif num == 6:                                                 # :482-524
        _inflow(bc, mask, slice(0, 2), ALL, win)
        if with_dye:
            dye[first2, :] = create_color_map(_RAMP, Y)[None]
        _outflow(bc, mask, -1, ALL, win)
        _floor_ceiling(bc, mask, X, Y, dye, win)
        path = obstacle_image or Path(__file__).resolve().parents[1] / "images" / "bc_mask" / "dragon.png"
        if not Path(path).exists():
            raise FileNotFoundError(f"scene 6 needs the reference's obstacle image (images/bc_mask/dragon.png); "
                                    f"not found at {path}")
        set_obstacle_fromfile(bc, mask, dye, Path(path), win)
    else:
        raise NotImplementedError
    return (bc, mask, dye) if with_dye else (bc, mask)


def scene_row_cost(num: int, x_res: int, y_res: int, n_sweeps: int, tile=(80, 112, 8, 8), tile_rows_per_chunk: int = 32) -> npt.NDArray:
    """Estimated cost (microseconds on a B200) of one time step per ROW of scene `num` with n_sweeps fused Jacobi iterations:
    the weight fs.distributed.balanced_bounds turns into strips of equal work.  The stencil kernels stream every cell
    (2.0 ms per 67 M cells); the fused Jacobi passes work by tiles of `tile` = (output rows, output columns, row halo, column
    halo): a tile inside a wall is skipped, an open-fluid tile costs one unit, a tile with BC cells about 2.2 (5.3 ns per unit
    and sweep: 323 us for the 7590 units of a T = 8 pass at bc2 8192^2, profiles/r02_*).  Built chunk by chunk (O(chunk)
    memory), identical on every rank."""
    X, Y = int(x_res), int(y_res)
    ti, tj, hi, hj = tile
    out = np.empty(X, dtype=np.float64)
    step = ti * tile_rows_per_chunk
    for a in range(0, X, step):
        b = min(a + step, X)
        lo, hi_row = max(a - hi, 0), min(b + hi, X)
        _, m = build_scene(num, X, Y, rows=(lo, hi_row))
        wall, fluid = (m == WALL), (m == FLUID)
        for r in range(a, b, ti):
            r0, r1 = max(r - hi, lo) - lo, min(r + ti + hi, hi_row) - lo       # loaded rows of this tile row, chunk coordinates
            units = 0.0
            for c in range(0, Y, tj):
                c0, c1 = max(c - hj, 0), min(c + tj + hj, Y)
                if wall[r0:r1, c0:c1].all():
                    continue                                   # nothing to relax: the pass drops the tile
                inside = r - hi >= 0 and r + ti + hi <= X and c - hj >= 0 and c + tj + hj <= Y
                units += 1.0 if inside and fluid[r0:r1, c0:c1].all() else 2.2
            rows = min(r + ti, b) - r
            out[r:r + rows] = (5.3e-3 * n_sweeps * units + 2.98e-5 * ti * Y) / ti
    return out


def _make(num: int, resolution: int, enable_dye: bool, obstacle_image=None, **kw) -> BoundaryCondition:
    if enable_dye:
        bc, mask, dye = build_scene(num, 2 * resolution, resolution, obstacle_image=obstacle_image, with_dye=True)
        return DyeBoundaryCondition(bc, dye, mask, **kw)
    bc, mask = build_scene(num, 2 * resolution, resolution, obstacle_image=obstacle_image)
    return BoundaryCondition(bc, mask, **kw)


def create_boundary_condition1(resolution: int, *, enable_dye: bool, **kw) -> BoundaryCondition:
    return _make(1, resolution, enable_dye, **kw)


def create_boundary_condition2(resolution: int, *, enable_dye: bool, **kw) -> BoundaryCondition:
    return _make(2, resolution, enable_dye, **kw)


def create_boundary_condition3(resolution: int, *, enable_dye: bool, **kw) -> BoundaryCondition:
    return _make(3, resolution, enable_dye, **kw)


def create_boundary_condition4(resolution: int, *, enable_dye: bool, **kw) -> BoundaryCondition:
    return _make(4, resolution, enable_dye, **kw)


def create_boundary_condition5(resolution: int, *, enable_dye: bool, **kw) -> BoundaryCondition:
    return _make(5, resolution, enable_dye, **kw)


def create_boundary_condition6(resolution: int, *, enable_dye: bool, **kw) -> BoundaryCondition:
    return _make(6, resolution, enable_dye, **kw)


def get_boundary_condition(num: int, resolution: int, *, enable_dye: bool, **kw) -> BoundaryCondition:
    """(:201-219) NotImplementedError for unknown scene numbers, like the reference."""
    if num not in (1, 2, 3, 4, 5, 6):
        raise NotImplementedError
    return _make(num, resolution, enable_dye, **kw)

"""Minimal Taichi-semantics shim (TEST INFRASTRUCTURE ONLY -- never imported by the product).

Purpose: execute the *unmodified* reference source (`/root/reference/fs/*.py`) in pure Python so
that golden fixtures for the hot path can be generated from the reference's own code in a
container where the real `taichi` wheel cannot be installed.  It is not a Taichi
re-implementation: it provides exactly the API surface `fs/*.py` touches and pins the lowering
rules SURVEY.md section 8(c) lists:

* every scalar is IEEE fp32 (`F32`), every binary op rounds to fp32, literal left-to-right order,
  no FMA contraction, no fast-math reassociation;
* Python-scope constants (`self.dt`, `self.dx**3`, `self.dt * self.weight`) are folded in double by
  the Python interpreter and cast to fp32 when they meet a device value;
* `x ** n` with integer n is a multiplication chain;
* `ti.min/ti.max` follow IEEE minNum/maxNum (`fminf/fmaxf`): a NaN operand loses;
* out-of-bounds *raw* field indexing clamps to the edge (SURVEY T3 pin);
* a struct-for (`for i, j in field`) can run in two modes:
    - "gather"  (default): all loads of an iteration see the pre-kernel state (plus that
      iteration's own stores); stores are applied when the loop ends, in (i, j) order, last writer
      wins.  This is the defined behaviour SURVEY T4 pins for the racy BC kernels.
    - "sequential": plain in-order execution with immediate stores (what one CPU thread does).
"""
from __future__ import annotations

import functools
import inspect

import numpy as np

_ERR = dict(over="ignore", invalid="ignore", divide="ignore", under="ignore")

f32 = np.float32
u8 = np.uint8
i32 = np.int32
cpu = "cpu"
gpu = "gpu"

_STATE = {"mode": "gather", "in_kernel": 0, "cur": {}, "pending": []}


def set_loop_mode(mode: str) -> None:
    assert mode in ("gather", "sequential")
    _STATE["mode"] = mode


def init(*args, **kwargs) -> None:  # noqa: ARG001
    return None


def template():
    return "template"


def static(x):
    return x


def data_oriented(cls):
    return cls


# --------------------------------------------------------------------------- scalars
class F32:
    """fp32 scalar with round-after-every-op semantics."""

    __slots__ = ("v",)

    def __init__(self, v) -> None:
        self.v = v.v if isinstance(v, F32) else np.float32(v)

    @staticmethod
    def _c(o):
        if isinstance(o, F32):
            return o.v
        if isinstance(o, (int, float, np.integer, np.floating)):
            return np.float32(o)
        return None

    def _bin(self, o, op, swap=False):
        if isinstance(o, Vector):
            return NotImplemented
        c = F32._c(o)
        if c is None:
            return NotImplemented
        with np.errstate(**_ERR):
            return F32(op(c, self.v) if swap else op(self.v, c))

    def __add__(self, o):
        return self._bin(o, np.add)

    def __radd__(self, o):
        return self._bin(o, np.add, True)

    def __sub__(self, o):
        return self._bin(o, np.subtract)

    def __rsub__(self, o):
        return self._bin(o, np.subtract, True)

    def __mul__(self, o):
        return self._bin(o, np.multiply)

    def __rmul__(self, o):
        return self._bin(o, np.multiply, True)

    def __truediv__(self, o):
        return self._bin(o, np.divide)

    def __rtruediv__(self, o):
        return self._bin(o, np.divide, True)

    def __neg__(self):
        return F32(-self.v)

    def __abs__(self):
        return F32(np.abs(self.v))

    def __pow__(self, n):
        assert isinstance(n, int) and n >= 1, "shim supports integer powers only"
        r = self
        for _ in range(n - 1):
            r = r * self
        return r

    def __lt__(self, o):
        return bool(self.v < F32._c(o))

    def __le__(self, o):
        return bool(self.v <= F32._c(o))

    def __gt__(self, o):
        return bool(self.v > F32._c(o))

    def __ge__(self, o):
        return bool(self.v >= F32._c(o))

    def __eq__(self, o):
        return bool(self.v == F32._c(o))

    def __ne__(self, o):
        return bool(self.v != F32._c(o))

    __hash__ = None

    def __float__(self):
        return float(self.v)

    def __int__(self):
        return int(self.v)

    def __repr__(self):
        return f"F32({self.v!r})"


def _is_int(x) -> bool:
    return isinstance(x, (int, np.integer)) and not isinstance(x, bool)


# --------------------------------------------------------------------------- vectors
class Vector:
    """Small dense vector (fp32 or int32 components)."""

    __slots__ = ("data", "_owner")

    def __init__(self, comps, _owner=None) -> None:
        if isinstance(comps, np.ndarray):
            self.data = comps
        else:
            comps = list(comps)
            if all(_is_int(c) for c in comps):
                self.data = np.array(comps, dtype=np.int32)
            else:
                self.data = np.array([F32._c(c) for c in comps], dtype=np.float32)
        self._owner = _owner

    # -- field constructors
    @staticmethod
    def field(n, dtype, shape):
        return Field(dtype, tuple(shape), n)

    # -- component access
    def _get(self, k):
        v = self.data[k]
        return int(v) if self.data.dtype == np.int32 else F32(v)

    def _set(self, k, val):
        self.data[k] = F32._c(val)
        if self._owner is not None:
            fld, idx = self._owner
            fld[idx] = Vector(self.data.copy())

    x = property(lambda s: s._get(0), lambda s, v: s._set(0, v))
    y = property(lambda s: s._get(1), lambda s, v: s._set(1, v))
    z = property(lambda s: s._get(2), lambda s, v: s._set(2, v))

    def __getitem__(self, k):
        return self._get(k)

    def __len__(self):
        return len(self.data)

    # -- arithmetic
    def _other(self, o):
        if isinstance(o, Vector):
            return o.data
        if isinstance(o, F32):
            return o.v
        if _is_int(o) and self.data.dtype == np.int32:
            return np.int32(o)
        if isinstance(o, (int, float, np.integer, np.floating)):
            return np.float32(o)
        return None

    def _bin(self, o, op, swap=False):
        b = self._other(o)
        if b is None:
            return NotImplemented
        a = self.data
        if a.dtype == np.int32 and not (isinstance(b, np.ndarray) and b.dtype == np.int32 or isinstance(b, np.int32)):
            a = a.astype(np.float32)
        if isinstance(b, np.ndarray) and b.dtype == np.int32 and a.dtype == np.float32:
            b = b.astype(np.float32)
        with np.errstate(**_ERR):
            r = op(b, a) if swap else op(a, b)
        if r.dtype not in (np.float32, np.int32):
            r = r.astype(np.float32)
        return Vector(r)

    def __add__(self, o):
        return self._bin(o, np.add)

    def __radd__(self, o):
        return self._bin(o, np.add, True)

    def __sub__(self, o):
        return self._bin(o, np.subtract)

    def __rsub__(self, o):
        return self._bin(o, np.subtract, True)

    def __mul__(self, o):
        return self._bin(o, np.multiply)

    def __rmul__(self, o):
        return self._bin(o, np.multiply, True)

    def __truediv__(self, o):
        a = self
        if self.data.dtype == np.int32:
            a = Vector(self.data.astype(np.float32))
        return a._bin(o, np.divide)

    def __neg__(self):
        return Vector(-self.data)

    def norm_sqr(self):
        acc = self._get(0) * self._get(0)
        for k in range(1, len(self.data)):
            acc = acc + self._get(k) * self._get(k)
        return acc

    def norm(self):
        return sqrt(self.norm_sqr())

    def __repr__(self):
        return f"Vector({self.data!r})"


class Matrix:
    __slots__ = ("cols_",)

    def __init__(self, cols_):
        self.cols_ = cols_

    @staticmethod
    def cols(cols_):
        return Matrix(list(cols_))

    def __matmul__(self, v: Vector):
        n_rows = len(self.cols_[0])
        out = []
        for r in range(n_rows):
            acc = self.cols_[0][r] * v[0]
            for k in range(1, len(self.cols_)):
                acc = acc + self.cols_[k][r] * v[k]
            out.append(acc)
        return Vector(out)


# --------------------------------------------------------------------------- fields
class Field:
    def __init__(self, dtype, shape, n=0) -> None:
        self.dtype = np.dtype(dtype)
        self.shape = tuple(int(s) for s in shape)
        self.n = n
        full = self.shape + ((n,) if n else ())
        self.arr = np.zeros(full, dtype=self.dtype)

    # iteration == struct-for
    def __iter__(self):
        for idx in np.ndindex(*self.shape):
            yield idx
            _end_iteration()
        _flush()

    def _idx(self, idx):
        if isinstance(idx, Vector):
            idx = tuple(int(c) for c in idx.data)
        elif not isinstance(idx, tuple):
            idx = (idx,)
        # OOB raw access clamps (SURVEY T3 pin)
        return tuple(min(max(int(k), 0), s - 1) for k, s in zip(idx, self.shape))

    def _wrap(self, idx, raw):
        if self.n:
            return Vector(np.array(raw, dtype=np.float32), _owner=(self, idx))
        if self.dtype == np.float32:
            return F32(raw)
        return int(raw)

    def __getitem__(self, idx):
        idx = self._idx(idx)
        if _STATE["mode"] == "gather" and _STATE["in_kernel"]:
            key = (id(self), idx)
            if key in _STATE["cur"]:
                return self._wrap(idx, _STATE["cur"][key][2])
        return self._wrap(idx, self.arr[idx])

    def __setitem__(self, idx, val):
        idx = self._idx(idx)
        if isinstance(val, Vector):
            raw = val.data.astype(self.dtype).copy()
        elif isinstance(val, F32):
            raw = val.v
        else:
            raw = self.dtype.type(val)
        if _STATE["mode"] == "gather" and _STATE["in_kernel"]:
            _STATE["cur"][(id(self), idx)] = (self, idx, raw)
        else:
            self.arr[idx] = raw

    def fill(self, v) -> None:
        self.arr[...] = v

    def from_numpy(self, a) -> None:
        self.arr[...] = np.asarray(a).astype(self.dtype)

    def to_numpy(self):
        return self.arr.copy()


def field(dtype, shape):
    return Field(dtype, tuple(shape) if not isinstance(shape, int) else (shape,), 0)


def _end_iteration() -> None:
    if _STATE["cur"]:
        _STATE["pending"].extend(_STATE["cur"].values())
        _STATE["cur"] = {}


def _flush() -> None:
    _end_iteration()
    for fld, idx, raw in _STATE["pending"]:
        fld.arr[idx] = raw
    _STATE["pending"] = []


# --------------------------------------------------------------------------- decorators
def _convert_args(fn):
    sig = inspect.signature(fn)
    params = list(sig.parameters.values())

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        args = list(args)
        for k, p in enumerate(params[: len(args)]):
            ann = p.annotation
            if ann in (float, "float", f32) and not isinstance(args[k], (F32, Vector)):
                args[k] = F32(args[k])
            elif ann in (int, "int") and not isinstance(args[k], int):
                args[k] = int(args[k])
        return fn(*args, **kwargs)

    return wrapper


def func(fn):
    return _convert_args(fn)


def kernel(fn):
    inner = _convert_args(fn)

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        _STATE["in_kernel"] += 1
        try:
            return inner(*args, **kwargs)
        finally:
            _STATE["in_kernel"] -= 1
            if _STATE["in_kernel"] == 0:
                _flush()

    return wrapper


# --------------------------------------------------------------------------- math
def _minmax(a, b, npop):
    if _is_int(a) and _is_int(b):
        return int(npop(a, b))
    if isinstance(a, Vector) or isinstance(b, Vector):
        va = a.data if isinstance(a, Vector) else F32._c(a)
        vb = b.data if isinstance(b, Vector) else F32._c(b)
        with np.errstate(**_ERR):
            return Vector(npop(va, vb).astype(np.float32))
    with np.errstate(**_ERR):
        return F32(npop(F32._c(a), F32._c(b)))


def max(a, b):  # noqa: A001
    return _minmax(a, b, np.fmax)  # fmaxf: NaN operand loses


def min(a, b):  # noqa: A001
    return _minmax(a, b, np.fmin)


def abs(a):  # noqa: A001
    if isinstance(a, Vector):
        return Vector(np.abs(a.data))
    if _is_int(a):
        return int(np.abs(a))
    return F32(np.abs(F32._c(a)))


def sqrt(a):
    with np.errstate(**_ERR):
        return F32(np.sqrt(F32._c(a)))


def floor(a):
    return F32(np.floor(F32._c(a)))


def atan2(a, b):
    return F32(np.arctan2(F32._c(a), F32._c(b)))

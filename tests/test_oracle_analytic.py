"""Analytic known-answer tests of the CPU oracle: answers that follow from the mathematics of each
reference kernel, independent of both the Taichi-semantics shim and the goldens generated through it.

The reference ships no tests; these pin the oracle from a second, independent side: polynomial exactness of the
difference stencils (fs/differentiation.py), of the Kawamura-Kuwahara scheme (fs/advection.py:27-60) and of the CIP
cubic (fs/solver.py:282-332), fixed points of the pressure iteration (fs/pressure_updater.py:23-38), the literal
(non-textbook) pressure source, and the pinned NaN rule of the confinement force.
The CUDA kernels are bit-compared with this oracle in tests/test_gpu_parity.py, so they inherit these properties.
"""
from __future__ import annotations

import numpy as np
import pytest

from oracle import oracle as orc

X, Y = 24, 20
DX = 1.0 / 16.0          # power of two, like every BASELINE config: scalings by dx are exact
FLUID = np.zeros((X, Y), dtype=np.uint8)


def grid():
    i, j = np.meshgrid(np.arange(X, dtype=np.float64), np.arange(Y, dtype=np.float64), indexing="ij")
    return i * DX, j * DX


def vec(a, b):
    return np.ascontiguousarray(np.stack([a, b], axis=-1), dtype=np.float32)


def interior(a, h=1):
    return a[h:X - h, h:Y - h]


# ------------------------------------------------------------------------------------------------ differences
def test_central_difference_is_exact_on_linear_fields_and_halved_at_clamped_edges():
    x, y = grid()
    f = vec(3.0 * x - 2.0 * y + 1.0, 0.5 * x + 4.0 * y)          # exactly representable values
    fx, fy = np.zeros_like(f), np.zeros_like(f)
    orc.set_grad(fx, fy, f, DX)                                   # fs/solver.py:207-211 -> diff_x / diff_y
    assert np.array_equal(interior(fx), np.broadcast_to(np.float32([3.0, 0.5]), interior(fx).shape))
    assert np.array_equal(interior(fy), np.broadcast_to(np.float32([-2.0, 4.0]), interior(fy).shape))
    # sample() clamps: at i = 0 the stencil is 0.5 * (f(1) - f(0)) / dx = half the slope (fs/differentiation.py:4-9, :41-44)
    assert np.array_equal(fx[0, 1:-1], np.broadcast_to(np.float32([1.5, 0.25]), fx[0, 1:-1].shape))
    assert np.array_equal(fy[1:-1, Y - 1], np.broadcast_to(np.float32([-1.0, 2.0]), fy[1:-1, Y - 1].shape))


def test_uniform_fields_have_zero_differences_everywhere():
    f = vec(np.full((X, Y), 0.7), np.full((X, Y), -1.3))
    fx, fy = np.ones_like(f), np.ones_like(f)
    orc.set_grad(fx, fy, f, DX)
    assert not fx.any() and not fy.any()
    w, wabs = np.ones((X, Y), np.float32), np.ones((X, Y), np.float32)
    orc.vort_calc(w, wabs, f, FLUID, DX)
    assert not w.any() and not wabs.any()


def test_curl_of_solid_body_rotation():
    x, y = grid()
    omega = 3.0
    v = vec(-omega * y, omega * x)
    w, wabs = np.zeros((X, Y), np.float32), np.zeros((X, Y), np.float32)
    orc.vort_calc(w, wabs, v, FLUID, DX)                          # fs/vorticity_confinement.py:27-32
    assert np.array_equal(interior(w), np.full_like(interior(w), 2.0 * omega))
    assert np.array_equal(wabs, np.abs(w))


# ------------------------------------------------------------------------------------------------ MAC update
def _cubic(x, y, c):
    return (c[0] + c[1] * x + c[2] * y + c[3] * x * x + c[4] * x * y + c[5] * y * y + c[6] * x ** 3 + c[7] * x * x * y
            + c[8] * x * y * y + c[9] * y ** 3)


def _cubic_dx(x, y, c):
    return c[1] + 2 * c[3] * x + c[4] * y + 3 * c[6] * x * x + 2 * c[7] * x * y + c[8] * y * y


def _cubic_dy(x, y, c):
    return c[2] + c[4] * x + 2 * c[5] * y + c[7] * x * x + 2 * c[8] * x * y + 3 * c[9] * y * y


@pytest.mark.parametrize("su,sv", [(1, 1), (-1, 1), (1, -1), (-1, -1)])
def test_kk_advection_is_exact_on_cubics(su, sv):
    """Kawamura-Kuwahara = 4th-order central difference + a 4th-difference dissipation term: both upwind branches
    differentiate cubics exactly (fs/advection.py:39-55).  vn = vc - dt * (u d/dx + v d/dy) vc with p = 0, 1/Re -> 0."""
    rng = np.random.default_rng(5)
    x, y = grid()
    cu, cv = rng.uniform(-1, 1, 10), rng.uniform(-1, 1, 10)
    cu[0], cv[0] = su * 8.0, sv * 8.0                             # fixes the sign of u and v on the whole grid
    u, w = _cubic(x, y, cu), _cubic(x, y, cv)
    assert (np.sign(u) == su).all() and (np.sign(w) == sv).all()
    vc = vec(u, w)
    vn = np.zeros_like(vc)
    dt = 1e-3
    orc.mac_update(vn, vc, np.zeros((X, Y), np.float32), FLUID, dt, DX, 1e30, "kk")     # fs/solver.py:94-107
    u32, w32 = vc[..., 0].astype(np.float64), vc[..., 1].astype(np.float64)
    want_u = u32 - dt * (u32 * _cubic_dx(x, y, cu) + w32 * _cubic_dy(x, y, cu))
    want_w = w32 - dt * (u32 * _cubic_dx(x, y, cv) + w32 * _cubic_dy(x, y, cv))
    got = vn.astype(np.float64)
    # fp32 cancellation in the 5-point sums: |values| ~ 10, / (6 dx) ~ 2.7 per unit -> a few 1e-5 absolute on the derivative
    np.testing.assert_allclose(interior(got[..., 0], 2), interior(want_u, 2), rtol=0, atol=dt * 10 * 2e-4)
    np.testing.assert_allclose(interior(got[..., 1], 2), interior(want_w, 2), rtol=0, atol=dt * 10 * 2e-4)


def test_upwind_advection_picks_the_upwind_side():
    """advect_upwind (fs/advection.py:12-24): u >= 0 uses the backward difference, u < 0 the forward one.  On
    phi = x^2 these differ by exactly 2 dx * ... so the side is observable."""
    x, y = grid()
    for s in (1.0, -1.0):
        vc = vec(np.full((X, Y), s), np.zeros((X, Y)))
        vc[..., 1] = (x * x).astype(np.float32)                   # advected component: phi = x^2 (exact in fp32)
        vn = np.zeros_like(vc)
        dt = 0.25
        orc.mac_update(vn, vc, np.zeros((X, Y), np.float32), FLUID, dt, DX, 1e30, "upwind")
        # d/dx by one-sided difference of x^2: backward 2x - dx, forward 2x + dx; v-component of velocity advects in y
        # with d/dy(x^2) = 0, so only u * d/dx remains
        want = x * x - dt * s * (2 * x - s * DX) - dt * (x * x) * 0.0
        np.testing.assert_array_equal(interior(vn[..., 1]), interior(want).astype(np.float32))


def test_viscous_term_is_the_five_point_laplacian_over_re():
    x, y = grid()
    vc = vec(x * x + 2.0 * y * y, np.zeros((X, Y)))             # Laplacian = 2 + 4 = 6 exactly
    fn = np.zeros_like(vc)
    dt, re = 0.5, 4.0
    orc.cip_nonadv(fn, vc, np.zeros((X, Y), np.float32), FLUID, dt, DX, re)            # fs/solver.py:229-240
    np.testing.assert_array_equal(interior(fn[..., 0]), interior(vc[..., 0]) + np.float32(dt * 6.0 / re))
    # pressure gradient: p = 3x - y  ->  fn = fc - dt * (3, -1)
    p = (3.0 * x - y).astype(np.float32)
    orc.cip_nonadv(fn, np.zeros_like(vc), p, FLUID, dt, DX, re)
    assert np.array_equal(interior(fn), np.broadcast_to(np.float32([-1.5, 0.5]), interior(fn).shape))


# ------------------------------------------------------------------------------------------------ CIP
def test_cip_with_zero_velocity_is_the_identity():
    rng = np.random.default_rng(1)
    f, fx, fy = (rng.uniform(-1, 1, (X, Y, 2)).astype(np.float32) for _ in range(3))
    fn, fxn, fyn = (np.full((X, Y, 2), 7.0, np.float32) for _ in range(3))
    orc.cip_advect(fn, fxn, fyn, f, fx, fy, np.zeros((X, Y, 2), np.float32), FLUID, 0.01, DX)   # fs/solver.py:282-332
    assert np.array_equal(fn, f) and np.array_equal(fxn, fx) and np.array_equal(fyn, fy)


@pytest.mark.parametrize("u0,v0", [(0.8, 0.5), (-0.6, 0.9), (0.7, -0.4), (-0.5, -0.3)])
def test_cip_with_constant_velocity_translates_cubics_exactly(u0, v0):
    """The CIP polynomial carries all ten monomials of a bivariate cubic (fs/solver.py:300-323), so a cubic profile with
    its analytic derivatives is transported exactly: f_new(x, y) = f(x - u0 dt, y - v0 dt), and the derivative fields
    follow (constant velocity: the stretching terms of :329-332 vanish)."""
    rng = np.random.default_rng(11)
    x, y = grid()
    c0, c1 = rng.uniform(-1, 1, 10), rng.uniform(-1, 1, 10)
    f = vec(_cubic(x, y, c0), _cubic(x, y, c1))
    fx = vec(_cubic_dx(x, y, c0), _cubic_dx(x, y, c1))
    fy = vec(_cubic_dy(x, y, c0), _cubic_dy(x, y, c1))
    v = vec(np.full((X, Y), u0), np.full((X, Y), v0))
    dt = 0.4 * DX                                                  # CFL 0.4 * max(|u0|, |v0|) < 1: departure point in the upwind cell
    fn, fxn, fyn = (np.zeros((X, Y, 2), np.float32) for _ in range(3))
    orc.cip_advect(fn, fxn, fyn, f, fx, fy, v, FLUID, dt, DX)
    xd, yd = x - np.float32(u0) * np.float32(dt), y - np.float32(v0) * np.float32(dt)
    for c, cf in enumerate((c0, c1)):
        # the coefficients divide differences of O(1) values by dx^3 = 2^-12: errors ~ 1e-7 * 4096 * X^3-weights stay < 2e-5
        np.testing.assert_allclose(interior(fn[..., c]), interior(_cubic(xd, yd, cf)), rtol=0, atol=3e-5)
        np.testing.assert_allclose(interior(fxn[..., c]), interior(_cubic_dx(xd, yd, cf)), rtol=0, atol=2e-3)
        np.testing.assert_allclose(interior(fyn[..., c]), interior(_cubic_dy(xd, yd, cf)), rtol=0, atol=2e-3)


# ------------------------------------------------------------------------------------------------ pressure
def test_constant_pressure_is_a_fixed_point_of_both_relaxations_when_v_is_zero():
    p = np.full((X, Y), 2.5, np.float32)
    pn = np.zeros_like(p)
    v0 = np.zeros((X, Y, 2), np.float32)
    orc.jacobi_sweep(pn, p, v0, FLUID, 0.01, DX)                  # fs/pressure_updater.py:62-66
    assert np.array_equal(pn, p)
    pn = p.copy()
    orc.rbsor_pass(pn, p, v0, FLUID, 0.01, DX, 1.3, 1)            # :98-114, odd then even
    orc.rbsor_pass(pn, pn, v0, FLUID, 0.01, DX, 1.3, 0)
    np.testing.assert_allclose(pn, p, rtol=3e-7)                  # (1 - w) p + w p rounds


def test_pressure_source_is_the_literal_non_textbook_expression():
    """predict_p (fs/pressure_updater.py:23-38): with s_x = v(i+1) - v(i-1), s_y = v(j+1) - v(j-1) the source is
    (s_x.x^2 + s_y.y^2 + s_y.x * s_x.y) / 8 - dx * (s_x.x + s_y.y) / (8 dt) -- NOT the textbook one (which carries
    2 * u_y * v_x and halves the squares).  Linear velocity field: every term is a known constant."""
    x, y = grid()
    a, b, c, d = 2.0, -1.0, 0.5, 3.0                              # u = a x + b y, v = c x + d y
    v = vec(a * x + b * y, c * x + d * y)
    dt = 1.0 / 64.0
    pn = np.zeros((X, Y), np.float32)
    orc.jacobi_sweep(pn, np.zeros((X, Y), np.float32), v, FLUID, dt, DX)
    sxx, syy, syx, sxy = 2 * DX * a, 2 * DX * d, 2 * DX * b, 2 * DX * c
    literal = (sxx ** 2 + syy ** 2 + syx * sxy) / 8.0 - DX * (sxx + syy) / (8.0 * dt)
    textbook = (sxx ** 2 + syy ** 2 + 2.0 * syx * sxy) / 16.0 - DX * (sxx + syy) / (8.0 * dt)
    assert abs(literal - textbook) > 1e-3
    np.testing.assert_array_equal(interior(pn), np.full_like(interior(pn), np.float32(literal)))


def test_pressure_bc_copies_the_fluid_neighbour_and_zeroes_the_outflow():
    mask = np.zeros((8, 6), np.uint8)
    mask[:, 0] = mask[:, -1] = 1            # walls on the j edges
    mask[0, 1:-1] = 2                       # inflow column
    mask[-1, 1:-1] = 3                      # outflow column
    mask[4, 3] = 1                          # an isolated wall cell: the first matching branch is the (i-1, j+1) corner average
    p = np.arange(48, dtype=np.float32).reshape(8, 6) + 1.0
    want = p.copy()
    want[1:-1, 0] = p[1:-1, 1]              # wall: p = p(i, j+1)  (fs/boundary_condition.py:52-53)
    want[1:-1, -1] = p[1:-1, -2]            # wall: p = p(i, j-1)  (:50-51)
    want[0, 1:-1] = p[1, 1:-1]              # inflow: p = p(i+1, j) (:62-63)
    want[-1, 1:-1] = 0.0                    # outflow: p = 0        (:64-65)
    want[4, 3] = (p[3, 3] + p[4, 4]) / 2.0  # elif order is the priority order (:54-55)
    got = p.copy()
    orc.p_bc(got, mask)
    assert np.array_equal(got[1:-1], want[1:-1]) and np.array_equal(got[0, 1:-1], want[0, 1:-1])
    assert np.array_equal(got[-1, 1:-1], want[-1, 1:-1])


# ------------------------------------------------------------------------------------------------ limiter / confinement
def test_limiter_rescales_only_fast_cells_and_lets_nan_through():
    v = np.zeros((4, 4, 2), np.float32)
    v[0, 0] = (30.0, 40.0)                  # norm 50 -> (6, 8)
    v[1, 1] = (6.0, 8.0)                    # norm exactly 10: untouched (strict >)
    v[2, 2] = (np.nan, 1.0)                 # NaN compares false: untouched
    v[3, 3] = (-3.0, 4.0)
    w = v.copy()
    orc.limit(w)                            # fs/solver.py:38-43
    np.testing.assert_allclose(w[0, 0], (6.0, 8.0), rtol=2e-7)
    assert np.array_equal(w[1, 1], v[1, 1]) and np.isnan(w[2, 2, 0]) and w[2, 2, 1] == 1.0
    assert np.array_equal(w[3, 3], v[3, 3])


def test_confinement_force_direction_clamp_and_the_pinned_nan_rule():
    x, y = grid()
    dt, eps = 0.01, 5.0
    dtw = np.float32(dt * eps)
    vc = np.zeros((X, Y, 2), np.float32)
    # |w| grows along +x: N = (1, 0), force = (N.y, -N.x) * w = (0, -w), clamped to [-0.1, 0.1]
    wabs = x.astype(np.float32)
    w = np.full((X, Y), 0.05, np.float32)
    vn = np.zeros_like(vc)
    orc.vort_add(vn, vc, w, wabs, FLUID, DX, dt, eps)             # fs/vorticity_confinement.py:34-55
    assert np.array_equal(interior(vn[..., 0]), np.zeros_like(interior(vn[..., 0])))
    assert np.array_equal(interior(vn[..., 1]), np.full_like(interior(vn[..., 1]), dtw * np.float32(-0.05)))
    w[:] = 7.0                                                     # |force| > 0.1 -> clamp
    orc.vort_add(vn, vc, w, wabs, FLUID, DX, dt, eps)
    assert np.array_equal(interior(vn[..., 1]), np.full_like(interior(vn[..., 1]), dtw * np.float32(-0.1)))
    # uniform |w| (every quiescent cell): grad = 0 -> 0/0 = NaN -> fminf/fmaxf rule -> +0.1 in BOTH components
    wabs[:] = 1.0
    orc.vort_add(vn, vc, w, wabs, FLUID, DX, dt, eps)
    assert np.array_equal(vn, np.full_like(vn, dtw * np.float32(0.1)))

/*
 * fs2d_oracle.c -- CPU restatement of the reference's per-step hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (`2d-fluid-simulator_b200/`) may call this;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Every function restates one Taichi kernel of takah29/2d-fluid-simulator (reference commit
 * ebd5d75, paths relative to /root/reference) with the lowering rules of SURVEY.md 8(c):
 *   - fp32 for every intermediate, literal left-to-right operation order, no FMA contraction
 *     (compile with -ffp-contract=off, no -ffast-math);
 *   - Python-scope constants folded in double on the host and passed in already cast to fp32
 *     (dx**2, dx**3, 2*dx, dt*weight, 1-omega): the caller (oracle/oracle.py) does the folding;
 *   - clamp-to-edge `sample()` (fs/differentiation.py:4-9); raw out-of-bounds indexing clamps too
 *     (SURVEY T3 pin);
 *   - min/max are fminf/fmaxf (SURVEY T2 pin);
 *   - in-place BC kernels read the pre-kernel state (gather form, SURVEY T4 pin); stores are
 *     applied in (i, j) order of the writing cell, last writer wins.
 *
 * Parity status: PINNED against golden fixtures produced by executing the unmodified reference
 * source under oracle/ti_shim (tests/golden/make_golden.py); not pinned against a real Taichi
 * runtime (none can be installed offline).
 *
 * Layout: grid X (slow, index i) x Y (fast, index j), row-major; vector fields AoS [i][j][c].
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define IDX(i, j) ((size_t)(i) * (size_t)Y + (size_t)(j))
static inline int clampi(int a, int lo, int hi) { return a < lo ? lo : (a > hi ? hi : a); }
#define CI(i) clampi((i), 0, X - 1)
#define CJ(j) clampi((j), 0, Y - 1)
/* fs/differentiation.py:4-9 sample(): clamp-to-edge load */
#define S1(f, i, j) ((f)[IDX(CI(i), CJ(j))])
#define S2(f, i, j, c) ((f)[2 * IDX(CI(i), CJ(j)) + (c)])
/* mask load with the same clamp policy (raw OOB mask reads) */
#define M(i, j) (mask[IDX(CI(i), CJ(j))])

/* fs/differentiation.py:12-14 */
static inline float signf_(float x) { return x < 0.0f ? -1.0f : 1.0f; }

/* ------------------------------------------------------------------------------------------
 * fs/boundary_condition.py:16-39  BoundaryCondition.set_velocity_boundary_condition (in place)
 * Literal form: snapshot, then every writer cell in (i,j) order stores to its target.
 * ---------------------------------------------------------------------------------------- */
void orc_vel_bc(float *v, const uint8_t *mask, const float *bc_const, int X, int Y) {
    size_t n = (size_t)X * Y * 2;
    float *s = (float *)malloc(n * sizeof(float));
    memcpy(s, v, n * sizeof(float));
    for (int i = 0; i < X; ++i)
        for (int j = 0; j < Y; ++j) {
            uint8_t m = mask[IDX(i, j)];
            if (m == 1 && 1 <= i && i < X - 1 && 1 <= j && j < Y - 1) {
                if (M(i - 1, j) == 0 && M(i, j - 1) == 1 && M(i, j + 1) == 1) {
                    v[2 * IDX(i + 1, j)] = -S2(s, i - 1, j, 0);
                    v[2 * IDX(i + 1, j) + 1] = -S2(s, i - 1, j, 1);
                } else if (M(i + 1, j) == 0 && M(i, j - 1) == 1 && M(i, j + 1) == 1) {
                    v[2 * IDX(i - 1, j)] = -S2(s, i + 1, j, 0);
                    v[2 * IDX(i - 1, j) + 1] = -S2(s, i + 1, j, 1);
                } else if (M(i, j - 1) == 0 && M(i - 1, j) == 1 && M(i + 1, j) == 1) {
                    v[2 * IDX(i, j + 1)] = -S2(s, i, j - 1, 0);
                    v[2 * IDX(i, j + 1) + 1] = -S2(s, i, j - 1, 1);
                } else if (M(i, j + 1) == 0 && M(i - 1, j) == 1 && M(i + 1, j) == 1) {
                    v[2 * IDX(i, j - 1)] = -S2(s, i, j + 1, 0);
                    v[2 * IDX(i, j - 1) + 1] = -S2(s, i, j + 1, 1);
                }
            } else if (m == 2) {
                v[2 * IDX(i, j)] = bc_const[2 * IDX(i, j)];
                v[2 * IDX(i, j) + 1] = bc_const[2 * IDX(i, j) + 1];
            } else if (m == 3) {
                /* vc[i,j].x = max(sample(vc,i-1,j).x, 0.05); the shim stores the whole vector
                 * (x new, y = pre-kernel y) */
                v[2 * IDX(i, j)] = fmaxf(S2(s, i - 1, j, 0), 0.05f);
                v[2 * IDX(i, j) + 1] = s[2 * IDX(i, j) + 1];
            }
        }
    free(s);
}

/* Which scatter branch (1..4) does wall cell (i,j) take, 0 if none.  boundary_condition.py:20-34 */
static inline int vel_branch(const uint8_t *mask, int X, int Y, int i, int j) {
    if (i < 0 || i >= X || j < 0 || j >= Y) return 0;
    if (mask[IDX(i, j)] != 1 || !(1 <= i && i < X - 1 && 1 <= j && j < Y - 1)) return 0;
    if (M(i - 1, j) == 0 && M(i, j - 1) == 1 && M(i, j + 1) == 1) return 1;
    if (M(i + 1, j) == 0 && M(i, j - 1) == 1 && M(i, j + 1) == 1) return 2;
    if (M(i, j - 1) == 0 && M(i - 1, j) == 1 && M(i + 1, j) == 1) return 3;
    if (M(i, j + 1) == 0 && M(i - 1, j) == 1 && M(i + 1, j) == 1) return 4;
    return 0;
}

/* Target-centric (race-free, parallel) form of the same kernel: for every target cell pick the
 * LAST writer in (i,j) order among its five candidate writers.  Must equal orc_vel_bc. */
void orc_vel_bc_tc(float *v, const uint8_t *mask, const float *bc_const, int X, int Y) {
    size_t n = (size_t)X * Y * 2;
    float *s = (float *)malloc(n * sizeof(float));
    memcpy(s, v, n * sizeof(float));
#pragma omp parallel for schedule(static)
    for (int i = 0; i < X; ++i)
        for (int j = 0; j < Y; ++j) {
            float *t = v + 2 * IDX(i, j);
            uint8_t m = mask[IDX(i, j)];
            /* candidates in DEscending writer order: (i+1,j) b2, (i,j+1) b4, own, (i,j-1) b3, (i-1,j) b1 */
            if (vel_branch(mask, X, Y, i + 1, j) == 2) {          /* writer W=(i+1,j): v[W.i-1] = -v[W.i+1] */
                t[0] = -S2(s, i + 2, j, 0); t[1] = -S2(s, i + 2, j, 1);
            } else if (vel_branch(mask, X, Y, i, j + 1) == 4) {   /* W=(i,j+1): v[W.j-1] = -v[W.j+1] */
                t[0] = -S2(s, i, j + 2, 0); t[1] = -S2(s, i, j + 2, 1);
            } else if (m == 2) {
                t[0] = bc_const[2 * IDX(i, j)]; t[1] = bc_const[2 * IDX(i, j) + 1];
            } else if (m == 3) {
                t[0] = fmaxf(S2(s, i - 1, j, 0), 0.05f);
            } else if (vel_branch(mask, X, Y, i, j - 1) == 3) {   /* W=(i,j-1): v[W.j+1] = -v[W.j-1] */
                t[0] = -S2(s, i, j - 2, 0); t[1] = -S2(s, i, j - 2, 1);
            } else if (vel_branch(mask, X, Y, i - 1, j) == 1) {   /* W=(i-1,j): v[W.i+1] = -v[W.i-1] */
                t[0] = -S2(s, i - 2, j, 0); t[1] = -S2(s, i - 2, j, 1);
            }
        }
    free(s);
}

/* ------------------------------------------------------------------------------------------
 * fs/boundary_condition.py:41-65  set_pressure_boundary_condition (in place, gather form)
 * compute-then-commit so that every load sees the pre-kernel state.
 * ---------------------------------------------------------------------------------------- */
static inline int p_bc_value(const float *p, const uint8_t *mask, int X, int Y, int i, int j, float *out) {
    uint8_t m = mask[IDX(i, j)];
    if (m == 1) {
        if (M(i - 1, j) == 0 && M(i, j - 1) == 1 && M(i, j + 1) == 1) { *out = S1(p, i - 1, j); return 1; }
        if (M(i + 1, j) == 0 && M(i, j - 1) == 1 && M(i, j + 1) == 1) { *out = S1(p, i + 1, j); return 1; }
        if (M(i, j - 1) == 0 && M(i - 1, j) == 1 && M(i + 1, j) == 1) { *out = S1(p, i, j - 1); return 1; }
        if (M(i, j + 1) == 0 && M(i - 1, j) == 1 && M(i + 1, j) == 1) { *out = S1(p, i, j + 1); return 1; }
        if (M(i - 1, j) == 0 && M(i, j + 1) == 0) { *out = (S1(p, i - 1, j) + S1(p, i, j + 1)) / 2.0f; return 1; }
        if (M(i + 1, j) == 0 && M(i, j + 1) == 0) { *out = (S1(p, i + 1, j) + S1(p, i, j + 1)) / 2.0f; return 1; }
        if (M(i - 1, j) == 0 && M(i, j - 1) == 0) { *out = (S1(p, i - 1, j) + S1(p, i, j - 1)) / 2.0f; return 1; }
        if (M(i + 1, j) == 0 && M(i, j - 1) == 0) { *out = (S1(p, i + 1, j) + S1(p, i, j - 1)) / 2.0f; return 1; }
        return 0;
    }
    if (m == 2) { *out = S1(p, i + 1, j); return 1; }
    if (m == 3) { *out = 0.0f; return 1; }
    return 0;
}

void orc_p_bc(float *p, const uint8_t *mask, int X, int Y, float *tmp /* X*Y scratch */) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < X; ++i)
        for (int j = 0; j < Y; ++j)
            if (mask[IDX(i, j)] != 0) {
                float val;
                tmp[IDX(i, j)] = p_bc_value(p, mask, X, Y, i, j, &val) ? val : p[IDX(i, j)];
            }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < X; ++i)
        for (int j = 0; j < Y; ++j)
            if (mask[IDX(i, j)] != 0) p[IDX(i, j)] = tmp[IDX(i, j)];
}

/* ------------------------------------------------------------------------------------------
 * fs/differentiation.py:41-60 helpers on component c of a vec2 field / on a scalar field
 * ---------------------------------------------------------------------------------------- */
#define DIFFX2(f, i, j, c) (0.5f * (S2(f, (i) + 1, j, c) - S2(f, (i) - 1, j, c)) / dx)
#define DIFFY2(f, i, j, c) (0.5f * (S2(f, i, (j) + 1, c) - S2(f, i, (j) - 1, c)) / dx)
#define DIFFX1(f, i, j) (0.5f * (S1(f, (i) + 1, j) - S1(f, (i) - 1, j)) / dx)
#define DIFFY1(f, i, j) (0.5f * (S1(f, i, (j) + 1) - S1(f, i, (j) - 1)) / dx)
#define DIFF2X2(f, i, j, c) ((S2(f, (i) + 1, j, c) - 2.0f * S2(f, i, j, c) + S2(f, (i) - 1, j, c)) / (dx * dx))
#define DIFF2Y2(f, i, j, c) ((S2(f, i, (j) + 1, c) - 2.0f * S2(f, i, j, c) + S2(f, i, (j) - 1, c)) / (dx * dx))

/* fs/advection.py:12-24 advect_upwind, component c of phi=vc */
static inline float adv_upwind(const float *vc, int X, int Y, int i, int j, int c, float dx) {
    float u = vc[2 * IDX(i, j)], w = vc[2 * IDX(i, j) + 1];
    int k = u < 0.0f ? i : i - 1;
    float a = u * ((S2(vc, k + 1, j, c) - S2(vc, k, j, c)) / dx);
    k = w < 0.0f ? j : j - 1;
    float b = w * ((S2(vc, i, k + 1, c) - S2(vc, i, k, c)) / dx);
    return a + b;
}

/* fs/advection.py:27-60 advect_kk_scheme, component c */
static inline float adv_kk(const float *vc, int X, int Y, int i, int j, int c, float dx) {
    static const float cneg[5] = {-2.0f, 10.0f, -9.0f, 2.0f, -1.0f}; /* coef            */
    static const float cpos[5] = {1.0f, -2.0f, 9.0f, -10.0f, 2.0f};  /* -coef[::-1]     */
    float u = vc[2 * IDX(i, j)], w = vc[2 * IDX(i, j) + 1];
    const float *k = u < 0.0f ? cneg : cpos;
    float acc = S2(vc, i + 2, j, c) * k[0];
    acc = acc + S2(vc, i + 1, j, c) * k[1];
    acc = acc + S2(vc, i, j, c) * k[2];
    acc = acc + S2(vc, i - 1, j, c) * k[3];
    acc = acc + S2(vc, i - 2, j, c) * k[4];
    float a = acc / (6.0f * dx);
    k = w < 0.0f ? cneg : cpos;
    acc = S2(vc, i, j + 2, c) * k[0];
    acc = acc + S2(vc, i, j + 1, c) * k[1];
    acc = acc + S2(vc, i, j, c) * k[2];
    acc = acc + S2(vc, i, j - 1, c) * k[3];
    acc = acc + S2(vc, i, j - 2, c) * k[4];
    float b = acc / (6.0f * dx);
    return u * a + w * b;
}

/* fs/solver.py:94-107 MacSolver._update_velocities; scheme 0=upwind 1=kk */
void orc_mac_update(float *vn, const float *vc, const float *pc, const uint8_t *mask, int X, int Y,
                    float dt, float dx, float re, int scheme) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < X; ++i)
        for (int j = 0; j < Y; ++j) {
            if (mask[IDX(i, j)] != 0) continue;
            float gp[2] = {DIFFX1(pc, i, j), DIFFY1(pc, i, j)};
            for (int c = 0; c < 2; ++c) {
                float adv = scheme == 0 ? adv_upwind(vc, X, Y, i, j, c, dx) : adv_kk(vc, X, Y, i, j, c, dx);
                float lap = (DIFF2X2(vc, i, j, c) + DIFF2Y2(vc, i, j, c)) / re;
                vn[2 * IDX(i, j) + c] = vc[2 * IDX(i, j) + c] + dt * (-adv - gp[c] + lap);
            }
        }
}

/* fs/solver.py:229-240 CipMacSolver._non_advection_phase (+ _calc_diffusion :263-265) */
void orc_cip_nonadv(float *fn, const float *fc, const float *pc, const uint8_t *mask, int X, int Y,
                    float dt, float dx, float re) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < X; ++i)
        for (int j = 0; j < Y; ++j) {
            if (mask[IDX(i, j)] == 1) continue;
            float gp[2] = {DIFFX1(pc, i, j), DIFFY1(pc, i, j)};
            for (int c = 0; c < 2; ++c) {
                float g = -gp[c] + (DIFF2X2(fc, i, j, c) + DIFF2Y2(fc, i, j, c)) / re;
                fn[2 * IDX(i, j) + c] = fc[2 * IDX(i, j) + c] + g * dt;
            }
        }
}

/* fs/solver.py:242-261 _non_advection_phase_grad; raw indexing -> clamp (T3); two_dx = f32(2.0*dx) */
void orc_cip_nonadv_grad(float *fxn, float *fyn, const float *fxc, const float *fyc, const float *fc,
                         const float *fn, const uint8_t *mask, int X, int Y, float two_dx) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < X; ++i)
        for (int j = 0; j < Y; ++j) {
            if (mask[IDX(i, j)] == 1) continue;
            for (int c = 0; c < 2; ++c) {
                fxn[2 * IDX(i, j) + c] = fxc[2 * IDX(i, j) + c] +
                    (S2(fn, i + 1, j, c) - S2(fc, i + 1, j, c) - S2(fn, i - 1, j, c) + S2(fc, i - 1, j, c)) / two_dx;
                fyn[2 * IDX(i, j) + c] = fyc[2 * IDX(i, j) + c] +
                    (S2(fn, i, j + 1, c) - S2(fc, i, j + 1, c) - S2(fn, i, j - 1, c) + S2(fc, i, j - 1, c)) / two_dx;
            }
        }
}

/* fs/solver.py:267-332 _advection_phase / _cip_advect.  v is the advecting velocity (== fc for
 * the velocity update).  dx2 = f32(dx**2), dx3 = f32(dx**3) folded in double by the caller. */
void orc_cip_advect(float *fn, float *fxn, float *fyn, const float *fc, const float *fxc, const float *fyc,
                    const float *v, const uint8_t *mask, int X, int Y, float dt, float dx, float dx2, float dx3) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < X; ++i)
        for (int j = 0; j < Y; ++j) {
            if (mask[IDX(i, j)] != 0) continue;
            float u = v[2 * IDX(i, j)], w = v[2 * IDX(i, j) + 1];
            float i_s = signf_(u), j_s = signf_(w);
            int i_m = i - (int)i_s, j_m = j - (int)j_s;
            float isd = i_s * dx3, jsd = j_s * dx3, isdx = i_s * dx;
            float Xd = -u * dt, Yd = -w * dt;
            float dxu = DIFFX2(v, i, j, 0), dxv = DIFFX2(v, i, j, 1);
            float dyu = DIFFY2(v, i, j, 0), dyv = DIFFY2(v, i, j, 1);
            for (int c = 0; c < 2; ++c) {
                float f00 = S2(fc, i, j, c), f0m = S2(fc, i, j_m, c), fm0 = S2(fc, i_m, j, c), fmm = S2(fc, i_m, j_m, c);
                float x00 = S2(fxc, i, j, c), x0m = S2(fxc, i, j_m, c), xm0 = S2(fxc, i_m, j, c);
                float y00 = S2(fyc, i, j, c), y0m = S2(fyc, i, j_m, c), ym0 = S2(fyc, i_m, j, c);
                float tmp1 = f00 - f0m - fm0 + fmm;
                float tmp2 = fm0 - f00;
                float tmp3 = f0m - f00;
                float a = (i_s * (xm0 + x00) * dx - 2.0f * (-tmp2)) / isd;
                float b = (j_s * (y0m + y00) * dx - 2.0f * (-tmp3)) / jsd;
                float cc = (-tmp1 - i_s * (x0m - x00) * dx) / jsd;
                float d = (-tmp1 - j_s * (ym0 - y00) * dx) / isd;
                float e = (3.0f * tmp2 + i_s * (xm0 + 2.0f * x00) * dx) / dx2;
                float f = (3.0f * tmp3 + j_s * (y0m + 2.0f * y00) * dx) / dx2;
                float g = (-(ym0 - y00) + cc * dx2) / isdx;
                fn[2 * IDX(i, j) + c] = ((a * Xd + cc * Yd + e) * Xd + g * Yd + x00) * Xd +
                                        ((b * Yd + d * Xd + f) * Yd + y00) * Yd + f00;
                float Fx = (3.0f * a * Xd + 2.0f * cc * Yd + 2.0f * e) * Xd + (d * Yd + g) * Yd + x00;
                float Fy = (3.0f * b * Yd + 2.0f * d * Xd + 2.0f * f) * Yd + (cc * Xd + g) * Xd + y00;
                fxn[2 * IDX(i, j) + c] = Fx - dt * (Fx * dxu + Fy * dxv) / 2.0f;
                fyn[2 * IDX(i, j) + c] = Fy - dt * (Fx * dyu + Fy * dyv) / 2.0f;
            }
        }
}

/* fs/solver.py:207-211 _set_grad (all cells) */
void orc_set_grad(float *fx, float *fy, const float *f, int X, int Y, float dx) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < X; ++i)
        for (int j = 0; j < Y; ++j)
            for (int c = 0; c < 2; ++c) {
                fx[2 * IDX(i, j) + c] = DIFFX2(f, i, j, c);
                fy[2 * IDX(i, j) + c] = DIFFY2(f, i, j, c);
            }
}

/* fs/vorticity_confinement.py:27-32 _calc_vorticity */
void orc_vort_calc(float *w, float *wabs, const float *vc, const uint8_t *mask, int X, int Y, float dx) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < X; ++i)
        for (int j = 0; j < Y; ++j) {
            if (mask[IDX(i, j)] != 0) continue;
            float o = DIFFX2(vc, i, j, 1) - DIFFY2(vc, i, j, 0);
            w[IDX(i, j)] = o;
            wabs[IDX(i, j)] = fabsf(o);
        }
}

/* fs/vorticity_confinement.py:34-55 _add_vorticity / _vorticity_vec; dtw = f32(dt*weight) */
void orc_vort_add(float *vn, const float *vc, const float *w, const float *wabs, const uint8_t *mask,
                  int X, int Y, float dx, float dtw) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < X; ++i)
        for (int j = 0; j < Y; ++j) {
            if (mask[IDX(i, j)] != 0) continue;
            float gx = DIFFX1(wabs, i, j), gy = DIFFY1(wabs, i, j);
            float nrm = sqrtf(gx * gx + gy * gy);
            float nx = gx / nrm, ny = gy / nrm;            /* 0/0 = NaN on quiescent cells (T2) */
            float o = w[IDX(i, j)];
            float fx = ny * o, fy = -nx * o;
            fx = fmaxf(fminf(fx, 0.1f), -0.1f);            /* NaN -> +0.1 (fminf/fmaxf rule)     */
            fy = fmaxf(fminf(fy, 0.1f), -0.1f);
            vn[2 * IDX(i, j)] = vc[2 * IDX(i, j)] + dtw * fx;
            vn[2 * IDX(i, j) + 1] = vc[2 * IDX(i, j) + 1] + dtw * fy;
        }
}

/* fs/pressure_updater.py:23-38 predict_p */
static inline float predict_p(const float *pc, const float *vc, int X, int Y, int i, int j, float dt, float dx) {
    float sxx = S2(vc, i + 1, j, 0) - S2(vc, i - 1, j, 0), sxy = S2(vc, i + 1, j, 1) - S2(vc, i - 1, j, 1);
    float syx = S2(vc, i, j + 1, 0) - S2(vc, i, j - 1, 0), syy = S2(vc, i, j + 1, 1) - S2(vc, i, j - 1, 1);
    return 0.25f * (S1(pc, i + 1, j) + S1(pc, i - 1, j) + S1(pc, i, j + 1) + S1(pc, i, j - 1)) +
           (sxx * sxx + syy * syy + (syx * sxy)) / 8.0f - dx * (sxx + syy) / (8.0f * dt);
}

/* fs/pressure_updater.py:62-66 JacobiPressureUpdater._update (one sweep, no BC) */
void orc_jacobi_sweep(float *pn, const float *pc, const float *vc, const uint8_t *mask, int X, int Y,
                      float dt, float dx) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < X; ++i)
        for (int j = 0; j < Y; ++j)
            if (mask[IDX(i, j)] != 1) pn[IDX(i, j)] = predict_p(pc, vc, X, Y, i, j, dt, dx);
}

/* fs/pressure_updater.py:98-114 one colour pass of RedBlackSorPressureUpdater.
 * parity 1 = _update_pressures_odd(pn, pc), parity 0 = _update_pressures_even(pn, pn): the
 * caller passes pc == pn for the even pass exactly as the reference does (:96). */
void orc_rbsor_pass(float *pn, const float *pc, const float *vc, const uint8_t *mask, int X, int Y,
                    float dt, float dx, float omega, float one_minus_omega, int parity) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < X; ++i)
        for (int j = 0; j < Y; ++j)
            if (((i + j) & 1) == parity && mask[IDX(i, j)] == 0)
                pn[IDX(i, j)] = one_minus_omega * pc[IDX(i, j)] + omega * predict_p(pc, vc, X, Y, i, j, dt, dx);
}

/* fs/solver.py:38-43 limit_field */
void orc_limit(float *v, int X, int Y, float limit) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < X; ++i)
        for (int j = 0; j < Y; ++j) {
            float x = v[2 * IDX(i, j)], y = v[2 * IDX(i, j) + 1];
            float nrm = sqrtf(x * x + y * y);
            if (nrm > limit) {
                v[2 * IDX(i, j)] = limit * (x / nrm);
                v[2 * IDX(i, j) + 1] = limit * (y / nrm);
            }
        }
}

/* ==========================================================================================
 * Dye transport: the same kernels on C-channel fields (C = 3 for dye).
 * SC(f, i, j, c): clamped sample of component c of a C-channel AoS field.
 * ======================================================================================== */
#define SC(f, i, j, c) ((f)[(size_t)C * IDX(CI(i), CJ(j)) + (c)])

/* fs/boundary_condition.py:94-99 DyeBoundaryCondition.set_dye_boundary_condition */
void orc_dye_bc(float *dye, const float *bc_dye, const uint8_t *mask, int X, int Y) {
    const int C = 3;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < X; ++i)
        for (int j = 0; j < Y; ++j)
            if (mask[IDX(i, j)] == 2)
                for (int c = 0; c < C; ++c) dye[C * IDX(i, j) + c] = bc_dye[C * IDX(i, j) + c];
}

/* fs/solver.py:46-49 clamp_field (all cells, all components) */
void orc_clamp(float *f, int X, int Y, int C, float low, float high) {
    size_t n = (size_t)X * Y * C;
#pragma omp parallel for schedule(static)
    for (size_t k = 0; k < n; ++k) f[k] = fminf(fmaxf(f[k], low), high);
}

/* fs/solver.py:157-161 DyeMacSolver._update_dye: dn = dc - dt * advect(vc, dc); scheme 0 upwind, 1 kk */
void orc_dye_mac(float *dn, const float *dc, const float *vc, const uint8_t *mask, int X, int Y, int C, float dt,
                 float dx, int scheme) {
    static const float cneg[5] = {-2.0f, 10.0f, -9.0f, 2.0f, -1.0f};
    static const float cpos[5] = {1.0f, -2.0f, 9.0f, -10.0f, 2.0f};
#pragma omp parallel for schedule(static)
    for (int i = 0; i < X; ++i)
        for (int j = 0; j < Y; ++j) {
            if (mask[IDX(i, j)] != 0) continue;
            float u = vc[2 * IDX(i, j)], w = vc[2 * IDX(i, j) + 1];
            for (int c = 0; c < C; ++c) {
                float adv;
                if (scheme == 0) { /* fs/advection.py:12-24 */
                    int k = u < 0.0f ? i : i - 1;
                    float a = u * ((SC(dc, k + 1, j, c) - SC(dc, k, j, c)) / dx);
                    k = w < 0.0f ? j : j - 1;
                    float b = w * ((SC(dc, i, k + 1, c) - SC(dc, i, k, c)) / dx);
                    adv = a + b;
                } else { /* fs/advection.py:27-60 */
                    const float *k = u < 0.0f ? cneg : cpos;
                    float acc = SC(dc, i + 2, j, c) * k[0];
                    acc = acc + SC(dc, i + 1, j, c) * k[1];
                    acc = acc + SC(dc, i, j, c) * k[2];
                    acc = acc + SC(dc, i - 1, j, c) * k[3];
                    acc = acc + SC(dc, i - 2, j, c) * k[4];
                    float a = acc / (6.0f * dx);
                    k = w < 0.0f ? cneg : cpos;
                    acc = SC(dc, i, j + 2, c) * k[0];
                    acc = acc + SC(dc, i, j + 1, c) * k[1];
                    acc = acc + SC(dc, i, j, c) * k[2];
                    acc = acc + SC(dc, i, j - 1, c) * k[3];
                    acc = acc + SC(dc, i, j - 2, c) * k[4];
                    float b = acc / (6.0f * dx);
                    adv = u * a + w * b;
                }
                dn[C * IDX(i, j) + c] = dc[C * IDX(i, j) + c] - dt * adv;
            }
        }
}

/* fs/solver.py:378-383 DyeCipMacSolver._non_advection_phase_dye: dn = dc + diffusion(dc) * dt (not-wall) */
void orc_dye_nonadv(float *dn, const float *dc, const uint8_t *mask, int X, int Y, int C, float dt, float dx, float re) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < X; ++i)
        for (int j = 0; j < Y; ++j) {
            if (mask[IDX(i, j)] == 1) continue;
            for (int c = 0; c < C; ++c) {
                float d2x = (SC(dc, i + 1, j, c) - 2.0f * SC(dc, i, j, c) + SC(dc, i - 1, j, c)) / (dx * dx);
                float d2y = (SC(dc, i, j + 1, c) - 2.0f * SC(dc, i, j, c) + SC(dc, i, j - 1, c)) / (dx * dx);
                dn[C * IDX(i, j) + c] = dc[C * IDX(i, j) + c] + (d2x + d2y) / re * dt;
            }
        }
}

/* fs/solver.py:242-261 on C channels */
void orc_cip_nonadv_grad_n(float *fxn, float *fyn, const float *fxc, const float *fyc, const float *fc, const float *fn,
                           const uint8_t *mask, int X, int Y, int C, float two_dx) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < X; ++i)
        for (int j = 0; j < Y; ++j) {
            if (mask[IDX(i, j)] == 1) continue;
            for (int c = 0; c < C; ++c) {
                fxn[C * IDX(i, j) + c] = fxc[C * IDX(i, j) + c] +
                    (SC(fn, i + 1, j, c) - SC(fc, i + 1, j, c) - SC(fn, i - 1, j, c) + SC(fc, i - 1, j, c)) / two_dx;
                fyn[C * IDX(i, j) + c] = fyc[C * IDX(i, j) + c] +
                    (SC(fn, i, j + 1, c) - SC(fc, i, j + 1, c) - SC(fn, i, j - 1, c) + SC(fc, i, j - 1, c)) / two_dx;
            }
        }
}

/* fs/solver.py:267-332 on C channels, advecting velocity v (2 channels) */
void orc_cip_advect_n(float *fn, float *fxn, float *fyn, const float *fc, const float *fxc, const float *fyc,
                      const float *v, const uint8_t *mask, int X, int Y, int C, float dt, float dx, float dx2, float dx3) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < X; ++i)
        for (int j = 0; j < Y; ++j) {
            if (mask[IDX(i, j)] != 0) continue;
            float u = v[2 * IDX(i, j)], w = v[2 * IDX(i, j) + 1];
            float i_s = signf_(u), j_s = signf_(w);
            int i_m = i - (int)i_s, j_m = j - (int)j_s;
            float isd = i_s * dx3, jsd = j_s * dx3, isdx = i_s * dx;
            float Xd = -u * dt, Yd = -w * dt;
            float dxu = DIFFX2(v, i, j, 0), dxv = DIFFX2(v, i, j, 1);
            float dyu = DIFFY2(v, i, j, 0), dyv = DIFFY2(v, i, j, 1);
            for (int c = 0; c < C; ++c) {
                float f00 = SC(fc, i, j, c), f0m = SC(fc, i, j_m, c), fm0 = SC(fc, i_m, j, c), fmm = SC(fc, i_m, j_m, c);
                float x00 = SC(fxc, i, j, c), x0m = SC(fxc, i, j_m, c), xm0 = SC(fxc, i_m, j, c);
                float y00 = SC(fyc, i, j, c), y0m = SC(fyc, i, j_m, c), ym0 = SC(fyc, i_m, j, c);
                float tmp1 = f00 - f0m - fm0 + fmm;
                float tmp2 = fm0 - f00;
                float tmp3 = f0m - f00;
                float a = (i_s * (xm0 + x00) * dx - 2.0f * (-tmp2)) / isd;
                float b = (j_s * (y0m + y00) * dx - 2.0f * (-tmp3)) / jsd;
                float cc = (-tmp1 - i_s * (x0m - x00) * dx) / jsd;
                float d = (-tmp1 - j_s * (ym0 - y00) * dx) / isd;
                float e = (3.0f * tmp2 + i_s * (xm0 + 2.0f * x00) * dx) / dx2;
                float f = (3.0f * tmp3 + j_s * (y0m + 2.0f * y00) * dx) / dx2;
                float g = (-(ym0 - y00) + cc * dx2) / isdx;
                fn[C * IDX(i, j) + c] = ((a * Xd + cc * Yd + e) * Xd + g * Yd + x00) * Xd +
                                        ((b * Yd + d * Xd + f) * Yd + y00) * Yd + f00;
                float Fx = (3.0f * a * Xd + 2.0f * cc * Yd + 2.0f * e) * Xd + (d * Yd + g) * Yd + x00;
                float Fy = (3.0f * b * Yd + 2.0f * d * Xd + 2.0f * f) * Yd + (cc * Xd + g) * Xd + y00;
                fxn[C * IDX(i, j) + c] = Fx - dt * (Fx * dxu + Fy * dxv) / 2.0f;
                fyn[C * IDX(i, j) + c] = Fy - dt * (Fx * dyu + Fy * dyv) / 2.0f;
            }
        }
}

/* fs/solver.py:207-211 on C channels */
void orc_set_grad_n(float *fx, float *fy, const float *f, int X, int Y, int C, float dx) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < X; ++i)
        for (int j = 0; j < Y; ++j)
            for (int c = 0; c < C; ++c) {
                fx[C * IDX(i, j) + c] = 0.5f * (SC(f, i + 1, j, c) - SC(f, i - 1, j, c)) / dx;
                fy[C * IDX(i, j) + c] = 0.5f * (SC(f, i, j + 1, c) - SC(f, i, j - 1, c)) / dx;
            }
}

/* ==========================================================================================
 * Render kernels: fs/fluid_simulator.py:38-58, :121-126 + fs/visualization.py:8-22.
 * mode 0 norm (+pressure tint), 1 pressure, 2 vorticity, 3 dye.  rgb: (X, Y, 3).
 * ======================================================================================== */
void orc_render(float *rgb, const float *v, const float *p, const float *dye, const uint8_t *mask, int X, int Y,
                float dx, int mode) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < X; ++i)
        for (int j = 0; j < Y; ++j) {
            float *o = rgb + 3 * IDX(i, j);
            if (mode == 0) {          /* _to_norm :38-44 */
                float x = v[2 * IDX(i, j)], y = v[2 * IDX(i, j) + 1];
                float c = sqrtf(x * x + y * y);
                float pv = p[IDX(i, j)];
                o[0] = 0.2f * c + 0.002f * fmaxf(pv, 0.0f);
                o[1] = 0.2f * c + 0.002f * 0.0f;
                o[2] = 0.2f * c + 0.002f * fmaxf(-pv, 0.0f);
            } else if (mode == 1) {   /* _to_pressure :46-51 */
                float pv = p[IDX(i, j)];
                o[0] = 0.04f * fmaxf(pv, 0.0f); o[1] = 0.04f * 0.0f; o[2] = 0.04f * fmaxf(-pv, 0.0f);
            } else if (mode == 2) {   /* _to_vorticity :53-58, visualization.py:19-22 */
                float val = DIFFX2(v, i, j, 1) - DIFFY2(v, i, j, 0);
                o[0] = 0.005f * fmaxf(val, 0.0f); o[1] = 0.005f * 0.0f; o[2] = 0.005f * fmaxf(-val, 0.0f);
            } else {                  /* _to_dye :121-126 */
                for (int c = 0; c < 3; ++c) o[c] = dye[3 * IDX(i, j) + c];
            }
            if (mask[IDX(i, j)] == 1) { o[0] = 0.5f; o[1] = 0.7f; o[2] = 0.5f; }   /* wall colour :17 */
        }
}

/* thread control for the timed CPU legs of bench.py: torchrun exports OMP_NUM_THREADS=1 to every rank, so the environment
 * cannot be trusted; the caller sets the team size explicitly */
void orc_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int orc_max_threads(void) { return omp_get_max_threads(); }

int orc_abi_version(void) { return 4; }

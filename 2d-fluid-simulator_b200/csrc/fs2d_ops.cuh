// fs2d_ops.cuh -- per-cell arithmetic of the CIP-path kernels, written against a neighbourhood accessor so that the
// same (bit-identical) expressions run from clamped global loads (GAt<true>, any cell), plain global offsets
// (GAt<false>, interior blocks) or a shared-memory tile filled by TMA (SAt, fs2d_stream.cu).
#pragma once
#include "fs2d_common.cuh"

namespace fs2d {

// neighbourhood of the cell (r, j) of a global array
template <bool CL>
struct GAt {
    static constexpr bool kCheapDependentLoads = false;   // a load that depends on a loaded value costs a DRAM latency
    const fs2d_dom &d;
    int r, j;
    __device__ __forceinline__ float ld1(const float *f, int dr, int dc) const { return fs2d::ld1<CL>(f, d, r + dr, j + dc); }
    __device__ __forceinline__ float2 ld2(const float *f, int dr, int dc) const { return fs2d::ld2<CL>(f, d, r + dr, j + dc); }
};
// neighbourhood of a cell inside shared-memory tiles: `f` points at the CELL ITSELF in a tile whose rows are
// `pitch1` floats (1-channel fields) / `pitch2` floats (2-channel fields) apart
struct SAt {
    static constexpr bool kCheapDependentLoads = true;
    int pitch1, pitch2;
    __device__ __forceinline__ float ld1(const float *f, int dr, int dc) const { return f[dr * pitch1 + dc]; }
    __device__ __forceinline__ float2 ld2(const float *f, int dr, int dc) const {
        return *reinterpret_cast<const float2 *>(f + dr * pitch2 + 2 * dc);
    }
};

// fs/solver.py:229-240  CipMacSolver._non_advection_phase
// loads of one cell of _non_advection_phase (all issued before any arithmetic: the IEEE divisions below contain
// branches the compiler does not hoist loads across)
struct NonadvIn { float2 c, ip, im, jp, jm; float pip, pim, pjp, pjm; };
template <class A>
__device__ __forceinline__ NonadvIn l_cip_nonadv(const A &at, const float *fc, const float *pc) {
    NonadvIn x;
    x.c = at.ld2(fc, 0, 0);
    x.ip = at.ld2(fc, +1, 0); x.im = at.ld2(fc, -1, 0);
    x.jp = at.ld2(fc, 0, +1); x.jm = at.ld2(fc, 0, -1);
    x.pip = at.ld1(pc, +1, 0); x.pim = at.ld1(pc, -1, 0);
    x.pjp = at.ld1(pc, 0, +1); x.pjm = at.ld1(pc, 0, -1);
    return x;
}
template <bool P2>
__device__ __forceinline__ float2 c_cip_nonadv(const NonadvIn &x, float dt, DivC<P2> ddx, DivC<P2> ddx2, float re) {
    const float2 gp = make_float2(ddx(0.5f * (x.pip - x.pim)), ddx(0.5f * (x.pjp - x.pjm)));   // (diff_x p, diff_y p)
    const float2 d2x = ddx2(x.ip - 2.0f * x.c + x.im), d2y = ddx2(x.jp - 2.0f * x.c + x.jm);
    const float2 g = -gp + (d2x + d2y) / re;
    return x.c + g * dt;
}

// fs/solver.py:267-332  _advection_phase / _cip_advect
struct CipOut { float2 f, fx, fy; };
template <bool P2, class A>
__device__ __forceinline__ CipOut c_cip_advect(const A &at, const float *fc, const float *fxc, const float *fyc, const float *v,
                                               float dt, float dx, DivC<P2> ddx, DivC<P2> ddx2, DivC<P2> ddx3) {
    const float dx2 = ddx2.c;
    float2 vel, dxv, dyv, f00, f0m, fm0, fmm, x00, x0m, xm0, y00, y0m, ym0;
    float i_s, j_s;
    if (A::kCheapDependentLoads) {
        // shared-memory tile: read the velocity first, then exactly the upwind neighbours (r_m, j_m) = (r - i_s, j - j_s)
        f00 = at.ld2(fc, 0, 0);
        vel = v == fc ? f00 : at.ld2(v, 0, 0);
        i_s = sign1(vel.x); j_s = sign1(vel.y);
        const int di = -(int)i_s, dj = -(int)j_s;
        dxv = ddx(0.5f * (at.ld2(v, +1, 0) - at.ld2(v, -1, 0)));   // diff_x(v) = (d/dx u, d/dx v)
        dyv = ddx(0.5f * (at.ld2(v, 0, +1) - at.ld2(v, 0, -1)));   // diff_y(v)
        f0m = at.ld2(fc, 0, dj); fm0 = at.ld2(fc, di, 0); fmm = at.ld2(fc, di, dj);
        x00 = at.ld2(fxc, 0, 0); x0m = at.ld2(fxc, 0, dj); xm0 = at.ld2(fxc, di, 0);
        y00 = at.ld2(fyc, 0, 0); y0m = at.ld2(fyc, 0, dj); ym0 = at.ld2(fyc, di, 0);
    } else {
        // Global memory: one round of independent loads.  Both candidates of every upwind neighbour are fetched (they are
        // neighbours of the cell, so L1/L2 hits) and the upwind one is selected afterwards -- loading only (r_m, j_m)
        // makes ten loads depend on the velocity load, i.e. two DRAM latencies per cell.
        f00 = at.ld2(fc, 0, 0);
        const float2 f0a = at.ld2(fc, 0, -1), f0b = at.ld2(fc, 0, +1);
        const float2 fa0 = at.ld2(fc, -1, 0), fb0 = at.ld2(fc, +1, 0);
        const float2 faa = at.ld2(fc, -1, -1), fab = at.ld2(fc, -1, +1);
        const float2 fba = at.ld2(fc, +1, -1), fbb = at.ld2(fc, +1, +1);
        x00 = at.ld2(fxc, 0, 0);
        const float2 x0a = at.ld2(fxc, 0, -1), x0b = at.ld2(fxc, 0, +1);
        const float2 xa0 = at.ld2(fxc, -1, 0), xb0 = at.ld2(fxc, +1, 0);
        y00 = at.ld2(fyc, 0, 0);
        const float2 y0a = at.ld2(fyc, 0, -1), y0b = at.ld2(fyc, 0, +1);
        const float2 ya0 = at.ld2(fyc, -1, 0), yb0 = at.ld2(fyc, +1, 0);
        const bool same = v == fc;   // the advecting velocity is the advected field itself in CipMacSolver (block-uniform)
        vel = same ? f00 : at.ld2(v, 0, 0);
        const float2 vb0 = same ? fb0 : at.ld2(v, +1, 0), va0 = same ? fa0 : at.ld2(v, -1, 0);
        const float2 v0b = same ? f0b : at.ld2(v, 0, +1), v0a = same ? f0a : at.ld2(v, 0, -1);
        i_s = sign1(vel.x); j_s = sign1(vel.y);
        const bool im = !(vel.x < 0.0f), jm = !(vel.y < 0.0f);   // upwind cell (r_m, j_m) = (r - i_s, j - j_s)
        dxv = ddx(0.5f * (vb0 - va0));  // diff_x(v) = (d/dx u, d/dx v)
        dyv = ddx(0.5f * (v0b - v0a));  // diff_y(v)
        f0m = jm ? f0a : f0b; fm0 = im ? fa0 : fb0;
        const float2 fma = im ? faa : fba, fmb = im ? fab : fbb;
        fmm = jm ? fma : fmb;
        x0m = jm ? x0a : x0b; xm0 = im ? xa0 : xb0;
        y0m = jm ? y0a : y0b; ym0 = im ? ya0 : yb0;
    }
    // +-dx^3, +-dx are exact sign flips; divisions by them are exact scalings when dx is 2^k
    const DivC<P2> disd = ddx3.signed_by(i_s), djsd = ddx3.signed_by(j_s), disdx = ddx.signed_by(i_s);
    const float Xd = -vel.x * dt, Yd = -vel.y * dt;

    const float2 tmp1 = f00 - f0m - fm0 + fmm;
    const float2 tmp2 = fm0 - f00;
    const float2 tmp3 = f0m - f00;

    const float2 a = disd(i_s * (xm0 + x00) * dx - 2.0f * (-tmp2));
    const float2 b = djsd(j_s * (y0m + y00) * dx - 2.0f * (-tmp3));
    const float2 c = djsd(-tmp1 - i_s * (x0m - x00) * dx);
    const float2 dd = disd(-tmp1 - j_s * (ym0 - y00) * dx);
    const float2 e = ddx2(3.0f * tmp2 + i_s * (xm0 + 2.0f * x00) * dx);
    const float2 f = ddx2(3.0f * tmp3 + j_s * (y0m + 2.0f * y00) * dx);
    const float2 g = disdx(-(ym0 - y00) + c * dx2);

    CipOut o;
    o.f = ((a * Xd + c * Yd + e) * Xd + g * Yd + x00) * Xd + ((b * Yd + dd * Xd + f) * Yd + y00) * Yd + f00;
    const float2 Fx = (3.0f * a * Xd + 2.0f * c * Yd + 2.0f * e) * Xd + (dd * Yd + g) * Yd + x00;
    const float2 Fy = (3.0f * b * Yd + 2.0f * dd * Xd + 2.0f * f) * Yd + (c * Xd + g) * Xd + y00;
    o.fx = Fx - dt * (Fx * dxv.x + Fy * dxv.y) / 2.0f;
    o.fy = Fy - dt * (Fx * dyv.x + Fy * dyv.y) / 2.0f;
    return o;
}

// fs/vorticity_confinement.py:34-55  _add_vorticity / _vorticity_vec
struct VortIn { float aip, aim, ajp, ajm, o; float2 c; };
template <class A>
__device__ __forceinline__ VortIn l_vort_add(const A &at, const float *vc, const float *w, const float *wabs) {
    VortIn x;
    x.aip = at.ld1(wabs, +1, 0); x.aim = at.ld1(wabs, -1, 0);
    x.ajp = at.ld1(wabs, 0, +1); x.ajm = at.ld1(wabs, 0, -1);
    x.o = at.ld1(w, 0, 0);
    x.c = at.ld2(vc, 0, 0);
    return x;
}
template <bool P2>
__device__ __forceinline__ float2 c_vort_add(const VortIn &x, DivC<P2> ddx, float dtw) {
    const float gx = ddx(0.5f * (x.aip - x.aim)), gy = ddx(0.5f * (x.ajp - x.ajm));
    float fx, fy;
    if (gx == 0.0f && gy == 0.0f) {
        // grad|w| = 0 exactly (every cell of a quiescent or uniform region): n = 0/0 = NaN in both components, NaN * w = NaN,
        // and the clamp turns NaN into +0.1 by the fminf/fmaxf rule whatever w is -- no sqrt, no divisions
        fx = fy = 0.1f;
    } else {
        const float n2 = gx * gx + gy * gy;
        const float nrm = n2 == 0.0f ? n2 : sqrtf(n2);                  // sqrt(+0) = +0 without the zero-operand slow path
        const float nx = fdiv_z(gx, nrm), ny = fdiv_z(gy, nrm);         // x/0 = +-inf when the squares underflowed
        fx = ny * x.o;
        fy = -nx * x.o;
        fx = fmaxf(fminf(fx, 0.1f), -0.1f);  // NaN -> +0.1 by the fminf/fmaxf rule
        fy = fmaxf(fminf(fy, 0.1f), -0.1f);
    }
    return make_float2(x.c.x + dtw * fx, x.c.y + dtw * fy);
}

}  // namespace fs2d

// fs2d_kernels.cu -- dense stencil kernels + sparse BC kernels of libfs2d.so (sm_100a).
//
// One kernel per reference Taichi kernel (file:line cited at each).  All kernels are pure
// 5-/9-/13-point fp32 stencils: HBM-bandwidth bound, no tensor cores.  Literal operation order of
// the reference source, compiled with -fmad=false, so outputs are bit-identical to the oracle.
#include <cstdarg>
#include <cstdio>
#include <cmath>

#include "fs2d_common.cuh"

namespace fs2d {

static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;  // kernels launched by this library (bench.py "gpu_launches")
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
bool is_pow2(float x) {
    int e;
    return x > 0.0f && std::isfinite(x) && std::frexp(x, &e) == 0.5f && (1.0f / x) * x == 1.0f;
}
int check_dom(const fs2d_dom &d) {
    FS2D_REQUIRE(d.rows > 0 && d.Y > 0, "empty domain");
    FS2D_REQUIRE(0 <= d.r0 && d.r0 <= d.r1 && d.r1 <= d.rows, "row range outside the local array");
    FS2D_REQUIRE(0 <= d.clo && d.clo <= d.chi && d.chi < d.rows, "clamp range outside the local array");
    return FS2D_OK;
}

// =============================================================================================
// sparse boundary conditions
// =============================================================================================
// fs/boundary_condition.py:16-39 -- phase 1: evaluate every entry from the pre-kernel state
__global__ void k_vel_bc_gather(const float2 *__restrict__ v, const float2 *__restrict__ bc_const,
                                const int32_t *__restrict__ tgt, const int32_t *__restrict__ src,
                                const uint8_t *__restrict__ kind, float2 *__restrict__ scratch, int n) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    uint8_t k = kind[e];
    float2 out;
    if (k == 0) {  // ghost cell two behind the face: v[tgt] = -v[src]  (:27-34)
        out = -v[src[e]];
    } else if (k == 1) {  // inflow: v = bc_const  (:35-36)
        out = bc_const[tgt[e]];
    } else {  // outflow: v.x = max(v(i-1,j).x, 0.05), y untouched  (:37-39)
        out = v[tgt[e]];
        out.x = fmaxf(v[src[e]].x, 0.05f);
    }
    scratch[e] = out;
}
__global__ void k_vel_bc_scatter(float2 *__restrict__ v, const int32_t *__restrict__ tgt,
                                 const float2 *__restrict__ scratch, int n) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) v[tgt[e]] = scratch[e];
}

// fs/boundary_condition.py:41-65
__global__ void k_p_bc_gather(const float *__restrict__ p, const int32_t *__restrict__ src0,
                              const int32_t *__restrict__ src1, const uint8_t *__restrict__ kind,
                              float *__restrict__ scratch, int n) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    uint8_t k = kind[e];
    float out;
    if (k == 0) out = p[src0[e]];
    else if (k == 1) out = (p[src0[e]] + p[src1[e]]) / 2.0f;
    else out = 0.0f;
    scratch[e] = out;
}
__global__ void k_p_bc_scatter(float *__restrict__ p, const int32_t *__restrict__ tgt,
                               const float *__restrict__ scratch, int n) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) p[tgt[e]] = scratch[e];
}

void launch_p_bc(float *p, const int32_t *tgt, const int32_t *src0, const int32_t *src1, const uint8_t *kind, float *scratch,
                 int n, cudaStream_t s) {
    if (n <= 0) return;
    g_launches += 2;
    k_p_bc_gather<<<nblk(n, 256), 256, 0, s>>>(p, src0, src1, kind, scratch, n);
    k_p_bc_scatter<<<nblk(n, 256), 256, 0, s>>>(p, tgt, scratch, n);
}

// =============================================================================================
// stencil building blocks (fs/differentiation.py:41-60), per component of a float2 field
// =============================================================================================
// CL: clamp-to-edge loads (true, any cell) or plain neighbour offsets (false, interior blocks; fs2d_common.cuh)
template <bool P2, bool CL = true>
__device__ __forceinline__ float2 diff_x2(const float *f, const fs2d_dom &d, int r, int j, DivC<P2> ddx) {
    return ddx(0.5f * (ld2<CL>(f, d, r + 1, j) - ld2<CL>(f, d, r - 1, j)));
}
template <bool P2, bool CL = true>
__device__ __forceinline__ float2 diff_y2(const float *f, const fs2d_dom &d, int r, int j, DivC<P2> ddx) {
    return ddx(0.5f * (ld2<CL>(f, d, r, j + 1) - ld2<CL>(f, d, r, j - 1)));
}
template <bool P2, bool CL = true>
__device__ __forceinline__ float diff_x1(const float *f, const fs2d_dom &d, int r, int j, DivC<P2> ddx) {
    return ddx(0.5f * (ld1<CL>(f, d, r + 1, j) - ld1<CL>(f, d, r - 1, j)));
}
template <bool P2, bool CL = true>
__device__ __forceinline__ float diff_y1(const float *f, const fs2d_dom &d, int r, int j, DivC<P2> ddx) {
    return ddx(0.5f * (ld1<CL>(f, d, r, j + 1) - ld1<CL>(f, d, r, j - 1)));
}
// (diff2_x + diff2_y) of a float2 field; ddx2 divides by dx*dx
template <bool P2, bool CL = true>
__device__ __forceinline__ float2 laplace2(const float *f, const fs2d_dom &d, int r, int j, float2 c, DivC<P2> ddx2) {
    float2 d2x = ddx2(ld2<CL>(f, d, r + 1, j) - 2.0f * c + ld2<CL>(f, d, r - 1, j));
    float2 d2y = ddx2(ld2<CL>(f, d, r, j + 1) - 2.0f * c + ld2<CL>(f, d, r, j - 1));
    return d2x + d2y;
}

// fs/advection.py:12-24
template <bool P2>
__device__ __forceinline__ float2 advect_upwind(const float *vc, const fs2d_dom &d, int r, int j, float2 c,
                                                DivC<P2> ddx) {
    int k = c.x < 0.0f ? r : r - 1;
    float2 a = c.x * ddx(ld2(vc, d, k + 1, j) - ld2(vc, d, k, j));
    k = c.y < 0.0f ? j : j - 1;
    float2 b = c.y * ddx(ld2(vc, d, r, k + 1) - ld2(vc, d, r, k));
    return a + b;
}
// fs/advection.py:27-60
__device__ __forceinline__ float2 advect_kk(const float *vc, const fs2d_dom &d, int r, int j, float2 c, float dx) {
    const float six_dx = 6.0f * dx;
    float k0, k1, k2, k3, k4;
    if (c.x < 0.0f) { k0 = -2.0f; k1 = 10.0f; k2 = -9.0f; k3 = 2.0f; k4 = -1.0f; }
    else { k0 = 1.0f; k1 = -2.0f; k2 = 9.0f; k3 = -10.0f; k4 = 2.0f; }
    float2 acc = ld2(vc, d, r + 2, j) * k0;
    acc = acc + ld2(vc, d, r + 1, j) * k1;
    acc = acc + c * k2;
    acc = acc + ld2(vc, d, r - 1, j) * k3;
    acc = acc + ld2(vc, d, r - 2, j) * k4;
    float2 a = acc / six_dx;
    if (c.y < 0.0f) { k0 = -2.0f; k1 = 10.0f; k2 = -9.0f; k3 = 2.0f; k4 = -1.0f; }
    else { k0 = 1.0f; k1 = -2.0f; k2 = 9.0f; k3 = -10.0f; k4 = 2.0f; }
    acc = ld2(vc, d, r, j + 2) * k0;
    acc = acc + ld2(vc, d, r, j + 1) * k1;
    acc = acc + c * k2;
    acc = acc + ld2(vc, d, r, j - 1) * k3;
    acc = acc + ld2(vc, d, r, j - 2) * k4;
    float2 b = acc / six_dx;
    return c.x * a + c.y * b;
}

// =============================================================================================
// fs/solver.py:94-107  MacSolver._update_velocities
// =============================================================================================
template <bool P2, int SCHEME>
__global__ void __launch_bounds__(TX *TY)
    k_mac_update(float *__restrict__ vn, const float *__restrict__ vc, const float *__restrict__ pc,
                 const uint8_t *__restrict__ mask, fs2d_dom d, float dt, float dx, DivC<P2> ddx, DivC<P2> ddx2,
                 float re) {
    FS2D_CELL(d, r, j)
    const size_t idx = IX(d, r, j);
    if (mask[idx] != 0) return;
    const float2 c = __ldg(reinterpret_cast<const float2 *>(vc) + idx);
    float2 adv = SCHEME == FS2D_SCHEME_UPWIND ? advect_upwind<P2>(vc, d, r, j, c, ddx) : advect_kk(vc, d, r, j, c, dx);
    float2 gp = make_float2(diff_x1<P2>(pc, d, r, j, ddx), diff_y1<P2>(pc, d, r, j, ddx));
    float2 lap = laplace2<P2>(vc, d, r, j, c, ddx2) / re;
    reinterpret_cast<float2 *>(vn)[idx] = c + dt * (-adv - gp + lap);
}

// =============================================================================================
// The per-step CIP-path kernels below process NU rows per thread (rows r, r+TY, ...): each row's
// value is computed by a side-effect-free function (clamped loads are safe for any cell, wall cells just
// compute an unused value) and stored under the kernel's write predicate afterwards, so the loads of all
// rows are in flight together.  One row per thread left these streaming kernels latency-bound
// (8 B in flight per thread: 2.8-3.5 TB/s).
// =============================================================================================
constexpr int NU_CIP_NONADV = 2;
constexpr int NU_CIP_NONADV_GRAD = 1;
constexpr int NU_CIP_ADVECT = 1;
constexpr int NU_VORT_CALC = 4;
constexpr int NU_VORT_ADD = 4;
constexpr int NU_LIMIT = 4;
#define FS2D_ROWS(d, j, r, ok)                                                       \
    const int j = FS2D_COLBLK * blockDim.x + threadIdx.x;                             \
    if (j >= (d).Y) return;                                                           \
    int r[NU];                                                                        \
    bool ok[NU];                                                                      \
    _Pragma("unroll") for (int u = 0; u < NU; ++u) {                                  \
        const int rr = (d).r0 + (FS2D_ROWBLK * NU + u) * blockDim.y + threadIdx.y;    \
        ok[u] = rr < (d).r1;                                                          \
        r[u] = ok[u] ? rr : (d).r1 - 1;                                               \
    }

// fs/solver.py:229-240  CipMacSolver._non_advection_phase
// loads of one cell of _non_advection_phase (all issued before any arithmetic: the IEEE divisions below contain
// branches the compiler does not hoist loads across)
struct NonadvIn { float2 c, ip, im, jp, jm; float pip, pim, pjp, pjm; };
template <bool CL>
__device__ __forceinline__ NonadvIn l_cip_nonadv(const float *fc, const float *pc, const fs2d_dom &d, int r, int j) {
    NonadvIn x;
    x.c = ld2<CL>(fc, d, r, j);
    x.ip = ld2<CL>(fc, d, r + 1, j); x.im = ld2<CL>(fc, d, r - 1, j);
    x.jp = ld2<CL>(fc, d, r, j + 1); x.jm = ld2<CL>(fc, d, r, j - 1);
    x.pip = ld1<CL>(pc, d, r + 1, j); x.pim = ld1<CL>(pc, d, r - 1, j);
    x.pjp = ld1<CL>(pc, d, r, j + 1); x.pjm = ld1<CL>(pc, d, r, j - 1);
    return x;
}
template <bool P2>
__device__ __forceinline__ float2 c_cip_nonadv(const NonadvIn &x, float dt, DivC<P2> ddx, DivC<P2> ddx2, float re) {
    const float2 gp = make_float2(ddx(0.5f * (x.pip - x.pim)), ddx(0.5f * (x.pjp - x.pjm)));   // (diff_x p, diff_y p)
    const float2 d2x = ddx2(x.ip - 2.0f * x.c + x.im), d2y = ddx2(x.jp - 2.0f * x.c + x.jm);
    const float2 g = -gp + (d2x + d2y) / re;
    return x.c + g * dt;
}
template <bool P2, bool CL>
__device__ __forceinline__ void b_cip_nonadv(float *__restrict__ fn, const float *__restrict__ fc, const float *__restrict__ pc,
                                             const uint8_t *__restrict__ mask, const fs2d_dom &d, float dt, DivC<P2> ddx,
                                             DivC<P2> ddx2, float re) {
    constexpr int NU = NU_CIP_NONADV;
    FS2D_ROWS(d, j, r, ok)
    NonadvIn in[NU];
    uint8_t m[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        m[u] = __ldg(mask + IX(d, r[u], j));
        in[u] = l_cip_nonadv<CL>(fc, pc, d, r[u], j);
    }
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        const float2 out = c_cip_nonadv<P2>(in[u], dt, ddx, ddx2, re);
        if (ok[u] && m[u] != 1) reinterpret_cast<float2 *>(fn)[IX(d, r[u], j)] = out;
    }
}
template <bool P2>
__global__ void __launch_bounds__(TX *TY)
    k_cip_nonadv(float *__restrict__ fn, const float *__restrict__ fc, const float *__restrict__ pc,
                 const uint8_t *__restrict__ mask, fs2d_dom d, float dt, DivC<P2> ddx, DivC<P2> ddx2, float re) {
    if (block_interior(d, TY * NU_CIP_NONADV, 1)) b_cip_nonadv<P2, false>(fn, fc, pc, mask, d, dt, ddx, ddx2, re);
    else b_cip_nonadv<P2, true>(fn, fc, pc, mask, d, dt, ddx, ddx2, re);
}

// fs/solver.py:242-261  _non_advection_phase_grad (raw indexing -> clamp)
template <bool P2, bool CL>
__device__ __forceinline__ void b_cip_nonadv_grad(float *__restrict__ fxn, float *__restrict__ fyn, const float *__restrict__ fxc,
                                                  const float *__restrict__ fyc, const float *__restrict__ fc,
                                                  const float *__restrict__ fn, const uint8_t *__restrict__ mask,
                                                  const fs2d_dom &d, DivC<P2> d2dx) {
    constexpr int NU = NU_CIP_NONADV_GRAD;
    FS2D_ROWS(d, j, r, ok)
    float2 ox[NU], oy[NU];
    uint8_t m[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        const int rr = r[u];
        m[u] = __ldg(mask + IX(d, rr, j));
        float2 gx = ld2<CL>(fn, d, rr + 1, j) - ld2<CL>(fc, d, rr + 1, j) - ld2<CL>(fn, d, rr - 1, j) + ld2<CL>(fc, d, rr - 1, j);
        float2 gy = ld2<CL>(fn, d, rr, j + 1) - ld2<CL>(fc, d, rr, j + 1) - ld2<CL>(fn, d, rr, j - 1) + ld2<CL>(fc, d, rr, j - 1);
        ox[u] = ld2<CL>(fxc, d, rr, j) + d2dx(gx);
        oy[u] = ld2<CL>(fyc, d, rr, j) + d2dx(gy);
    }
#pragma unroll
    for (int u = 0; u < NU; ++u)
        if (ok[u] && m[u] != 1) {
            reinterpret_cast<float2 *>(fxn)[IX(d, r[u], j)] = ox[u];
            reinterpret_cast<float2 *>(fyn)[IX(d, r[u], j)] = oy[u];
        }
}
template <bool P2>
__global__ void __launch_bounds__(TX *TY)
    k_cip_nonadv_grad(float *__restrict__ fxn, float *__restrict__ fyn, const float *__restrict__ fxc,
                      const float *__restrict__ fyc, const float *__restrict__ fc, const float *__restrict__ fn,
                      const uint8_t *__restrict__ mask, fs2d_dom d, DivC<P2> d2dx) {
    if (block_interior(d, TY * NU_CIP_NONADV_GRAD, 1)) b_cip_nonadv_grad<P2, false>(fxn, fyn, fxc, fyc, fc, fn, mask, d, d2dx);
    else b_cip_nonadv_grad<P2, true>(fxn, fyn, fxc, fyc, fc, fn, mask, d, d2dx);
}

// =============================================================================================
// fs/solver.py:267-332  _advection_phase / _cip_advect
// =============================================================================================
struct CipOut { float2 f, fx, fy; };
template <bool P2, bool CL>
__device__ __forceinline__ CipOut c_cip_advect(const float *fc, const float *fxc, const float *fyc, const float *v,
                                               const fs2d_dom &d, int r, int j, float dt, float dx, DivC<P2> ddx,
                                               DivC<P2> ddx2, DivC<P2> ddx3) {
    const float dx2 = ddx2.c;
    // One round of independent loads: both candidates of every upwind neighbour are fetched (they are neighbours of
    // the cell, so L1/L2 hits) and the upwind one is selected afterwards -- loading only (r_m, j_m) makes ten loads
    // depend on the velocity load, i.e. two DRAM latencies per cell, and left the kernel latency-bound.
    const float2 f00 = ld2<CL>(fc, d, r, j);
    const float2 f0a = ld2<CL>(fc, d, r, j - 1), f0b = ld2<CL>(fc, d, r, j + 1);
    const float2 fa0 = ld2<CL>(fc, d, r - 1, j), fb0 = ld2<CL>(fc, d, r + 1, j);
    const float2 faa = ld2<CL>(fc, d, r - 1, j - 1), fab = ld2<CL>(fc, d, r - 1, j + 1);
    const float2 fba = ld2<CL>(fc, d, r + 1, j - 1), fbb = ld2<CL>(fc, d, r + 1, j + 1);
    const float2 x00 = ld2<CL>(fxc, d, r, j), x0a = ld2<CL>(fxc, d, r, j - 1), x0b = ld2<CL>(fxc, d, r, j + 1);
    const float2 xa0 = ld2<CL>(fxc, d, r - 1, j), xb0 = ld2<CL>(fxc, d, r + 1, j);
    const float2 y00 = ld2<CL>(fyc, d, r, j), y0a = ld2<CL>(fyc, d, r, j - 1), y0b = ld2<CL>(fyc, d, r, j + 1);
    const float2 ya0 = ld2<CL>(fyc, d, r - 1, j), yb0 = ld2<CL>(fyc, d, r + 1, j);
    const bool same = v == fc;   // the advecting velocity is the advected field itself in CipMacSolver (block-uniform)
    const float2 vel = same ? f00 : ld2<CL>(v, d, r, j);
    const float2 vb0 = same ? fb0 : ld2<CL>(v, d, r + 1, j), va0 = same ? fa0 : ld2<CL>(v, d, r - 1, j);
    const float2 v0b = same ? f0b : ld2<CL>(v, d, r, j + 1), v0a = same ? f0a : ld2<CL>(v, d, r, j - 1);

    const float i_s = sign1(vel.x), j_s = sign1(vel.y);
    const bool im = !(vel.x < 0.0f), jm = !(vel.y < 0.0f);   // upwind cell (r_m, j_m) = (r - i_s, j - j_s)
    // +-dx^3, +-dx are exact sign flips; divisions by them are exact scalings when dx is 2^k
    const DivC<P2> disd = ddx3.signed_by(i_s), djsd = ddx3.signed_by(j_s), disdx = ddx.signed_by(i_s);
    const float Xd = -vel.x * dt, Yd = -vel.y * dt;
    const float2 dxv = ddx(0.5f * (vb0 - va0));  // diff_x(v) = (d/dx u, d/dx v)
    const float2 dyv = ddx(0.5f * (v0b - v0a));  // diff_y(v)

    const float2 f0m = jm ? f0a : f0b, fm0 = im ? fa0 : fb0;
    const float2 fma = im ? faa : fba, fmb = im ? fab : fbb;
    const float2 fmm = jm ? fma : fmb;
    const float2 x0m = jm ? x0a : x0b, xm0 = im ? xa0 : xb0;
    const float2 y0m = jm ? y0a : y0b, ym0 = im ? ya0 : yb0;

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
template <bool P2, bool CL>
__device__ __forceinline__ void b_cip_advect(float *__restrict__ fn, float *__restrict__ fxn, float *__restrict__ fyn,
                                             const float *__restrict__ fc, const float *__restrict__ fxc,
                                             const float *__restrict__ fyc, const float *__restrict__ v,
                                             const uint8_t *__restrict__ mask, const fs2d_dom &d, float dt, float dx,
                                             DivC<P2> ddx, DivC<P2> ddx2, DivC<P2> ddx3) {
    constexpr int NU = NU_CIP_ADVECT;
    FS2D_ROWS(d, j, r, ok)
    CipOut o[NU];
    uint8_t m[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        m[u] = __ldg(mask + IX(d, r[u], j));
        o[u] = c_cip_advect<P2, CL>(fc, fxc, fyc, v, d, r[u], j, dt, dx, ddx, ddx2, ddx3);
    }
#pragma unroll
    for (int u = 0; u < NU; ++u)
        if (ok[u] && m[u] == 0) {
            const size_t idx = IX(d, r[u], j);
            reinterpret_cast<float2 *>(fn)[idx] = o[u].f;
            reinterpret_cast<float2 *>(fxn)[idx] = o[u].fx;
            reinterpret_cast<float2 *>(fyn)[idx] = o[u].fy;
        }
}
template <bool P2>
__global__ void __launch_bounds__(TX *TY)
    k_cip_advect(float *__restrict__ fn, float *__restrict__ fxn, float *__restrict__ fyn,
                 const float *__restrict__ fc, const float *__restrict__ fxc, const float *__restrict__ fyc,
                 const float *__restrict__ v, const uint8_t *__restrict__ mask, fs2d_dom d, float dt, float dx,
                 DivC<P2> ddx, DivC<P2> ddx2, DivC<P2> ddx3) {
    if (block_interior(d, TY * NU_CIP_ADVECT, 1)) b_cip_advect<P2, false>(fn, fxn, fyn, fc, fxc, fyc, v, mask, d, dt, dx, ddx, ddx2, ddx3);
    else b_cip_advect<P2, true>(fn, fxn, fyn, fc, fxc, fyc, v, mask, d, dt, dx, ddx, ddx2, ddx3);
}

// fs/solver.py:207-211  _set_grad
template <bool P2>
__global__ void __launch_bounds__(TX *TY)
    k_set_grad(float *__restrict__ fx, float *__restrict__ fy, const float *__restrict__ f, fs2d_dom d, DivC<P2> ddx) {
    FS2D_CELL(d, r, j)
    const size_t idx = IX(d, r, j);
    reinterpret_cast<float2 *>(fx)[idx] = diff_x2<P2>(f, d, r, j, ddx);
    reinterpret_cast<float2 *>(fy)[idx] = diff_y2<P2>(f, d, r, j, ddx);
}

// =============================================================================================
// fs/vorticity_confinement.py:27-32 / :34-55
// =============================================================================================
template <bool P2, bool CL>
__device__ __forceinline__ void b_vort_calc(float *__restrict__ w, float *__restrict__ wabs, const float *__restrict__ vc,
                                            const uint8_t *__restrict__ mask, const fs2d_dom &d, DivC<P2> ddx) {
    constexpr int NU = NU_VORT_CALC;
    FS2D_ROWS(d, j, r, ok)
    float o[NU];
    uint8_t m[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        m[u] = __ldg(mask + IX(d, r[u], j));
        o[u] = diff_x2<P2, CL>(vc, d, r[u], j, ddx).y - diff_y2<P2, CL>(vc, d, r[u], j, ddx).x;
    }
#pragma unroll
    for (int u = 0; u < NU; ++u)
        if (ok[u] && m[u] == 0) {
            w[IX(d, r[u], j)] = o[u];
            wabs[IX(d, r[u], j)] = fabsf(o[u]);
        }
}
template <bool P2>
__global__ void __launch_bounds__(TX *TY)
    k_vort_calc(float *__restrict__ w, float *__restrict__ wabs, const float *__restrict__ vc,
                const uint8_t *__restrict__ mask, fs2d_dom d, DivC<P2> ddx) {
    if (block_interior(d, TY * NU_VORT_CALC, 1)) b_vort_calc<P2, false>(w, wabs, vc, mask, d, ddx);
    else b_vort_calc<P2, true>(w, wabs, vc, mask, d, ddx);
}
struct VortIn { float aip, aim, ajp, ajm, o; float2 c; };
template <bool CL>
__device__ __forceinline__ VortIn l_vort_add(const float *vc, const float *w, const float *wabs, const fs2d_dom &d, int r, int j) {
    VortIn x;
    x.aip = ld1<CL>(wabs, d, r + 1, j); x.aim = ld1<CL>(wabs, d, r - 1, j);
    x.ajp = ld1<CL>(wabs, d, r, j + 1); x.ajm = ld1<CL>(wabs, d, r, j - 1);
    x.o = ld1<CL>(w, d, r, j);
    x.c = ld2<CL>(vc, d, r, j);
    return x;
}
template <bool P2>
__device__ __forceinline__ float2 c_vort_add(const VortIn &x, DivC<P2> ddx, float dtw) {
    const float gx = ddx(0.5f * (x.aip - x.aim)), gy = ddx(0.5f * (x.ajp - x.ajm));
    const float n2 = gx * gx + gy * gy;
    const float nrm = n2 == 0.0f ? n2 : sqrtf(n2);                  // sqrt(+0) = +0 without the zero-operand slow path
    const float nx = fdiv_z(gx, nrm), ny = fdiv_z(gy, nrm);         // 0/0 = NaN on quiescent cells
    float fx = ny * x.o, fy = -nx * x.o;
    fx = fmaxf(fminf(fx, 0.1f), -0.1f);  // NaN -> +0.1 by the fminf/fmaxf rule
    fy = fmaxf(fminf(fy, 0.1f), -0.1f);
    return make_float2(x.c.x + dtw * fx, x.c.y + dtw * fy);
}
template <bool P2, bool CL>
__device__ __forceinline__ void b_vort_add(float *__restrict__ vn, const float *__restrict__ vc, const float *__restrict__ w,
                                           const float *__restrict__ wabs, const uint8_t *__restrict__ mask, const fs2d_dom &d,
                                           DivC<P2> ddx, float dtw) {
    constexpr int NU = NU_VORT_ADD;
    FS2D_ROWS(d, j, r, ok)
    VortIn in[NU];
    uint8_t m[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) {   // all loads of the thread's rows first (the divisions below contain branches)
        m[u] = __ldg(mask + IX(d, r[u], j));
        in[u] = l_vort_add<CL>(vc, w, wabs, d, r[u], j);
    }
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        const float2 o = c_vort_add<P2>(in[u], ddx, dtw);
        if (ok[u] && m[u] == 0) reinterpret_cast<float2 *>(vn)[IX(d, r[u], j)] = o;
    }
}
template <bool P2>
__global__ void __launch_bounds__(TX *TY)
    k_vort_add(float *__restrict__ vn, const float *__restrict__ vc, const float *__restrict__ w,
               const float *__restrict__ wabs, const uint8_t *__restrict__ mask, fs2d_dom d, DivC<P2> ddx, float dtw) {
    if (block_interior(d, TY * NU_VORT_ADD, 1)) b_vort_add<P2, false>(vn, vc, w, wabs, mask, d, ddx, dtw);
    else b_vort_add<P2, true>(vn, vc, w, wabs, mask, d, ddx, dtw);
}

// fs/solver.py:38-43  limit_field
__global__ void __launch_bounds__(TX *TY) k_limit(float *__restrict__ v, fs2d_dom d, float limit) {
    constexpr int NU = NU_LIMIT;
    FS2D_ROWS(d, j, r, ok)
    float2 c[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) c[u] = reinterpret_cast<const float2 *>(v)[IX(d, r[u], j)];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        const float nrm = sqrtf(c[u].x * c[u].x + c[u].y * c[u].y);
        if (ok[u] && nrm > limit)
            reinterpret_cast<float2 *>(v)[IX(d, r[u], j)] = make_float2(limit * (c[u].x / nrm), limit * (c[u].y / nrm));
    }
}

}  // namespace fs2d

// =============================================================================================
// C ABI
// =============================================================================================
using namespace fs2d;

extern "C" {

const char *fs2d_last_error(void) { return g_err; }
int fs2d_version(void) { return 1; }
unsigned long long fs2d_launch_count(void) { return g_launches; }
int fs2d_device_ok(void) {
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
        set_error("no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
        return 0;
    }
    if (prop.major != 10) {
        set_error("libfs2d is built for sm_100a only; device is sm_%d%d", prop.major, prop.minor);
        return 0;
    }
    return 1;
}

#define STREAM ((cudaStream_t)stream)

int fs2d_vel_bc(float *v, const float *bc_const, const int32_t *tgt, const int32_t *src, const uint8_t *kind,
                float *scratch, int n, void *stream) {
    if (n == 0) return FS2D_OK;
    FS2D_REQUIRE(v && bc_const && tgt && src && kind && scratch && n > 0, "null table/field pointer");
    ++g_launches; k_vel_bc_gather<<<nblk(n, 256), 256, 0, STREAM>>>((const float2 *)v, (const float2 *)bc_const, tgt, src, kind,
                                                       (float2 *)scratch, n);
    ++g_launches; k_vel_bc_scatter<<<nblk(n, 256), 256, 0, STREAM>>>((float2 *)v, tgt, (const float2 *)scratch, n);
    FS2D_LAUNCH_CHECK();
    return FS2D_OK;
}

int fs2d_pressure_bc(float *p, const int32_t *tgt, const int32_t *src0, const int32_t *src1, const uint8_t *kind,
                     float *scratch, int n, void *stream) {
    if (n == 0) return FS2D_OK;
    FS2D_REQUIRE(p && tgt && src0 && src1 && kind && scratch && n > 0, "null table/field pointer");
    launch_p_bc(p, tgt, src0, src1, kind, scratch, n, STREAM);
    FS2D_LAUNCH_CHECK();
    return FS2D_OK;
}

#define DISPATCH_P2(p2, CALL_T, CALL_F) \
    do {                                \
        if (p2) { CALL_T; } else { CALL_F; } \
    } while (0)

int fs2d_mac_update(float *vn, const float *vc, const float *pc, const uint8_t *mask, fs2d_dom d, float dt, float dx,
                    float re, int scheme, void *stream) {
    FS2D_REQUIRE(vn && vc && pc && mask, "null field pointer");
    FS2D_REQUIRE(scheme == FS2D_SCHEME_UPWIND || scheme == FS2D_SCHEME_KK, "unknown advection scheme");
    if (int e = check_dom(d)) return e;
    if (d.r1 == d.r0) return FS2D_OK;
    const bool p2 = is_pow2(dx);
    const float dx2 = dx * dx;
#define MAC(P2, S) \
    ++g_launches; k_mac_update<P2, S><<<dense_grid(d), dense_block(), 0, STREAM>>>(vn, vc, pc, mask, d, dt, dx, DivC<P2>(dx), DivC<P2>(dx2), re)
    if (scheme == FS2D_SCHEME_UPWIND) DISPATCH_P2(p2, MAC(true, FS2D_SCHEME_UPWIND), MAC(false, FS2D_SCHEME_UPWIND));
    else DISPATCH_P2(p2, MAC(true, FS2D_SCHEME_KK), MAC(false, FS2D_SCHEME_KK));
#undef MAC
    FS2D_LAUNCH_CHECK();
    return FS2D_OK;
}

int fs2d_cip_nonadv(float *fn, const float *fc, const float *pc, const uint8_t *mask, fs2d_dom d, float dt, float dx,
                    float re, void *stream) {
    FS2D_REQUIRE(fn && fc && pc && mask, "null field pointer");
    if (int e = check_dom(d)) return e;
    if (d.r1 == d.r0) return FS2D_OK;
    const float dx2 = dx * dx;
#define NA(P2) ++g_launches, k_cip_nonadv<P2><<<dense_grid_nu(d, NU_CIP_NONADV), dense_block(), 0, STREAM>>>(fn, fc, pc, mask, d, dt, DivC<P2>(dx), DivC<P2>(dx2), re)
    DISPATCH_P2(is_pow2(dx), NA(true), NA(false));
#undef NA
    FS2D_LAUNCH_CHECK();
    return FS2D_OK;
}

int fs2d_cip_nonadv_grad(float *fxn, float *fyn, const float *fxc, const float *fyc, const float *fc, const float *fn,
                         const uint8_t *mask, fs2d_dom d, float two_dx, void *stream) {
    FS2D_REQUIRE(fxn && fyn && fxc && fyc && fc && fn && mask, "null field pointer");
    if (int e = check_dom(d)) return e;
    if (d.r1 == d.r0) return FS2D_OK;
#define NG(P2) ++g_launches, k_cip_nonadv_grad<P2><<<dense_grid_nu(d, NU_CIP_NONADV_GRAD), dense_block(), 0, STREAM>>>(fxn, fyn, fxc, fyc, fc, fn, mask, d, DivC<P2>(two_dx))
    DISPATCH_P2(is_pow2(two_dx), NG(true), NG(false));
#undef NG
    FS2D_LAUNCH_CHECK();
    return FS2D_OK;
}

int fs2d_cip_advect(float *fn, float *fxn, float *fyn, const float *fc, const float *fxc, const float *fyc,
                    const float *v, const uint8_t *mask, fs2d_dom d, float dt, float dx, float dx2, float dx3,
                    void *stream) {
    FS2D_REQUIRE(fn && fxn && fyn && fc && fxc && fyc && v && mask, "null field pointer");
    if (int e = check_dom(d)) return e;
    if (d.r1 == d.r0) return FS2D_OK;
    const bool p2 = is_pow2(dx) && is_pow2(dx2) && is_pow2(dx3);
#define CA(P2) ++g_launches, k_cip_advect<P2><<<dense_grid_nu(d, NU_CIP_ADVECT), dense_block(), 0, STREAM>>>(fn, fxn, fyn, fc, fxc, fyc, v, mask, d, dt, dx, DivC<P2>(dx), DivC<P2>(dx2), DivC<P2>(dx3))
    DISPATCH_P2(p2, CA(true), CA(false));
#undef CA
    FS2D_LAUNCH_CHECK();
    return FS2D_OK;
}

int fs2d_set_grad(float *fx, float *fy, const float *f, fs2d_dom d, float dx, void *stream) {
    FS2D_REQUIRE(fx && fy && f, "null field pointer");
    if (int e = check_dom(d)) return e;
    if (d.r1 == d.r0) return FS2D_OK;
#define SG(P2) ++g_launches, k_set_grad<P2><<<dense_grid(d), dense_block(), 0, STREAM>>>(fx, fy, f, d, DivC<P2>(dx))
    DISPATCH_P2(is_pow2(dx), SG(true), SG(false));
#undef SG
    FS2D_LAUNCH_CHECK();
    return FS2D_OK;
}

int fs2d_vort_calc(float *w, float *wabs, const float *vc, const uint8_t *mask, fs2d_dom d, float dx, void *stream) {
    FS2D_REQUIRE(w && wabs && vc && mask, "null field pointer");
    if (int e = check_dom(d)) return e;
    if (d.r1 == d.r0) return FS2D_OK;
#define VC(P2) ++g_launches, k_vort_calc<P2><<<dense_grid_nu(d, NU_VORT_CALC), dense_block(), 0, STREAM>>>(w, wabs, vc, mask, d, DivC<P2>(dx))
    DISPATCH_P2(is_pow2(dx), VC(true), VC(false));
#undef VC
    FS2D_LAUNCH_CHECK();
    return FS2D_OK;
}

int fs2d_vort_add(float *vn, const float *vc, const float *w, const float *wabs, const uint8_t *mask, fs2d_dom d,
                  float dx, float dtw, void *stream) {
    FS2D_REQUIRE(vn && vc && w && wabs && mask, "null field pointer");
    if (int e = check_dom(d)) return e;
    if (d.r1 == d.r0) return FS2D_OK;
#define VA(P2) ++g_launches, k_vort_add<P2><<<dense_grid_nu(d, NU_VORT_ADD), dense_block(), 0, STREAM>>>(vn, vc, w, wabs, mask, d, DivC<P2>(dx), dtw)
    DISPATCH_P2(is_pow2(dx), VA(true), VA(false));
#undef VA
    FS2D_LAUNCH_CHECK();
    return FS2D_OK;
}

int fs2d_limit(float *v, fs2d_dom d, float limit, void *stream) {
    FS2D_REQUIRE(v, "null field pointer");
    if (int e = check_dom(d)) return e;
    if (d.r1 == d.r0) return FS2D_OK;
    ++g_launches; k_limit<<<dense_grid_nu(d, NU_LIMIT), dense_block(), 0, STREAM>>>(v, d, limit);
    FS2D_LAUNCH_CHECK();
    return FS2D_OK;
}

}  // extern "C"

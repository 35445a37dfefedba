// fs2d_kernels.cu -- dense stencil kernels + sparse BC kernels of libfs2d.so (sm_100a).
//
// One kernel per reference Taichi kernel (file:line cited at each).  All kernels are pure
// 5-/9-/13-point fp32 stencils: HBM-bandwidth bound, no tensor cores.  Literal operation order of
// the reference source, compiled with -fmad=false, so outputs are bit-identical to the oracle.
#include <cstdarg>
#include <cstdio>
#include <cmath>

#include "fs2d_ops.cuh"

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
        in[u] = l_cip_nonadv(GAt<CL>{d, r[u], j}, fc, pc);
    }
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        const float2 out = c_cip_nonadv<P2>(in[u], dt, ddx, ddx2, re);
        if (ok[u] && m[u] != 1) reinterpret_cast<float2 *>(fn)[IX(d, r[u], j)] = out;
    }
}
// (asking for 6 / 8 resident blocks -- 40 / 32 registers -- measured 330 / 356 us against 330 us: left to the compiler)
template <bool P2>
__global__ void __launch_bounds__(TX *TY)
    k_cip_nonadv(float *__restrict__ fn, const float *__restrict__ fc, const float *__restrict__ pc,
                 const uint8_t *__restrict__ mask, fs2d_dom d, float dt, DivC<P2> ddx, DivC<P2> ddx2, float re) {
    if (block_interior(d, TY * NU_CIP_NONADV, 1)) b_cip_nonadv<P2, false>(fn, fc, pc, mask, d, dt, ddx, ddx2, re);
    else b_cip_nonadv<P2, true>(fn, fc, pc, mask, d, dt, ddx, ddx2, re);
}

// The same on FOUR CELLS PER THREAD: a warp covers 128 columns of one row; the rows r-1, r, r+1 of v are two 128-bit loads
// each and those of p one, the j-neighbours of the quad's end cells come from the adjacent lanes by shuffle (the warp's first /
// last lane loads them, or takes its own cell on a grid edge, as sample() clamps).  9 LDG.128 + 6 SHFL per 4 cells instead of
// 40 loads: the one-cell kernel spends its issue slots on loads and their addresses.  Per-cell arithmetic: c_cip_nonadv, the
// very function the other paths call.  Requires Y % 4 == 0 and 16-byte aligned fields (fs2d_set_tuning(6, 0): never used).
constexpr int NA4_WARPS = 8;
int g_nonadv_vec = 1;
template <bool P2>
__global__ void __launch_bounds__(32 * NA4_WARPS)
    k_cip_nonadv4(float *__restrict__ fn, const float *__restrict__ fc, const float *__restrict__ pc,
                  const uint8_t *__restrict__ mask, fs2d_dom d, float dt, DivC<P2> ddx, DivC<P2> ddx2, float re) {
    constexpr uint32_t FULL = 0xffffffffu;
    const int lane = threadIdx.x;
    const int r = d.r0 + FS2D_ROWBLK * NA4_WARPS + threadIdx.y;
    if (r >= d.r1) return;   // warp-uniform
    const int j0 = FS2D_COLBLK * 128 + 4 * lane;
    const bool active = j0 < d.Y;
    const int jc = active ? j0 : d.Y - 4;   // lanes past the grid read a valid quad (their values feed no active lane's result)
    auto vquad = [&](int row, float2 (&v)[4]) {
        const float4 *q = reinterpret_cast<const float4 *>(fc + 2 * IX(d, row, jc));
        const float4 a = __ldg(q), b = __ldg(q + 1);
        v[0] = make_float2(a.x, a.y); v[1] = make_float2(a.z, a.w); v[2] = make_float2(b.x, b.y); v[3] = make_float2(b.z, b.w);
    };
    auto pquad = [&](int row, float (&v)[4]) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(pc + IX(d, row, jc)));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    };
    float2 C[4], U[4], D[4];
    float pC[4], pU[4], pD[4];
    const int ru = CR(d, r + 1), rd = CR(d, r - 1);
    vquad(r, C); vquad(ru, U); vquad(rd, D);
    pquad(r, pC); pquad(ru, pU); pquad(rd, pD);
    const uchar4 mk = __ldg(reinterpret_cast<const uchar4 *>(mask + IX(d, r, jc)));
    float2 L, R;
    L.x = __shfl_up_sync(FULL, C[3].x, 1); L.y = __shfl_up_sync(FULL, C[3].y, 1);
    R.x = __shfl_down_sync(FULL, C[0].x, 1); R.y = __shfl_down_sync(FULL, C[0].y, 1);
    float pL = __shfl_up_sync(FULL, pC[3], 1), pR = __shfl_down_sync(FULL, pC[0], 1);
    if (lane == 0) {   // cell j0 - 1: outside the warp's span, or the cell itself on the grid's first column
        if (j0 == 0) { L = C[0]; pL = pC[0]; }
        else { L = __ldg(reinterpret_cast<const float2 *>(fc) + IX(d, r, j0 - 1)); pL = __ldg(pc + IX(d, r, j0 - 1)); }
    }
    if (lane == 31 || j0 + 4 >= d.Y) {   // cell j0 + 4
        if (j0 + 4 >= d.Y) { R = C[3]; pR = pC[3]; }
        else { R = __ldg(reinterpret_cast<const float2 *>(fc) + IX(d, r, j0 + 4)); pR = __ldg(pc + IX(d, r, j0 + 4)); }
    }
    if (!active) return;
    float2 out[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        NonadvIn x;
        x.c = C[q]; x.ip = U[q]; x.im = D[q];
        x.jp = q < 3 ? C[q < 3 ? q + 1 : 3] : R;
        x.jm = q > 0 ? C[q > 0 ? q - 1 : 0] : L;
        x.pip = pU[q]; x.pim = pD[q];
        x.pjp = q < 3 ? pC[q < 3 ? q + 1 : 3] : pR;
        x.pjm = q > 0 ? pC[q > 0 ? q - 1 : 0] : pL;
        out[q] = c_cip_nonadv<P2>(x, dt, ddx, ddx2, re);
    }
    float *dst = fn + 2 * IX(d, r, j0);
    if (mk.x != 1 && mk.y != 1 && mk.z != 1 && mk.w != 1) {
        float4 *q = reinterpret_cast<float4 *>(dst);
        q[0] = make_float4(out[0].x, out[0].y, out[1].x, out[1].y);
        q[1] = make_float4(out[2].x, out[2].y, out[3].x, out[3].y);
    } else {
        const uint8_t m[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (m[q] != 1) reinterpret_cast<float2 *>(dst)[q] = out[q];
    }
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
        o[u] = c_cip_advect<P2>(GAt<CL>{d, r[u], j}, fc, fxc, fyc, v, dt, dx, ddx, ddx2, ddx3);
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
        in[u] = l_vort_add(GAt<CL>{d, r[u], j}, vc, w, wabs);
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

// fs/vorticity_confinement.py:57-59  apply() = _calc_vorticity + _add_vorticity in ONE pass over HBM.
// The force at a fluid cell reads |curl v| at its four (clamped) neighbours; for a fluid neighbour that is the value
// _calc_vorticity is about to store there, so it is recomputed from v (same expression, same inputs: bit-identical);
// a non-fluid neighbour keeps its stored value (never written by _calc_vorticity), which is loaded.
// Traffic: v 8 + mask 1 read, vorticity 4 + |vorticity| 4 + v.next 8 written = 25 B/cell instead of 17 + 25.
// Shared-memory tiling: a block first evaluates the curl once for every cell of its (VA_ROWS + 2) x (TX + 2) tile
// (clamped coordinates, so the ring repeats the edge cells exactly like sample()), then the force from the tile.
constexpr int VA_ROWS = 16;               // tile rows per block (TY threads x 4 rows each)
constexpr int VA_PITCH = TX + 2 + 1;      // odd pitch: the column walk of phase 1 stays conflict-free
template <bool P2, bool CL>
__device__ __forceinline__ float c_curl(const float *vc, const fs2d_dom &d, int r, int j, DivC<P2> ddx) {
    return diff_x2<P2, CL>(vc, d, r, j, ddx).y - diff_y2<P2, CL>(vc, d, r, j, ddx).x;
}
template <bool P2, bool CL>
__device__ __forceinline__ void b_vort_apply(float *__restrict__ vn, float *__restrict__ w, float *__restrict__ wabs,
                                             const float *__restrict__ vc, const uint8_t *__restrict__ mask,
                                             const fs2d_dom &d, DivC<P2> ddx, float dtw, float (*s_w)[VA_PITCH],
                                             float (*s_a)[VA_PITCH]) {
    const int rb = d.r0 + FS2D_ROWBLK * VA_ROWS, jb = FS2D_COLBLK * TX;
    const int tid = threadIdx.y * TX + threadIdx.x;
    // phase 1: signed curl and effective |curl| (stored value at non-fluid cells) of the tile and its ring
    for (int e = tid; e < (VA_ROWS + 2) * (TX + 2); e += TX * TY) {
        const int lr = e / (TX + 2), lc = e - lr * (TX + 2);
        const int r = CL ? CR(d, rb - 1 + lr) : rb - 1 + lr, j = CL ? CJ(d, jb - 1 + lc) : jb - 1 + lc;
        const size_t idx = IX(d, r, j);
        // The curl is evaluated for every tile cell, fluid or not (its loads are clamped or interior, and a value that is
        // not wanted is dropped), so the velocity loads do not wait for the mask load: the kernel was bound by these two
        // serialised DRAM latencies per trip (431 -> 378 us at 8192^2).  Measured on top of it and dropped: unrolling the
        // trips (no gain), issuing phase 2's mask / velocity loads in one round or before phase 1 (404-425 us).
        const uint8_t m = __ldg(mask + idx);
        const float cv = c_curl<P2, CL>(vc, d, r, j, ddx);
        float cw = 0.0f, ca;
        if (m == 0) {
            cw = cv;
            ca = fabsf(cv);
        } else {
            ca = wabs[idx];   // never written by _calc_vorticity
        }
        s_w[lr][lc] = cw;
        s_a[lr][lc] = ca;
    }
    __syncthreads();
    // phase 2: the confinement force on the block's cells
    const int j = jb + threadIdx.x, lc = threadIdx.x + 1;
    if (j >= d.Y) return;
#pragma unroll
    for (int u = 0; u < VA_ROWS / TY; ++u) {
        const int lr = threadIdx.y + u * TY + 1, r = rb + lr - 1;
        if (r >= d.r1) break;
        const size_t idx = IX(d, r, j);
        if (__ldg(mask + idx) != 0) continue;
        VortIn x;
        x.aip = s_a[lr + 1][lc]; x.aim = s_a[lr - 1][lc]; x.ajp = s_a[lr][lc + 1]; x.ajm = s_a[lr][lc - 1];
        x.o = s_w[lr][lc];
        x.c = __ldg(reinterpret_cast<const float2 *>(vc) + idx);
        const float2 out = c_vort_add<P2>(x, ddx, dtw);
        w[idx] = x.o;
        wabs[idx] = fabsf(x.o);
        reinterpret_cast<float2 *>(vn)[idx] = out;
    }
}
template <bool P2>
__global__ void __launch_bounds__(TX *TY)
    k_vort_apply(float *__restrict__ vn, float *__restrict__ w, float *__restrict__ wabs, const float *__restrict__ vc,
                 const uint8_t *__restrict__ mask, fs2d_dom d, DivC<P2> ddx, float dtw) {
    __shared__ float s_w[VA_ROWS + 2][VA_PITCH], s_a[VA_ROWS + 2][VA_PITCH];
    // interior: the ring (1 cell) and the curl stencil of the ring (1 more) stay inside the clamp window
    const int rb = d.r0 + FS2D_ROWBLK * VA_ROWS, jb = FS2D_COLBLK * TX;
    const bool interior = rb - 2 >= d.clo && rb + VA_ROWS + 1 <= d.chi && jb - 2 >= 0 && jb + TX + 1 <= d.Y - 1;
    if (interior) b_vort_apply<P2, false>(vn, w, wabs, vc, mask, d, ddx, dtw, s_w, s_a);
    else b_vort_apply<P2, true>(vn, w, wabs, vc, mask, d, ddx, dtw, s_w, s_a);
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
    {
        const void *ptrs[] = {fn, fc, pc, mask};
        if (g_stream == 2 && stream_ok(d, ptrs, 4)) {   // measured no faster than the direct kernel (330 us): off by default
            if (int e = stream_cip_nonadv(fn, fc, pc, mask, d, dt, dx, re, is_pow2(dx), STREAM)) return e;
            FS2D_LAUNCH_CHECK();
            return FS2D_OK;
        }
    }
    const float dx2 = dx * dx;
    if (g_nonadv_vec && d.Y % 4 == 0 && (uintptr_t)fn % 16 == 0 && (uintptr_t)fc % 16 == 0 && (uintptr_t)pc % 16 == 0 &&
        (uintptr_t)mask % 4 == 0) {
        const dim3 g4((unsigned)((d.Y + 127) / 128), (unsigned)((d.r1 - d.r0 + NA4_WARPS - 1) / NA4_WARPS), 1), b4(32, NA4_WARPS, 1);
#define NA4(P2) ++g_launches, k_cip_nonadv4<P2><<<g4, b4, 0, STREAM>>>(fn, fc, pc, mask, d, dt, DivC<P2>(dx), DivC<P2>(dx2), re)
        DISPATCH_P2(is_pow2(dx), NA4(true), NA4(false));
#undef NA4
        FS2D_LAUNCH_CHECK();
        return FS2D_OK;
    }
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
    {
        const void *ptrs[] = {fn, fxn, fyn, fc, fxc, fyc, mask};
        if (v == fc && stream_ok(d, ptrs, 7)) {
            if (int e = stream_cip_advect(fn, fxn, fyn, fc, fxc, fyc, mask, d, dt, dx, dx2, dx3, p2, STREAM)) return e;
            FS2D_LAUNCH_CHECK();
            return FS2D_OK;
        }
    }
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

int fs2d_vort_apply(float *vn, float *w, float *wabs, const float *vc, const uint8_t *mask, fs2d_dom d, float dx, float dtw,
                    void *stream) {
    FS2D_REQUIRE(vn && w && wabs && vc && mask, "null field pointer");
    FS2D_REQUIRE(vn != vc, "the confinement force cannot be added in place");
    if (int e = check_dom(d)) return e;
    if (d.r1 == d.r0) return FS2D_OK;
    {
        const void *ptrs[] = {vn, w, wabs, vc, mask};
        if (g_stream == 2 && stream_ok(d, ptrs, 5)) {   // measured slower (534 us) than the tiled kernel below (465 us): opt-in
            if (int e = stream_vort_apply(vn, w, wabs, vc, mask, d, dx, dtw, is_pow2(dx), STREAM)) return e;
            FS2D_LAUNCH_CHECK();
            return FS2D_OK;
        }
    }
#define VP(P2) ++g_launches, k_vort_apply<P2><<<dense_grid_nu(d, VA_ROWS / TY), dense_block(), 0, STREAM>>>(vn, w, wabs, vc, mask, d, DivC<P2>(dx), dtw)
    DISPATCH_P2(is_pow2(dx), VP(true), VP(false));
#undef VP
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

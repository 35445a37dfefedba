// fs2d_dye.cu -- dye transport: the reference's Dye solvers run the same stencils on
// 3-channel AoS fields (fs/solver.py:110-161, :335-401).  Kernels here are generic in the channel count C
// (scalar loads per component, same per-component operation order as the float2 kernels in
// fs2d_kernels.cu), instantiated for C = 3.  Dye is off the benchmark path (`-no_dye`) but on the path main.py runs by
// default.  One cell (all C channels) per thread.  Measured and dropped in round 2: one CHANNEL of one cell per thread, threads
// ordered like the floats of the AoS row so that every warp access is 128 contiguous bytes -- bit-identical and 1.5-2.5x
// SLOWER (CIP advection 750 -> 1098 us, gradient update 416 -> 682, non-advection 230 -> 570 us at 8192 x 4096): the kernels
// are bound by instruction issue, and the per-thread index / mask / velocity work is then paid per float instead of per cell.
#include "fs2d_common.cuh"

namespace fs2d {

// CL = true: clamp-to-edge like sample(); CL = false: the caller guarantees that the access lies inside the clamp window
// (block_interior), so the neighbour is a plain offset -- the per-load min/max and 64-bit index arithmetic was a third of
// the dye kernels' instructions
template <int C, bool CL = true>
__device__ __forceinline__ float ldc(const float *f, const fs2d_dom &d, int r, int j, int c) {
    if (CL) return __ldg(f + (size_t)C * IX(d, CR(d, r), CJ(d, j)) + c);
    return __ldg(f + (ptrdiff_t)C * ((ptrdiff_t)r * d.Y + j) + c);
}

// fs/boundary_condition.py:94-99  set_dye_boundary_condition: dye = bc_dye on inflow cells (sparse list)
__global__ void k_dye_bc(float *__restrict__ dye, const float *__restrict__ bc_dye, const int32_t *__restrict__ tgt, int n) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const size_t o = 3 * (size_t)tgt[e];
    dye[o] = bc_dye[o];
    dye[o + 1] = bc_dye[o + 1];
    dye[o + 2] = bc_dye[o + 2];
}

// fs/solver.py:46-49  clamp_field (all cells, all components): ti.min(ti.max(f, low), high).  A value that is already inside
// [low, high] is not written back (same bits either way): after the first step almost every dye value is, so the pass is a
// read of the field (12 B/cell) instead of a read and a write ("already inside" = the clamped value has the same BITS, so NaN -> low and -0 -> what fmaxf makes of it are still stored).
__global__ void __launch_bounds__(256) k_clamp(float *__restrict__ f, size_t begin, size_t end, float low, float high) {
    const size_t k = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < end) {
        const float v = f[k], c = fminf(fmaxf(v, low), high);
        if (__float_as_uint(c) != __float_as_uint(v)) f[k] = c;
    }
}
// the same on 16-byte chunks [begin4, end4) of the array
__global__ void __launch_bounds__(256) k_clamp4(float4 *__restrict__ f, size_t begin4, size_t end4, float low, float high) {
    const size_t k = begin4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= end4) return;
    const float4 v = f[k];
    const float4 c = make_float4(fminf(fmaxf(v.x, low), high), fminf(fmaxf(v.y, low), high), fminf(fmaxf(v.z, low), high),
                                 fminf(fmaxf(v.w, low), high));
    if (__float_as_uint(c.x) != __float_as_uint(v.x) || __float_as_uint(c.y) != __float_as_uint(v.y) ||
        __float_as_uint(c.z) != __float_as_uint(v.z) || __float_as_uint(c.w) != __float_as_uint(v.w))
        f[k] = c;
}

// fs/solver.py:157-161  DyeMacSolver._update_dye: dn = dc - dt * advect(vc, dc)   (fluid cells)
template <bool P2, int SCHEME, int C>
__global__ void __launch_bounds__(TX *TY)
    k_dye_mac(float *__restrict__ dn, const float *__restrict__ dc, const float *__restrict__ vc,
              const uint8_t *__restrict__ mask, fs2d_dom d, float dt, float dx, DivC<P2> ddx) {
    FS2D_CELL(d, r, j)
    const size_t idx = IX(d, r, j);
    if (mask[idx] != 0) return;
    const float2 vel = __ldg(reinterpret_cast<const float2 *>(vc) + idx);
    const float six_dx = 6.0f * dx;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        float adv;
        if (SCHEME == FS2D_SCHEME_UPWIND) {  // fs/advection.py:12-24
            int k = vel.x < 0.0f ? r : r - 1;
            const float a = vel.x * ddx(ldc<C>(dc, d, k + 1, j, c) - ldc<C>(dc, d, k, j, c));
            k = vel.y < 0.0f ? j : j - 1;
            const float b = vel.y * ddx(ldc<C>(dc, d, r, k + 1, c) - ldc<C>(dc, d, r, k, c));
            adv = a + b;
        } else {  // fs/advection.py:27-60
            float k0, k1, k2, k3, k4;
            if (vel.x < 0.0f) { k0 = -2.0f; k1 = 10.0f; k2 = -9.0f; k3 = 2.0f; k4 = -1.0f; }
            else { k0 = 1.0f; k1 = -2.0f; k2 = 9.0f; k3 = -10.0f; k4 = 2.0f; }
            float acc = ldc<C>(dc, d, r + 2, j, c) * k0;
            acc = acc + ldc<C>(dc, d, r + 1, j, c) * k1;
            acc = acc + ldc<C>(dc, d, r, j, c) * k2;
            acc = acc + ldc<C>(dc, d, r - 1, j, c) * k3;
            acc = acc + ldc<C>(dc, d, r - 2, j, c) * k4;
            const float a = acc / six_dx;
            if (vel.y < 0.0f) { k0 = -2.0f; k1 = 10.0f; k2 = -9.0f; k3 = 2.0f; k4 = -1.0f; }
            else { k0 = 1.0f; k1 = -2.0f; k2 = 9.0f; k3 = -10.0f; k4 = 2.0f; }
            acc = ldc<C>(dc, d, r, j + 2, c) * k0;
            acc = acc + ldc<C>(dc, d, r, j + 1, c) * k1;
            acc = acc + ldc<C>(dc, d, r, j, c) * k2;
            acc = acc + ldc<C>(dc, d, r, j - 1, c) * k3;
            acc = acc + ldc<C>(dc, d, r, j - 2, c) * k4;
            const float b = acc / six_dx;
            adv = vel.x * a + vel.y * b;
        }
        dn[C * idx + c] = __ldg(dc + C * idx + c) - dt * adv;
    }
}

// fs/solver.py:378-383  _non_advection_phase_dye: dn = dc + diffusion(dc) * dt   (not-wall cells)
template <bool P2, int C, bool CL>
__device__ __forceinline__ void b_dye_nonadv(float *__restrict__ dn, const float *__restrict__ dc, const uint8_t *__restrict__ mask,
                                             const fs2d_dom &d, float dt, DivC<P2> ddx2, float re) {
    FS2D_CELL(d, r, j)
    const size_t idx = IX(d, r, j);
    const uint8_t m = __ldg(mask + idx);   // consulted at the stores: the field loads do not wait for it
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float cc = __ldg(dc + C * idx + c);
        const float d2x = ddx2(ldc<C, CL>(dc, d, r + 1, j, c) - 2.0f * cc + ldc<C, CL>(dc, d, r - 1, j, c));
        const float d2y = ddx2(ldc<C, CL>(dc, d, r, j + 1, c) - 2.0f * cc + ldc<C, CL>(dc, d, r, j - 1, c));
        const float out = cc + fdiv_z(d2x + d2y, re) * dt;
        if (m != 1) dn[C * idx + c] = out;
    }
}
template <bool P2, int C>
__global__ void __launch_bounds__(TX *TY)
    k_dye_nonadv(float *__restrict__ dn, const float *__restrict__ dc, const uint8_t *__restrict__ mask, fs2d_dom d,
                 float dt, DivC<P2> ddx2, float re) {
    if (block_interior(d, TY, 1)) b_dye_nonadv<P2, C, false>(dn, dc, mask, d, dt, ddx2, re);
    else b_dye_nonadv<P2, C, true>(dn, dc, mask, d, dt, ddx2, re);
}

// The same on FOUR CELLS PER THREAD (3 channels x 4 cells = 12 consecutive floats of the AoS row = three 16-byte words): a warp
// covers 128 columns of one row, the rows above / below are three 128-bit loads each, the j-neighbours of the quad's end cells
// come from the adjacent lanes by shuffle (the warp's first / last lane loads them, or takes its own cell on a grid edge, as
// sample() clamps).  9 LDG.128 + 6 SHFL per 4 cells instead of 60 scalar loads: the one-cell kernel is bound by instruction
// issue.  Same expression and order per value.  Requires Y % 4 == 0 and 16-byte aligned fields.
constexpr int DN4_WARPS = 8;
int g_dye_vec = 1;   // fs2d_set_tuning(5, v): 0 = always the one-cell-per-thread non-advection kernel
template <bool P2>
__global__ void __launch_bounds__(32 * DN4_WARPS)
    k_dye_nonadv4(float *__restrict__ dn, const float *__restrict__ dc, const uint8_t *__restrict__ mask, fs2d_dom d, float dt,
                  DivC<P2> ddx2, float re) {
    constexpr uint32_t FULL = 0xffffffffu;
    const int lane = threadIdx.x;
    const int r = d.r0 + FS2D_ROWBLK * DN4_WARPS + threadIdx.y;
    if (r >= d.r1) return;   // warp-uniform
    const int j0 = FS2D_COLBLK * 128 + 4 * lane;
    const bool active = j0 < d.Y;
    const int jc = active ? j0 : d.Y - 4;   // lanes past the grid read a valid quad (their values feed no active lane's result)
    auto quad = [&](int row, float (&v)[12]) {
        const float4 *q = reinterpret_cast<const float4 *>(dc + 3 * IX(d, row, jc));
        const float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w;
    };
    float C[12], U[12], D[12], L[3], R[3];
    quad(r, C);
    quad(CR(d, r + 1), U);
    quad(CR(d, r - 1), D);
    const uchar4 mk = __ldg(reinterpret_cast<const uchar4 *>(mask + IX(d, r, jc)));
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        L[c] = __shfl_up_sync(FULL, C[9 + c], 1);
        R[c] = __shfl_down_sync(FULL, C[c], 1);
    }
    if (lane == 0) {   // cell j0 - 1: outside the warp's span, or the cell itself on the grid's first column
#pragma unroll
        for (int c = 0; c < 3; ++c) L[c] = j0 == 0 ? C[c] : __ldg(dc + 3 * IX(d, r, j0 - 1) + c);
    }
    if (lane == 31 || j0 + 4 >= d.Y) {   // cell j0 + 4
#pragma unroll
        for (int c = 0; c < 3; ++c) R[c] = j0 + 4 >= d.Y ? C[9 + c] : __ldg(dc + 3 * IX(d, r, j0 + 4) + c);
    }
    if (!active) return;
    float out[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) {
        const float cc = C[e];
        const float jp = e < 9 ? C[e + 3] : R[e - 9], jm = e >= 3 ? C[e - 3] : L[e];
        const float d2x = ddx2(U[e] - 2.0f * cc + D[e]);
        const float d2y = ddx2(jp - 2.0f * cc + jm);
        out[e] = cc + fdiv_z(d2x + d2y, re) * dt;
    }
    float *dst = dn + 3 * IX(d, r, j0);
    if (mk.x != 1 && mk.y != 1 && mk.z != 1 && mk.w != 1) {
        float4 *q = reinterpret_cast<float4 *>(dst);
        q[0] = make_float4(out[0], out[1], out[2], out[3]);
        q[1] = make_float4(out[4], out[5], out[6], out[7]);
        q[2] = make_float4(out[8], out[9], out[10], out[11]);
    } else {
        const uint8_t m[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (m[q] != 1) {
                dst[3 * q] = out[3 * q];
                dst[3 * q + 1] = out[3 * q + 1];
                dst[3 * q + 2] = out[3 * q + 2];
            }
    }
}

// fs/solver.py:242-261  _non_advection_phase_grad on C channels
template <bool P2, int C, bool CL>
__device__ __forceinline__ void b_nonadv_grad_n(float *__restrict__ fxn, float *__restrict__ fyn, const float *__restrict__ fxc,
                                                const float *__restrict__ fyc, const float *__restrict__ fc,
                                                const float *__restrict__ fn, const uint8_t *__restrict__ mask, const fs2d_dom &d,
                                                DivC<P2> d2dx) {
    FS2D_CELL(d, r, j)
    const size_t idx = IX(d, r, j);
    const uint8_t m = __ldg(mask + idx);   // consulted at the stores: the field loads do not wait for it
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float gx = ldc<C, CL>(fn, d, r + 1, j, c) - ldc<C, CL>(fc, d, r + 1, j, c) - ldc<C, CL>(fn, d, r - 1, j, c) + ldc<C, CL>(fc, d, r - 1, j, c);
        const float gy = ldc<C, CL>(fn, d, r, j + 1, c) - ldc<C, CL>(fc, d, r, j + 1, c) - ldc<C, CL>(fn, d, r, j - 1, c) + ldc<C, CL>(fc, d, r, j - 1, c);
        const float ox = __ldg(fxc + C * idx + c) + d2dx(gx), oy = __ldg(fyc + C * idx + c) + d2dx(gy);
        if (m != 1) {
            fxn[C * idx + c] = ox;
            fyn[C * idx + c] = oy;
        }
    }
}
template <bool P2, int C>
__global__ void __launch_bounds__(TX *TY)
    k_nonadv_grad_n(float *__restrict__ fxn, float *__restrict__ fyn, const float *__restrict__ fxc,
                    const float *__restrict__ fyc, const float *__restrict__ fc, const float *__restrict__ fn,
                    const uint8_t *__restrict__ mask, fs2d_dom d, DivC<P2> d2dx) {
    if (block_interior(d, TY, 1)) b_nonadv_grad_n<P2, C, false>(fxn, fyn, fxc, fyc, fc, fn, mask, d, d2dx);
    else b_nonadv_grad_n<P2, C, true>(fxn, fyn, fxc, fyc, fc, fn, mask, d, d2dx);
}

// fs/solver.py:267-332  _advection_phase/_cip_advect on C channels, advecting velocity v (float2)
template <bool P2, int C, bool CL>
__device__ __forceinline__ void b_cip_advect_n(float *__restrict__ fn, float *__restrict__ fxn, float *__restrict__ fyn,
                                               const float *__restrict__ fc, const float *__restrict__ fxc,
                                               const float *__restrict__ fyc, const float *__restrict__ v,
                                               const uint8_t *__restrict__ mask, const fs2d_dom &d, float dt, float dx,
                                               DivC<P2> ddx, DivC<P2> ddx2, DivC<P2> ddx3) {
    FS2D_CELL(d, r, j)
    const size_t idx = IX(d, r, j);
    const uint8_t m = __ldg(mask + idx);   // consulted at the stores: the velocity load (which every other load waits for) does not wait for it
    const float2 vel = __ldg(reinterpret_cast<const float2 *>(v) + idx);
    const float i_s = sign1(vel.x), j_s = sign1(vel.y);
    const int r_m = r - (int)i_s, j_m = j - (int)j_s;
    const DivC<P2> disd = ddx3.signed_by(i_s), djsd = ddx3.signed_by(j_s), disdx = ddx.signed_by(i_s);   // / (i_s * dx3) etc.
    const float dx2 = ddx2.c;
    const float Xd = -vel.x * dt, Yd = -vel.y * dt;
    const float2 dxv = ddx(0.5f * (ld2<CL>(v, d, r + 1, j) - ld2<CL>(v, d, r - 1, j)));
    const float2 dyv = ddx(0.5f * (ld2<CL>(v, d, r, j + 1) - ld2<CL>(v, d, r, j - 1)));
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float f00 = ldc<C, CL>(fc, d, r, j, c), f0m = ldc<C, CL>(fc, d, r, j_m, c), fm0 = ldc<C, CL>(fc, d, r_m, j, c), fmm = ldc<C, CL>(fc, d, r_m, j_m, c);
        const float x00 = ldc<C, CL>(fxc, d, r, j, c), x0m = ldc<C, CL>(fxc, d, r, j_m, c), xm0 = ldc<C, CL>(fxc, d, r_m, j, c);
        const float y00 = ldc<C, CL>(fyc, d, r, j, c), y0m = ldc<C, CL>(fyc, d, r, j_m, c), ym0 = ldc<C, CL>(fyc, d, r_m, j, c);
        const float tmp1 = f00 - f0m - fm0 + fmm;
        const float tmp2 = fm0 - f00;
        const float tmp3 = f0m - f00;
        const float a = disd(i_s * (xm0 + x00) * dx - 2.0f * (-tmp2));
        const float b = djsd(j_s * (y0m + y00) * dx - 2.0f * (-tmp3));
        const float cc = djsd(-tmp1 - i_s * (x0m - x00) * dx);
        const float dd = disd(-tmp1 - j_s * (ym0 - y00) * dx);
        const float e = ddx2(3.0f * tmp2 + i_s * (xm0 + 2.0f * x00) * dx);
        const float f = ddx2(3.0f * tmp3 + j_s * (y0m + 2.0f * y00) * dx);
        const float g = disdx(-(ym0 - y00) + cc * dx2);
        const float F = ((a * Xd + cc * Yd + e) * Xd + g * Yd + x00) * Xd + ((b * Yd + dd * Xd + f) * Yd + y00) * Yd + f00;
        const float Fx = (3.0f * a * Xd + 2.0f * cc * Yd + 2.0f * e) * Xd + (dd * Yd + g) * Yd + x00;
        const float Fy = (3.0f * b * Yd + 2.0f * dd * Xd + 2.0f * f) * Yd + (cc * Xd + g) * Xd + y00;
        if (m == 0) {
            fn[C * idx + c] = F;
            fxn[C * idx + c] = Fx - dt * (Fx * dxv.x + Fy * dxv.y) / 2.0f;
            fyn[C * idx + c] = Fy - dt * (Fx * dyv.x + Fy * dyv.y) / 2.0f;
        }
    }
}
template <bool P2, int C>
__global__ void __launch_bounds__(TX *TY, 5)
    k_cip_advect_n(float *__restrict__ fn, float *__restrict__ fxn, float *__restrict__ fyn,
                   const float *__restrict__ fc, const float *__restrict__ fxc, const float *__restrict__ fyc,
                   const float *__restrict__ v, const uint8_t *__restrict__ mask, fs2d_dom d, float dt, float dx,
                   DivC<P2> ddx, DivC<P2> ddx2, DivC<P2> ddx3) {
    if (block_interior(d, TY, 1)) b_cip_advect_n<P2, C, false>(fn, fxn, fyn, fc, fxc, fyc, v, mask, d, dt, dx, ddx, ddx2, ddx3);
    else b_cip_advect_n<P2, C, true>(fn, fxn, fyn, fc, fxc, fyc, v, mask, d, dt, dx, ddx, ddx2, ddx3);
}

// fs/solver.py:207-211  _set_grad on C channels
template <bool P2, int C>
__global__ void __launch_bounds__(TX *TY)
    k_set_grad_n(float *__restrict__ fx, float *__restrict__ fy, const float *__restrict__ f, fs2d_dom d, DivC<P2> ddx) {
    FS2D_CELL(d, r, j)
    const size_t idx = IX(d, r, j);
#pragma unroll
    for (int c = 0; c < C; ++c) {
        fx[C * idx + c] = ddx(0.5f * (ldc<C>(f, d, r + 1, j, c) - ldc<C>(f, d, r - 1, j, c)));
        fy[C * idx + c] = ddx(0.5f * (ldc<C>(f, d, r, j + 1, c) - ldc<C>(f, d, r, j - 1, c)));
    }
}

// fs/fluid_simulator.py:38-58, :121-126 + fs/visualization.py:8-22: field -> RGB image, wall colour override
template <bool P2, int MODE>
__global__ void __launch_bounds__(TX *TY)
    k_render(float *__restrict__ rgb, const float *__restrict__ v, const float *__restrict__ p, const float *__restrict__ dye,
             const uint8_t *__restrict__ mask, fs2d_dom d, DivC<P2> ddx) {
    FS2D_CELL(d, r, j)
    const size_t idx = IX(d, r, j);
    float o0, o1, o2;
    if (MODE == 0) {
        const float2 c = __ldg(reinterpret_cast<const float2 *>(v) + idx);
        const float n = sqrtf(c.x * c.x + c.y * c.y), pv = __ldg(p + idx);
        o0 = 0.2f * n + 0.002f * fmaxf(pv, 0.0f);
        o1 = 0.2f * n + 0.002f * 0.0f;
        o2 = 0.2f * n + 0.002f * fmaxf(-pv, 0.0f);
    } else if (MODE == 1) {
        const float pv = __ldg(p + idx);
        o0 = 0.04f * fmaxf(pv, 0.0f); o1 = 0.04f * 0.0f; o2 = 0.04f * fmaxf(-pv, 0.0f);
    } else if (MODE == 2) {
        const float val = ddx(0.5f * (ld2(v, d, r + 1, j) - ld2(v, d, r - 1, j))).y - ddx(0.5f * (ld2(v, d, r, j + 1) - ld2(v, d, r, j - 1))).x;
        o0 = 0.005f * fmaxf(val, 0.0f); o1 = 0.005f * 0.0f; o2 = 0.005f * fmaxf(-val, 0.0f);
    } else {
        o0 = __ldg(dye + 3 * idx); o1 = __ldg(dye + 3 * idx + 1); o2 = __ldg(dye + 3 * idx + 2);
    }
    if (mask[idx] == 1) { o0 = 0.5f; o1 = 0.7f; o2 = 0.5f; }
    rgb[3 * idx] = o0; rgb[3 * idx + 1] = o1; rgb[3 * idx + 2] = o2;
}

}  // namespace fs2d

using namespace fs2d;
#define STREAM ((cudaStream_t)stream)
#define P2_DISPATCH(p2, T, F) \
    do {                      \
        ++g_launches;         \
        if (p2) { T; } else { F; } \
    } while (0)

extern "C" {

int fs2d_dye_bc(float *dye, const float *bc_dye, const int32_t *tgt, int n, void *stream) {
    if (n == 0) return FS2D_OK;
    FS2D_REQUIRE(dye && bc_dye && tgt && n > 0, "null table/field pointer");
    ++g_launches;
    k_dye_bc<<<nblk(n, 256), 256, 0, STREAM>>>(dye, bc_dye, tgt, n);
    FS2D_LAUNCH_CHECK();
    return FS2D_OK;
}

int fs2d_clamp(float *f, fs2d_dom d, int channels, float low, float high, void *stream) {
    FS2D_REQUIRE(f && channels >= 1, "null field pointer");
    if (int e = check_dom(d)) return e;
    if (d.r1 == d.r0) return FS2D_OK;
    const size_t begin = (size_t)d.r0 * d.Y * channels, end = (size_t)d.r1 * d.Y * channels;
    ++g_launches;
    if ((uintptr_t)f % 16 == 0 && begin % 4 == 0 && end % 4 == 0)
        k_clamp4<<<(unsigned)(((end - begin) / 4 + 255) / 256), 256, 0, STREAM>>>(reinterpret_cast<float4 *>(f), begin / 4, end / 4, low, high);
    else
        k_clamp<<<(unsigned)((end - begin + 255) / 256), 256, 0, STREAM>>>(f, begin, end, low, high);
    FS2D_LAUNCH_CHECK();
    return FS2D_OK;
}

int fs2d_dye_mac(float *dn, const float *dc, const float *vc, const uint8_t *mask, fs2d_dom d, float dt, float dx, int scheme,
                 void *stream) {
    FS2D_REQUIRE(dn && dc && vc && mask, "null field pointer");
    FS2D_REQUIRE(scheme == FS2D_SCHEME_UPWIND || scheme == FS2D_SCHEME_KK, "unknown advection scheme");
    if (int e = check_dom(d)) return e;
    if (d.r1 == d.r0) return FS2D_OK;
#define DM(P2, S) k_dye_mac<P2, S, 3><<<dense_grid(d), dense_block(), 0, STREAM>>>(dn, dc, vc, mask, d, dt, dx, DivC<P2>(dx))
    if (scheme == FS2D_SCHEME_UPWIND) P2_DISPATCH(is_pow2(dx), DM(true, FS2D_SCHEME_UPWIND), DM(false, FS2D_SCHEME_UPWIND));
    else P2_DISPATCH(is_pow2(dx), DM(true, FS2D_SCHEME_KK), DM(false, FS2D_SCHEME_KK));
#undef DM
    FS2D_LAUNCH_CHECK();
    return FS2D_OK;
}

int fs2d_dye_nonadv(float *dn, const float *dc, const uint8_t *mask, fs2d_dom d, float dt, float dx, float re, void *stream) {
    FS2D_REQUIRE(dn && dc && mask, "null field pointer");
    if (int e = check_dom(d)) return e;
    if (d.r1 == d.r0) return FS2D_OK;
    const float dx2 = dx * dx;
    const bool vec = g_dye_vec && d.Y % 4 == 0 && (uintptr_t)dn % 16 == 0 && (uintptr_t)dc % 16 == 0 && (uintptr_t)mask % 4 == 0;
    const dim3 g4((unsigned)((d.Y + 127) / 128), (unsigned)((d.r1 - d.r0 + DN4_WARPS - 1) / DN4_WARPS), 1), b4(32, DN4_WARPS, 1);
#define DN(P2)                                                                                                 \
    do {                                                                                                       \
        if (vec) k_dye_nonadv4<P2><<<g4, b4, 0, STREAM>>>(dn, dc, mask, d, dt, DivC<P2>(dx2), re);               \
        else k_dye_nonadv<P2, 3><<<dense_grid(d), dense_block(), 0, STREAM>>>(dn, dc, mask, d, dt, DivC<P2>(dx2), re); \
    } while (0)
    P2_DISPATCH(is_pow2(dx), DN(true), DN(false));
#undef DN
    FS2D_LAUNCH_CHECK();
    return FS2D_OK;
}

int fs2d_dye_nonadv_grad(float *fxn, float *fyn, const float *fxc, const float *fyc, const float *fc, const float *fn,
                         const uint8_t *mask, fs2d_dom d, float two_dx, void *stream) {
    FS2D_REQUIRE(fxn && fyn && fxc && fyc && fc && fn && mask, "null field pointer");
    if (int e = check_dom(d)) return e;
    if (d.r1 == d.r0) return FS2D_OK;
#define NG(P2) k_nonadv_grad_n<P2, 3><<<dense_grid(d), dense_block(), 0, STREAM>>>(fxn, fyn, fxc, fyc, fc, fn, mask, d, DivC<P2>(two_dx))
    P2_DISPATCH(is_pow2(two_dx), NG(true), NG(false));
#undef NG
    FS2D_LAUNCH_CHECK();
    return FS2D_OK;
}

int fs2d_dye_cip_advect(float *fn, float *fxn, float *fyn, const float *fc, const float *fxc, const float *fyc,
                        const float *v, const uint8_t *mask, fs2d_dom d, float dt, float dx, float dx2, float dx3,
                        void *stream) {
    FS2D_REQUIRE(fn && fxn && fyn && fc && fxc && fyc && v && mask, "null field pointer");
    if (int e = check_dom(d)) return e;
    if (d.r1 == d.r0) return FS2D_OK;
    const bool p2 = is_pow2(dx) && is_pow2(dx2) && is_pow2(dx3);
#define CA(P2) k_cip_advect_n<P2, 3><<<dense_grid(d), dense_block(), 0, STREAM>>>(fn, fxn, fyn, fc, fxc, fyc, v, mask, d, dt, dx, DivC<P2>(dx), DivC<P2>(dx2), DivC<P2>(dx3))
    P2_DISPATCH(p2, CA(true), CA(false));
#undef CA
    FS2D_LAUNCH_CHECK();
    return FS2D_OK;
}

int fs2d_dye_set_grad(float *fx, float *fy, const float *f, fs2d_dom d, float dx, void *stream) {
    FS2D_REQUIRE(fx && fy && f, "null field pointer");
    if (int e = check_dom(d)) return e;
    if (d.r1 == d.r0) return FS2D_OK;
#define SG(P2) k_set_grad_n<P2, 3><<<dense_grid(d), dense_block(), 0, STREAM>>>(fx, fy, f, d, DivC<P2>(dx))
    P2_DISPATCH(is_pow2(dx), SG(true), SG(false));
#undef SG
    FS2D_LAUNCH_CHECK();
    return FS2D_OK;
}

int fs2d_render(float *rgb, const float *v, const float *p, const float *dye, const uint8_t *mask, fs2d_dom d, float dx,
                int mode, void *stream) {
    FS2D_REQUIRE(rgb && mask && mode >= 0 && mode <= 3, "bad render arguments");
    FS2D_REQUIRE((mode == 0 && v && p) || (mode == 1 && p) || (mode == 2 && v) || (mode == 3 && dye), "missing input field");
    if (int e = check_dom(d)) return e;
    if (d.r1 == d.r0) return FS2D_OK;
#define RD(P2, M) k_render<P2, M><<<dense_grid(d), dense_block(), 0, STREAM>>>(rgb, v, p, dye, mask, d, DivC<P2>(dx))
    const bool p2 = is_pow2(dx);
    switch (mode) {
        case 0: P2_DISPATCH(p2, RD(true, 0), RD(false, 0)); break;
        case 1: P2_DISPATCH(p2, RD(true, 1), RD(false, 1)); break;
        case 2: P2_DISPATCH(p2, RD(true, 2), RD(false, 2)); break;
        default: P2_DISPATCH(p2, RD(true, 3), RD(false, 3)); break;
    }
#undef RD
    FS2D_LAUNCH_CHECK();
    return FS2D_OK;
}

}  // extern "C"

// fs2d_vort_march.cu -- EXPERIMENTAL (off by default, fs2d_set_tuning(5, 1)): VorticityConfinement.apply()
// (fs/vorticity_confinement.py:57-59 = _calc_vorticity :27-32 + _add_vorticity :34-55) as a MARCHING kernel, sm_100a.
//
// k_vort_apply (fs2d_kernels.cu) evaluates the curl on a shared-memory tile + ring and is issue-bound: ~230 instructions
// per cell (tile index arithmetic, 1.16 curl evaluations per cell, shared-memory traffic, two barriers) for 25 B/cell.
// Here a warp owns 32 consecutive columns (lanes 1..30 produce output, lanes 0 and 31 are halo lanes of the neighbouring
// warps' columns) and walks down VM_ROWS rows keeping v(r), v(r+1), v(r+2) and the curl / |curl| of rows r-1, r, r+1 in
// registers: ONE curl evaluation per cell, j-neighbours by warp shuffles, no shared memory, no barriers.
//
// Exactly the arithmetic of k_vort_apply (same c_vort_add, same curl expression): bit-identical results.
//   * sample() clamping (fs/differentiation.py:4-9) is explicit: a neighbour outside the clamp window [clo, chi] x [0, Y-1]
//     is the cell itself, so its |curl| is the cell's own; loads of v use clamped indices.
//   * a non-fluid neighbour contributes the value stored in vorticity_abs (never written by _calc_vorticity).
#include "fs2d_ops.cuh"

namespace fs2d {

constexpr int VM_ROWS = 32;    // rows marched by one warp
constexpr int VM_WARPS = 4;    // warps per block (each on its own row chunk)
constexpr int VM_COLS = 30;    // output columns per warp (32 lanes - 2 halo lanes)

int g_vort_march = 0;          // fs2d_set_tuning(5, v)

template <bool P2>
__global__ void __launch_bounds__(32 * VM_WARPS)
    k_vort_march(float *__restrict__ vn, float *__restrict__ w, float *__restrict__ wabs, const float *__restrict__ vc,
                 const uint8_t *__restrict__ mask, fs2d_dom d, DivC<P2> ddx, float dtw) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x;
    const int j = (int)blockIdx.x * VM_COLS - 1 + lane;          // this lane's column (may lie outside the grid: halo lanes)
    const int jc = CJ(d, j);                                     // column it loads
    const bool col_out = lane >= 1 && lane <= VM_COLS && j >= 0 && j < d.Y;   // lane produces output
    const int rs = d.r0 + ((int)blockIdx.y * VM_WARPS + (int)threadIdx.y) * VM_ROWS;
    const int re = min(rs + VM_ROWS, d.r1);
    if (rs >= re) return;                                        // warp-uniform
    const float2 *v2 = reinterpret_cast<const float2 *>(vc);

    // |curl| that a neighbour sees at cell (q, jc): the value _calc_vorticity stores there if the cell is fluid, else the
    // stored one.  vm / vq / vp = v(CR(q-1)), v(q), v(CR(q+1)) at column jc.  All lanes call this together (shuffles).
    auto curl_at = [&](int q, float2 vm, float2 vq, float2 vp, float &c_out, float &a_out) {
        float xl = __shfl_up_sync(FULL, vq.x, 1), xr = __shfl_down_sync(FULL, vq.x, 1);
        if (lane == 0) xl = __ldg(v2 + IX(d, q, CJ(d, jc - 1))).x;        // no lane to the left / right: load
        if (lane == 31) xr = __ldg(v2 + IX(d, q, CJ(d, jc + 1))).x;
        if (j - 1 < 0) xl = vq.x;                                         // sample() clamps: the neighbour is the cell itself
        if (j + 1 > d.Y - 1) xr = vq.x;
        const float c = ddx(0.5f * (vp.y - vm.y)) - ddx(0.5f * (xr - xl));   // diff_x(v).y - diff_y(v).x
        const size_t idx = IX(d, q, jc);
        c_out = c;
        a_out = __ldg(mask + idx) == 0 ? fabsf(c) : wabs[idx];
    };

    // prologue: rows rs-1 (if inside the clamp window) and rs
    float2 v_m = __ldg(v2 + IX(d, CR(d, rs - 1), jc)), v_c = __ldg(v2 + IX(d, rs, jc)), v_p = __ldg(v2 + IX(d, CR(d, rs + 1), jc));
    float c_c, a_c, a_m = 0.0f, c_dummy;
    if (rs - 1 >= d.clo) {   // warp-uniform
        const float2 v_mm = __ldg(v2 + IX(d, CR(d, rs - 2), jc));
        curl_at(rs - 1, v_mm, v_m, v_c, c_dummy, a_m);
    }
    curl_at(rs, v_m, v_c, v_p, c_c, a_c);

    for (int r = rs; r < re; ++r) {
        // row r + 1 (if inside the clamp window): its v rows, curl and effective |curl|
        float c_p = 0.0f, a_p = 0.0f;
        float2 v_pp = v_p;
        if (r + 1 <= d.chi) {   // warp-uniform
            v_pp = __ldg(v2 + IX(d, CR(d, r + 2), jc));
            curl_at(r + 1, v_c, v_p, v_pp, c_p, a_p);
        }
        // neighbours of (r, j) as sample() sees them
        float a_jm = __shfl_up_sync(FULL, a_c, 1), a_jp = __shfl_down_sync(FULL, a_c, 1);
        if (j - 1 < 0) a_jm = a_c;
        if (j + 1 > d.Y - 1) a_jp = a_c;
        const float a_im = r - 1 < d.clo ? a_c : a_m, a_ip = r + 1 > d.chi ? a_c : a_p;
        if (col_out) {
            const size_t idx = IX(d, r, j);
            if (__ldg(mask + idx) == 0) {
                VortIn x;
                x.aip = a_ip; x.aim = a_im; x.ajp = a_jp; x.ajm = a_jm;
                x.o = c_c;
                x.c = v_c;
                const float2 out = c_vort_add<P2>(x, ddx, dtw);
                w[idx] = c_c;
                wabs[idx] = fabsf(c_c);
                reinterpret_cast<float2 *>(vn)[idx] = out;
            }
        }
        a_m = a_c; a_c = a_p; c_c = c_p;
        v_m = v_c; v_c = v_p; v_p = v_pp;
    }
}

int vort_march(float *vn, float *w, float *wabs, const float *vc, const uint8_t *mask, const fs2d_dom &d, float dx, float dtw,
               cudaStream_t s) {
    const dim3 block(32, VM_WARPS, 1);
    const dim3 grid(nblk(d.Y, VM_COLS), nblk(d.r1 - d.r0, VM_ROWS * VM_WARPS), 1);
    ++g_launches;
    if (is_pow2(dx)) k_vort_march<true><<<grid, block, 0, s>>>(vn, w, wabs, vc, mask, d, DivC<true>(dx), dtw);
    else k_vort_march<false><<<grid, block, 0, s>>>(vn, w, wabs, vc, mask, d, DivC<false>(dx), dtw);
    return FS2D_OK;
}

}  // namespace fs2d

// fs2d_pressure.cu -- pressure relaxation kernels of libfs2d.so (sm_100a).
//
// predict_p (fs/pressure_updater.py:23-38):
//     p' = 0.25*(p(i+1,j) + p(i-1,j) + p(i,j+1) + p(i,j-1)) + t2 - t3
//     t2 = (sx.x^2 + sy.y^2 + sy.x*sx.y)/8,  t3 = dx*(sx.x + sy.y)/(8*dt),  sx = v(i+1,j)-v(i-1,j), sy = v(i,j+1)-v(i,j-1)
// The velocity does not change between the sweeps of one pressure update (:56-60), so (t2, t3) is
// computed ONCE per update into a float2 "source" array by k_p_source and every sweep evaluates the
// literal expression (t1 + t2) - t3 from it: bit-identical to recomputing it from v per sweep, but the
// sweep itself shrinks to a 5-point stencil with no divisions.
//
// Bytes per cell per sweep: p 4 + src 8 + pcode 1 read, p 4 write = 17 (algorithmic figure used for the
// roofline is 12).  The fused multi-sweep kernel (k_jacobi_fused) reads the same 17 B once
// per T sweeps.
#include "fs2d_common.cuh"

namespace fs2d {

// ---------------------------------------------------------------------------------------------
// source terms
// ---------------------------------------------------------------------------------------------
constexpr int NU_P_SOURCE = 4;   // rows per thread
template <bool CL>
__device__ __forceinline__ void b_p_source(float *__restrict__ src, const float *__restrict__ vc, const fs2d_dom &d, float dt,
                                           float dx) {
    const int j = FS2D_COLBLK * blockDim.x + threadIdx.x;
    if (j >= d.Y) return;
    float2 sx[NU_P_SOURCE], sy[NU_P_SOURCE];
    int r[NU_P_SOURCE];
    bool ok[NU_P_SOURCE];
#pragma unroll
    for (int u = 0; u < NU_P_SOURCE; ++u) {   // all loads of the thread's rows first (memory-level parallelism)
        const int rr = d.r0 + (FS2D_ROWBLK * NU_P_SOURCE + u) * blockDim.y + threadIdx.y;
        ok[u] = rr < d.r1;
        r[u] = ok[u] ? rr : d.r1 - 1;
        sx[u] = ld2<CL>(vc, d, r[u] + 1, j) - ld2<CL>(vc, d, r[u] - 1, j);
        sy[u] = ld2<CL>(vc, d, r[u], j + 1) - ld2<CL>(vc, d, r[u], j - 1);
    }
#pragma unroll
    for (int u = 0; u < NU_P_SOURCE; ++u) {
        const float t2 = (sx[u].x * sx[u].x + sy[u].y * sy[u].y + (sy[u].x * sx[u].y)) / 8.0f;
        const float t3 = fdiv_z(dx * (sx[u].x + sy[u].y), 8.0f * dt);
        if (ok[u]) reinterpret_cast<float2 *>(src)[IX(d, r[u], j)] = make_float2(t2, t3);
    }
}
// 6 resident blocks (40 registers, 8 bytes spilled): 213 -> 196 us at 8192^2; 8 blocks (32 registers): 207 us
__global__ void __launch_bounds__(TX *TY, 6)
    k_p_source(float *__restrict__ src, const float *__restrict__ vc, fs2d_dom d, float dt, float dx) {
    if (block_interior(d, TY * NU_P_SOURCE, 1)) b_p_source<false>(src, vc, d, dt, dx);
    else b_p_source<true>(src, vc, d, dt, dx);
}

// ---------------------------------------------------------------------------------------------
// inline pressure BC: post-BC value of cell (r, j) recomputed from the pre-BC field and pcode
// (fs/boundary_condition.py:41-65); r, j already clamped
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool pc_is_wall(uint8_t c) { return c >= FS2D_PC_W_IM && c <= FS2D_PC_W_NONE; }
__device__ __forceinline__ bool pc_plain(uint8_t c) { return c == FS2D_PC_FLUID || c == FS2D_PC_W_NONE; }

__device__ __noinline__ float p_post(const float *pc, const uint8_t *pcode, const fs2d_dom &d, int r, int j) {
    const size_t idx = IX(d, r, j);
    const uint8_t c = __ldg(pcode + idx) & 15;  // low nibble = FS2D_PC_* code
    switch (c) {
        case FS2D_PC_FLUID:
        case FS2D_PC_W_NONE: return __ldg(pc + idx);
        case FS2D_PC_W_IM: return ld1(pc, d, r - 1, j);
        case FS2D_PC_W_IP: return ld1(pc, d, r + 1, j);
        case FS2D_PC_W_JM: return ld1(pc, d, r, j - 1);
        case FS2D_PC_W_JP: return ld1(pc, d, r, j + 1);
        case FS2D_PC_W_IM_JP: return (ld1(pc, d, r - 1, j) + ld1(pc, d, r, j + 1)) / 2.0f;
        case FS2D_PC_W_IP_JP: return (ld1(pc, d, r + 1, j) + ld1(pc, d, r, j + 1)) / 2.0f;
        case FS2D_PC_W_IM_JM: return (ld1(pc, d, r - 1, j) + ld1(pc, d, r, j - 1)) / 2.0f;
        case FS2D_PC_W_IP_JM: return (ld1(pc, d, r + 1, j) + ld1(pc, d, r, j - 1)) / 2.0f;
        case FS2D_PC_INFLOW: return ld1(pc, d, r + 1, j);
        default: return 0.0f;  // FS2D_PC_OUTFLOW
    }
}

// ---------------------------------------------------------------------------------------------
// one Jacobi sweep (fs/pressure_updater.py:62-66), any Y
// ---------------------------------------------------------------------------------------------
template <bool INLINE_BC>
__global__ void __launch_bounds__(TX *TY)
    k_jacobi_scalar(float *__restrict__ pn, const float *__restrict__ pc, const float *__restrict__ src,
                    const uint8_t *__restrict__ pcode, fs2d_dom d) {
    FS2D_CELL(d, r, j)
    const size_t idx = IX(d, r, j);
    if (pc_is_wall(__ldg(pcode + idx) & 15)) return;
    float pe, pw, pn_, ps;
    if (INLINE_BC) {
        pe = p_post(pc, pcode, d, CR(d, r + 1), j);
        pw = p_post(pc, pcode, d, CR(d, r - 1), j);
        pn_ = p_post(pc, pcode, d, r, CJ(d, j + 1));
        ps = p_post(pc, pcode, d, r, CJ(d, j - 1));
    } else {
        pe = ld1(pc, d, r + 1, j);
        pw = ld1(pc, d, r - 1, j);
        pn_ = ld1(pc, d, r, j + 1);
        ps = ld1(pc, d, r, j - 1);
    }
    const float2 s = __ldg(reinterpret_cast<const float2 *>(src) + idx);
    pn[idx] = 0.25f * (pe + pw + pn_ + ps) + s.x - s.y;
}

// vectorised marching sweep: each warp owns a 128-column x JM_ROWS-row chunk, 4 cells / lane along j
// (128-bit loads/stores), and walks down the rows keeping the (i-1, i, i+1) pressure rows in registers,
// so every p row is fetched once per chunk (+2 halo rows); j-neighbours come from warp shuffles.
// pcode byte = FS2D_PC_* code | (neighbour-is-a-BC-cell bits << 4), so no neighbour codes are loaded.
// Requires Y % 4 == 0.
constexpr int JM_WARPS = 8;   // warps per block (each on its own row chunk)
int g_jm_rows = 4;            // rows marched by one warp (tunable: fs2d_set_tuning(0, rows))
template <bool INLINE_BC, int JM_ROWS>
__global__ void __launch_bounds__(32 * JM_WARPS, 6)
    k_jacobi_march(float *__restrict__ pn, const float *__restrict__ pc, const float *__restrict__ src,
                   const uint8_t *__restrict__ pcode, fs2d_dom d) {
    const int lane = threadIdx.x;
    const int j0 = 4 * (FS2D_COLBLK * 32 + lane);
    const bool active = j0 < d.Y;
    const int jc = active ? j0 : 0;  // inactive lanes read column 0 (values unused) and never store
    const int r_begin = d.r0 + (FS2D_ROWBLK * JM_WARPS + threadIdx.y) * JM_ROWS;
    const int r_end = min(r_begin + JM_ROWS, d.r1);
    if (r_begin >= r_end) return;  // warp-uniform
    const int jl = CJ(d, jc - 1), jr = CJ(d, jc + 4);
    float4 up = __ldg(reinterpret_cast<const float4 *>(pc + IX(d, CR(d, r_begin - 1), jc)));
    float4 cen = __ldg(reinterpret_cast<const float4 *>(pc + IX(d, r_begin, jc)));
    for (int r = r_begin; r < r_end; ++r) {
        const size_t idx = IX(d, r, jc);
        const float4 dn = __ldg(reinterpret_cast<const float4 *>(pc + IX(d, CR(d, r + 1), jc)));
        const uchar4 cc = __ldg(reinterpret_cast<const uchar4 *>(pcode + idx));
        // j-neighbours: lane l-1's .w / lane l+1's .x; the warp's edge lanes fetch them from memory
        float pl = __shfl_up_sync(0xffffffffu, cen.w, 1), pr = __shfl_down_sync(0xffffffffu, cen.x, 1);
        if (lane == 0) pl = __ldg(pc + IX(d, r, jl));
        if (lane == 31 || jc + 4 >= d.Y) pr = __ldg(pc + IX(d, r, jr));
        const uint8_t k0 = cc.x, k1 = cc.y, k2 = cc.z, k3 = cc.w;
        const bool w0 = pc_is_wall(k0 & 15), w1 = pc_is_wall(k1 & 15), w2 = pc_is_wall(k2 & 15), w3 = pc_is_wall(k3 & 15);
        if (active && !(w0 && w1 && w2 && w3)) {
            const float4 s01 = __ldg(reinterpret_cast<const float4 *>(src + 2 * idx));
            const float4 s23 = __ldg(reinterpret_cast<const float4 *>(src + 2 * idx) + 1);
            float pw[4] = {up.x, up.y, up.z, up.w};        // (i-1, j)
            float pe[4] = {dn.x, dn.y, dn.z, dn.w};        // (i+1, j)
            float ps[4] = {pl, cen.x, cen.y, cen.z};       // (i, j-1)
            float pq[4] = {cen.y, cen.z, cen.w, pr};       // (i, j+1)
            if (INLINE_BC && ((k0 | k1 | k2 | k3) & 0xF0)) {  // rare: some neighbour is a BC cell
                const uint8_t kk[4] = {k0, k1, k2, k3};
                const int ru = CR(d, r - 1), rd = CR(d, r + 1);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (kk[k] & 0x10) pw[k] = p_post(pc, pcode, d, ru, jc + k);
                    if (kk[k] & 0x20) pe[k] = p_post(pc, pcode, d, rd, jc + k);
                    if (kk[k] & 0x40) ps[k] = p_post(pc, pcode, d, r, CJ(d, jc + k - 1));
                    if (kk[k] & 0x80) pq[k] = p_post(pc, pcode, d, r, CJ(d, jc + k + 1));
                }
            }
            const float t2[4] = {s01.x, s01.z, s23.x, s23.z}, t3[4] = {s01.y, s01.w, s23.y, s23.w};
            float out[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) out[k] = 0.25f * (pe[k] + pw[k] + pq[k] + ps[k]) + t2[k] - t3[k];
            if (!(w0 || w1 || w2 || w3)) {
                *reinterpret_cast<float4 *>(pn + idx) = make_float4(out[0], out[1], out[2], out[3]);
            } else {
                if (!w0) pn[idx] = out[0];
                if (!w1) pn[idx + 1] = out[1];
                if (!w2) pn[idx + 2] = out[2];
                if (!w3) pn[idx + 3] = out[3];
            }
        }
        up = cen;
        cen = dn;
    }
}

// fs/pressure_updater.py:98-114  one colour pass of red-black SOR (pc may alias pn)
__global__ void __launch_bounds__(TX *TY)
    k_rbsor_pass(float *pn, const float *pc, const float *__restrict__ src, const uint8_t *__restrict__ mask, fs2d_dom d,
                 float omega, float one_minus_omega, int parity) {
    FS2D_CELL(d, r, j)
    const size_t idx = IX(d, r, j);
    if (((d.gi0 + r + j) & 1) != parity || mask[idx] != 0) return;
    // plain loads: pc may alias pn (even pass) -- neighbours have the other colour, never written here
    const float pe = pc[IX(d, CR(d, r + 1), j)], pw = pc[IX(d, CR(d, r - 1), j)];
    const float pq = pc[IX(d, r, CJ(d, j + 1))], ps = pc[IX(d, r, CJ(d, j - 1))];
    const float2 s = __ldg(reinterpret_cast<const float2 *>(src) + idx);
    const float pred = 0.25f * (pe + pw + pq + ps) + s.x - s.y;
    pn[idx] = one_minus_omega * pc[idx] + omega * pred;
}

// fs/pressure_updater.py:86-96  BOTH colour passes of one red-black SOR iteration in one pass over HBM:
//     odd cells:   pn = (1 - w) pc + w predict(pc)          (_update_pressures_odd,  :98-102)
//     even cells:  pn = (1 - w) pn + w predict(pn)          (_update_pressures_even, :104-108, reads the odd cells just written)
// A warp owns 120 columns x RB_ROWS rows (lane l: columns J0 + 4l .. 4l+3; lanes 0 and 31 only supply j-neighbours) and
// marches down the rows: the "odd-updated" row O[q] (odd fluid cells relaxed from pc, every other cell = what pn holds) is
// formed in registers, and as soon as O[q-2], O[q-1], O[q] exist the even cells of row q-1 are relaxed from them and the
// row is stored -- only its fluid cells, so never-written cells keep their values.  O of the rows just
// outside the chunk is recomputed by both neighbours (it depends on pc and on never-written cells of pn only), so chunks
// are independent.  21 B/cell (pc 4, pn 4 + 4, source 8, mask 1) instead of two passes over all five arrays; same
// expression and order per cell: bit-identical to the two k_rbsor_pass launches.  Requires Y % 4 == 0, pn != pc.
constexpr int RB_ROWS = 32, RB_WARPS = 8, RB_COLS = 120;
// 3 resident blocks (80 registers instead of 94, 20 bytes spilled): 164 -> 154 us per iteration at 8192 x 4096; 4 blocks
// (64 registers, 120 bytes spilled): 188 us; 5 blocks: 304 us.  The kernel stays bound by its registers (three pc rows, three
// O rows and the source of the previous row live across the march): occupancy 37 %, 0.64 of the HBM peak on its 21 B/cell.
__global__ void __launch_bounds__(32 * RB_WARPS, 3)
    k_rbsor_fused(float *__restrict__ pn, const float *__restrict__ pc, const float *__restrict__ src,
                  const uint8_t *__restrict__ mask, fs2d_dom d, float omega, float one_minus_omega) {
    constexpr uint32_t FULL = 0xffffffffu;
    const int lane = threadIdx.x;
    const int jc0 = FS2D_COLBLK * RB_COLS - 4 + 4 * lane;
    const bool in_grid = jc0 >= 0 && jc0 < d.Y;
    const int jc = in_grid ? jc0 : (jc0 < 0 ? 0 : d.Y - 4);   // lanes outside the grid read a valid address, their values are unused
    const bool owner = in_grid && lane >= 1 && lane <= 30;
    const bool left_edge = jc0 == 0, right_edge = jc0 + 4 >= d.Y;
    const int rb = d.r0 + (FS2D_ROWBLK * RB_WARPS + threadIdx.y) * RB_ROWS;
    const int re = min(rb + RB_ROWS, d.r1);
    if (rb >= re) return;   // warp-uniform
    auto row4 = [&](const float *f, int r) { return __ldg(reinterpret_cast<const float4 *>(f + IX(d, CR(d, r), jc))); };
    float4 pc_m = row4(pc, rb - 2), pc_0 = row4(pc, rb - 1), pc_p;
    float Om2[4] = {0, 0, 0, 0}, Om1[4] = {0, 0, 0, 0}, Oq[4];
    float t2m[4] = {0, 0, 0, 0}, t3m[4] = {0, 0, 0, 0};
    uint32_t fluid_m = 0;   // bit h: cell h of row q-1 is a fluid cell
    for (int q = rb - 1; q <= re; ++q) {
        pc_p = row4(pc, q + 1);
        const int qc = CR(d, q);   // rows outside the clamp window are formed from a valid row; they are never used (see up / down below)
        const size_t idx = IX(d, qc, jc);
        const float4 old = *reinterpret_cast<const float4 *>(pn + idx);   // plain load: pn is written by this kernel
        const float4 s01 = __ldg(reinterpret_cast<const float4 *>(src + 2 * idx)), s23 = __ldg(reinterpret_cast<const float4 *>(src + 2 * idx) + 1);
        const uchar4 mk = __ldg(reinterpret_cast<const uchar4 *>(mask + idx));
        const float t2[4] = {s01.x, s01.z, s23.x, s23.z}, t3[4] = {s01.y, s01.w, s23.y, s23.w};
        const uint32_t fluid = (uint32_t)(mk.x == 0) | ((uint32_t)(mk.y == 0) << 1) | ((uint32_t)(mk.z == 0) << 2) | ((uint32_t)(mk.w == 0) << 3);
        // ---- odd cells of row q from pc ----
        float pl = __shfl_up_sync(FULL, pc_0.w, 1), pr = __shfl_down_sync(FULL, pc_0.x, 1);
        if (left_edge) pl = pc_0.x;      // sample() clamps to the cell itself
        if (right_edge) pr = pc_0.w;
        const float c0[4] = {pc_0.x, pc_0.y, pc_0.z, pc_0.w};
        const float e0[4] = {pc_p.x, pc_p.y, pc_p.z, pc_p.w}, w0[4] = {pc_m.x, pc_m.y, pc_m.z, pc_m.w};
        const float q0[4] = {pc_0.y, pc_0.z, pc_0.w, pr}, z0[4] = {pl, pc_0.x, pc_0.y, pc_0.z};
        const float o4[4] = {old.x, old.y, old.z, old.w};
        const int par0 = (d.gi0 + qc + jc) & 1;   // parity of the lane's first cell in row q
        const bool in_window = q >= d.r0 && q < d.r1;   // the odd pass only touches the rows [r0, r1): a row outside them is what pn holds
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const float pred = 0.25f * (e0[h] + w0[h] + q0[h] + z0[h]) + t2[h] - t3[h];
            const bool odd = ((par0 + h) & 1) == 1;
            Oq[h] = (in_window && odd && ((fluid >> h) & 1u)) ? one_minus_omega * c0[h] + omega * pred : o4[h];
        }
        // ---- even cells of row q-1 from O[q-2], O[q-1], O[q]; store the row ----
        if (q - 1 >= rb) {
            const int r = q - 1;
            float ol = __shfl_up_sync(FULL, Om1[3], 1), orr = __shfl_down_sync(FULL, Om1[0], 1);
            if (left_edge) ol = Om1[0];
            if (right_edge) orr = Om1[3];
            const bool has_up = r - 1 >= d.clo, has_dn = r + 1 <= d.chi;
            const float qn[4] = {Om1[1], Om1[2], Om1[3], orr}, zn[4] = {ol, Om1[0], Om1[1], Om1[2]};
            const int par1 = (d.gi0 + r + jc) & 1;
            float out[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const float pe = has_dn ? Oq[h] : Om1[h], pw = has_up ? Om2[h] : Om1[h];
                const float pred = 0.25f * (pe + pw + qn[h] + zn[h]) + t2m[h] - t3m[h];
                const bool even = ((par1 + h) & 1) == 0;
                out[h] = (even && ((fluid_m >> h) & 1u)) ? one_minus_omega * Om1[h] + omega * pred : Om1[h];
            }
            if (owner && fluid_m) {
                float *dst = pn + IX(d, r, jc);
                if (fluid_m == 15u) {
                    *reinterpret_cast<float4 *>(dst) = make_float4(out[0], out[1], out[2], out[3]);
                } else {
#pragma unroll
                    for (int h = 0; h < 4; ++h)
                        if ((fluid_m >> h) & 1u) dst[h] = out[h];
                }
            }
        }
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            Om2[h] = Om1[h];
            Om1[h] = Oq[h];
            t2m[h] = t2[h];
            t3m[h] = t3[h];
        }
        fluid_m = fluid;
        pc_m = pc_0;
        pc_0 = pc_p;
    }
}

static void launch_jacobi(float *pn, const float *pc, const float *src, const uint8_t *pcode, const fs2d_dom &d,
                          int inline_bc, cudaStream_t s) {
    const bool vec = (d.Y % 4 == 0) && ((uintptr_t)pn % 16 == 0) && ((uintptr_t)pc % 16 == 0) &&
                     ((uintptr_t)src % 16 == 0) && ((uintptr_t)pcode % 4 == 0);
    ++g_launches;
    if (vec) {
#define JM_LAUNCH(R)                                                                          \
    do {                                                                                      \
        dim3 blk(32, JM_WARPS, 1), grd(nblk(d.Y, 128), nblk(d.r1 - d.r0, R * JM_WARPS), 1);    \
        if (inline_bc) k_jacobi_march<true, R><<<grd, blk, 0, s>>>(pn, pc, src, pcode, d);    \
        else k_jacobi_march<false, R><<<grd, blk, 0, s>>>(pn, pc, src, pcode, d);             \
    } while (0)
        switch (g_jm_rows) {
            case 1: JM_LAUNCH(1); break;
            case 2: JM_LAUNCH(2); break;
            case 8: JM_LAUNCH(8); break;
            case 16: JM_LAUNCH(16); break;
            default: JM_LAUNCH(4); break;
        }
#undef JM_LAUNCH
    } else {
        if (inline_bc) k_jacobi_scalar<true><<<dense_grid(d), dense_block(), 0, s>>>(pn, pc, src, pcode, d);
        else k_jacobi_scalar<false><<<dense_grid(d), dense_block(), 0, s>>>(pn, pc, src, pcode, d);
    }
}

}  // namespace fs2d

using namespace fs2d;
#define STREAM ((cudaStream_t)stream)

namespace fs2d { extern int g_dye_vec, g_nonadv_vec; }   // fs2d_dye.cu, fs2d_kernels.cu
extern "C" {

int fs2d_set_tuning(int key, int value) {
    if (key == 0 && (value == 1 || value == 2 || value == 4 || value == 8 || value == 16)) { g_jm_rows = value; return FS2D_OK; }
    if (key == 2 && value >= 0 && value <= 2) { fs2d::g_stream = value; return FS2D_OK; }
    if (key == 3 && value >= 0 && value <= 3) { fs2d::g_stream_cfg = value; return FS2D_OK; }
    if (key == 4 && (value == 0 || value == 1)) { fs2d::g_tail_emit = value; return FS2D_OK; }
    if (key == 5 && (value == 0 || value == 1)) { fs2d::g_dye_vec = value; return FS2D_OK; }
    if (key == 6 && (value == 0 || value == 1)) { fs2d::g_nonadv_vec = value; return FS2D_OK; }
    set_error("unknown tuning key %d / value %d", key, value);
    return FS2D_E_BADARG;
}

int fs2d_pressure_source(float *src, const float *vc, fs2d_dom d, float dt, float dx, void *stream) {
    FS2D_REQUIRE(src && vc, "null field pointer");
    if (int e = check_dom(d)) return e;
    if (d.r1 == d.r0) return FS2D_OK;
    ++g_launches;
    k_p_source<<<dense_grid_nu(d, NU_P_SOURCE), dense_block(), 0, STREAM>>>(src, vc, d, dt, dx);
    FS2D_LAUNCH_CHECK();
    return FS2D_OK;
}

int fs2d_jacobi_sweep(float *pn, const float *pc, const float *src, const uint8_t *pcode, fs2d_dom d, int inline_bc,
                      void *stream) {
    FS2D_REQUIRE(pn && pc && src && pcode, "null field pointer");
    FS2D_REQUIRE(pn != pc, "Jacobi sweep cannot run in place");
    if (int e = check_dom(d)) return e;
    if (d.r1 == d.r0) return FS2D_OK;
    launch_jacobi(pn, pc, src, pcode, d, inline_bc, STREAM);
    FS2D_LAUNCH_CHECK();
    return FS2D_OK;
}

// Decompose n_sweeps reference iterations into fused passes (size t >= 1, only sizes whose bit is set in
// fuse_mask) and literal single iterations (size 0), cheapest first by a cost table measured on B200 at
// 8192^2 cells (scripts/sweep_bench.py, us): the last `tail_literal` iterations are literal and the
// number of buffer flips (= number of entries) must have the parity of n_sweeps so that the two PHYSICAL
// buffers end up exactly as in the reference.
static int plan_jacobi(int n_sweeps, int fuse_mask, int *out, int cap, int tail_literal = 2) {
    static const float pass_cost[13] = {0, 166, 178, 184, 184, 223, 248, 283, 323, 379, 411, 463, 496};   // r02 (profiles/r02_fused_sweep_bench_v3.txt)
    const float lit_cost = 190.0f;
    const int n_lit = n_sweeps < tail_literal ? n_sweeps : tail_literal, n_f = n_sweeps - n_lit;
    int n = 0;
    if (n_f > 0 && (fuse_mask & 0x1FFE) && n_f < 4096) {
        // dp[i][par]: cheapest way to do i iterations with an entry count of parity par
        static thread_local float dp[4097][2];
        static thread_local short choice[4097][2];
        const float INF = 1e30f;
        dp[0][0] = 0; dp[0][1] = INF;
        for (int i = 1; i <= n_f; ++i)
            for (int par = 0; par < 2; ++par) {
                float best = dp[i - 1][par ^ 1] + lit_cost;   // one literal iteration
                short ch = 0;
                for (int t = 1; t <= 12 && t <= i; ++t)
                    if ((fuse_mask >> t) & 1) {
                        const float cst = dp[i - t][par ^ 1] + pass_cost[t];
                        if (cst < best) { best = cst; ch = (short)t; }
                    }
                dp[i][par] = best;
                choice[i][par] = ch;
            }
        int par = n_f & 1, i = n_f;
        if (dp[i][par] >= INF) return -1;
        while (i > 0) {
            const int t = choice[i][par];
            if (n >= cap) return -1;
            out[n++] = t;
            i -= t ? t : 1;
            par ^= 1;
        }
    } else {
        for (int i = 0; i < n_f; ++i) { if (n >= cap) return -1; out[n++] = 0; }
    }
    for (int i = 0; i < n_lit; ++i) { if (n >= cap) return -1; out[n++] = 0; }
    return n;
}

// The schedule of one update: it ends {..., fused pass that also emits the BC values of its penultimate state, ONE literal
// iteration} (see jacobi_fused_body<EMIT>; measured 56 us faster per 80-iteration update at 8192^2) -- or, with
// fs2d_set_tuning(4, 0), {..., two literal iterations}.  *tail = 1 if entry n - 2 is that emitting pass (if the entry
// before the last one is itself a literal iteration, the usual reasoning applies unchanged).
static int plan_with_tail(int n_sweeps, int fuse_mask, int *out, int cap, bool *tail) {
    int n = -1;
    if (fs2d::g_tail_emit && fuse_mask != 0 && n_sweeps >= 3) n = plan_jacobi(n_sweeps, fuse_mask, out, cap, 1);
    *tail = n >= 2 && out[n - 2] > 0;
    if (n < 0) n = plan_jacobi(n_sweeps, fuse_mask, out, cap);
    return n;
}

int fs2d_jacobi_plan(int n_sweeps, int fuse_mask, int *sizes, int cap, int *n_entries) {
    FS2D_REQUIRE(n_sweeps >= 0 && fuse_mask >= 0 && sizes && n_entries && cap > 0, "bad plan arguments");
    bool tail = false;
    const int n = plan_with_tail(n_sweeps, fuse_mask, sizes, cap, &tail);
    FS2D_REQUIRE(n >= 0, "plan does not fit the output array");
    *n_entries = n;
    return FS2D_OK;
}

int fs2d_jacobi_update(float *pa, float *pb, const float *src, const uint8_t *pcode, fs2d_dom d, int n_sweeps,
                       const int32_t *tgt, const int32_t *src0, const int32_t *src1, const uint8_t *kind, float *scratch,
                       int n_bc, int fuse_mask, const int32_t *const *orders, const int *n_orders, int *final_in_b, void *stream) {
    FS2D_REQUIRE(pa && pb && src && pcode && pa != pb, "null/aliased field pointer");
    FS2D_REQUIRE(n_sweeps >= 0 && fuse_mask >= 0, "negative sweep count");
    FS2D_REQUIRE(n_bc == 0 || (tgt && src0 && src1 && kind && scratch), "null BC table");
    FS2D_REQUIRE(!orders || n_orders, "tile lists without their lengths");
    if (int e = check_dom(d)) return e;
    float *cur = pa, *nxt = pb;
    if (d.r1 == d.r0 || !fused_supported(pa, pb, src, pcode, d)) fuse_mask = 0;
    static thread_local int plan[4200];
    bool tail = false;
    const int n = plan_with_tail(n_sweeps, fuse_mask, plan, 4200, &tail);
    FS2D_REQUIRE(n >= 0, "iteration count too large");
    for (int k = 0; k < n; ++k) {
        if (plan[k] > 0) {
            // fused pass: plan[k] iterations in shared memory (fs2d_fused.cu)
            const int t = plan[k];
            const int32_t *order = orders ? orders[t] : nullptr;
            if (int e = fused_pass(cur, nxt, src, pcode, d, t, STREAM, 0, 0, tail && k == n - 2, order, order ? n_orders[t] : 0)) return e;
        } else {
            // literal reference iteration (fs/pressure_updater.py:57-60): in-place sparse BC, then the plain sweep
            launch_p_bc(cur, tgt, src0, src1, kind, scratch, n_bc, STREAM);
            if (d.r1 > d.r0) launch_jacobi(nxt, cur, src, pcode, d, 0, STREAM);
        }
        float *x = cur; cur = nxt; nxt = x;
    }
    FS2D_LAUNCH_CHECK();
    if (final_in_b) *final_in_b = (cur == pb);
    return FS2D_OK;
}

int fs2d_rbsor_iteration(float *pn, const float *pc, const float *src, const uint8_t *mask, fs2d_dom d, float omega,
                         float one_minus_omega, void *stream) {
    FS2D_REQUIRE(pn && pc && src && mask && pn != pc, "null/aliased field pointer");
    if (int e = check_dom(d)) return e;
    if (d.r1 == d.r0) return FS2D_OK;
    const bool vec = (d.Y % 4 == 0) && ((uintptr_t)pn % 16 == 0) && ((uintptr_t)pc % 16 == 0) && ((uintptr_t)src % 16 == 0) &&
                     ((uintptr_t)mask % 4 == 0);
    if (!vec) {   // the two colour passes of the reference, one kernel each
        if (int e = fs2d_rbsor_pass(pn, pc, src, mask, d, omega, one_minus_omega, 1, stream)) return e;
        return fs2d_rbsor_pass(pn, pn, src, mask, d, omega, one_minus_omega, 0, stream);
    }
    ++g_launches;
    const dim3 blk(32, RB_WARPS, 1), grd(nblk(d.Y, RB_COLS), nblk(d.r1 - d.r0, RB_ROWS * RB_WARPS), 1);
    k_rbsor_fused<<<grd, blk, 0, STREAM>>>(pn, pc, src, mask, d, omega, one_minus_omega);
    FS2D_LAUNCH_CHECK();
    return FS2D_OK;
}

int fs2d_rbsor_pass(float *pn, const float *pc, const float *src, const uint8_t *mask, fs2d_dom d, float omega,
                    float one_minus_omega, int parity, void *stream) {
    FS2D_REQUIRE(pn && pc && src && mask, "null field pointer");
    FS2D_REQUIRE(parity == 0 || parity == 1, "parity must be 0 or 1");
    if (int e = check_dom(d)) return e;
    if (d.r1 == d.r0) return FS2D_OK;
    ++g_launches;
    k_rbsor_pass<<<dense_grid(d), dense_block(), 0, STREAM>>>(pn, pc, src, mask, d, omega, one_minus_omega, parity);
    FS2D_LAUNCH_CHECK();
    return FS2D_OK;
}

}  // extern "C"

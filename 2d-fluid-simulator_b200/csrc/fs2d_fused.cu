// fs2d_fused.cu -- T Jacobi iterations (pressure BC + sweep, fs/pressure_updater.py:56-66) per pass
// over HBM, in shared memory (temporal blocking), for sm_100a.
//
// Why it is exact: one reference iteration is local -- a relaxed cell reads the post-BC values of its 4
// neighbours, and a BC cell's value is a function of cells at most one step further
// (fs/boundary_condition.py:41-65).  A CTA therefore loads a (SI x SJ) tile of p, the velocity source
// terms (t2, t3) and pcode, and runs T iterations on it while the region in which values are still
// correct shrinks by one cell per iteration from every side that is not a global edge (the host verifies
// this bound for the actual mask, fs/_bc_tables.py:fused_reach_ok); it then stores the inner
// (SI-2T) x (SJ-2T) cells.  Every cell-iteration evaluates the literal expression
// 0.25*(p(i+1,j)+p(i-1,j)+p(i,j+1)+p(i,j-1)) + t2 - t3, so results are bit-identical to T separate sweeps.
//
// Structure (B200): persistent CTAs (one per SM) loop over tiles; one elected thread feeds a staging buffer
// with three TMA 2-D box loads (cp.async.bulk.tensor -> UTMALDG; out-of-grid parts are zero-filled by the
// TMA unit) signalled through an mbarrier; the next tile's loads are issued as soon as iteration 0 has
// consumed the staging buffer, so HBM traffic overlaps iterations 1..T-1 and the store phase.  Thread
// (tx, ty) owns column tx and K consecutive rows of the tile: its K pressures, K (t2, t3) pairs and its
// update/slow bit masks live in registers for all T iterations; per iteration it writes its K values to a
// ping-pong smem plane and reads only the j-neighbours (and the two i-neighbours outside its own rows) back.
//
// HBM bytes per cell per iteration: (4 + 8 + 1.1) * (SI*SJ)/((SI-2T)(SJ-2HJ)) / T + 4/T  (T=4: 5.0 B vs 17 B).
#include <cuda.h>

#include "fs2d_common.cuh"

namespace fs2d {

constexpr int FK = 8;            // rows per thread
constexpr int FNTY = 8;          // row blocks per tile
constexpr int FSI = FK * FNTY;   // 64 tile rows
constexpr int FSJ = 128;         // tile columns = threads per row block
constexpr int F_THREADS = FSJ * FNTY;
constexpr int F_TMAX = 12;
constexpr int FCW = FSJ + 16;   // columns of the staged pcode box (its start is rounded down to 16 bytes)
// TMA (measured on B200, scripts/probes/tma_probe.cu): the box start must be 16-byte aligned in the innermost
// dimension (an unaligned column coordinate raises "illegal instruction"), rows are free.  So the column halo
// HJ is T rounded up to a multiple of 4 floats and the 1-byte pcode box starts at the previous multiple of 16.

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, P1;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

struct FusedGeom {
    int T;        // iterations in this pass == row halo
    int HJ;       // column halo: T rounded up to a multiple of 4 (TMA alignment)
    int TI, TJ;   // output tile rows / cols = FSI - 2T, FSJ - 2HJ
    int tiles_i, tiles_j;
    int skip_from, skip_n;   // tile rows [skip_from, skip_from + skip_n) of the tiling are left to another launch
};
// tile row of the tiling for the launch's `row`-th tile row
__device__ __forceinline__ int f_tile_row(const FusedGeom &g, int row) { return row < g.skip_from ? row : row + g.skip_n; }

// post-BC pressure of tile cell (r, c) read from plane `pl` (same rule as p_post in fs2d_pressure.cu);
// rlo..rhi / clo..chi: tile coordinates of the clamp bounds of sample()
__device__ __forceinline__ float f_post(const float *pl, const uint8_t *code, int r, int c, int rlo, int rhi, int clo,
                                        int chi) {
    const int rm = max(r - 1, rlo), rp = min(r + 1, rhi), cm = max(c - 1, clo), cp = min(c + 1, chi);
    int a = r * FSJ + c, b = a, mode = 0;  // mode 0: value of cell a; 1: (a + b) / 2; 2: zero
    switch (code[r * FSJ + c] & 15) {
        case FS2D_PC_W_IM: a = rm * FSJ + c; break;
        case FS2D_PC_W_IP: a = rp * FSJ + c; break;
        case FS2D_PC_W_JM: a = r * FSJ + cm; break;
        case FS2D_PC_W_JP: a = r * FSJ + cp; break;
        case FS2D_PC_W_IM_JP: a = rm * FSJ + c; b = r * FSJ + cp; mode = 1; break;
        case FS2D_PC_W_IP_JP: a = rp * FSJ + c; b = r * FSJ + cp; mode = 1; break;
        case FS2D_PC_W_IM_JM: a = rm * FSJ + c; b = r * FSJ + cm; mode = 1; break;
        case FS2D_PC_W_IP_JM: a = rp * FSJ + c; b = r * FSJ + cm; mode = 1; break;
        case FS2D_PC_INFLOW: a = rp * FSJ + c; break;
        case FS2D_PC_OUTFLOW: mode = 2; break;
        default: break;  // FLUID / W_NONE: the stored value
    }
    const float va = pl[a], vb = pl[b];
    return mode == 0 ? va : (mode == 1 ? (va + vb) / 2.0f : 0.0f);
}

// Shared-memory layout (float index unless noted).  All planes are addressed as offsets from ONE base pointer
// so that the compiler keeps them in the shared address space (LDS/STS, not generic LD/ST).
constexpr int FN = FSI * FSJ;          // cells per plane
constexpr int OFF_P0 = 0;              // TMA destination of p, plane of iteration 0
constexpr int OFF_SRC = FN;            // TMA destination of (t2, t3), 2*FN floats
constexpr int OFF_W0 = 3 * FN;         // working plane
constexpr int OFF_W1 = 4 * FN;         // working plane
constexpr int OFF_BYTES = 5 * FN;      // byte area: staged pcode (FSI*FCW), tile pcode (FN), slow list (2*FN)
constexpr size_t F_SMEM = (size_t)OFF_BYTES * 4 + (size_t)FSI * FCW + FN + 2 * FN;

__global__ void __launch_bounds__(F_THREADS, 1)
    k_jacobi_fused(const __grid_constant__ CUtensorMap map_p, const __grid_constant__ CUtensorMap map_src,
                   const __grid_constant__ CUtensorMap map_code, float *__restrict__ p_out, unsigned int *tile_ctr,
                   fs2d_dom d, FusedGeom g) {
    extern __shared__ __align__(1024) float sm[];
    uint8_t *stg_code = reinterpret_cast<uint8_t *>(sm + OFF_BYTES);
    uint8_t *wcode = stg_code + FSI * FCW;
    uint16_t *slow_list = reinterpret_cast<uint16_t *>(wcode + FN);
    __shared__ __align__(8) uint64_t bar;
    __shared__ int n_slow[2];   // slow cells of the current / next tile (ping-pong)
    __shared__ int s_next;      // tile index fetched by the leader for the next round

    const int tid = threadIdx.y * FSJ + threadIdx.x;
    const int c = threadIdx.x;             // tile column
    const int lr0 = threadIdx.y * FK;      // first tile row of this thread
    const int o0 = lr0 * FSJ + c;          // plane offset of this thread's first cell
    const bool leader = tid == 0;
    const int n_tiles = g.tiles_i * g.tiles_j;
    constexpr uint32_t TX_BYTES = FN * (4 + 8) + FSI * FCW;

    // NOTE: the descriptors must be addressed in the kernel-parameter space (the TMA unit cannot read a copy
    // that the compiler spilled to local memory), so take their addresses here, not through a lambda capture.
    const CUtensorMap *mp = &map_p, *ms = &map_src, *mc = &map_code;
#define FS2D_ISSUE(tile)                                                            \
    do {                                                                            \
        const int R0_ = d.r0 + f_tile_row(g, (tile) / g.tiles_j) * g.TI - g.T;                   \
        const int C0_ = ((tile) % g.tiles_j) * g.TJ - g.HJ;                         \
        mbar_expect_tx(&bar, TX_BYTES);                                             \
        tma_load_2d(sm + OFF_P0, mp, C0_, R0_, &bar);                               \
        tma_load_2d(sm + OFF_SRC, ms, 2 * C0_, R0_, &bar);                          \
        tma_load_2d(stg_code, mc, C0_ & ~15, R0_, &bar);                            \
    } while (0)

    if (leader) {
        mbar_init(&bar, 1);
        n_slow[0] = n_slow[1] = 0;
    }
    __syncthreads();
    int t = blockIdx.x;
    if (leader && t < n_tiles) FS2D_ISSUE(t);
    uint32_t parity = 0;
    const int dl = max(c - 1, 0) - c, dr = min(c + 1, FSJ - 1) - c;   // j-neighbour offsets, clamped inside the tile
    const int o_up = max(lr0 - 1, 0) * FSJ + c, o_dn = min(lr0 + FK, FSI - 1) * FSJ + c;

    while (t < n_tiles) {
        const int R0 = d.r0 + f_tile_row(g, t / g.tiles_j) * g.TI - g.T;   // local-array row of tile row 0
        const int C0 = (t % g.tiles_j) * g.TJ - g.HJ;         // column of tile column 0
        const int coff = C0 - (C0 & ~15);                     // where tile column 0 sits inside the staged pcode box
        // clamp bounds of sample() in tile coordinates (global edges only)
        const int rlo = max(0, d.clo - R0), rhi = min(FSI - 1, d.chi - R0);
        const int clo = max(0, -C0), chi = min(FSJ - 1, d.Y - 1 - C0);
        const int par = parity;

        mbar_wait(&bar, parity);
        parity ^= 1;

        // ---- per-thread state from the staging buffer -------------------------------------------
        float p[FK], t2[FK], t3[FK];
        uint32_t upd = 0, slow = 0;
#pragma unroll
        for (int k = 0; k < FK; ++k) {
            const int lr = lr0 + k, o = o0 + k * FSJ;
            p[k] = sm[OFF_P0 + o];
            const float2 s2 = reinterpret_cast<const float2 *>(sm + OFF_SRC)[o];
            t2[k] = s2.x;
            t3[k] = s2.y;
            const uint8_t pc = stg_code[lr * FCW + coff + c];
            wcode[o] = pc;
            const int code = pc & 15;
            const bool inside = lr >= rlo && lr <= rhi && c >= clo && c <= chi;
            const bool relaxed = code == FS2D_PC_FLUID || code == FS2D_PC_INFLOW || code == FS2D_PC_OUTFLOW;
            const bool e_rl = R0 + lr == d.clo, e_rh = R0 + lr == d.chi;
            const bool e_cl = C0 + c == 0, e_ch = C0 + c == d.Y - 1;
            // Cells on the tile rim lack a neighbour, so what they compute is garbage -- harmlessly: a rim value is
            // consumed by its inner neighbour only in the iteration in which it still holds the loaded state, and
            // that neighbour is outside the valid region from then on anyway.  Not freezing them lets tiles
            // without walls skip the per-cell update predicate altogether (all_upd below).
            const bool u = inside && relaxed;
            const bool sl = u && ((pc >> 4) != 0 || e_rl || e_rh || e_cl || e_ch);
            upd |= (uint32_t)u << k;
            slow |= (uint32_t)sl << k;
            if (sl) slow_list[atomicAdd(&n_slow[par], 1)] = (uint16_t)o;   // cells that need post-BC neighbour values
        }

        const bool all_upd = __syncthreads_and(upd == (1u << FK) - 1u) != 0;   // block-uniform: open-fluid tile

        // ---- T iterations --------------------------------------------------------------------------
        int cur = OFF_P0, nxt = OFF_W0;
        for (int s = 0; s < g.T; ++s) {
            __syncthreads();  // (A) plane `cur`, wcode and the slow list are complete; readers of `nxt` are done
            if (leader && s == (g.T > 1 ? 1 : 0)) {
                n_slow[par ^ 1] = 0;
                if (g.T > 1) {  // staging is free: fetch the next tile index and start its loads
                    const int tn = (int)atomicAdd(tile_ctr, 1u) + (int)gridDim.x;
                    s_next = tn;
                    if (tn < n_tiles) FS2D_ISSUE(tn);
                }
            }
            const int ns = n_slow[par];
            if (ns > 0) {  // block-uniform: the tile has cells next to BC cells / global edges
                // Balanced fix-up: all threads share the slow cells and leave, in plane `nxt`, the SUM of the four
                // post-BC neighbour values (the reference's order) for the owning thread to pick up.
                for (int e = tid; e < ns; e += F_THREADS) {
                    const int o = slow_list[e], r = o / FSJ, cc = o % FSJ;
                    float sum = f_post(sm + cur, wcode, min(r + 1, rhi), cc, rlo, rhi, clo, chi);
                    sum = sum + f_post(sm + cur, wcode, max(r - 1, rlo), cc, rlo, rhi, clo, chi);
                    sum = sum + f_post(sm + cur, wcode, r, min(cc + 1, chi), rlo, rhi, clo, chi);
                    sum = sum + f_post(sm + cur, wcode, r, max(cc - 1, clo), rlo, rhi, clo, chi);
                    sm[nxt + o] = sum;
                }
                __syncthreads();  // (B)
            }
            const float upx = sm[cur + o_up], dnx = sm[cur + o_dn];
            float prev_old = upx;
            if (all_upd) {
#pragma unroll
                for (int k = 0; k < FK; ++k) {
                    const int o = o0 + k * FSJ;
                    const float lf = sm[cur + o + dl], rt = sm[cur + o + dr];
                    const float dnv = k < FK - 1 ? p[k + 1] : dnx;
                    const float sum = dnv + prev_old + rt + lf;  // (i+1) + (i-1) + (j+1) + (j-1), the reference's order
                    prev_old = p[k];
                    p[k] = 0.25f * sum + t2[k] - t3[k];
                }
            } else {
#pragma unroll
                for (int k = 0; k < FK; ++k) {
                    const int o = o0 + k * FSJ;
                    const float lf = sm[cur + o + dl], rt = sm[cur + o + dr];
                    const float dnv = k < FK - 1 ? p[k + 1] : dnx;
                    const float sum = dnv + prev_old + rt + lf;
                    const float v = 0.25f * sum + t2[k] - t3[k];
                    prev_old = p[k];
                    p[k] = ((upd >> k) & 1u) ? v : p[k];
                }
            }
            if (slow) {
#pragma unroll
                for (int k = 0; k < FK; ++k)
                    if ((slow >> k) & 1u) p[k] = 0.25f * sm[nxt + o0 + k * FSJ] + t2[k] - t3[k];
            }
#pragma unroll
            for (int k = 0; k < FK; ++k) sm[nxt + o0 + k * FSJ] = p[k];
            cur = nxt;
            nxt = (nxt == OFF_W0) ? OFF_W1 : OFF_W0;
        }

        // ---- store the inner (TI x TJ) cells that were updated and belong to rows [r0, r1) ----------
        if (c >= g.HJ && c < g.HJ + g.TJ && C0 + c < d.Y) {
#pragma unroll
            for (int k = 0; k < FK; ++k) {
                const int lr = lr0 + k, gr = R0 + lr;
                if (lr >= g.T && lr < g.T + g.TI && gr < d.r1 && ((upd >> k) & 1u)) p_out[(size_t)gr * d.Y + (C0 + c)] = p[k];
            }
        }
        __syncthreads();  // all reads of the working planes are done before the next tile overwrites them
        if (g.T == 1) {   // the staging plane doubled as the only `cur` plane: release it only now
            if (leader) {
                const int tn = (int)atomicAdd(tile_ctr, 1u) + (int)gridDim.x;
                s_next = tn;
                if (tn < n_tiles) FS2D_ISSUE(tn);
            }
            __syncthreads();
        }
        t = s_next;
    }
#undef FS2D_ISSUE
}

// ---------------------------------------------------------------------------------------------
// Variant 3: register tile + warp shuffles.  A warp owns HK consecutive tile rows and ALL 128 tile columns;
// lane l owns columns 4l..4l+3 of those rows, i.e. a thread keeps a 8 x 4 block of p, t2 and t3 in registers
// for the whole pass.  Per iteration the j-neighbours outside the thread's block come from the adjacent
// lanes by two shuffles per row (a warp spans the tile width, so no other warp is involved) and the
// i-neighbours outside the block are the adjacent warps' edge rows, exchanged through the ping-pong planes
// with one LDS.128 + one STS.128 per edge row.  Open-fluid tiles therefore cost ~6 FP32 + 0.7 other
// instructions per cell-iteration (variant 1: 13) and touch shared memory only for the two edge rows of each
// warp; tiles with cells next to BC cells / global edges additionally mirror all rows into the plane every
// iteration so that the cooperative slow-cell fix-up of variant 1 (f_post) works unchanged.  Same tile,
// TMA staging, tile scheduler, validity rules and arithmetic (literal order) as variant 1.
// ---------------------------------------------------------------------------------------------
static_assert(FSJ == 128, "a warp (32 lanes x 4 columns) must span the tile width");

__device__ __forceinline__ float4 lds4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void sts4(float *p, float a, float b, float c, float d) {
    *reinterpret_cast<float4 *>(p) = make_float4(a, b, c, d);
}

// One Jacobi iteration on a thread's HK x 4 register block, in place.  up/dn: old values of the rows above / below
// the block.  The rows are software-pipelined so that row k-1 is overwritten only after row k has consumed its old
// values: no second register copy of the block and no end-of-iteration moves.  ALL: every cell is relaxed (no
// per-cell select).  Every cell evaluates 0.25*((p(i+1,j)+p(i-1,j))+p(i,j+1))+p(i,j-1)) + t2 - t3, the reference's order.
template <int HK, bool ALL>
__device__ __forceinline__ void jacobi_rows(float (&p)[HK][4], const float (&t2)[HK][4], const float (&t3)[HK][4],
                                            uint32_t upd, const float4 up, const float4 dn) {
    constexpr uint32_t FULL = 0xffffffffu;
    float lf[HK], rt[HK];
#pragma unroll
    for (int k = 0; k < HK; ++k) {   // all shuffles first: their latency is paid once per iteration
        lf[k] = __shfl_up_sync(FULL, p[k][3], 1);
        rt[k] = __shfl_down_sync(FULL, p[k][0], 1);
    }
    float S[4], A[4];
    S[0] = p[1][0] + up.x + p[0][1] + lf[0];
    S[1] = p[1][1] + up.y + p[0][2] + p[0][0];
    S[2] = p[1][2] + up.z + p[0][3] + p[0][1];
    S[3] = p[1][3] + up.w + rt[0] + p[0][2];
#pragma unroll
    for (int k = 1; k <= HK; ++k) {
        if (k < HK) {   // last use of the old row k-1
            A[0] = (k < HK - 1 ? p[k + 1][0] : dn.x) + p[k - 1][0];
            A[1] = (k < HK - 1 ? p[k + 1][1] : dn.y) + p[k - 1][1];
            A[2] = (k < HK - 1 ? p[k + 1][2] : dn.z) + p[k - 1][2];
            A[3] = (k < HK - 1 ? p[k + 1][3] : dn.w) + p[k - 1][3];
        }
#pragma unroll
        for (int h = 0; h < 4; ++h) {   // finish row k-1
            const float v = 0.25f * S[h] + t2[k - 1][h] - t3[k - 1][h];
            if (ALL) p[k - 1][h] = v;
            else p[k - 1][h] = ((upd >> (4 * (k - 1) + h)) & 1u) ? v : p[k - 1][h];
        }
        if (k < HK) {
            S[0] = A[0] + p[k][1] + lf[k];
            S[1] = A[1] + p[k][2] + p[k][0];
            S[2] = A[2] + p[k][3] + p[k][1];
            S[3] = A[3] + rt[k] + p[k][2];
        }
    }
}

template <int HK>
__global__ void __launch_bounds__(32 * (FSI / HK), 1)
    k_jacobi_fused3(const __grid_constant__ CUtensorMap map_p, const __grid_constant__ CUtensorMap map_src,
                    const __grid_constant__ CUtensorMap map_code, float *__restrict__ p_out, unsigned int *tile_ctr,
                    fs2d_dom d, FusedGeom g) {
    constexpr int H_THREADS = 32 * (FSI / HK);
    static_assert(HK >= 2 && FSI % HK == 0 && 4 * HK <= 32, "rows per thread");
    extern __shared__ __align__(1024) float sm[];
    uint8_t *stg_code = reinterpret_cast<uint8_t *>(sm + OFF_BYTES);
    uint8_t *wcode = stg_code + FSI * FCW;
    uint16_t *slow_list = reinterpret_cast<uint16_t *>(wcode + FN);
    __shared__ __align__(8) uint64_t bar;
    __shared__ int n_slow[2];
    __shared__ int s_next;

    const int lane = threadIdx.x;
    const int tid = threadIdx.y * 32 + lane;
    const int c = 4 * lane;                 // first tile column of this thread
    const int swz = (lane >> 2) & 1;        // order in which the two (t2, t3) chunks of a row are read
    const int lr0 = threadIdx.y * HK;       // first tile row of this thread
    const int o0 = lr0 * FSJ + c;           // plane offset of the thread's first cell
    const bool leader = tid == 0;
    const int n_tiles = g.tiles_i * g.tiles_j;
    constexpr uint32_t TX_BYTES = FN * (4 + 8) + FSI * FCW;
    constexpr uint32_t BLOCK_BITS = HK == 8 ? 0xffffffffu : (1u << (4 * HK)) - 1u;   // one bit per cell of the block
    const CUtensorMap *mp = &map_p, *ms = &map_src, *mc = &map_code;
#define FS2D_ISSUE(tile)                                                            \
    do {                                                                            \
        const int R0_ = d.r0 + f_tile_row(g, (tile) / g.tiles_j) * g.TI - g.T;                   \
        const int C0_ = ((tile) % g.tiles_j) * g.TJ - g.HJ;                         \
        mbar_expect_tx(&bar, TX_BYTES);                                             \
        tma_load_2d(sm + OFF_P0, mp, C0_, R0_, &bar);                               \
        tma_load_2d(sm + OFF_SRC, ms, 2 * C0_, R0_, &bar);                          \
        tma_load_2d(stg_code, mc, C0_ & ~15, R0_, &bar);                            \
    } while (0)

    if (leader) {
        mbar_init(&bar, 1);
        n_slow[0] = n_slow[1] = 0;
    }
    __syncthreads();
    int t = blockIdx.x;
    if (leader && t < n_tiles) FS2D_ISSUE(t);
    uint32_t parity = 0;
    // edge rows of the adjacent warps (clamped inside the tile: rim rows compute harmless garbage, see variant 1)
    const int o_up = max(lr0 - 1, 0) * FSJ + c, o_dn = min(lr0 + HK, FSI - 1) * FSJ + c;
    constexpr uint32_t FULL = 0xffffffffu;

    while (t < n_tiles) {
        const int R0 = d.r0 + f_tile_row(g, t / g.tiles_j) * g.TI - g.T;
        const int C0 = (t % g.tiles_j) * g.TJ - g.HJ;
        const int coff = C0 - (C0 & ~15);   // multiple of 4: C0 is a multiple of 4
        const int rlo = max(0, d.clo - R0), rhi = min(FSI - 1, d.chi - R0);
        const int clo = max(0, -C0), chi = min(FSJ - 1, d.Y - 1 - C0);
        const int par = parity;
        // per-thread column flags: inside the grid / on a global edge column
        uint32_t col_in = 0, col_edge = 0;
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            col_in |= (uint32_t)(c + h >= clo && c + h <= chi) << h;
            col_edge |= (uint32_t)(C0 + c + h == 0 || C0 + c + h == d.Y - 1) << h;
        }

        mbar_wait(&bar, parity);
        parity ^= 1;

        // ---- per-thread state from the staging buffer: HK rows x 4 columns --------------------------
        float p[HK][4], t2[HK][4], t3[HK][4];
        uint32_t upd = 0, slow = 0;   // bit 4k + h: row k, column c + h
#pragma unroll
        for (int k = 0; k < HK; ++k) {
            const int lr = lr0 + k, o = o0 + k * FSJ;
            const float4 pv = lds4(sm + OFF_P0 + o);
            // (t2, t3) of the thread's 4 columns = two 16-byte chunks at a 32-byte lane stride: read in the order
            // (even, odd) by lanes 0-3 of every 8 and (odd, even) by lanes 4-7, so each quarter-warp wavefront touches
            // all 8 bank groups once (a plain read is 2-way bank conflicted), then put them back in order
            const float4 sa = lds4(sm + OFF_SRC + 2 * o + 4 * swz), sb = lds4(sm + OFF_SRC + 2 * o + 4 * (1 - swz));
            const float4 s01 = swz ? sb : sa, s23 = swz ? sa : sb;
            p[k][0] = pv.x; p[k][1] = pv.y; p[k][2] = pv.z; p[k][3] = pv.w;
            t2[k][0] = s01.x; t3[k][0] = s01.y; t2[k][1] = s01.z; t3[k][1] = s01.w;
            t2[k][2] = s23.x; t3[k][2] = s23.y; t2[k][3] = s23.z; t3[k][3] = s23.w;
            const uint32_t cw = *reinterpret_cast<const uint32_t *>(stg_code + lr * FCW + coff + c);   // 4 pcode bytes
            *reinterpret_cast<uint32_t *>(wcode + o) = cw;
            const bool row_in = lr >= rlo && lr <= rhi;
            const bool row_edge = R0 + lr == d.clo || R0 + lr == d.chi;
            if (cw == 0u && !row_edge && col_edge == 0u) {   // four open-fluid cells without BC neighbours (the common case)
                if (row_in) upd |= col_in << (4 * k);
            } else {
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const uint32_t pc = (cw >> (8 * h)) & 0xffu, code = pc & 15u;
                    const bool relaxed = code == FS2D_PC_FLUID || code == FS2D_PC_INFLOW || code == FS2D_PC_OUTFLOW;
                    const bool u = row_in && ((col_in >> h) & 1u) && relaxed;   // rim cells may compute garbage, see variant 1
                    const bool sl = u && ((pc >> 4) != 0u || row_edge || ((col_edge >> h) & 1u));
                    upd |= (uint32_t)u << (4 * k + h);
                    slow |= (uint32_t)sl << (4 * k + h);
                    if (sl) slow_list[atomicAdd(&n_slow[par], 1)] = (uint16_t)(o + h);
                }
            }
        }
        const bool all_upd = __all_sync(FULL, upd == BLOCK_BITS) != 0;   // warp-uniform: open fluid in all of this warp's rows
        __syncthreads();  // (A) wcode and the slow list are complete
        const int ns = n_slow[par];   // block-uniform; > 0: the tile has cells next to BC cells / global edges

        // ---- T iterations --------------------------------------------------------------------------
        int cur = OFF_P0, nxt = OFF_W0;
        for (int s = 0; s < g.T; ++s) {
            if (leader && s == (g.T > 1 ? 1 : 0)) {
                n_slow[par ^ 1] = 0;
                if (g.T > 1) {  // staging is free: fetch the next tile index and start its loads
                    const int tn = (int)atomicAdd(tile_ctr, 1u) + (int)gridDim.x;
                    s_next = tn;
                    if (tn < n_tiles) FS2D_ISSUE(tn);
                }
            }
            if (ns > 0) {
                // Balanced fix-up: all threads share the slow cells and leave, in plane `nxt`, the SUM of the four
                // post-BC neighbour values (the reference's order) for the owning thread to pick up.
                for (int e = tid; e < ns; e += H_THREADS) {
                    const int o = slow_list[e], r = o / FSJ, cc = o % FSJ;
                    float sum = f_post(sm + cur, wcode, min(r + 1, rhi), cc, rlo, rhi, clo, chi);
                    sum = sum + f_post(sm + cur, wcode, max(r - 1, rlo), cc, rlo, rhi, clo, chi);
                    sum = sum + f_post(sm + cur, wcode, r, min(cc + 1, chi), rlo, rhi, clo, chi);
                    sum = sum + f_post(sm + cur, wcode, r, max(cc - 1, clo), rlo, rhi, clo, chi);
                    sm[nxt + o] = sum;
                }
                __syncthreads();  // (B)
            }
            const float4 upv = lds4(sm + cur + o_up), dnv = lds4(sm + cur + o_dn);
            if (all_upd) jacobi_rows<HK, true>(p, t2, t3, upd, upv, dnv);
            else jacobi_rows<HK, false>(p, t2, t3, upd, upv, dnv);
            if (slow) {   // pick up the post-BC neighbour sums left in plane `nxt` by the fix-up
#pragma unroll
                for (int k = 0; k < HK; ++k) {
                    if ((slow >> (4 * k)) & 15u) {
#pragma unroll
                        for (int h = 0; h < 4; ++h)
                            if ((slow >> (4 * k + h)) & 1u) p[k][h] = 0.25f * sm[nxt + o0 + k * FSJ + h] + t2[k][h] - t3[k][h];
                    }
                }
            }
            if (ns > 0) {   // slow tile: mirror the whole block so that f_post can read any cell
#pragma unroll
                for (int k = 0; k < HK; ++k) sts4(sm + nxt + o0 + k * FSJ, p[k][0], p[k][1], p[k][2], p[k][3]);
            } else {        // only the block's edge rows are read by other warps
                sts4(sm + nxt + o0, p[0][0], p[0][1], p[0][2], p[0][3]);
                sts4(sm + nxt + o0 + (HK - 1) * FSJ, p[HK - 1][0], p[HK - 1][1], p[HK - 1][2], p[HK - 1][3]);
            }
            __syncthreads();  // plane `nxt` is complete; after the last iteration: all plane/list reads of this tile are done
            cur = nxt;
            nxt = (nxt == OFF_W0) ? OFF_W1 : OFF_W0;
        }

        // ---- store the inner (TI x TJ) cells that were updated and belong to rows [r0, r1) ----------
        // HJ, TJ, C0 and Y are multiples of 4, so a thread's four columns are inside or outside together
        if (c >= g.HJ && c < g.HJ + g.TJ && C0 + c < d.Y) {
#pragma unroll
            for (int k = 0; k < HK; ++k) {
                const int lr = lr0 + k, gr = R0 + lr;
                const uint32_t m = (upd >> (4 * k)) & 15u;
                if (lr >= g.T && lr < g.T + g.TI && gr < d.r1 && m) {
                    float *dst = p_out + (size_t)gr * d.Y + (C0 + c);
                    if (m == 15u) {
                        *reinterpret_cast<float4 *>(dst) = make_float4(p[k][0], p[k][1], p[k][2], p[k][3]);
                    } else {
#pragma unroll
                        for (int h = 0; h < 4; ++h)
                            if ((m >> h) & 1u) dst[h] = p[k][h];
                    }
                }
            }
        }
        if (g.T == 1) {   // the staging plane doubled as the only `cur` plane: release it only now
            if (leader) {
                const int tn = (int)atomicAdd(tile_ctr, 1u) + (int)gridDim.x;
                s_next = tn;
                if (tn < n_tiles) FS2D_ISSUE(tn);
            }
            __syncthreads();
        }
        t = s_next;
    }
#undef FS2D_ISSUE
}

// ---------------------------------------------------------------------------------------------
// Variant 5: the register-tile kernel on a TALLER tile (96 x 128 cells, 12 warps x 8 rows, 3 warps per scheduler).
// 96 rows raise the useful fraction of a tile (T = 8: 0.73 vs 0.66), amortise the per-tile load/store phases over
// 1.5x more cells and give the schedulers a third warp to hide shuffle / LDS latencies.  The shared memory no longer
// holds two full working planes next to the staging buffers (that would need 246 KB), so
//   * open-fluid tiles exchange only the edge rows of each warp through a small ping-pong buffer (EX), and
//   * tiles with cells next to BC cells / global edges ("slow" tiles, 8 % of bc2 at 8192^2) re-use the (t2, t3)
//     staging area -- free once every thread has its source terms in registers -- as the second working plane and the
//     slow-cell list, read pcode straight from its staging buffer, and therefore start the next tile's TMA loads
//     only when they are done (no overlap for these tiles).
// ---------------------------------------------------------------------------------------------
constexpr int VSI = 96;                       // tile rows
constexpr int VN = VSI * FSJ;                 // cells per plane
constexpr int V_WARPS = VSI / 8;              // 12
constexpr int V_THREADS = 32 * V_WARPS;       // 384
constexpr int VOFF_P0 = 0;                    // TMA destination of p; slow tiles: working plane A
constexpr int VOFF_SRC = VN;                  // TMA destination of (t2, t3), 2*VN floats; slow tiles: working plane B (VN floats) ...
constexpr int VOFF_LIST = 2 * VN;             // ... and the slow-cell list (VN uint16)
constexpr int VOFF_EX = 3 * VN;               // edge-row exchange: [2][V_WARPS][2 rows][FSJ] floats
constexpr int VEX_PLANE = V_WARPS * 2 * FSJ;
constexpr int VOFF_BYTES = VOFF_EX + 2 * VEX_PLANE;   // staged pcode, VSI x FCW bytes
constexpr size_t V_SMEM = (size_t)VOFF_BYTES * 4 + (size_t)VSI * FCW;
static_assert((VOFF_BYTES * 4) % 128 == 0 && (VOFF_SRC * 4) % 128 == 0, "TMA destinations must be 128-byte aligned");

// f_post for a code array with its own row pitch (the staged pcode box)
__device__ __forceinline__ float f_post_p(const float *pl, const uint8_t *code, int cpitch, int r, int c, int rlo, int rhi,
                                          int clo, int chi) {
    const int rm = max(r - 1, rlo), rp = min(r + 1, rhi), cm = max(c - 1, clo), cp = min(c + 1, chi);
    int a = r * FSJ + c, b = a, mode = 0;  // mode 0: value of cell a; 1: (a + b) / 2; 2: zero
    switch (code[r * cpitch + c] & 15) {
        case FS2D_PC_W_IM: a = rm * FSJ + c; break;
        case FS2D_PC_W_IP: a = rp * FSJ + c; break;
        case FS2D_PC_W_JM: a = r * FSJ + cm; break;
        case FS2D_PC_W_JP: a = r * FSJ + cp; break;
        case FS2D_PC_W_IM_JP: a = rm * FSJ + c; b = r * FSJ + cp; mode = 1; break;
        case FS2D_PC_W_IP_JP: a = rp * FSJ + c; b = r * FSJ + cp; mode = 1; break;
        case FS2D_PC_W_IM_JM: a = rm * FSJ + c; b = r * FSJ + cm; mode = 1; break;
        case FS2D_PC_W_IP_JM: a = rp * FSJ + c; b = r * FSJ + cm; mode = 1; break;
        case FS2D_PC_INFLOW: a = rp * FSJ + c; break;
        case FS2D_PC_OUTFLOW: mode = 2; break;
        default: break;  // FLUID / W_NONE: the stored value
    }
    const float va = pl[a], vb = pl[b];
    return mode == 0 ? va : (mode == 1 ? (va + vb) / 2.0f : 0.0f);
}

// f_post_p split into its index part and its value part.  FASTSLOW tiles resolve, ONCE per tile, which plane cells
// the post-BC value of each neighbour of a slow cell reads (two 14-bit plane offsets + a 2-bit mode in one word); the
// fix-up of every iteration is then four table look-ups instead of four walks through the pcode switch.
__device__ __forceinline__ uint32_t f_resolve_p(const uint8_t *code, int cpitch, int r, int c, int rlo, int rhi, int clo, int chi) {
    const int rm = max(r - 1, rlo), rp = min(r + 1, rhi), cm = max(c - 1, clo), cp = min(c + 1, chi);
    int a = r * FSJ + c, b = a, mode = 0;  // mode 0: value of cell a; 1: (a + b) / 2; 2: zero
    switch (code[r * cpitch + c] & 15) {
        case FS2D_PC_W_IM: a = rm * FSJ + c; break;
        case FS2D_PC_W_IP: a = rp * FSJ + c; break;
        case FS2D_PC_W_JM: a = r * FSJ + cm; break;
        case FS2D_PC_W_JP: a = r * FSJ + cp; break;
        case FS2D_PC_W_IM_JP: a = rm * FSJ + c; b = r * FSJ + cp; mode = 1; break;
        case FS2D_PC_W_IP_JP: a = rp * FSJ + c; b = r * FSJ + cp; mode = 1; break;
        case FS2D_PC_W_IM_JM: a = rm * FSJ + c; b = r * FSJ + cm; mode = 1; break;
        case FS2D_PC_W_IP_JM: a = rp * FSJ + c; b = r * FSJ + cm; mode = 1; break;
        case FS2D_PC_INFLOW: a = rp * FSJ + c; break;
        case FS2D_PC_OUTFLOW: mode = 2; break;
        default: break;  // FLUID / W_NONE: the stored value
    }
    return (uint32_t)a | ((uint32_t)b << 14) | ((uint32_t)mode << 28);
}
__device__ __forceinline__ float f_resolved_value(const float *pl, uint32_t x) {
    const float va = pl[x & 0x3fffu], vb = pl[(x >> 14) & 0x3fffu];
    const uint32_t mode = x >> 28;
    return mode == 0u ? va : (mode == 1u ? (va + vb) / 2.0f : 0.0f);   // the expression of f_post_p
}
constexpr int FS_CAP = VN / 8;   // slow cells per tile the resolved table has room for (behind the slow-cell list): 1536
static_assert(VN <= (1 << 14), "plane offsets must fit 14 bits");

// bar.sync on a named barrier shared by `count` threads (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
// Variant 6 (EXPERIMENTAL, PAIR = true): in open-fluid tiles a warp's iteration depends only on the edge rows of the warps
// directly above and below it, so after iteration 0 the CTA-wide barrier per iteration is replaced by two 64-thread named
// barriers (with the upper and with the lower neighbour; even-indexed pairs first, so the waits cannot form a cycle).
// Warps may then drift by an iteration against their neighbours: one warp's shuffle / LDS phase overlaps another's FP32
// phase instead of all twelve stalling together.  The ping-pong exchange buffer is deep enough: warp w reads the rows
// its neighbours wrote for iteration s at the start of iteration s + 1 and meets them at the pair barrier after its own
// write, before they can write iteration s + 2.  Slow tiles keep the CTA barriers.  Same arithmetic: bit-identical.
//
// EMIT (EXPERIMENTAL, the "tail" pass of fs2d_jacobi_update when fs2d_set_tuning(4, 1)): besides its output the pass
// stores, into the wall-BC cells of its INPUT array `emit` (= p_in; those cells are never read, their values are
// recomputed from pcode), the BC values of its PENULTIMATE state.  The reference leaves exactly these values in the
// wall cells of the buffer its last sweep writes (fs/pressure_updater.py:56-60: BC'd in place one iteration earlier,
// SURVEY T1), so a pass of T iterations ending at iteration n - 1, followed by ONE literal iteration, reproduces both
// physical buffers -- instead of ending every update with two literal iterations.  A wall-BC cell has a fluid
// neighbour, which is a slow cell of the same loaded tile, so only slow tiles emit.
//
// FASTSLOW (EXPERIMENTAL, variants 7 / 8): slow tiles with at most FS_CAP slow cells use the resolved table above.
template <bool PAIR, bool EMIT, bool FASTSLOW>
__device__ __forceinline__ void jacobi_fused5_body(const CUtensorMap *mp, const CUtensorMap *ms, const CUtensorMap *mc,
                                                   float *__restrict__ p_out, unsigned int *tile_ctr, const fs2d_dom &d,
                                                   const FusedGeom &g, float *emit) {
    constexpr int HK = 8;
    extern __shared__ __align__(1024) float sm[];
    uint8_t *stg_code = reinterpret_cast<uint8_t *>(sm + VOFF_BYTES);
    uint16_t *slow_list = reinterpret_cast<uint16_t *>(sm + VOFF_LIST);
    __shared__ __align__(8) uint64_t bar;
    __shared__ int n_slow;
    __shared__ int s_next;

    const int lane = threadIdx.x, w = threadIdx.y;
    const int tid = w * 32 + lane;
    const int c = 4 * lane;                 // first tile column of this thread
    const int swz = (lane >> 2) & 1;        // order in which the two (t2, t3) chunks of a row are read
    const int lr0 = w * HK;                 // first tile row of this thread
    const int o0 = lr0 * FSJ + c;           // plane offset of the thread's first cell
    const bool leader = tid == 0;
    const int n_tiles = g.tiles_i * g.tiles_j;
    constexpr uint32_t TX_BYTES = VN * (4 + 8) + VSI * FCW;
#define FS2D_ISSUE(tile)                                                            \
    do {                                                                            \
        const int R0_ = d.r0 + f_tile_row(g, (tile) / g.tiles_j) * g.TI - g.T;                   \
        const int C0_ = ((tile) % g.tiles_j) * g.TJ - g.HJ;                         \
        mbar_expect_tx(&bar, TX_BYTES);                                             \
        tma_load_2d(sm + VOFF_P0, mp, C0_, R0_, &bar);                              \
        tma_load_2d(sm + VOFF_SRC, ms, 2 * C0_, R0_, &bar);                         \
        tma_load_2d(stg_code, mc, C0_ & ~15, R0_, &bar);                            \
    } while (0)
#define FS2D_NEXT_TILE()                                                            \
    do {                                                                            \
        const int tn = (int)atomicAdd(tile_ctr, 1u) + (int)gridDim.x;               \
        s_next = tn;                                                                \
        if (tn < n_tiles) FS2D_ISSUE(tn);                                           \
    } while (0)

    if (leader) mbar_init(&bar, 1);
    __syncthreads();
    int t = blockIdx.x;
    if (leader && t < n_tiles) FS2D_ISSUE(t);
    uint32_t parity = 0;
    // rows adjacent to the thread's block inside a full plane (clamped inside the tile: rim rows compute harmless
    // garbage, see variant 1) and inside the exchange buffer (edge rows of the adjacent warps)
    const int o_up = max(lr0 - 1, 0) * FSJ + c, o_dn = min(lr0 + HK, VSI - 1) * FSJ + c;
    const int x_own = VOFF_EX + (w * 2) * FSJ + c;                                  // this warp's top row; + FSJ: bottom row
    const int x_up = w > 0 ? VOFF_EX + ((w - 1) * 2 + 1) * FSJ + c : x_own;         // bottom row of the warp above
    const int x_dn = w < V_WARPS - 1 ? VOFF_EX + ((w + 1) * 2) * FSJ + c : x_own + FSJ;
    constexpr uint32_t FULL = 0xffffffffu;

    while (t < n_tiles) {
        const int R0 = d.r0 + f_tile_row(g, t / g.tiles_j) * g.TI - g.T;
        const int C0 = (t % g.tiles_j) * g.TJ - g.HJ;
        const int coff = C0 - (C0 & ~15);   // multiple of 4: C0 is a multiple of 4
        const int rlo = max(0, d.clo - R0), rhi = min(VSI - 1, d.chi - R0);
        const int clo = max(0, -C0), chi = min(FSJ - 1, d.Y - 1 - C0);
        uint32_t col_in = 0, col_edge = 0;   // per-thread column flags: inside the grid / on a global edge column
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            col_in |= (uint32_t)(c + h >= clo && c + h <= chi) << h;
            col_edge |= (uint32_t)(C0 + c + h == 0 || C0 + c + h == d.Y - 1) << h;
        }
        if (leader) n_slow = 0;   // ordered before the list is built by the barrier below

        mbar_wait(&bar, parity);
        parity ^= 1;

        // ---- per-thread state from the staging buffer: HK rows x 4 columns --------------------------
        float p[HK][4], t2[HK][4], t3[HK][4];
        uint32_t upd = 0, slow = 0;   // bit 4k + h: row k, column c + h
#pragma unroll
        for (int k = 0; k < HK; ++k) {
            const int lr = lr0 + k, o = o0 + k * FSJ;
            const float4 pv = lds4(sm + VOFF_P0 + o);
            // two 16-byte chunks at a 32-byte lane stride, read in swizzled order (conflict-free, see variant 3)
            const float4 sa = lds4(sm + VOFF_SRC + 2 * o + 4 * swz), sb = lds4(sm + VOFF_SRC + 2 * o + 4 * (1 - swz));
            const float4 s01 = swz ? sb : sa, s23 = swz ? sa : sb;
            p[k][0] = pv.x; p[k][1] = pv.y; p[k][2] = pv.z; p[k][3] = pv.w;
            t2[k][0] = s01.x; t3[k][0] = s01.y; t2[k][1] = s01.z; t3[k][1] = s01.w;
            t2[k][2] = s23.x; t3[k][2] = s23.y; t2[k][3] = s23.z; t3[k][3] = s23.w;
            const uint32_t cw = *reinterpret_cast<const uint32_t *>(stg_code + lr * FCW + coff + c);   // 4 pcode bytes
            const bool row_in = lr >= rlo && lr <= rhi;
            const bool row_edge = R0 + lr == d.clo || R0 + lr == d.chi;
            if (cw == 0u && !row_edge && col_edge == 0u) {   // four open-fluid cells without BC neighbours (the common case)
                if (row_in) upd |= col_in << (4 * k);
            } else {
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const uint32_t pc = (cw >> (8 * h)) & 0xffu, code = pc & 15u;
                    const bool relaxed = code == FS2D_PC_FLUID || code == FS2D_PC_INFLOW || code == FS2D_PC_OUTFLOW;
                    const bool u = row_in && ((col_in >> h) & 1u) && relaxed;   // rim cells may compute garbage, see variant 1
                    const bool sl = u && ((pc >> 4) != 0u || row_edge || ((col_edge >> h) & 1u));
                    upd |= (uint32_t)u << (4 * k + h);
                    slow |= (uint32_t)sl << (4 * k + h);
                }
            }
        }
        const bool all_upd = __all_sync(FULL, upd == FULL) != 0;   // warp-uniform: open fluid in all of this warp's rows
        // every thread has left the staging buffers; block-uniform verdict: does the tile have slow cells?
        const bool tile_slow = __syncthreads_or(slow != 0u) != 0;

        if (!tile_slow) {
            // ---- open-fluid tile: T iterations, only the warps' edge rows go through shared memory ---------------
            for (int s = 0; s < g.T; ++s) {
                if (leader && s == 1) FS2D_NEXT_TILE();   // staging is free (T > 1): start the next tile's loads
                const int xs = ((s + 1) & 1) * VEX_PLANE;   // exchange plane written by iteration s - 1
                const float4 upv = lds4(sm + (s == 0 ? VOFF_P0 + o_up : x_up + xs));
                const float4 dnv = lds4(sm + (s == 0 ? VOFF_P0 + o_dn : x_dn + xs));
                if (all_upd) jacobi_rows<HK, true>(p, t2, t3, upd, upv, dnv);
                else jacobi_rows<HK, false>(p, t2, t3, upd, upv, dnv);
                if (s + 1 < g.T) {
                    const int xw = (s & 1) * VEX_PLANE;
                    sts4(sm + x_own + xw, p[0][0], p[0][1], p[0][2], p[0][3]);
                    sts4(sm + x_own + xw + FSJ, p[HK - 1][0], p[HK - 1][1], p[HK - 1][2], p[HK - 1][3]);
                    if (!PAIR || s == 0) {
                        __syncthreads();   // edge rows of iteration s are complete (and s_next is visible after s == 1)
                    } else {               // (after s == 0 the CTA barrier also frees the staged plane for the prefetch)
                        // pair barrier id 1 + (index of the pair's upper warp); even-indexed pairs first
                        if ((w & 1) == 0) {
                            if (w < V_WARPS - 1) named_bar_sync(1 + w, 64);
                            if (w > 0) named_bar_sync(w, 64);
                        } else {
                            named_bar_sync(w, 64);
                            if (w < V_WARPS - 1) named_bar_sync(1 + w, 64);
                        }
                    }
                }
            }
            if (PAIR && g.T > 2) __syncthreads();   // s_next (written at s == 1) becomes visible to every warp
        } else {
            // ---- slow tile: full working planes in P0 / the (t2, t3) staging area, cooperative fix-up ------------
#pragma unroll
            for (int k = 0; k < HK; ++k) {
                if ((slow >> (4 * k)) & 15u) {
#pragma unroll
                    for (int h = 0; h < 4; ++h)
                        if ((slow >> (4 * k + h)) & 1u) slow_list[atomicAdd(&n_slow, 1)] = (uint16_t)(o0 + k * FSJ + h);
                }
            }
            __syncthreads();
            const int ns = n_slow;
            const uint8_t *code = stg_code + coff;   // tile pcode, row pitch FCW (the staging buffer stays intact)
            // resolved neighbour table: 4 words per slow cell, behind the list (VN uint16 = VN / 2 floats); each thread reads
            // back only the entries it wrote (same e -> thread mapping), so no barrier is needed
            uint32_t *res = reinterpret_cast<uint32_t *>(sm + VOFF_LIST + VN / 2);
            const bool fast = FASTSLOW && ns <= FS_CAP;   // block-uniform
            if (fast) {
                for (int e = tid; e < ns; e += V_THREADS) {
                    const int o = slow_list[e], r = o / FSJ, cc = o % FSJ;
                    res[4 * e + 0] = f_resolve_p(code, FCW, min(r + 1, rhi), cc, rlo, rhi, clo, chi);
                    res[4 * e + 1] = f_resolve_p(code, FCW, max(r - 1, rlo), cc, rlo, rhi, clo, chi);
                    res[4 * e + 2] = f_resolve_p(code, FCW, r, min(cc + 1, chi), rlo, rhi, clo, chi);
                    res[4 * e + 3] = f_resolve_p(code, FCW, r, max(cc - 1, clo), rlo, rhi, clo, chi);
                }
            }
            int cur = VOFF_P0, nxt = VOFF_SRC;
            for (int s = 0; s < g.T; ++s) {
                if (EMIT && s == g.T - 1) {
                    // plane `cur` holds the state after T - 1 iterations: its BC values go to the wall-BC cells of the
                    // tile's output region in the input array
                    for (int e = tid; e < g.TI * g.TJ; e += V_THREADS) {
                        const int r = g.T + e / g.TJ, cc = g.HJ + e % g.TJ, gr = R0 + r, gc = C0 + cc;
                        if (gr >= d.r1 || gc >= d.Y) continue;
                        const int cd = code[r * FCW + cc] & 15;
                        if (cd >= FS2D_PC_W_IM && cd <= FS2D_PC_W_IP_JM)
                            emit[(size_t)gr * d.Y + gc] = f_post_p(sm + cur, code, FCW, r, cc, rlo, rhi, clo, chi);
                    }
                }
                // all threads share the slow cells and leave, in plane `nxt`, the SUM of the four post-BC neighbour
                // values (the reference's order) for the owning thread to pick up
                if (fast) {
                    for (int e = tid; e < ns; e += V_THREADS) {
                        const uint4 q = *reinterpret_cast<const uint4 *>(res + 4 * e);
                        float sum = f_resolved_value(sm + cur, q.x);
                        sum = sum + f_resolved_value(sm + cur, q.y);
                        sum = sum + f_resolved_value(sm + cur, q.z);
                        sum = sum + f_resolved_value(sm + cur, q.w);
                        sm[nxt + slow_list[e]] = sum;
                    }
                } else {
                    for (int e = tid; e < ns; e += V_THREADS) {
                        const int o = slow_list[e], r = o / FSJ, cc = o % FSJ;
                        float sum = f_post_p(sm + cur, code, FCW, min(r + 1, rhi), cc, rlo, rhi, clo, chi);
                        sum = sum + f_post_p(sm + cur, code, FCW, max(r - 1, rlo), cc, rlo, rhi, clo, chi);
                        sum = sum + f_post_p(sm + cur, code, FCW, r, min(cc + 1, chi), rlo, rhi, clo, chi);
                        sum = sum + f_post_p(sm + cur, code, FCW, r, max(cc - 1, clo), rlo, rhi, clo, chi);
                        sm[nxt + o] = sum;
                    }
                }
                __syncthreads();
                const float4 upv = lds4(sm + cur + o_up), dnv = lds4(sm + cur + o_dn);
                if (all_upd) jacobi_rows<HK, true>(p, t2, t3, upd, upv, dnv);
                else jacobi_rows<HK, false>(p, t2, t3, upd, upv, dnv);
                if (slow) {
#pragma unroll
                    for (int k = 0; k < HK; ++k) {
                        if ((slow >> (4 * k)) & 15u) {
#pragma unroll
                            for (int h = 0; h < 4; ++h)
                                if ((slow >> (4 * k + h)) & 1u) p[k][h] = 0.25f * sm[nxt + o0 + k * FSJ + h] + t2[k][h] - t3[k][h];
                        }
                    }
                }
                if (s + 1 < g.T) {   // mirror the whole block so that f_post can read any cell
#pragma unroll
                    for (int k = 0; k < HK; ++k) sts4(sm + nxt + o0 + k * FSJ, p[k][0], p[k][1], p[k][2], p[k][3]);
                }
                __syncthreads();   // plane `nxt` complete; after the last iteration: all plane / list / pcode reads are done
                const int x = cur; cur = nxt; nxt = x;
            }
        }
        const bool deferred = tile_slow || g.T == 1;   // the next tile's loads could not be started during the iterations
        if (deferred) {
            if (!tile_slow) __syncthreads();           // T == 1: the halo rows of iteration 0 were read from the staged plane
            if (leader) FS2D_NEXT_TILE();
        }

        // ---- store the inner (TI x TJ) cells that were updated and belong to rows [r0, r1) ----------
        // HJ, TJ, C0 and Y are multiples of 4, so a thread's four columns are inside or outside together
        if (c >= g.HJ && c < g.HJ + g.TJ && C0 + c < d.Y) {
#pragma unroll
            for (int k = 0; k < HK; ++k) {
                const int lr = lr0 + k, gr = R0 + lr;
                const uint32_t m = (upd >> (4 * k)) & 15u;
                if (lr >= g.T && lr < g.T + g.TI && gr < d.r1 && m) {
                    float *dst = p_out + (size_t)gr * d.Y + (C0 + c);
                    if (m == 15u) {
                        *reinterpret_cast<float4 *>(dst) = make_float4(p[k][0], p[k][1], p[k][2], p[k][3]);
                    } else {
#pragma unroll
                        for (int h = 0; h < 4; ++h)
                            if ((m >> h) & 1u) dst[h] = p[k][h];
                    }
                }
            }
        }
        if (deferred || g.T == 2) __syncthreads();   // s_next visible to all (T >= 3: a barrier followed the s == 1 write)
        t = s_next;
    }
#undef FS2D_ISSUE
#undef FS2D_NEXT_TILE
}

__global__ void __launch_bounds__(V_THREADS, 1)
    k_jacobi_fused5(const __grid_constant__ CUtensorMap map_p, const __grid_constant__ CUtensorMap map_src,
                    const __grid_constant__ CUtensorMap map_code, float *__restrict__ p_out, unsigned int *tile_ctr,
                    fs2d_dom d, FusedGeom g) {
    // the descriptors must be addressed in the kernel-parameter space: take their addresses here
    jacobi_fused5_body<false, false, false>(&map_p, &map_src, &map_code, p_out, tile_ctr, d, g, nullptr);
}
__global__ void __launch_bounds__(V_THREADS, 1)
    k_jacobi_fused5e(const __grid_constant__ CUtensorMap map_p, const __grid_constant__ CUtensorMap map_src,
                     const __grid_constant__ CUtensorMap map_code, float *__restrict__ p_out, unsigned int *tile_ctr,
                     fs2d_dom d, FusedGeom g, float *emit) {
    jacobi_fused5_body<false, true, false>(&map_p, &map_src, &map_code, p_out, tile_ctr, d, g, emit);
}
__global__ void __launch_bounds__(V_THREADS, 1)
    k_jacobi_fused6e(const __grid_constant__ CUtensorMap map_p, const __grid_constant__ CUtensorMap map_src,
                     const __grid_constant__ CUtensorMap map_code, float *__restrict__ p_out, unsigned int *tile_ctr,
                     fs2d_dom d, FusedGeom g, float *emit) {
    jacobi_fused5_body<true, true, false>(&map_p, &map_src, &map_code, p_out, tile_ctr, d, g, emit);
}
__global__ void __launch_bounds__(V_THREADS, 1)
    k_jacobi_fused6(const __grid_constant__ CUtensorMap map_p, const __grid_constant__ CUtensorMap map_src,
                    const __grid_constant__ CUtensorMap map_code, float *__restrict__ p_out, unsigned int *tile_ctr,
                    fs2d_dom d, FusedGeom g) {
    jacobi_fused5_body<true, false, false>(&map_p, &map_src, &map_code, p_out, tile_ctr, d, g, nullptr);
}

// variants 7 / 8 (EXPERIMENTAL): variants 5 / 6 with the resolved slow-cell table; *e: with the emitting tail
#define FS2D_FUSED_KERNEL(NAME, PAIR, EMIT)                                                                                     \
    __global__ void __launch_bounds__(V_THREADS, 1)                                                                             \
        NAME(const __grid_constant__ CUtensorMap map_p, const __grid_constant__ CUtensorMap map_src,                            \
             const __grid_constant__ CUtensorMap map_code, float *__restrict__ p_out, unsigned int *tile_ctr, fs2d_dom d,       \
             FusedGeom g, float *emit) {                                                                                        \
        jacobi_fused5_body<PAIR, EMIT, true>(&map_p, &map_src, &map_code, p_out, tile_ctr, d, g, emit);                         \
    }
FS2D_FUSED_KERNEL(k_jacobi_fused7, false, false)
FS2D_FUSED_KERNEL(k_jacobi_fused7e, false, true)
FS2D_FUSED_KERNEL(k_jacobi_fused8, true, false)
FS2D_FUSED_KERNEL(k_jacobi_fused8e, true, true)
#undef FS2D_FUSED_KERNEL

// (A packed fp32x2 variant -- FADD2/FFMA2, column-pair ownership -- was measured at 915 us/pass vs 795 us for
// variant 1 at 8192^2, T=8: bank-conflicted scalar j-neighbour loads and pack/unpack moves; removed.)
int g_tail_emit = 0;       // fs2d_set_tuning(4, v): 1 = end fs2d_jacobi_update with {emitting pass, ONE literal iteration} (experimental)
int g_fused_variant = 5;   // fs2d_set_tuning(1, v): 1 = one column per thread (64 x 128 tile, smem planes); 3 = register tile +
                           // shuffles (64 x 128 tile); 5 = register tile on a 96 x 128 tile; 6 = 5 with pair barriers (experimental)

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

int make_map(CUtensorMap *m, CUtensorMapDataType dt, size_t esz, const void *base, uint64_t cols, uint64_t rows,
                    uint32_t box_cols, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return FS2D_E_CUDA;
    }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * esz};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, dt, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (base %p, %llu x %llu)", (int)r, base,
                  (unsigned long long)rows, (unsigned long long)cols);
        return FS2D_E_CUDA;
    }
    return FS2D_OK;
}

int fused_tile_rows() { return g_fused_variant >= 5 ? VSI : FSI; }

bool fused_supported(const float *pa, const float *pb, const float *src, const uint8_t *pcode, const fs2d_dom &d) {
    return d.Y % 16 == 0 && ((uintptr_t)pa % 16 == 0) && ((uintptr_t)pb % 16 == 0) && ((uintptr_t)src % 16 == 0) &&
           ((uintptr_t)pcode % 16 == 0);
}

int fused_pass(const float *p_in, float *p_out, const float *src, const uint8_t *pcode, const fs2d_dom &d, int T,
               cudaStream_t s, int skip_from, int skip_n, bool emit) {
    static int n_sm = 0;
    static bool attr_set = false;
    if (!n_sm) {
        int dev = 0;
        FS2D_CUDA_CHECK(cudaGetDevice(&dev));
        FS2D_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    if (!attr_set) {
        FS2D_CUDA_CHECK(cudaFuncSetAttribute(k_jacobi_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F_SMEM));
        FS2D_CUDA_CHECK(cudaFuncSetAttribute(k_jacobi_fused3<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F_SMEM));
        FS2D_CUDA_CHECK(cudaFuncSetAttribute(k_jacobi_fused5, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V_SMEM));
        FS2D_CUDA_CHECK(cudaFuncSetAttribute(k_jacobi_fused6, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V_SMEM));
        FS2D_CUDA_CHECK(cudaFuncSetAttribute(k_jacobi_fused5e, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V_SMEM));
        FS2D_CUDA_CHECK(cudaFuncSetAttribute(k_jacobi_fused6e, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V_SMEM));
        FS2D_CUDA_CHECK(cudaFuncSetAttribute(k_jacobi_fused7, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V_SMEM));
        FS2D_CUDA_CHECK(cudaFuncSetAttribute(k_jacobi_fused7e, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V_SMEM));
        FS2D_CUDA_CHECK(cudaFuncSetAttribute(k_jacobi_fused8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V_SMEM));
        FS2D_CUDA_CHECK(cudaFuncSetAttribute(k_jacobi_fused8e, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V_SMEM));
        attr_set = true;
    }
    CUtensorMap mp, ms, mc;
    const int tile_rows = fused_tile_rows();
    if (int e = make_map(&mp, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, p_in, d.Y, d.rows, FSJ, tile_rows)) return e;
    if (int e = make_map(&ms, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, src, 2ull * d.Y, d.rows, 2 * FSJ, tile_rows)) return e;
    if (int e = make_map(&mc, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, pcode, d.Y, d.rows, FCW, tile_rows)) return e;
    FusedGeom g;
    g.T = T;
    g.HJ = (T + 3) & ~3;
    g.TI = tile_rows - 2 * T;
    g.TJ = FSJ - 2 * g.HJ;
    const int all_rows = (d.r1 - d.r0 + g.TI - 1) / g.TI;
    if (skip_n < 0 || skip_from < 0 || skip_from + skip_n > all_rows) {
        set_error("bad argument: skipped tile rows [%d, %d) outside the %d tile rows of the pass", skip_from, skip_from + skip_n,
                  all_rows);
        return FS2D_E_BADARG;
    }
    g.skip_from = skip_from;
    g.skip_n = skip_n;
    g.tiles_i = all_rows - skip_n;
    g.tiles_j = (d.Y + g.TJ - 1) / g.TJ;
    if (g.tiles_i == 0) return FS2D_OK;
    const int n_tiles = g.tiles_i * g.tiles_j;
    const int grid = n_tiles < n_sm ? n_tiles : n_sm;
    static unsigned int *ctr = nullptr;   // dynamic tile scheduler: tiles next to walls cost more than open-fluid tiles
    if (!ctr) FS2D_CUDA_CHECK(cudaMalloc(&ctr, sizeof(unsigned int)));
    FS2D_CUDA_CHECK(cudaMemsetAsync(ctr, 0, sizeof(unsigned int), s));
    ++g_launches;
    if (emit) {
        if (g_fused_variant < 5) {
            set_error("the emitting tail pass exists for the fused variants 5 and 6 only");
            return FS2D_E_BADARG;
        }
        float *em = const_cast<float *>(p_in);
        const dim3 blk(32, V_WARPS, 1);
        if (g_fused_variant == 6) k_jacobi_fused6e<<<grid, blk, V_SMEM, s>>>(mp, ms, mc, p_out, ctr, d, g, em);
        else if (g_fused_variant == 7) k_jacobi_fused7e<<<grid, blk, V_SMEM, s>>>(mp, ms, mc, p_out, ctr, d, g, em);
        else if (g_fused_variant == 8) k_jacobi_fused8e<<<grid, blk, V_SMEM, s>>>(mp, ms, mc, p_out, ctr, d, g, em);
        else k_jacobi_fused5e<<<grid, blk, V_SMEM, s>>>(mp, ms, mc, p_out, ctr, d, g, em);
        return FS2D_OK;
    }
    if (g_fused_variant == 3) k_jacobi_fused3<8><<<grid, dim3(32, FSI / 8, 1), F_SMEM, s>>>(mp, ms, mc, p_out, ctr, d, g);
    else if (g_fused_variant == 5) k_jacobi_fused5<<<grid, dim3(32, V_WARPS, 1), V_SMEM, s>>>(mp, ms, mc, p_out, ctr, d, g);
    else if (g_fused_variant == 6) k_jacobi_fused6<<<grid, dim3(32, V_WARPS, 1), V_SMEM, s>>>(mp, ms, mc, p_out, ctr, d, g);
    else if (g_fused_variant == 7) k_jacobi_fused7<<<grid, dim3(32, V_WARPS, 1), V_SMEM, s>>>(mp, ms, mc, p_out, ctr, d, g, nullptr);
    else if (g_fused_variant == 8) k_jacobi_fused8<<<grid, dim3(32, V_WARPS, 1), V_SMEM, s>>>(mp, ms, mc, p_out, ctr, d, g, nullptr);
    else k_jacobi_fused<<<grid, dim3(FSJ, FNTY, 1), F_SMEM, s>>>(mp, ms, mc, p_out, ctr, d, g);
    return FS2D_OK;
}

}  // namespace fs2d

using namespace fs2d;

extern "C" {

int fs2d_fused_tile(int T, int *rows, int *cols, int *halo_rows, int *halo_cols, int *t_max) {
    if (rows) *rows = fused_tile_rows();
    if (cols) *cols = FSJ;
    if (halo_rows) *halo_rows = T;
    if (halo_cols) *halo_cols = (T + 3) & ~3;
    if (t_max) *t_max = F_TMAX;
    return FS2D_OK;
}

int fs2d_jacobi_fused(float *p_out, const float *p_in, const float *src, const uint8_t *pcode, fs2d_dom d, int T,
                      void *stream) {
    FS2D_REQUIRE(p_out && p_in && src && pcode && p_out != p_in, "null/aliased field pointer");
    FS2D_REQUIRE(T >= 1 && T <= F_TMAX, "fused iteration count out of range");
    FS2D_REQUIRE(d.Y % 16 == 0 && ((uintptr_t)p_in % 16 == 0) && ((uintptr_t)p_out % 16 == 0) && ((uintptr_t)src % 16 == 0) &&
                     ((uintptr_t)pcode % 16 == 0),
                 "fused Jacobi needs Y % 16 == 0 and 16-byte aligned fields (TMA row pitch / base alignment)");
    if (int e = check_dom(d)) return e;
    if (d.r1 == d.r0) return FS2D_OK;
    if (int e = fused_pass(p_in, p_out, src, pcode, d, T, (cudaStream_t)stream, 0, 0)) return e;
    FS2D_LAUNCH_CHECK();
    return FS2D_OK;
}

int fs2d_jacobi_fused_part(float *p_out, const float *p_in, const float *src, const uint8_t *pcode, fs2d_dom d, int T,
                           int skip_from, int skip_n, void *stream) {
    FS2D_REQUIRE(p_out && p_in && src && pcode && p_out != p_in, "null/aliased field pointer");
    FS2D_REQUIRE(T >= 1 && T <= F_TMAX, "fused iteration count out of range");
    FS2D_REQUIRE(d.Y % 16 == 0 && ((uintptr_t)p_in % 16 == 0) && ((uintptr_t)p_out % 16 == 0) && ((uintptr_t)src % 16 == 0) &&
                     ((uintptr_t)pcode % 16 == 0),
                 "fused Jacobi needs Y % 16 == 0 and 16-byte aligned fields (TMA row pitch / base alignment)");
    if (int e = check_dom(d)) return e;
    if (d.r1 == d.r0) return FS2D_OK;
    if (int e = fused_pass(p_in, p_out, src, pcode, d, T, (cudaStream_t)stream, skip_from, skip_n)) return e;
    FS2D_LAUNCH_CHECK();
    return FS2D_OK;
}

int fs2d_jacobi_fused_tail(float *p_out, float *p_in, const float *src, const uint8_t *pcode, fs2d_dom d, int T, int skip_from,
                           int skip_n, void *stream) {
    FS2D_REQUIRE(p_out && p_in && src && pcode && p_out != p_in, "null/aliased field pointer");
    FS2D_REQUIRE(T >= 1 && T <= F_TMAX, "fused iteration count out of range");
    FS2D_REQUIRE(d.Y % 16 == 0 && ((uintptr_t)p_in % 16 == 0) && ((uintptr_t)p_out % 16 == 0) && ((uintptr_t)src % 16 == 0) &&
                     ((uintptr_t)pcode % 16 == 0),
                 "fused Jacobi needs Y % 16 == 0 and 16-byte aligned fields (TMA row pitch / base alignment)");
    if (int e = check_dom(d)) return e;
    if (d.r1 == d.r0) return FS2D_OK;
    if (int e = fused_pass(p_in, p_out, src, pcode, d, T, (cudaStream_t)stream, skip_from, skip_n, true)) return e;
    FS2D_LAUNCH_CHECK();
    return FS2D_OK;
}

}  // extern "C"

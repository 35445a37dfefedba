// fs2d_fused.cu -- T Jacobi iterations (pressure BC + sweep, fs/pressure_updater.py:56-66) per pass over HBM
// (temporal blocking) for sm_100a.
//
// Why it is exact: one reference iteration is local -- a relaxed cell reads the post-BC values of its 4 neighbours, and a
// BC cell's value is a function of cells at most one step further (fs/boundary_condition.py:41-65).  A CTA therefore
// loads a 96 x 128 tile of p and of the velocity source terms (t2, t3), runs T iterations on it while the region in
// which values are still correct shrinks by one cell per iteration from every side that is not a global edge (the host
// verifies this bound for the actual mask, fs/_bc_tables.py:fused_reach_ok), and stores the inner (96-2T) x (128-2HJ)
// cells.  Every cell-iteration evaluates the literal expression 0.25*(p(i+1,j)+p(i-1,j)+p(i,j+1)+p(i,j-1)) + t2 - t3
// (FMUL + FADD, no FMA), so results are bit-identical to T separate sweeps.
//
// Structure (B200), one persistent CTA of 12 warps per SM:
//   * REGISTER TILE.  Warp w owns tile rows 8w..8w+7 and all 128 columns, lane l the columns 4l..4l+3: a thread keeps an
//     8 x 4 block of p, t2 and t3 in registers for the whole pass.  j-neighbours outside the block come from the adjacent
//     lanes by two shuffles per row, i-neighbours outside it are the edge rows of the adjacent warps.
//   * AUTONOMOUS WARPS (tiles of open fluid, "PURE").  Nothing in such a tile needs the whole CTA: every warp has its own
//     slice of the staging buffer, its own mbarrier and issues its own TMA box loads (cp.async.bulk.tensor -> UTMALDG) for
//     the NEXT tile as soon as its registers hold the current one; neighbouring warps hand their edge rows over through a
//     small ping-pong exchange buffer guarded by per-warp progress counters (st.release / ld.acquire in shared memory)
//     instead of CTA barriers.  The warps of a CTA therefore drift apart by up to an iteration per warp, so one warp's
//     load / store phase overlaps the arithmetic of the others -- with CTA barriers (round 1) all twelve warps went through
//     the latency-bound phases together and the kernel spent half its time outside the iteration body.
//   * Tiles with cells next to BC cells / global edges ("SLOW"; 8 % of bc2 at 8192^2) keep the cooperative scheme: full
//     working planes in the staging area, a per-tile list of the cells whose neighbours need post-BC values, their BC
//     source cells resolved once per tile into a table, CTA barriers.
//   * TILE ORDER.  The class of every tile of a pass (PURE / SLOW / nothing to do) depends only on pcode and the pass
//     geometry, so it is computed once (fs2d_fused_order: slow tiles first, dealt round-robin to the CTAs, tiles without a
//     relaxed cell dropped) and the kernel just reads its next tile from that list: no tile-scheduler atomics, no
//     CTA-wide verdict, balanced load.  Without a list the library classifies on the fly (k_fused_classify, unsorted).
//
// HBM bytes per cell per iteration: (4 + 8) * (96*128)/((96-2T)(128-2HJ)) / T + 4/T  (T = 8: 2.6 B vs 17 B for a sweep).
#include <cuda.h>

#include <map>
#include <mutex>
#include <utility>

#include "fs2d_common.cuh"

namespace fs2d {

constexpr int FSJ = 128;         // tile columns: a warp (32 lanes x 4 columns) spans the tile width
constexpr int F_TMAX = 12;
constexpr int FCW = FSJ + 16;    // columns of the staged pcode box (its start is rounded down to 16 bytes)
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
// The iteration loop of the autonomous warps addresses shared memory by 32-bit shared-window addresses computed ONCE
// (ld/st.shared): through generic pointers the compiler re-derived the window base (S2R SR_CgaCtaId + LEA) and the slot /
// row offsets in every iteration -- a quarter of the loop's non-arithmetic instructions.
#ifdef FS2D_EMU
typedef uintptr_t saddr_t;   // CPU emulation (tests/cuda_emu): a plain pointer
#else
typedef uint32_t saddr_t;
#endif
__device__ __forceinline__ saddr_t smem_addr(const void *p) { return (saddr_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds4_s(saddr_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts4_s(saddr_t a, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
// progress counters of the warps in shared memory: publish with release, poll with acquire (CTA scope)
__device__ __forceinline__ void st_release_s(saddr_t a, int v) {
    asm volatile("st.release.cta.shared::cta.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
// the smaller of two counters (both neighbours of a warp), one look.  `after`: a value the look must not be scheduled ahead
// of (the compiler otherwise hoists the loads to right behind the previous shared-memory access, where they are useless)
__device__ __forceinline__ int flag_peek2(saddr_t a, saddr_t b, float after = 0.0f) {
    int x, y;
    asm volatile("ld.acquire.cta.shared::cta.b32 %0, [%1];" : "=r"(x) : "r"(a), "f"(after) : "memory");
    asm volatile("ld.acquire.cta.shared::cta.b32 %0, [%1];" : "=r"(y) : "r"(b), "f"(after) : "memory");
    return min(x, y);
}
// keeps a loop-invariant value in its register: without it the compiler, short of registers, rebuilds such values from
// threadIdx / SR_CgaCtaId in every iteration
template <class T>
__device__ __forceinline__ T keep_in_register(T v) {
    asm volatile("" : "+r"(v));
    return v;
}
// wait until both counters have reached v; `seen` = an earlier flag_peek2 of the same counters (they only grow)
__device__ __forceinline__ void flag_wait2(saddr_t a, saddr_t b, int v, int seen) {
    while (seen < v) seen = flag_peek2(a, b);
}
// generic-proxy accesses of a staging buffer are ordered before the async-proxy (TMA) writes that refill it
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct FusedGeom {
    int T;        // iterations in this pass == row halo
    int HJ;       // column halo: T rounded up to a multiple of 4 (TMA alignment)
    int TI, TJ;   // output tile rows / cols = 96 - 2T, 128 - 2HJ
    int tiles_i, tiles_j;
    int skip_from, skip_n;   // tile rows [skip_from, skip_from + skip_n) of the tiling are left to another launch
};
// tile row of the tiling for the launch's `row`-th tile row
__host__ __device__ __forceinline__ int f_tile_row(const FusedGeom &g, int row) { return row < g.skip_from ? row : row + g.skip_n; }

// entries of a tile list: tile column | tile row of the launch << 14 | class << 28 (no division in the kernel)
constexpr int FC_PURE = 0;   // every loaded cell is an open-fluid cell inside the grid, away from BC cells and global edges
constexpr int FC_SLOW = 1;   // anything else that has work to do
constexpr int FC_SKIP = 2;   // no cell of the output region is relaxed or takes a BC value: nothing to store
constexpr int FC_SHIFT = 28;
constexpr int FC_ROW_SHIFT = 14;
constexpr int FC_COL_MASK = (1 << FC_ROW_SHIFT) - 1;   // also the largest tile row / column count of a launch
__host__ __device__ __forceinline__ int fc_row(int e) { return (e >> FC_ROW_SHIFT) & FC_COL_MASK; }
__host__ __device__ __forceinline__ int fc_col(int e) { return e & FC_COL_MASK; }

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

// The same iteration in two phases, for the autonomous warps of open-fluid tiles: phase A needs no other warp (the rows
// 1..HK-2 of the block, in place, and the old values of the rows 1 and HK-2, which the rows 0 and HK-1 still need);
// phase B finishes the rows 0 and HK-1 from the neighbouring warps' edge rows.  A warp publishes its own edge rows, runs
// phase A -- three quarters of the arithmetic -- and only then looks whether its neighbours have published theirs, so it
// hardly ever waits and the latency of the look-up is off the critical path.  Same expression and order per cell.
template <int HK>
__device__ __forceinline__ void jacobi_rows_shuffles(const float (&p)[HK][4], float (&lf)[HK], float (&rt)[HK]) {
    constexpr uint32_t FULL = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < HK; ++k) {   // j-neighbours of the block's first / last column, all rows: own data only
        lf[k] = __shfl_up_sync(FULL, p[k][3], 1);
        rt[k] = __shfl_down_sync(FULL, p[k][0], 1);
    }
}
template <int HK, class Hook>
__device__ __forceinline__ void jacobi_rows_inner(float (&p)[HK][4], const float (&t2)[HK][4], const float (&t3)[HK][4],
                                                  const float (&lf)[HK], const float (&rt)[HK], float (&a1)[4], float (&a6)[4],
                                                  Hook hook) {
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        a1[h] = p[1][h];
        a6[h] = p[HK - 2][h];
    }
    float S[4], A[4];
    S[0] = p[2][0] + p[0][0] + p[1][1] + lf[1];
    S[1] = p[2][1] + p[0][1] + p[1][2] + p[1][0];
    S[2] = p[2][2] + p[0][2] + p[1][3] + p[1][1];
    S[3] = p[2][3] + p[0][3] + rt[1] + p[1][2];
#pragma unroll
    for (int k = 2; k <= HK - 1; ++k) {
        if (k == HK - 1) hook(p[k - 3][0]);   // (a load whose result is wanted after the last row; issued once row k-3 is done)
        if (k < HK - 1) {   // last use of the old row k-1
#pragma unroll
            for (int h = 0; h < 4; ++h) A[h] = p[k + 1][h] + p[k - 1][h];
        }
#pragma unroll
        for (int h = 0; h < 4; ++h) p[k - 1][h] = 0.25f * S[h] + t2[k - 1][h] - t3[k - 1][h];   // finish row k-1
        if (k < HK - 1) {
            S[0] = A[0] + p[k][1] + lf[k];
            S[1] = A[1] + p[k][2] + p[k][0];
            S[2] = A[2] + p[k][3] + p[k][1];
            S[3] = A[3] + rt[k] + p[k][2];
        }
    }
}
template <int HK>
__device__ __forceinline__ void jacobi_rows_outer(float (&p)[HK][4], const float (&t2)[HK][4], const float (&t3)[HK][4],
                                                  const float (&lf)[HK], const float (&rt)[HK], const float (&a1)[4],
                                                  const float (&a6)[4], const float4 up, const float4 dn) {
    float S[4], Z[4];
    S[0] = a1[0] + up.x + p[0][1] + lf[0];
    S[1] = a1[1] + up.y + p[0][2] + p[0][0];
    S[2] = a1[2] + up.z + p[0][3] + p[0][1];
    S[3] = a1[3] + up.w + rt[0] + p[0][2];
    Z[0] = dn.x + a6[0] + p[HK - 1][1] + lf[HK - 1];
    Z[1] = dn.y + a6[1] + p[HK - 1][2] + p[HK - 1][0];
    Z[2] = dn.z + a6[2] + p[HK - 1][3] + p[HK - 1][1];
    Z[3] = dn.w + a6[3] + rt[HK - 1] + p[HK - 1][2];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        p[0][h] = 0.25f * S[h] + t2[0][h] - t3[0][h];
        p[HK - 1][h] = 0.25f * Z[h] + t2[HK - 1][h] - t3[HK - 1][h];
    }
}

// ---------------------------------------------------------------------------------------------
// tile geometry and shared-memory layout (float index unless noted).  All planes are addressed as offsets from ONE
// base pointer so that the compiler keeps them in the shared address space (LDS/STS, not generic LD/ST).
// ---------------------------------------------------------------------------------------------
constexpr int HK = 8;                         // rows per thread
constexpr int VSI = 96;                       // tile rows
constexpr int VN = VSI * FSJ;                 // cells per plane
constexpr int V_WARPS = VSI / HK;             // 12
constexpr int V_THREADS = 32 * V_WARPS;       // 384
constexpr int VOFF_P0 = 0;                    // TMA destination of p (warp w: rows 8w..8w+7); slow tiles: working plane A
constexpr int VOFF_SRC = VN;                  // TMA destination of (t2, t3), 2*VN floats; slow tiles: working plane B (VN floats) ...
constexpr int VOFF_LIST = 2 * VN;             // ... the slow-cell list (VN uint16) and behind it the resolved neighbour table
constexpr int VOFF_EX = 3 * VN;               // edge-row exchange: [2][V_WARPS][2 rows][FSJ] floats
constexpr int VEX_PLANE = V_WARPS * 2 * FSJ;
constexpr int VOFF_BYTES = VOFF_EX + 2 * VEX_PLANE;   // staged pcode, VSI x FCW bytes (warp w: rows 8w..8w+7)
constexpr size_t V_SMEM = (size_t)VOFF_BYTES * 4 + (size_t)VSI * FCW;
static_assert((VOFF_BYTES * 4) % 128 == 0 && (VOFF_SRC * 4) % 128 == 0, "TMA destinations must be 128-byte aligned");
static_assert((HK * FSJ * 4) % 128 == 0 && (HK * FCW) % 128 == 0, "the warps' staging slices must be 128-byte aligned");
static_assert(FSJ == 128, "a warp (32 lanes x 4 columns) must span the tile width");

// post-BC pressure of tile cell (r, c) read from plane `pl` (same rule as p_post in fs2d_pressure.cu), split into its
// index part and its value part: slow tiles resolve, ONCE per tile, which plane cells the post-BC value of each neighbour
// of a slow cell reads (two 14-bit plane offsets + a 2-bit mode in one word); the fix-up of every iteration is then four
// table look-ups.  rlo..rhi / clo..chi: tile coordinates of the clamp bounds of sample(); cpitch: row pitch of `code`.
__device__ __forceinline__ uint32_t f_resolve(const uint8_t *code, int cpitch, int r, int c, int rlo, int rhi, int clo, int chi) {
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
    return mode == 0u ? va : (mode == 1u ? (va + vb) / 2.0f : 0.0f);
}
constexpr int FS_CAP = VN / 8;   // slow cells per tile the resolved table has room for (behind the slow-cell list): 1536
static_assert(VN <= (1 << 14), "plane offsets must fit 14 bits");

// EMIT (the "tail" pass of fs2d_jacobi_update): besides its output the pass stores, into the wall-BC cells of its INPUT
// array `emit` (= p_in; those cells are never read, their values are recomputed from pcode), the BC values of its
// PENULTIMATE state.  The reference leaves exactly these values in the wall cells of the buffer its last sweep writes
// (fs/pressure_updater.py:56-60: BC'd in place one iteration earlier), so a pass of T iterations ending at
// iteration n - 1, followed by ONE literal iteration, reproduces both physical buffers -- instead of ending every update
// with two literal iterations.  A wall-BC cell is never part of a PURE tile, so only slow tiles emit.
template <bool EMIT>
__device__ __forceinline__ void jacobi_fused_body(const CUtensorMap *mp, const CUtensorMap *ms, const CUtensorMap *mc,
                                                  float *__restrict__ p_out, const int *__restrict__ order, int n_order,
                                                  const int *__restrict__ n_order_dev, const fs2d_dom &d, const FusedGeom &g,
                                                  float *emit) {
    extern __shared__ __align__(1024) float sm[];
    uint8_t *stg_code = reinterpret_cast<uint8_t *>(sm + VOFF_BYTES);
    uint16_t *slow_list = reinterpret_cast<uint16_t *>(sm + VOFF_LIST);
    __shared__ __align__(8) uint64_t full[V_WARPS];   // TMA arrival of each warp's staging slice
    __shared__ int prog[V_WARPS];                     // edge-row sets each warp has published (monotone over the launch)
    __shared__ int n_slow;

    const int lane = threadIdx.x, w = threadIdx.y;
    const int tid = w * 32 + lane;
    const int c = 4 * lane;                 // first tile column of this thread
    const int swz = (lane >> 2) & 1;        // order in which the two (t2, t3) chunks of a row are read
    const int lr0 = w * HK;                 // first tile row of this thread
    const int o0 = lr0 * FSJ + c;           // plane offset of the thread's first cell
    const bool issuer = lane == 0;          // issues this warp's TMA loads
    constexpr uint32_t FULL = 0xffffffffu;
    constexpr uint32_t TX_PURE = HK * FSJ * (4 + 8), TX_SLOW = TX_PURE + HK * FCW;
    // one warp's slice (8 rows) of the boxes of tile entry `e`
#define FS2D_ISSUE(e)                                                                                  \
    do {                                                                                               \
        const bool slow_ = ((e) >> FC_SHIFT) != FC_PURE;                                               \
        const int R0_ = d.r0 + f_tile_row(g, fc_row(e)) * g.TI - g.T + lr0;                            \
        const int C0_ = fc_col(e) * g.TJ - g.HJ;                                                       \
        mbar_expect_tx(&full[w], slow_ ? TX_SLOW : TX_PURE);                                           \
        tma_load_2d(sm + VOFF_P0 + lr0 * FSJ, mp, C0_, R0_, &full[w]);                                 \
        tma_load_2d(sm + VOFF_SRC + 2 * lr0 * FSJ, ms, 2 * C0_, R0_, &full[w]);                        \
        if (slow_) tma_load_2d(stg_code + lr0 * FCW, mc, C0_ & ~15, R0_, &full[w]);                    \
    } while (0)

    if (tid < V_WARPS) {
        mbar_init(&full[tid], 1);
        prog[tid] = 0;
    }
    __syncthreads();
    if (n_order_dev) n_order = __ldg(n_order_dev);   // list built on the launch stream (k_fused_compact): its length lives on the device
    // the CTA walks the entries blockIdx.x + k * gridDim.x; an entry is read two tiles before it is processed (one tile
    // before its loads are issued), so the latency of the list is never exposed
    const int step = (int)gridDim.x;
    int i = blockIdx.x;
    int entry = i < n_order ? __ldg(order + i) : -1;
    int entry_next = i + step < n_order ? __ldg(order + i + step) : -1;
    if (issuer && entry >= 0) FS2D_ISSUE(entry);
    uint32_t parity = 0;
    int lv = 0;   // edge-row sets published so far by every warp of this CTA (all warps count alike)
    // rows adjacent to the thread's block inside a full plane (clamped inside the tile: rim rows compute harmless
    // garbage -- a rim value is consumed by its inner neighbour only while it still holds the loaded state, and that
    // neighbour is outside the valid region from then on anyway) and inside the exchange buffer
    const int o_up = max(lr0 - 1, 0) * FSJ + c, o_dn = min(lr0 + HK, VSI - 1) * FSJ + c;
    const int x_own = VOFF_EX + (w * 2) * FSJ + c;                                  // this warp's top row; + FSJ: bottom row
    const int x_up = w > 0 ? VOFF_EX + ((w - 1) * 2 + 1) * FSJ + c : x_own;         // bottom row of the warp above
    const int x_dn = w < V_WARPS - 1 ? VOFF_EX + ((w + 1) * 2) * FSJ + c : x_own + FSJ;
    // the same as shared-window addresses of exchange plane 0 (plane 1: + EX_PLANE_BYTES), and the neighbours' counters (a rim
    // warp has one neighbour: it looks at that one twice)
    const saddr_t xa_own = keep_in_register(smem_addr(sm + x_own));   // the others are fixed distances away (x_up, x_dn above)
    const saddr_t fa_own = keep_in_register(smem_addr(&prog[w]));
    const bool w_first = w == 0, w_last = w == V_WARPS - 1;
    constexpr saddr_t EX_PLANE_BYTES = VEX_PLANE * 4, ROW_BYTES = FSJ * 4;

    while (entry >= 0) {
        const bool tile_slow = (entry >> FC_SHIFT) != FC_PURE;
        const int entry_next2 = i + 2 * step < n_order ? __ldg(order + i + 2 * step) : -1;
        const int R0 = d.r0 + f_tile_row(g, fc_row(entry)) * g.TI - g.T;   // local-array row of tile row 0
        const int C0 = fc_col(entry) * g.TJ - g.HJ;                          // column of tile column 0
        float p[HK][4], t2[HK][4], t3[HK][4];

        if (!tile_slow) {
            // ---- open-fluid tile: this warp on its own ---------------------------------------------------------------
            mbar_wait(&full[w], parity);
#pragma unroll
            for (int k = 0; k < HK; ++k) {
                const int o = o0 + k * FSJ;
                const float4 pv = lds4(sm + VOFF_P0 + o);
                // (t2, t3) of the thread's 4 columns = two 16-byte chunks at a 32-byte lane stride: read in the order
                // (even, odd) by lanes 0-3 of every 8 and (odd, even) by lanes 4-7, so each quarter-warp wavefront touches
                // all 8 bank groups once (a plain read is 2-way bank conflicted), then put them back in order
                const float4 sa = lds4(sm + VOFF_SRC + 2 * o + 4 * swz), sb = lds4(sm + VOFF_SRC + 2 * o + 4 * (1 - swz));
                const float4 s01 = swz ? sb : sa, s23 = swz ? sa : sb;
                p[k][0] = pv.x; p[k][1] = pv.y; p[k][2] = pv.z; p[k][3] = pv.w;
                t2[k][0] = s01.x; t3[k][0] = s01.y; t2[k][1] = s01.z; t3[k][1] = s01.w;
                t2[k][2] = s23.x; t3[k][2] = s23.y; t2[k][3] = s23.z; t3[k][3] = s23.w;
            }
            __syncwarp();   // the slice is in registers: refill it with the next tile while this one is iterated
            if (issuer) {
                fence_proxy_async();
                if (entry_next >= 0) FS2D_ISSUE(entry_next);
            }
            // edge rows of state s are published as set number lv + s into exchange plane (lv + s) & 1; a warp computes
            // state s + 1 once both neighbours have published set lv + s.  A warp cannot overwrite a set its neighbour
            // still reads: to produce set n + 2 it needs the neighbour's set n + 1, made after the neighbour read set n.
            for (int s = 0; s < g.T; ++s) {
                const int set = lv + s;
                const saddr_t xo = xa_own + ((set & 1) ? EX_PLANE_BYTES : 0);
                sts4_s(xo, p[0][0], p[0][1], p[0][2], p[0][3]);
                sts4_s(xo + ROW_BYTES, p[HK - 1][0], p[HK - 1][1], p[HK - 1][2], p[HK - 1][3]);
                float lf[HK], rt[HK], a1[4], a6[4];
                jacobi_rows_shuffles<HK>(p, lf, rt);   // (the edge-row stores drain meanwhile: the release below does not wait)
                __syncwarp();
                if (issuer) st_release_s(fa_own, set + 1);
                const saddr_t fa_up = w_first ? fa_own + 4 : fa_own - 4, fa_dn = w_last ? fa_own - 4 : fa_own + 4;
                int seen = 0;
                // rows 1..6 need no other warp; three rows before they are done, take a first look at the neighbours' counters
                jacobi_rows_inner<HK>(p, t2, t3, lf, rt, a1, a6, [&](float after) { seen = flag_peek2(fa_up, fa_dn, after); });
                flag_wait2(fa_up, fa_dn, set + 1, seen);
                const float4 upv = lds4_s(w_first ? xo : xo - ROW_BYTES), dnv = lds4_s(w_last ? xo + ROW_BYTES : xo + 2 * ROW_BYTES);
                jacobi_rows_outer<HK>(p, t2, t3, lf, rt, a1, a6, upv, dnv);
            }
            lv += g.T;
            // ---- store the inner (TI x TJ) cells that belong to rows [r0, r1) -----------------------------------------
            if (c >= g.HJ && c < g.HJ + g.TJ) {
#pragma unroll
                for (int k = 0; k < HK; ++k) {
                    const int lr = lr0 + k, gr = R0 + lr;
                    if (lr >= g.T && lr < g.T + g.TI && gr < d.r1)
                        *reinterpret_cast<float4 *>(p_out + (size_t)gr * d.Y + (C0 + c)) = make_float4(p[k][0], p[k][1], p[k][2], p[k][3]);
                }
            }
        } else {
            // ---- tile with BC cells / global edges / cells outside the grid: the whole CTA together ---------------------
            const int coff = C0 - (C0 & ~15);   // multiple of 4: C0 is a multiple of 4
            // clamp bounds of sample() in tile coordinates (global edges only)
            const int rlo = max(0, d.clo - R0), rhi = min(VSI - 1, d.chi - R0);
            const int clo = max(0, -C0), chi = min(FSJ - 1, d.Y - 1 - C0);
            uint32_t col_in = 0, col_edge = 0;   // per-thread column flags: inside the grid / on a global edge column
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                col_in |= (uint32_t)(c + h >= clo && c + h <= chi) << h;
                col_edge |= (uint32_t)(C0 + c + h == 0 || C0 + c + h == d.Y - 1) << h;
            }
            if (tid == 0) n_slow = 0;   // ordered before the list is built by the barrier below
            mbar_wait(&full[w], parity);
            uint32_t upd = 0, slow = 0;   // bit 4k + h: row k, column c + h
#pragma unroll
            for (int k = 0; k < HK; ++k) {
                const int lr = lr0 + k, o = o0 + k * FSJ;
                const float4 pv = lds4(sm + VOFF_P0 + o);
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
                        const bool u = row_in && ((col_in >> h) & 1u) && relaxed;   // rim cells may compute garbage, see above
                        const bool sl = u && ((pc >> 4) != 0u || row_edge || ((col_edge >> h) & 1u));
                        upd |= (uint32_t)u << (4 * k + h);
                        slow |= (uint32_t)sl << (4 * k + h);
                    }
                }
            }
            const bool all_upd = __all_sync(FULL, upd == FULL) != 0;   // warp-uniform: open fluid in all of this warp's rows
            __syncthreads();   // every warp's slice has landed and is in registers: the staging buffers become working planes
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
            const uint8_t *code = stg_code + coff;   // tile pcode, row pitch FCW (its staging buffer stays intact)
            // resolved neighbour table: 4 words per slow cell, behind the list (VN uint16 = VN / 2 floats); each thread reads
            // back only the entries it wrote (same e -> thread mapping), so no barrier is needed
            uint32_t *res = reinterpret_cast<uint32_t *>(sm + VOFF_LIST + VN / 2);
            const bool fast = ns <= FS_CAP;   // block-uniform
            if (fast) {
                for (int e = tid; e < ns; e += V_THREADS) {
                    const int o = slow_list[e], r = o / FSJ, cc = o % FSJ;
                    res[4 * e + 0] = f_resolve(code, FCW, min(r + 1, rhi), cc, rlo, rhi, clo, chi);
                    res[4 * e + 1] = f_resolve(code, FCW, max(r - 1, rlo), cc, rlo, rhi, clo, chi);
                    res[4 * e + 2] = f_resolve(code, FCW, r, min(cc + 1, chi), rlo, rhi, clo, chi);
                    res[4 * e + 3] = f_resolve(code, FCW, r, max(cc - 1, clo), rlo, rhi, clo, chi);
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
                            emit[(size_t)gr * d.Y + gc] = f_resolved_value(sm + cur, f_resolve(code, FCW, r, cc, rlo, rhi, clo, chi));
                    }
                }
                // all threads share the slow cells and leave, in plane `nxt`, the SUM of the four post-BC neighbour
                // values (the reference's order) for the owning thread to pick up
                for (int e = tid; e < ns; e += V_THREADS) {
                    const int o = slow_list[e];
                    uint4 q;
                    if (fast) {
                        q = *reinterpret_cast<const uint4 *>(res + 4 * e);
                    } else {
                        const int r = o / FSJ, cc = o % FSJ;
                        q.x = f_resolve(code, FCW, min(r + 1, rhi), cc, rlo, rhi, clo, chi);
                        q.y = f_resolve(code, FCW, max(r - 1, rlo), cc, rlo, rhi, clo, chi);
                        q.z = f_resolve(code, FCW, r, min(cc + 1, chi), rlo, rhi, clo, chi);
                        q.w = f_resolve(code, FCW, r, max(cc - 1, clo), rlo, rhi, clo, chi);
                    }
                    float sum = f_resolved_value(sm + cur, q.x);
                    sum = sum + f_resolved_value(sm + cur, q.y);
                    sum = sum + f_resolved_value(sm + cur, q.z);
                    sum = sum + f_resolved_value(sm + cur, q.w);
                    sm[nxt + o] = sum;
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
                if (s + 1 < g.T) {   // mirror the whole block so that the fix-up can read any cell
#pragma unroll
                    for (int k = 0; k < HK; ++k) sts4(sm + nxt + o0 + k * FSJ, p[k][0], p[k][1], p[k][2], p[k][3]);
                }
                __syncthreads();   // plane `nxt` complete; after the last iteration: all plane / list / pcode reads are done
                const int x = cur; cur = nxt; nxt = x;
            }
            if (issuer) {   // the staging buffers are free again: this warp's slice of the next tile
                fence_proxy_async();
                if (entry_next >= 0) FS2D_ISSUE(entry_next);
            }
            lv += g.T;   // keeps the set numbering of the exchange buffer in step across the warps (nothing was published)
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
        }
        parity ^= 1;
        i += step;
        entry = entry_next;
        entry_next = entry_next2;
    }
#undef FS2D_ISSUE
}

__global__ void __launch_bounds__(V_THREADS, 1)
    k_jacobi_fused(const __grid_constant__ CUtensorMap map_p, const __grid_constant__ CUtensorMap map_src,
                   const __grid_constant__ CUtensorMap map_code, float *__restrict__ p_out, const int *__restrict__ order,
                   int n_order, const int *__restrict__ n_order_dev, fs2d_dom d, FusedGeom g) {
    // the descriptors must be addressed in the kernel-parameter space (the TMA unit cannot read a copy that the compiler
    // spilled to local memory): take their addresses here
    jacobi_fused_body<false>(&map_p, &map_src, &map_code, p_out, order, n_order, n_order_dev, d, g, nullptr);
}
__global__ void __launch_bounds__(V_THREADS, 1)
    k_jacobi_fused_emit(const __grid_constant__ CUtensorMap map_p, const __grid_constant__ CUtensorMap map_src,
                        const __grid_constant__ CUtensorMap map_code, float *__restrict__ p_out, const int *__restrict__ order,
                        int n_order, const int *__restrict__ n_order_dev, fs2d_dom d, FusedGeom g, float *emit) {
    jacobi_fused_body<true>(&map_p, &map_src, &map_code, p_out, order, n_order, n_order_dev, d, g, emit);
}

// Class of every tile of a launch: one CTA of 128 threads per tile, thread = tile column.  out[t] = t | class << 28.
__global__ void __launch_bounds__(FSJ)
    k_fused_classify(const uint8_t *__restrict__ pcode, fs2d_dom d, FusedGeom g, int *__restrict__ out) {
    const int t = blockIdx.x;
    const int R0 = d.r0 + f_tile_row(g, t / g.tiles_j) * g.TI - g.T;
    const int C0 = (t % g.tiles_j) * g.TJ - g.HJ;
    const int tc = threadIdx.x, col = C0 + tc;
    const bool col_in = col >= 0 && col < d.Y;
    const bool col_edge = col == 0 || col == d.Y - 1;
    const bool col_out = tc >= g.HJ && tc < g.HJ + g.TJ && col < d.Y;
    bool pure = col_in && !col_edge, skip = true;
    for (int lr = 0; lr < VSI; ++lr) {
        const int r = R0 + lr;
        const bool row_in = r >= d.clo && r <= d.chi;
        if (!row_in || r == d.clo || r == d.chi) pure = false;
        if (!row_in || !col_in) continue;
        const uint8_t pc = __ldg(pcode + (size_t)r * d.Y + col);
        if (pc != 0) pure = false;
        if (col_out && lr >= g.T && lr < g.T + g.TI && r < d.r1 && (pc & 15) != FS2D_PC_W_NONE) skip = false;
    }
    const int all_pure = __syncthreads_and(pure ? 1 : 0), all_skip = __syncthreads_and(skip ? 1 : 0);
    if (tc == 0) {
        const int cls = all_skip ? FC_SKIP : (all_pure ? FC_PURE : FC_SLOW);
        out[t] = (t % g.tiles_j) | ((t / g.tiles_j) << FC_ROW_SHIFT) | (cls << FC_SHIFT);
    }
}

// The tile list of a launch from the classes: the slow tiles first (dealt round-robin to the persistent CTAs, so each gets
// its share of the expensive ones and the cheap tiles fill the tail), then the open-fluid tiles in natural order (CTAs that
// run side by side then work on neighbouring tiles, whose halos overlap in L2); tiles with nothing to store are dropped.
// One CTA; counts = {entries, slow entries, dropped}.
constexpr int CP_THREADS = 1024;
__global__ void __launch_bounds__(CP_THREADS) k_fused_compact(const int *__restrict__ cls, int n, int *__restrict__ out, int *counts) {
    __shared__ int s_slow[CP_THREADS], s_pure[CP_THREADS];
    const int tid = threadIdx.x;
    const int chunk = (n + CP_THREADS - 1) / CP_THREADS;
    const int lo = min(tid * chunk, n), hi = min(lo + chunk, n);
    int n_slow = 0, n_pure = 0;
    for (int k = lo; k < hi; ++k) {
        const int c = cls[k] >> FC_SHIFT;
        n_slow += c == FC_SLOW;
        n_pure += c == FC_PURE;
    }
    s_slow[tid] = n_slow;
    s_pure[tid] = n_pure;
    __syncthreads();
    for (int off = 1; off < CP_THREADS; off <<= 1) {   // inclusive scan
        const int a = tid >= off ? s_slow[tid - off] : 0, b = tid >= off ? s_pure[tid - off] : 0;
        __syncthreads();
        s_slow[tid] += a;
        s_pure[tid] += b;
        __syncthreads();
    }
    const int tot_slow = s_slow[CP_THREADS - 1], tot_pure = s_pure[CP_THREADS - 1];
    int o_slow = s_slow[tid] - n_slow, o_pure = tot_slow + s_pure[tid] - n_pure;
    for (int k = lo; k < hi; ++k) {
        const int e = cls[k], c = e >> FC_SHIFT;
        if (c == FC_SLOW) out[o_slow++] = e;
        else if (c == FC_PURE) out[o_pure++] = e;
    }
    if (tid == 0) {
        counts[0] = tot_slow + tot_pure;
        counts[1] = tot_slow;
        counts[2] = n - tot_slow - tot_pure;
    }
}

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

bool fused_supported(const float *pa, const float *pb, const float *src, const uint8_t *pcode, const fs2d_dom &d) {
    return d.Y % 16 == 0 && ((uintptr_t)pa % 16 == 0) && ((uintptr_t)pb % 16 == 0) && ((uintptr_t)src % 16 == 0) &&
           ((uintptr_t)pcode % 16 == 0);
}

// per-device facts and per-(device, stream) scratch of the on-the-fly tile classification.  Two passes in flight on
// different streams (or devices) never share a buffer; one host thread per stream (include/fs2d.h).
struct DeviceInfo {
    int n_sm = 0;
    bool attr_set = false;
};
struct Scratch {
    int *buf = nullptr;
    int cap = 0;
};
static std::mutex g_fused_mutex;
static std::map<int, DeviceInfo> g_devices;
static std::map<std::pair<int, cudaStream_t>, Scratch> g_scratch;

static int device_info(int *dev_out, int *n_sm) {
    int dev = 0;
    FS2D_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_fused_mutex);
    DeviceInfo &di = g_devices[dev];
    if (!di.n_sm) FS2D_CUDA_CHECK(cudaDeviceGetAttribute(&di.n_sm, cudaDevAttrMultiProcessorCount, dev));
    if (!di.attr_set) {
        FS2D_CUDA_CHECK(cudaFuncSetAttribute(k_jacobi_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V_SMEM));
        FS2D_CUDA_CHECK(cudaFuncSetAttribute(k_jacobi_fused_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V_SMEM));
        di.attr_set = true;
    }
    *dev_out = dev;
    *n_sm = di.n_sm;
    return FS2D_OK;
}

static int make_geom(FusedGeom *g, const fs2d_dom &d, int T, int skip_from, int skip_n) {
    g->T = T;
    g->HJ = (T + 3) & ~3;
    g->TI = VSI - 2 * T;
    g->TJ = FSJ - 2 * g->HJ;
    const int all_rows = (d.r1 - d.r0 + g->TI - 1) / g->TI;
    if (skip_n < 0 || skip_from < 0 || skip_from + skip_n > all_rows) {
        set_error("bad argument: skipped tile rows [%d, %d) outside the %d tile rows of the pass", skip_from, skip_from + skip_n,
                  all_rows);
        return FS2D_E_BADARG;
    }
    g->skip_from = skip_from;
    g->skip_n = skip_n;
    g->tiles_i = all_rows - skip_n;
    g->tiles_j = (d.Y + g->TJ - 1) / g->TJ;
    if (g->tiles_i > FC_COL_MASK || g->tiles_j > FC_COL_MASK) {
        set_error("bad argument: more than 2^14 tile rows or columns in one fused pass");
        return FS2D_E_BADARG;
    }
    return FS2D_OK;
}

// classes -> tile list on stream s: classes into scratch, list into `out` (n_tiles ints), counts into counts_dev (3 ints)
static int build_order(const uint8_t *pcode, const fs2d_dom &d, const FusedGeom &g, int *cls_tmp, int *out, int *counts_dev,
                       cudaStream_t s) {
    const int n_tiles = g.tiles_i * g.tiles_j;
    g_launches += 2;
    k_fused_classify<<<n_tiles, FSJ, 0, s>>>(pcode, d, g, cls_tmp);
    k_fused_compact<<<1, CP_THREADS, 0, s>>>(cls_tmp, n_tiles, out, counts_dev);
    return FS2D_OK;
}

// scratch of (device, stream) with room for n_tiles classes + n_tiles entries + 4 counters
static int get_scratch(int dev, cudaStream_t s, int n_tiles, int **buf) {
    std::lock_guard<std::mutex> lock(g_fused_mutex);
    Scratch &sc = g_scratch[std::make_pair(dev, s)];
    const int need = 2 * n_tiles + 4;
    if (sc.cap < need) {
        if (sc.buf) FS2D_CUDA_CHECK(cudaFree(sc.buf));   // (cudaFree waits for the passes still using it)
        sc.buf = nullptr;
        sc.cap = 0;
        FS2D_CUDA_CHECK(cudaMalloc(&sc.buf, sizeof(int) * (size_t)need));
        sc.cap = need;
    }
    *buf = sc.buf;
    return FS2D_OK;
}

int fused_pass(const float *p_in, float *p_out, const float *src, const uint8_t *pcode, const fs2d_dom &d, int T,
               cudaStream_t s, int skip_from, int skip_n, bool emit, const int *order, int n_order) {
    int dev = 0, n_sm = 0;
    if (int e = device_info(&dev, &n_sm)) return e;
    FusedGeom g;
    if (int e = make_geom(&g, d, T, skip_from, skip_n)) return e;
    if (g.tiles_i == 0) return FS2D_OK;
    const int n_tiles = g.tiles_i * g.tiles_j;
    CUtensorMap mp, ms, mc;
    if (int e = make_map(&mp, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, p_in, d.Y, d.rows, FSJ, HK)) return e;
    if (int e = make_map(&ms, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, src, 2ull * d.Y, d.rows, 2 * FSJ, HK)) return e;
    if (int e = make_map(&mc, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, pcode, d.Y, d.rows, FCW, HK)) return e;
    const int *n_order_dev = nullptr;
    if (!order) {   // no precomputed tile list (fs2d_fused_order): build it now, on the launch stream; its length stays on the device
        int *buf = nullptr;
        if (int e = get_scratch(dev, s, n_tiles, &buf)) return e;
        if (int e = build_order(pcode, d, g, buf, buf + n_tiles, buf + 2 * n_tiles, s)) return e;
        order = buf + n_tiles;
        n_order_dev = buf + 2 * n_tiles;
        n_order = n_tiles;   // upper bound (grid size); the kernel reads the true length
    }
    if (n_order <= 0) return FS2D_OK;
    if (n_order > n_tiles) {
        set_error("bad argument: tile list of %d entries for a pass of %d tiles", n_order, n_tiles);
        return FS2D_E_BADARG;
    }
    const int grid = n_order < n_sm ? n_order : n_sm;
    ++g_launches;
    if (emit) k_jacobi_fused_emit<<<grid, dim3(32, V_WARPS, 1), V_SMEM, s>>>(mp, ms, mc, p_out, order, n_order, n_order_dev, d, g, const_cast<float *>(p_in));
    else k_jacobi_fused<<<grid, dim3(32, V_WARPS, 1), V_SMEM, s>>>(mp, ms, mc, p_out, order, n_order, n_order_dev, d, g);
    return FS2D_OK;
}

int g_tail_emit = 1;   // fs2d_set_tuning(4, v): 0 = end fs2d_jacobi_update with two literal iterations instead of {emitting pass, one}

}  // namespace fs2d

using namespace fs2d;

extern "C" {

int fs2d_fused_tile(int T, int *rows, int *cols, int *halo_rows, int *halo_cols, int *t_max) {
    if (rows) *rows = VSI;
    if (cols) *cols = FSJ;
    if (halo_rows) *halo_rows = T;
    if (halo_cols) *halo_cols = (T + 3) & ~3;
    if (t_max) *t_max = F_TMAX;
    return FS2D_OK;
}

#define FS2D_FUSED_ARGS_OK()                                                                                                   \
    FS2D_REQUIRE(p_out && p_in && src && pcode && p_out != p_in, "null/aliased field pointer");                                \
    FS2D_REQUIRE(T >= 1 && T <= F_TMAX, "fused iteration count out of range");                                                 \
    FS2D_REQUIRE(fused_supported(p_in, p_out, src, pcode, d),                                                                  \
                 "fused Jacobi needs Y % 16 == 0 and 16-byte aligned fields (TMA row pitch / base alignment)");               \
    FS2D_REQUIRE(n_order >= 0, "negative tile-list length");                                                                   \
    if (int e = check_dom(d)) return e;                                                                                        \
    if (d.r1 == d.r0) return FS2D_OK

int fs2d_fused_order(const uint8_t *pcode, fs2d_dom d, int T, int skip_from, int skip_n, int32_t *order, int cap, int *counts,
                     void *stream) {
    FS2D_REQUIRE(pcode && order && counts, "null pointer");
    FS2D_REQUIRE(T >= 1 && T <= F_TMAX, "fused iteration count out of range");
    if (int e = check_dom(d)) return e;
    counts[0] = counts[1] = counts[2] = 0;
    if (d.r1 == d.r0) return FS2D_OK;
    FusedGeom g;
    if (int e = make_geom(&g, d, T, skip_from, skip_n)) return e;
    const int n_tiles = g.tiles_i * g.tiles_j;
    if (n_tiles == 0) return FS2D_OK;
    FS2D_REQUIRE(cap >= n_tiles, "tile-list buffer too small (needs one entry per tile of the pass)");
    cudaStream_t s = (cudaStream_t)stream;
    int dev = 0, n_sm = 0, *buf = nullptr;
    if (int e = device_info(&dev, &n_sm)) return e;
    if (int e = get_scratch(dev, s, n_tiles, &buf)) return e;
    if (int e = build_order(pcode, d, g, buf, order, buf + 2 * n_tiles, s)) return e;
    FS2D_LAUNCH_CHECK();
    FS2D_CUDA_CHECK(cudaMemcpyAsync(counts, buf + 2 * n_tiles, 3 * sizeof(int), cudaMemcpyDeviceToHost, s));
    FS2D_CUDA_CHECK(cudaStreamSynchronize(s));
    return FS2D_OK;
}

int fs2d_jacobi_fused(float *p_out, const float *p_in, const float *src, const uint8_t *pcode, fs2d_dom d, int T,
                      const int32_t *order, int n_order, void *stream) {
    FS2D_FUSED_ARGS_OK();
    if (int e = fused_pass(p_in, p_out, src, pcode, d, T, (cudaStream_t)stream, 0, 0, false, order, n_order)) return e;
    FS2D_LAUNCH_CHECK();
    return FS2D_OK;
}

int fs2d_jacobi_fused_part(float *p_out, const float *p_in, const float *src, const uint8_t *pcode, fs2d_dom d, int T,
                           int skip_from, int skip_n, const int32_t *order, int n_order, void *stream) {
    FS2D_FUSED_ARGS_OK();
    if (int e = fused_pass(p_in, p_out, src, pcode, d, T, (cudaStream_t)stream, skip_from, skip_n, false, order, n_order)) return e;
    FS2D_LAUNCH_CHECK();
    return FS2D_OK;
}

int fs2d_jacobi_fused_tail(float *p_out, float *p_in, const float *src, const uint8_t *pcode, fs2d_dom d, int T, int skip_from,
                           int skip_n, const int32_t *order, int n_order, void *stream) {
    FS2D_FUSED_ARGS_OK();
    if (int e = fused_pass(p_in, p_out, src, pcode, d, T, (cudaStream_t)stream, skip_from, skip_n, true, order, n_order)) return e;
    FS2D_LAUNCH_CHECK();
    return FS2D_OK;
}

}  // extern "C"

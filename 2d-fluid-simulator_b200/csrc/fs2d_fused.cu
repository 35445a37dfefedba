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
constexpr size_t F_SMEM = (size_t)FSI * FSJ * (4 + 8 + 4 + 4 + 1) + (size_t)FSI * FCW + 128;
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
};

// post-BC pressure of tile cell (r, c) from plane `cur` (same rule as p_post in fs2d_pressure.cu);
// rlo..rhi / clo..chi: tile coordinates of the clamp bounds
__device__ __noinline__ float f_post(const float *cur, const uint8_t *code, int r, int c, int rlo, int rhi, int clo,
                                     int chi) {
    const int rm = max(r - 1, rlo), rp = min(r + 1, rhi), cm = max(c - 1, clo), cp = min(c + 1, chi);
    switch (code[r * FSJ + c] & 15) {
        case FS2D_PC_FLUID:
        case FS2D_PC_W_NONE: return cur[r * FSJ + c];
        case FS2D_PC_W_IM: return cur[rm * FSJ + c];
        case FS2D_PC_W_IP: return cur[rp * FSJ + c];
        case FS2D_PC_W_JM: return cur[r * FSJ + cm];
        case FS2D_PC_W_JP: return cur[r * FSJ + cp];
        case FS2D_PC_W_IM_JP: return (cur[rm * FSJ + c] + cur[r * FSJ + cp]) / 2.0f;
        case FS2D_PC_W_IP_JP: return (cur[rp * FSJ + c] + cur[r * FSJ + cp]) / 2.0f;
        case FS2D_PC_W_IM_JM: return (cur[rm * FSJ + c] + cur[r * FSJ + cm]) / 2.0f;
        case FS2D_PC_W_IP_JM: return (cur[rp * FSJ + c] + cur[r * FSJ + cm]) / 2.0f;
        case FS2D_PC_INFLOW: return cur[rp * FSJ + c];
        default: return 0.0f;  // FS2D_PC_OUTFLOW
    }
}
// slow path of one cell: all four neighbours through f_post with clamping
__device__ __noinline__ float f_slow_sum(const float *cur, const uint8_t *code, int r, int c, int rlo, int rhi, int clo,
                                         int chi) {
    const float pe = f_post(cur, code, min(r + 1, rhi), c, rlo, rhi, clo, chi);
    const float pw = f_post(cur, code, max(r - 1, rlo), c, rlo, rhi, clo, chi);
    const float pq = f_post(cur, code, r, min(c + 1, chi), rlo, rhi, clo, chi);
    const float ps = f_post(cur, code, r, max(c - 1, clo), rlo, rhi, clo, chi);
    return pe + pw + pq + ps;
}

__global__ void __launch_bounds__(F_THREADS, 1)
    k_jacobi_fused(const __grid_constant__ CUtensorMap map_p, const __grid_constant__ CUtensorMap map_src,
                   const __grid_constant__ CUtensorMap map_code, float *__restrict__ p_out, fs2d_dom d, FusedGeom g) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    float *stg_p = reinterpret_cast<float *>(smem);                          // FSI*FSJ floats   (TMA dst)
    float *stg_src = stg_p + FSI * FSJ;                                      // FSI*FSJ float2   (TMA dst)
    uint8_t *stg_code = reinterpret_cast<uint8_t *>(stg_src + 2 * FSI * FSJ);  // FSI*FCW bytes (TMA dst)
    float *w0 = reinterpret_cast<float *>(stg_code + FSI * FCW);
    float *w1 = w0 + FSI * FSJ;
    uint8_t *wcode = reinterpret_cast<uint8_t *>(w1 + FSI * FSJ);
    __shared__ __align__(8) uint64_t bar;

    const int c = threadIdx.x;             // tile column
    const int lr0 = threadIdx.y * FK;      // first tile row of this thread
    const bool leader = (threadIdx.x == 0 && threadIdx.y == 0);
    const int n_tiles = g.tiles_i * g.tiles_j;
    constexpr uint32_t TX_BYTES = FSI * FSJ * (4 + 8) + FSI * FCW;

    // NOTE: the descriptors must be addressed in the kernel-parameter space (the TMA unit cannot read a copy
    // that the compiler spilled to local memory), so take their addresses here, not through a lambda capture.
    const CUtensorMap *mp = &map_p, *ms = &map_src, *mc = &map_code;
#define FS2D_ISSUE(tile)                                                            \
    do {                                                                            \
        const int R0_ = d.r0 + ((tile) / g.tiles_j) * g.TI - g.T;                   \
        const int C0_ = ((tile) % g.tiles_j) * g.TJ - g.HJ;                         \
        mbar_expect_tx(&bar, TX_BYTES);                                             \
        tma_load_2d(stg_p, mp, C0_, R0_, &bar);                                     \
        tma_load_2d(stg_src, ms, 2 * C0_, R0_, &bar);                               \
        tma_load_2d(stg_code, mc, C0_ & ~15, R0_, &bar);                            \
    } while (0)

    if (leader) mbar_init(&bar, 1);
    __syncthreads();
    int t = blockIdx.x;
    if (leader && t < n_tiles) FS2D_ISSUE(t);
    uint32_t parity = 0;
    const int cl = max(c - 1, 0), cr = min(c + 1, FSJ - 1);

    for (; t < n_tiles; t += gridDim.x) {
        const int R0 = d.r0 + (t / g.tiles_j) * g.TI - g.T;   // local-array row of tile row 0
        const int C0 = (t % g.tiles_j) * g.TJ - g.HJ;         // column of tile column 0
        const int coff = C0 - (C0 & ~15);                     // where tile column 0 sits inside the staged pcode box
        // clamp bounds of sample() in tile coordinates (global edges only)
        const int rlo = max(0, d.clo - R0), rhi = min(FSI - 1, d.chi - R0);
        const int clo = max(0, -C0), chi = min(FSJ - 1, d.Y - 1 - C0);

        mbar_wait(&bar, parity);
        parity ^= 1;

        // ---- per-thread state from the staging buffer -------------------------------------------
        float p[FK], t2[FK], t3[FK];
        uint32_t upd = 0, slow = 0;
#pragma unroll
        for (int k = 0; k < FK; ++k) {
            const int lr = lr0 + k, o = lr * FSJ + c;
            p[k] = stg_p[o];
            const float2 s = reinterpret_cast<const float2 *>(stg_src)[o];
            t2[k] = s.x;
            t3[k] = s.y;
            const uint8_t pc = stg_code[lr * FCW + coff + c];
            wcode[o] = pc;
            const int code = pc & 15;
            const bool inside = lr >= rlo && lr <= rhi && c >= clo && c <= chi;
            const bool relaxed = code == FS2D_PC_FLUID || code == FS2D_PC_INFLOW || code == FS2D_PC_OUTFLOW;
            const bool e_rl = lr == rlo && R0 + lr == d.clo, e_rh = lr == rhi && R0 + lr == d.chi;
            const bool e_cl = c == clo && C0 + c == 0, e_ch = c == chi && C0 + c == d.Y - 1;
            // a cell on the tile rim whose missing neighbour is NOT a global edge cannot be updated
            const bool frozen = (lr == 0 && !e_rl) || (lr == FSI - 1 && !e_rh) || (c == 0 && !e_cl) || (c == FSJ - 1 && !e_ch);
            const bool u = inside && relaxed && !frozen;
            const bool sl = u && ((pc >> 4) != 0 || e_rl || e_rh || e_cl || e_ch);
            upd |= (uint32_t)u << k;
            slow |= (uint32_t)sl << k;
        }

        // ---- T iterations --------------------------------------------------------------------------
        const float *cur = stg_p;
        float *nxt = w0;
        for (int s = 0; s < g.T; ++s) {
            __syncthreads();  // plane `cur` (and wcode) complete; previous readers of `nxt` done
            if (s == 1 && leader && t + (int)gridDim.x < n_tiles) FS2D_ISSUE(t + (int)gridDim.x);  // staging is free now
            const float upx = cur[max(lr0 - 1, 0) * FSJ + c];
            const float dnx = cur[min(lr0 + FK, FSI - 1) * FSJ + c];
            float np[FK];
#pragma unroll
            for (int k = 0; k < FK; ++k) {
                const int lr = lr0 + k;
                const float lf = cur[lr * FSJ + cl], rt = cur[lr * FSJ + cr];
                const float upv = k > 0 ? p[k - 1] : upx;
                const float dnv = k < FK - 1 ? p[k + 1] : dnx;
                float sum = dnv + upv + rt + lf;  // (i+1) + (i-1) + (j+1) + (j-1), the reference's order
                if ((slow >> k) & 1u) sum = f_slow_sum(cur, wcode, lr, c, rlo, rhi, clo, chi);
                const float v = 0.25f * sum + t2[k] - t3[k];
                np[k] = ((upd >> k) & 1u) ? v : p[k];
            }
#pragma unroll
            for (int k = 0; k < FK; ++k) {
                p[k] = np[k];
                nxt[(lr0 + k) * FSJ + c] = np[k];
            }
            cur = nxt;
            nxt = (nxt == w0) ? w1 : w0;
        }
        if (g.T == 1) {  // staging was never released inside the loop
            __syncthreads();
            if (leader && t + (int)gridDim.x < n_tiles) FS2D_ISSUE(t + (int)gridDim.x);
        }

        // ---- store the inner (TI x TJ) cells that were updated and belong to rows [r0, r1) ----------
        if (c >= g.HJ && c < g.HJ + g.TJ && C0 + c < d.Y) {
#pragma unroll
            for (int k = 0; k < FK; ++k) {
                const int lr = lr0 + k, gr = R0 + lr;
                if (lr >= g.T && lr < g.T + g.TI && gr < d.r1 && ((upd >> k) & 1u)) p_out[(size_t)gr * d.Y + (C0 + c)] = p[k];
            }
        }
        __syncthreads();  // all reads of the working planes done before the next tile overwrites them
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

static int make_map(CUtensorMap *m, CUtensorMapDataType dt, size_t esz, const void *base, uint64_t cols, uint64_t rows,
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

int fused_pass(const float *p_in, float *p_out, const float *src, const uint8_t *pcode, const fs2d_dom &d, int T,
               cudaStream_t s) {
    static int n_sm = 0;
    static bool attr_set = false;
    if (!n_sm) {
        int dev = 0;
        FS2D_CUDA_CHECK(cudaGetDevice(&dev));
        FS2D_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    if (!attr_set) {
        FS2D_CUDA_CHECK(cudaFuncSetAttribute(k_jacobi_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F_SMEM));
        attr_set = true;
    }
    CUtensorMap mp, ms, mc;
    if (int e = make_map(&mp, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, p_in, d.Y, d.rows, FSJ, FSI)) return e;
    if (int e = make_map(&ms, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, src, 2ull * d.Y, d.rows, 2 * FSJ, FSI)) return e;
    if (int e = make_map(&mc, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, pcode, d.Y, d.rows, FCW, FSI)) return e;
    FusedGeom g;
    g.T = T;
    g.HJ = (T + 3) & ~3;
    g.TI = FSI - 2 * T;
    g.TJ = FSJ - 2 * g.HJ;
    g.tiles_i = (d.r1 - d.r0 + g.TI - 1) / g.TI;
    g.tiles_j = (d.Y + g.TJ - 1) / g.TJ;
    const int n_tiles = g.tiles_i * g.tiles_j;
    const int grid = n_tiles < n_sm ? n_tiles : n_sm;
    ++g_launches;
    k_jacobi_fused<<<grid, dim3(FSJ, FNTY, 1), F_SMEM, s>>>(mp, ms, mc, p_out, d, g);
    return FS2D_OK;
}

}  // namespace fs2d

using namespace fs2d;

extern "C" {

int fs2d_fused_tile(int T, int *rows, int *cols, int *halo_rows, int *halo_cols, int *t_max) {
    if (rows) *rows = FSI;
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
    FS2D_REQUIRE(d.Y % 16 == 0 && ((uintptr_t)p_in % 16 == 0) && ((uintptr_t)src % 16 == 0) && ((uintptr_t)pcode % 16 == 0),
                 "fused Jacobi needs Y % 16 == 0 and 16-byte aligned fields (TMA row pitch / base alignment)");
    if (int e = check_dom(d)) return e;
    if (d.r1 == d.r0) return FS2D_OK;
    if (int e = fused_pass(p_in, p_out, src, pcode, d, T, (cudaStream_t)stream)) return e;
    FS2D_LAUNCH_CHECK();
    return FS2D_OK;
}

}  // extern "C"

// fs2d_stream.cu -- TMA-fed streaming versions of the CIP-path stencil kernels (sm_100a).
//
// The one-cell-per-thread kernels in fs2d_kernels.cu are latency-bound: what a thread block has in flight is what
// its threads hold in registers, and the FP-heavy kernels (CIP advection: ~300 instructions per cell) cannot spare
// the registers.  Here the loads are decoupled from the threads: persistent CTAs walk over 16 x 64 cell tiles in
// row-major order; one elected thread keeps a ring of ST_STAGES shared-memory stages filled with
// cp.async.bulk.tensor (TMA) box loads of every input field (tile + halo) signalled through mbarriers, so two further
// tiles (~64 KB per CTA) are always in flight while the 256 threads compute the current tile from shared memory.
// The arithmetic is the accessor-templated code of fs2d_ops.cuh: bit-identical to the direct kernels.
//
// Clamp-to-edge sample(): TMA zero-fills what lies outside the array; tiles that touch the clamp window
// [clo, chi] x [0, Y-1] repair their halo in shared memory (edge row / column replicated outwards, rows first so that
// the corners follow) before computing -- the same values clamped loads would have fetched.
#include <cuda.h>

#include "fs2d_ops.cuh"

namespace fs2d {

int make_map(CUtensorMap *m, CUtensorMapDataType dt, size_t esz, const void *base, uint64_t cols, uint64_t rows,
             uint32_t box_cols, uint32_t box_rows);

constexpr int ST_TR = 16;        // tile rows
constexpr int ST_TC = 64;        // tile columns
constexpr int ST_HC = 4;         // column halo loaded on each side (cells): keeps every box start 16-byte aligned
constexpr int ST_BC = ST_TC + 2 * ST_HC;   // box columns (cells)

__device__ __forceinline__ uint32_t st_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(st_smem(bar)), "r"(count));
}
__device__ __forceinline__ void st_mbar_expect(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(st_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, P1;\n\t}"
            : "=r"(done)
            : "r"(st_smem(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void st_tma_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            st_smem(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(st_smem(bar))
        : "memory");
}

// Box rows / columns that lie outside the clamp window take the value of the nearest row / column inside it.
// bx: box of `rows` rows, `P` elements per row; br0 / bc0: global row / column (in cells) of box element (0, 0).
template <typename T>
__device__ __forceinline__ void st_repair_rows(T *bx, int P, int rows, int br0, const fs2d_dom &d, int tid, int nthr) {
    const int lo = max(0, d.clo - br0), hi = min(rows, d.chi - br0 + 1);   // valid box rows [lo, hi)
    const int n_bad = lo + (rows - hi);
    for (int e = tid; e < n_bad * P; e += nthr) {
        const int q = e / P, x = e - q * P;
        const int row = q < lo ? q : hi + (q - lo);
        bx[row * P + x] = bx[(q < lo ? lo : hi - 1) * P + x];
    }
}
template <typename T>
__device__ __forceinline__ void st_repair_cols(T *bx, int P, int C, int rows, int cols, int bc0, const fs2d_dom &d, int tid,
                                               int nthr) {
    const int lo = max(0, -bc0), hi = min(cols, d.Y - bc0);   // valid box columns [lo, hi) in cells
    const int n_bad = lo + (cols - hi);
    for (int e = tid; e < n_bad * rows * C; e += nthr) {
        const int row = e / (n_bad * C), rem = e - row * (n_bad * C);
        const int q = rem / C, ch = rem - q * C;
        const int col = q < lo ? q : hi + (q - lo);
        bx[row * P + col * C + ch] = bx[row * P + (q < lo ? lo : hi - 1) * C + ch];
    }
}

// Layout of one stage for an operator with NF float fields of chan(k) channels, row halo H and a centre-only mask box.
template <class Op>
struct StageLayout {
    static constexpr int ROWS = ST_TR + 2 * Op::H;
    __host__ __device__ static constexpr int pitch(int k) { return Op::chan(k) * ST_BC; }   // floats per box row
    __host__ __device__ static constexpr int box_bytes(int k) { return ((pitch(k) * ROWS * 4 + 127) / 128) * 128; }
    __host__ __device__ static constexpr int field_off(int k) { return k == 0 ? 0 : field_off(k - 1) + box_bytes(k - 1); }
    // mask box: the tile itself, or (Op::MH = 1) one more row and 16 more columns on each side (16-byte aligned start)
    static constexpr int MASK_OFF = field_off(Op::NF);
    static constexpr int MASK_HC = Op::MH ? 16 : 0;
    static constexpr int MASK_ROWS = ST_TR + 2 * Op::MH, MASK_COLS = ST_TC + 2 * MASK_HC;
    static constexpr int STAGE_BYTES = MASK_OFF + ((MASK_ROWS * MASK_COLS + 127) / 128) * 128;
    __host__ __device__ static constexpr uint32_t tx_bytes(int k = 0) {
        return k == Op::NF ? (uint32_t)(MASK_ROWS * MASK_COLS) : (uint32_t)(pitch(k) * ROWS * 4) + tx_bytes(k + 1);
    }
};

template <int NF>
struct StreamMaps {
    CUtensorMap f[NF];
    CUtensorMap mask;
};
struct StreamGeom {
    int tiles_j, n_tiles;
};

// thread (tx, ty) = (tid % 64, tid / 64) computes column tx of rows ty, ty + ST_THREADS / 64, ...
template <class Op, int ST_STAGES, int MIN_CTAS, int ST_THREADS>
__global__ void __launch_bounds__(ST_THREADS, MIN_CTAS)
    k_stream(const __grid_constant__ StreamMaps<Op::NF> maps, const Op op, const fs2d_dom d, const StreamGeom g) {
    using L = StageLayout<Op>;
    extern __shared__ __align__(1024) uint8_t st_sm[];
    __shared__ __align__(8) uint64_t full[ST_STAGES];
    const int tid = threadIdx.x;
    const int tx = tid % ST_TC, ty = tid / ST_TC;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < ST_STAGES; ++s) st_mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int first = blockIdx.x, step = gridDim.x;

#define ST_ISSUE(tile_, stage_)                                                                                       \
    do {                                                                                                              \
        const int R0_ = d.r0 + ((tile_) / g.tiles_j) * ST_TR, C0_ = ((tile_) % g.tiles_j) * ST_TC;                    \
        uint8_t *base_ = st_sm + (size_t)(stage_) * L::STAGE_BYTES;                                                   \
        st_mbar_expect(&full[stage_], L::tx_bytes());                                                                 \
        _Pragma("unroll") for (int k_ = 0; k_ < Op::NF; ++k_)                                                         \
            st_tma_2d(base_ + L::field_off(k_), &maps.f[k_], Op::chan(k_) * (C0_ - ST_HC), R0_ - Op::H, &full[stage_]); \
        st_tma_2d(base_ + L::MASK_OFF, &maps.mask, C0_ - L::MASK_HC, R0_ - Op::MH, &full[stage_]);                    \
    } while (0)

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < ST_STAGES - 1; ++s)
            if (first + s * step < g.n_tiles) ST_ISSUE(first + s * step, s);
    }
    int k = 0;
    for (int tile = first; tile < g.n_tiles; tile += step, ++k) {
        const int stage = k % ST_STAGES;
        if (tid == 0) {   // refill the stage the previous tile has released (all threads passed its closing barrier)
            const int ahead = tile + (ST_STAGES - 1) * step;
            const int astage = (k + ST_STAGES - 1) % ST_STAGES;
            if (ahead < g.n_tiles) ST_ISSUE(ahead, astage);
        }
        const int R0 = d.r0 + (tile / g.tiles_j) * ST_TR, C0 = (tile % g.tiles_j) * ST_TC;   // first cell of the tile
        uint8_t *base = st_sm + (size_t)stage * L::STAGE_BYTES;
        st_mbar_wait(&full[stage], (uint32_t)((k / ST_STAGES) & 1));

        // ---- clamp repair for tiles that reach outside the clamp window (block-uniform, rare) ----------------
        if (R0 - Op::H < d.clo || R0 + ST_TR + Op::H - 1 > d.chi || C0 - ST_HC < 0 || C0 + ST_TC + ST_HC > d.Y) {
#pragma unroll
            for (int f = 0; f < Op::NF; ++f)   // rows first
                st_repair_rows(reinterpret_cast<float *>(base + L::field_off(f)), L::pitch(f), L::ROWS, R0 - Op::H, d, tid,
                               ST_THREADS);
            if (Op::MH) st_repair_rows(base + L::MASK_OFF, L::MASK_COLS, L::MASK_ROWS, R0 - Op::MH, d, tid, ST_THREADS);
            __syncthreads();
#pragma unroll
            for (int f = 0; f < Op::NF; ++f)   // then columns, over all rows (corners follow)
                st_repair_cols(reinterpret_cast<float *>(base + L::field_off(f)), L::pitch(f), Op::chan(f), L::ROWS, ST_BC,
                               C0 - ST_HC, d, tid, ST_THREADS);
            if (Op::MH) st_repair_cols(base + L::MASK_OFF, L::MASK_COLS, 1, L::MASK_ROWS, L::MASK_COLS, C0 - L::MASK_HC, d, tid, ST_THREADS);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // these generic-proxy writes precede the next TMA refill
            __syncthreads();
        }

        // ---- compute: 4 cells per thread ---------------------------------------------------------------------
        const int j = C0 + tx;
        if (j < d.Y) {
            const uint8_t *mk = base + L::MASK_OFF;
#pragma unroll
            for (int u = 0; u < ST_TR / (ST_THREADS / ST_TC); ++u) {
                const int lr = ty + u * (ST_THREADS / ST_TC), r = R0 + lr;
                if (r < d.r1) {
                    const float *ctr[Op::NF];
#pragma unroll
                    for (int f = 0; f < Op::NF; ++f)
                        ctr[f] = reinterpret_cast<const float *>(base + L::field_off(f)) + (lr + Op::H) * L::pitch(f) +
                                 Op::chan(f) * (tx + ST_HC);
                    op.cell(ctr, mk + (lr + Op::MH) * L::MASK_COLS + tx + L::MASK_HC, L::MASK_COLS, r, j, (size_t)r * d.Y + j, d);
                }
            }
        }
        __syncthreads();   // the stage may be refilled
    }
#undef ST_ISSUE
}

// ---------------------------------------------------------------------------------------------
// operators
// ---------------------------------------------------------------------------------------------
// fs/solver.py:267-332  _advection_phase with the advecting velocity == the advected field (CipMacSolver)
template <bool P2>
struct OpAdvect {
    static constexpr int NF = 3, H = 1, MH = 0;
    __host__ __device__ static constexpr int chan(int) { return 2; }
    float *fn, *fxn, *fyn;
    float dt, dx;
    DivC<P2> ddx, ddx2, ddx3;
    __device__ __forceinline__ void cell(const float *const (&c)[NF], const uint8_t *mk, int, int, int, size_t idx,
                                         const fs2d_dom &) const {
        if (*mk != 0) return;
        const SAt at{ST_BC, 2 * ST_BC};
        const CipOut o = c_cip_advect<P2>(at, c[0], c[1], c[2], c[0], dt, dx, ddx, ddx2, ddx3);
        reinterpret_cast<float2 *>(fn)[idx] = o.f;
        reinterpret_cast<float2 *>(fxn)[idx] = o.fx;
        reinterpret_cast<float2 *>(fyn)[idx] = o.fy;
    }
};
// fs/solver.py:229-240  _non_advection_phase
template <bool P2>
struct OpNonadv {
    static constexpr int NF = 2, H = 1, MH = 0;
    __host__ __device__ static constexpr int chan(int k) { return k == 0 ? 2 : 1; }
    float *fn;
    float dt, re;
    DivC<P2> ddx, ddx2;
    __device__ __forceinline__ void cell(const float *const (&c)[NF], const uint8_t *mk, int, int, int, size_t idx,
                                         const fs2d_dom &) const {
        if (*mk == 1) return;
        const SAt at{ST_BC, 2 * ST_BC};
        reinterpret_cast<float2 *>(fn)[idx] = c_cip_nonadv<P2>(l_cip_nonadv(at, c[0], c[1]), dt, ddx, ddx2, re);
    }
};

// fs/vorticity_confinement.py:57-59  apply() = _calc_vorticity + _add_vorticity (see k_vort_apply in fs2d_kernels.cu):
// the curl of the cell and of its four clamped neighbours from the v tile (halo 2); non-fluid neighbours keep their
// stored |vorticity|, read from global memory (rare).
template <bool P2>
struct OpVort {
    static constexpr int NF = 1, H = 2, MH = 1;
    __host__ __device__ static constexpr int chan(int) { return 2; }
    float *vn, *w, *wabs;
    float dtw;
    DivC<P2> ddx;
    __device__ __forceinline__ float curl(const SAt &at, const float *c, int dr, int dc) const {
        const float2 gx = ddx(0.5f * (at.ld2(c, dr + 1, dc) - at.ld2(c, dr - 1, dc)));   // diff_x(v) at the neighbour
        const float2 gy = ddx(0.5f * (at.ld2(c, dr, dc + 1) - at.ld2(c, dr, dc - 1)));   // diff_y(v)
        return gx.y - gy.x;
    }
    __device__ __forceinline__ void cell(const float *const (&c)[NF], const uint8_t *mk, int mp, int r, int j, size_t idx,
                                         const fs2d_dom &d) const {
        if (*mk != 0) return;
        const SAt at{ST_BC, 2 * ST_BC};
        // clamped neighbours (sample()): offsets of (CR(r+-1), j) and (r, CJ(j+-1)) relative to the cell
        const int dr[4] = {min(r + 1, d.chi) - r, max(r - 1, d.clo) - r, 0, 0};
        const int dc[4] = {0, 0, min(j + 1, d.Y - 1) - j, max(j - 1, 0) - j};
        float a[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            a[k] = mk[dr[k] * mp + dc[k]] == 0 ? fabsf(curl(at, c[0], dr[k], dc[k])) : wabs[idx + (ptrdiff_t)dr[k] * d.Y + dc[k]];
        VortIn x;
        x.aip = a[0]; x.aim = a[1]; x.ajp = a[2]; x.ajm = a[3];
        x.o = curl(at, c[0], 0, 0);
        x.c = at.ld2(c[0], 0, 0);
        const float2 out = c_vort_add<P2>(x, ddx, dtw);
        w[idx] = x.o;
        wabs[idx] = fabsf(x.o);
        reinterpret_cast<float2 *>(vn)[idx] = out;
    }
};

int g_stream = 1;   // fs2d_set_tuning(2, v): 0 = always the direct kernels of fs2d_kernels.cu
// TMA needs 16-byte aligned bases and row pitches (Y % 16 covers the 1-byte mask rows)
bool stream_ok(const fs2d_dom &d, const void *const *ptrs, int n) {
    if (!g_stream || d.Y % 16 != 0 || (long long)d.rows * d.Y >= (1ll << 31)) return false;
    for (int i = 0; i < n; ++i)
        if ((uintptr_t)ptrs[i] % 16 != 0) return false;
    return true;
}

int g_stream_cfg = 1;   // fs2d_set_tuning(3, v): {stages x CTAs/SM x threads}: 0 = 3x2x256, 1 = 2x3x256, 2 = 3x2x512, 3 = 2x2x512
template <class Op, int ST_STAGES, int MIN_CTAS, int ST_THREADS>
static int launch_stream_cfg(const Op &op, const float *const *fields, const uint8_t *mask, const fs2d_dom &d, cudaStream_t s) {
    using L = StageLayout<Op>;
    constexpr int SMEM = L::STAGE_BYTES * ST_STAGES;
    // per DEVICE (a process may drive several): SM count, and the opt-in shared-memory size of this instantiation
    static int n_sm_of[FS2D_MAX_DEVICES] = {};
    static bool attr_set_on[FS2D_MAX_DEVICES] = {};
    int dev = 0;
    FS2D_CUDA_CHECK(cudaGetDevice(&dev));
    FS2D_REQUIRE(dev >= 0 && dev < FS2D_MAX_DEVICES, "device ordinal out of range");
    if (!n_sm_of[dev]) FS2D_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm_of[dev], cudaDevAttrMultiProcessorCount, dev));
    if (!attr_set_on[dev]) {
        FS2D_CUDA_CHECK(cudaFuncSetAttribute(k_stream<Op, ST_STAGES, MIN_CTAS, ST_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_set_on[dev] = true;
    }
    const int n_sm = n_sm_of[dev];
    StreamMaps<Op::NF> maps;
    for (int k = 0; k < Op::NF; ++k)
        if (int e = make_map(&maps.f[k], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, fields[k], (uint64_t)Op::chan(k) * d.Y, d.rows,
                             Op::chan(k) * ST_BC, L::ROWS))
            return e;
    if (int e = make_map(&maps.mask, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, mask, d.Y, d.rows, L::MASK_COLS, L::MASK_ROWS)) return e;
    StreamGeom g;
    g.tiles_j = (d.Y + ST_TC - 1) / ST_TC;
    g.n_tiles = g.tiles_j * ((d.r1 - d.r0 + ST_TR - 1) / ST_TR);
    const int grid = g.n_tiles < MIN_CTAS * n_sm ? g.n_tiles : MIN_CTAS * n_sm;
    ++g_launches;
    k_stream<Op, ST_STAGES, MIN_CTAS, ST_THREADS><<<grid, ST_THREADS, SMEM, s>>>(maps, op, d, g);
    return FS2D_OK;
}
template <class Op>
static int launch_stream(const Op &op, const float *const *fields, const uint8_t *mask, const fs2d_dom &d, cudaStream_t s) {
    if (g_stream_cfg == 1) return launch_stream_cfg<Op, 2, 3, 256>(op, fields, mask, d, s);
    if (g_stream_cfg == 2) return launch_stream_cfg<Op, 3, 2, 512>(op, fields, mask, d, s);
    if (g_stream_cfg == 3) return launch_stream_cfg<Op, 2, 2, 512>(op, fields, mask, d, s);
    return launch_stream_cfg<Op, 3, 2, 256>(op, fields, mask, d, s);
}

int stream_cip_advect(float *fn, float *fxn, float *fyn, const float *fc, const float *fxc, const float *fyc,
                      const uint8_t *mask, const fs2d_dom &d, float dt, float dx, float dx2, float dx3, bool p2,
                      cudaStream_t s) {
    const float *fields[3] = {fc, fxc, fyc};
    if (p2)
        return launch_stream(OpAdvect<true>{fn, fxn, fyn, dt, dx, DivC<true>(dx), DivC<true>(dx2), DivC<true>(dx3)}, fields,
                             mask, d, s);
    return launch_stream(OpAdvect<false>{fn, fxn, fyn, dt, dx, DivC<false>(dx), DivC<false>(dx2), DivC<false>(dx3)}, fields,
                         mask, d, s);
}

int stream_cip_nonadv(float *fn, const float *fc, const float *pc, const uint8_t *mask, const fs2d_dom &d, float dt, float dx,
                      float re, bool p2, cudaStream_t s) {
    const float *fields[2] = {fc, pc};
    const float dx2 = dx * dx;
    if (p2) return launch_stream(OpNonadv<true>{fn, dt, re, DivC<true>(dx), DivC<true>(dx2)}, fields, mask, d, s);
    return launch_stream(OpNonadv<false>{fn, dt, re, DivC<false>(dx), DivC<false>(dx2)}, fields, mask, d, s);
}

int stream_vort_apply(float *vn, float *w, float *wabs, const float *vc, const uint8_t *mask, const fs2d_dom &d, float dx,
                      float dtw, bool p2, cudaStream_t s) {
    const float *fields[1] = {vc};
    if (p2) return launch_stream(OpVort<true>{vn, w, wabs, dtw, DivC<true>(dx)}, fields, mask, d, s);
    return launch_stream(OpVort<false>{vn, w, wabs, dtw, DivC<false>(dx)}, fields, mask, d, s);
}

}  // namespace fs2d

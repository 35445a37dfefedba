// fs2d_common.cuh -- shared device helpers for the sm_100a kernels of libfs2d.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fs2d.h"

namespace fs2d {

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char *fmt, ...);
#define FS2D_CUDA_CHECK(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            fs2d::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return FS2D_E_CUDA;                                                            \
        }                                                                                  \
    } while (0)
#define FS2D_REQUIRE(cond, msg)                                   \
    do {                                                          \
        if (!(cond)) {                                            \
            fs2d::set_error("bad argument: %s (%s)", msg, #cond); \
            return FS2D_E_BADARG;                                 \
        }                                                         \
    } while (0)
#define FS2D_LAUNCH_CHECK() FS2D_CUDA_CHECK(cudaGetLastError())

extern unsigned long long g_launches;  // every kernel launch of the library bumps this
constexpr int FS2D_MAX_DEVICES = 64;   // per-device launch facts (SM count, opt-in shared-memory attributes) are kept per ordinal
int check_dom(const fs2d_dom &d);
// in-place sparse pressure BC (gather then scatter); no-op for n <= 0
void launch_p_bc(float *p, const int32_t *tgt, const int32_t *src0, const int32_t *src1, const uint8_t *kind, float *scratch,
                 int n, cudaStream_t s);
inline unsigned nblk(int n, int b) { return (unsigned)((n + b - 1) / b); }
bool is_pow2(float x);
// one fused pass of T Jacobi iterations p_in -> p_out (fs2d_fused.cu); order / n_order: tile list of fs2d_fused_order for
// exactly this geometry, or nullptr (classified on the fly)
int fused_pass(const float *p_in, float *p_out, const float *src, const uint8_t *pcode, const fs2d_dom &d, int T,
               cudaStream_t s, int skip_from, int skip_n, bool emit, const int *order, int n_order);
extern int g_tail_emit;      // fs2d_set_tuning(4, v), see fs2d_pressure.cu
// TMA-fed streaming versions of the stencil kernels (fs2d_stream.cu); g_stream: fs2d_set_tuning(2, 0/1)
extern int g_stream, g_stream_cfg;
bool stream_ok(const fs2d_dom &d, const void *const *ptrs, int n);
int stream_cip_advect(float *fn, float *fxn, float *fyn, const float *fc, const float *fxc, const float *fyc,
                      const uint8_t *mask, const fs2d_dom &d, float dt, float dx, float dx2, float dx3, bool p2,
                      cudaStream_t s);
int stream_cip_nonadv(float *fn, const float *fc, const float *pc, const uint8_t *mask, const fs2d_dom &d, float dt, float dx,
                      float re, bool p2, cudaStream_t s);
int stream_vort_apply(float *vn, float *w, float *wabs, const float *vc, const uint8_t *mask, const fs2d_dom &d, float dx,
                      float dtw, bool p2, cudaStream_t s);
bool fused_supported(const float *pa, const float *pb, const float *src, const uint8_t *pcode, const fs2d_dom &d);

// ---- indexing (clamp-to-edge sample(), fs/differentiation.py:4-9) -----------------------------
__device__ __forceinline__ int CR(const fs2d_dom &d, int r) { return min(max(r, d.clo), d.chi); }
__device__ __forceinline__ int CJ(const fs2d_dom &d, int j) { return min(max(j, 0), d.Y - 1); }
__device__ __forceinline__ size_t IX(const fs2d_dom &d, int r, int j) { return (size_t)r * (size_t)d.Y + (size_t)j; }

__device__ __forceinline__ float ld1(const float *f, const fs2d_dom &d, int r, int j) {
    return __ldg(f + IX(d, CR(d, r), CJ(d, j)));
}
__device__ __forceinline__ float2 ld2(const float *f, const fs2d_dom &d, int r, int j) {
    return __ldg(reinterpret_cast<const float2 *>(f) + IX(d, CR(d, r), CJ(d, j)));
}

// Templated variants.  CL = true: clamp-to-edge as above.  CL = false: the caller guarantees that the access lies
// inside the clamp window (see block_interior), so the neighbour is a plain offset from the cell -- no min/max and
// one address computation shared by all fields.  The streaming kernels were issue-bound on this index arithmetic.
template <bool CL>
__device__ __forceinline__ float ld1(const float *f, const fs2d_dom &d, int r, int j) {
    if (CL) return ld1(f, d, r, j);
    return __ldg(f + ((ptrdiff_t)r * d.Y + j));
}
template <bool CL>
__device__ __forceinline__ float2 ld2(const float *f, const fs2d_dom &d, int r, int j) {
    if (CL) return ld2(f, d, r, j);
    return __ldg(reinterpret_cast<const float2 *>(f) + ((ptrdiff_t)r * d.Y + j));
}
// Dense-kernel grids put the COLUMN blocks on blockIdx.x (fastest-varying in the block scheduler): the blocks resident
// at any time then cover a few complete rows (64 KB contiguous each at Y = 8192) instead of a 512-byte wide column
// of thousands of rows (64 KB stride: one DRAM page activation per 512 bytes).
#define FS2D_COLBLK ((int)blockIdx.x)
#define FS2D_ROWBLK ((int)blockIdx.y)
// True if every cell of this thread block (block_rows rows from the block's first row, blockDim.x columns) and every
// neighbour within `halo` cells of it lies inside the clamp window [clo, chi] x [0, Y-1], and all its rows are
// updated rows (< r1).  Block-uniform.
__device__ __forceinline__ bool block_interior(const fs2d_dom &d, int block_rows, int halo) {
    const int rb = d.r0 + FS2D_ROWBLK * block_rows, jb = FS2D_COLBLK * blockDim.x;
    return rb - halo >= d.clo && rb + block_rows - 1 + halo <= d.chi && rb + block_rows <= d.r1 && jb - halo >= 0 &&
           jb + (int)blockDim.x - 1 + halo <= d.Y - 1;
}

// ---- float2 arithmetic with the reference's per-component order ----------------------------------
// Multiplications use the Blackwell packed instruction (mul.rn.f32x2 -> FMUL2: both components in one issue slot;
// each lane rounds exactly like mul.rn.f32).  Additions stay scalar on purpose: ptxas 12.9 contracts a
// mul.rn.f32x2 feeding an add/sub.rn.f32x2 into FFMA2 even under --fmad false (measured, scripts/probes/README.md),
// which would change the rounding; it does not contract FMUL2 with a scalar FADD.  The vector kernels (CIP advection:
// ~210 fp32 operations per cell) are issue-bound, so this removes ~30 % of their FP instructions.
#ifndef FS2D_NO_F32X2
__device__ __forceinline__ uint64_t f2_pack(float2 a) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}
__device__ __forceinline__ float2 f2_unpack(uint64_t v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ float2 operator*(float2 a, float2 b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b)));
    return f2_unpack(r);
}
__device__ __forceinline__ float2 operator*(float2 a, float s) { return a * make_float2(s, s); }
__device__ __forceinline__ float2 operator*(float s, float2 a) { return make_float2(s, s) * a; }
#else
__device__ __forceinline__ float2 operator*(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
__device__ __forceinline__ float2 operator*(float s, float2 a) { return make_float2(s * a.x, s * a.y); }
__device__ __forceinline__ float2 operator*(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
#endif
__device__ __forceinline__ float2 operator+(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 operator-(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 operator-(float2 a) { return make_float2(-a.x, -a.y); }
// x / y, IEEE-exact (div.rn.f32 semantics), that never sends a ZERO DIVIDEND into the division: the correctly
// rounded fp32 division is a reciprocal + Newton sequence with a range check (FCHK) that diverts zero / denormal
// operands to a ~100-instruction slow path.  Exactly uniform regions (free stream, quiescent start: Laplacians and
// vorticity gradients are exactly 0 there) made that slow path the common case.  0 / y = +-0 (sign = XOR of the
// signs) for every y but 0 and NaN, where it is NaN.
__device__ __forceinline__ float fdiv_z(float x, float y) {
#ifdef FS2D_FDIV_BRANCHLESS
    const bool z = x == 0.0f;
    const bool ybad = !(fabsf(y) > 0.0f);   // y is 0 or NaN
    const float q = (z ? 1.0f : x) / ((z && ybad) ? 1.0f : y);
    const float sz = __int_as_float((__float_as_int(x) ^ __float_as_int(y)) & (int)0x80000000);
    return z ? (ybad ? __int_as_float(0x7fffffff) : sz) : q;
#else
    // a real branch: warps whose dividends are all zero (uniform regions: the whole grid of a quiescent start) skip the
    // division sequence altogether; in mixed warps the non-zero lanes divide as usual
    if (x == 0.0f) {
        const bool ybad = !(fabsf(y) > 0.0f);   // y is 0 or NaN
        return ybad ? __int_as_float(0x7fffffff) : __int_as_float((__float_as_int(x) ^ __float_as_int(y)) & (int)0x80000000);
    }
    return x / y;
#endif
}
__device__ __forceinline__ float2 operator/(float2 a, float s) { return make_float2(fdiv_z(a.x, s), fdiv_z(a.y, s)); }
__device__ __forceinline__ float2 operator/(float2 a, float2 b) { return make_float2(fdiv_z(a.x, b.x), fdiv_z(a.y, b.y)); }

// Division by a grid constant c.  When c is a power of two, x / c == x * (1/c) bit-for-bit (exact
// scaling), so the pow2 instantiation avoids the IEEE division sequence without changing results.
template <bool P2>
struct DivC {
    float c, inv;
    __host__ __device__ DivC(float c_) : c(c_), inv(1.0f / c_) {}
    // division by s*c for s = +-1: both the divisor and its reciprocal just change sign (no device-side division)
    __device__ __forceinline__ DivC signed_by(float s) const {
        DivC r(*this);
        r.c = c * s;
        r.inv = inv * s;
        return r;
    }
    __device__ __forceinline__ float operator()(float x) const { return P2 ? x * inv : fdiv_z(x, c); }
    __device__ __forceinline__ float2 operator()(float2 x) const {
        return P2 ? x * inv : make_float2(fdiv_z(x.x, c), fdiv_z(x.y, c));
    }
};

// fs/differentiation.py:12-14
__device__ __forceinline__ float sign1(float x) { return x < 0.0f ? -1.0f : 1.0f; }

// dense-kernel launch geometry: block = (TX, TY) threads, one cell per thread
constexpr int TX = 64;
constexpr int TY = 4;
inline dim3 dense_block() { return dim3(TX, TY, 1); }
inline dim3 dense_grid(const fs2d_dom &d) {
    return dim3((unsigned)((d.Y + TX - 1) / TX), (unsigned)((d.r1 - d.r0 + TY - 1) / TY), 1);
}
// launch grid of the streaming kernels that process `nu` rows per thread (fs2d_kernels.cu)
inline dim3 dense_grid_nu(const fs2d_dom &d, int nu) {
    return dim3((unsigned)((d.Y + TX - 1) / TX), (unsigned)((d.r1 - d.r0 + TY * nu - 1) / (TY * nu)), 1);
}
#define FS2D_CELL(d, r, j)                                   \
    const int j = FS2D_COLBLK * blockDim.x + threadIdx.x;     \
    const int r = (d).r0 + FS2D_ROWBLK * blockDim.y + threadIdx.y; \
    if (j >= (d).Y || r >= (d).r1) return;

}  // namespace fs2d

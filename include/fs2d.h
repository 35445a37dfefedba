/*
 * fs2d.h -- C ABI of libfs2d.so: the B200 (sm_100a) implementation of the per-step hot path of
 * takah29/2d-fluid-simulator (reference commit ebd5d75).
 *
 * The reference has no native/FFI boundary (pure Python + Taichi JIT); each entry point below
 * replaces one Taichi kernel and is what a ctypes/cffi binding of that kernel would bind.  All
 * pointers are DEVICE pointers owned by the caller (PyTorch tensors in our host layer); the
 * library never allocates or frees field memory.  `stream` is a cudaStream_t passed as void*.
 * Every function returns 0 on success or a negative FS2D_E_* code; fs2d_last_error() returns a
 * thread-local message for the last failure.  Not thread-safe per stream; one host thread per GPU.
 *
 * Grid: X rows (slow axis, index i) x Y columns (fast axis, index j), row-major, fp32.
 * Vector fields are AoS float2 [i][j][2] exactly as Taichi's Vector.field / to_numpy()
 * (fs/double_buffer.py:10-11).  Cell types (uint8): 0 fluid, 1 wall, 2 inflow, 3 outflow
 * (fs/boundary_condition.py:82,225).
 *
 * Row-strip domains: every dense kernel takes an fs2d_dom describing the LOCAL array of one rank:
 *   rows      allocated rows of the local array (owned rows + halo rows)
 *   Y         columns
 *   r0, r1    the kernel updates local rows [r0, r1)
 *   clo, chi  clamp-to-edge bounds of sample() (fs/differentiation.py:4-9) in local row indices:
 *             a read of row r uses row min(max(r, clo), chi).  Single GPU: clo=0, chi=rows-1.
 *             Interior ranks: clo=0, chi=rows-1 as well (halo rows exist); edge ranks clamp to
 *             their first/last owned row.
 *   gi0       global row index of local row 0 (red/black parity)
 *
 * Floating point: kernels are compiled with -fmad=false and use IEEE div/sqrt, literal
 * left-to-right operation order of the reference source, fminf/fmaxf NaN rule, so
 * results are bit-identical to the CPU oracle (oracle/fs2d_oracle.c).
 */
#ifndef FS2D_H
#define FS2D_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FS2D_OK 0
#define FS2D_E_BADARG (-1)
#define FS2D_E_CUDA (-2)
#define FS2D_E_NCCL (-3)

typedef struct {
    int rows, Y, r0, r1, clo, chi, gi0;
} fs2d_dom;

/* advection scheme selector of fs2d_mac_update (fs/advection.py:12-24, :27-60) */
#define FS2D_SCHEME_UPWIND 0
#define FS2D_SCHEME_KK 1

/* pressure-BC cell codes (derived from the mask by the host layer; see DESIGN.md "pcode").
 * A pcode byte = code (low nibble) | neighbour bits (high nibble): 0x10/0x20/0x40/0x80 set when the
 * clamped neighbour at (i-1,j)/(i+1,j)/(i,j-1)/(i,j+1) is a BC cell (code not FLUID / W_NONE), i.e. when
 * its post-BC value differs from its stored value. */
#define FS2D_PC_FLUID 0   /* mask 0                                                          */
#define FS2D_PC_W_IM 1    /* wall: p = p(i-1,j)          boundary_condition.py:46-47         */
#define FS2D_PC_W_IP 2    /* wall: p = p(i+1,j)          :48-49                              */
#define FS2D_PC_W_JM 3    /* wall: p = p(i,j-1)          :50-51                              */
#define FS2D_PC_W_JP 4    /* wall: p = p(i,j+1)          :52-53                              */
#define FS2D_PC_W_IM_JP 5 /* wall: (p(i-1,j)+p(i,j+1))/2 :54-55                              */
#define FS2D_PC_W_IP_JP 6 /* wall: (p(i+1,j)+p(i,j+1))/2 :56-57                              */
#define FS2D_PC_W_IM_JM 7 /* wall: (p(i-1,j)+p(i,j-1))/2 :58-59                              */
#define FS2D_PC_W_IP_JM 8 /* wall: (p(i+1,j)+p(i,j-1))/2 :60-61                              */
#define FS2D_PC_W_NONE 9  /* wall, no branch taken: value is never written ("stale", T1)     */
#define FS2D_PC_INFLOW 10 /* mask 2: BC value p(i+1,j)   :62-63 ; cell is relaxed afterwards */
#define FS2D_PC_OUTFLOW 11 /* mask 3: BC value 0         :64-65 ; cell is relaxed afterwards */

const char *fs2d_last_error(void);
int fs2d_version(void);
/* number of kernels this library has launched in this process (bench.py gpu_launches) */
unsigned long long fs2d_launch_count(void);
/* performance knobs (never change results):
 * key 0 = rows marched per warp by the single-iteration Jacobi sweep {1, 2, 4 (default), 8, 16};
 * key 2 = TMA-fed streaming versions of the CIP-path stencil kernels {0: off, 1: CIP advection (default), 2: also the
 *         non-advection phase and the vorticity confinement};
 * key 3 = streaming-kernel shape {stages x CTAs/SM x threads}: {0: 3x2x256, 1: 2x3x256 (default), 2: 3x2x512, 3: 2x2x512}.
 * key 4 = tail of fs2d_jacobi_update {1 (default): the last fused pass also emits the BC values of its penultimate state, so
 *         ONE literal iteration ends the update; 0: two literal iterations}.
 * key 5 = dye non-advection phase {1 (default): four cells per thread with 128-bit accesses when Y % 4 == 0 and the fields are
 *         16-byte aligned; 0: always one cell per thread}.
 * key 6 = CIP non-advection phase of the velocity, the same choice {1 (default): four cells per thread; 0: one cell per thread}.
 * Unknown keys / values return FS2D_E_BADARG and change nothing. */
int fs2d_set_tuning(int key, int value);
/* 1 if the library was built for sm_100a and a device of compute capability 10.x is current */
int fs2d_device_ok(void);

/* ---- sparse in-place boundary conditions (tables built once from the static mask) ---------- */
/* BoundaryCondition.set_velocity_boundary_condition, fs/boundary_condition.py:16-39 (gather form,
 * SURVEY T4).  Entry e: kind 0: v[tgt] = -v[src]; kind 1: v[tgt] = bc_const[tgt];
 * kind 2: v[tgt].x = max(v[src].x, 0.05).  tgt/src are linear cell indices into the local array.
 * scratch: >= 2*n floats. */
int fs2d_vel_bc(float *v, const float *bc_const, const int32_t *tgt, const int32_t *src, const uint8_t *kind,
                float *scratch, int n, void *stream);
/* BoundaryCondition.set_pressure_boundary_condition, fs/boundary_condition.py:41-65 (gather form).
 * kind 0: p[tgt] = p[src0]; kind 1: p[tgt] = (p[src0] + p[src1]) / 2; kind 2: p[tgt] = 0.
 * scratch: >= n floats. */
int fs2d_pressure_bc(float *p, const int32_t *tgt, const int32_t *src0, const int32_t *src1, const uint8_t *kind,
                     float *scratch, int n, void *stream);

/* ---- dense stencil kernels ------------------------------------------------------------------ */
/* MacSolver._update_velocities, fs/solver.py:94-107 (fluid cells) */
int fs2d_mac_update(float *vn, const float *vc, const float *pc, const uint8_t *mask, fs2d_dom d, float dt, float dx,
                    float re, int scheme, void *stream);
/* CipMacSolver._non_advection_phase, fs/solver.py:229-240 (not-wall cells) */
int fs2d_cip_nonadv(float *fn, const float *fc, const float *pc, const uint8_t *mask, fs2d_dom d, float dt, float dx,
                    float re, void *stream);
/* CipMacSolver._non_advection_phase_grad, fs/solver.py:242-261; two_dx = (float)(2.0*dx) */
int fs2d_cip_nonadv_grad(float *fxn, float *fyn, const float *fxc, const float *fyc, const float *fc, const float *fn,
                         const uint8_t *mask, fs2d_dom d, float two_dx, void *stream);
/* CipMacSolver._advection_phase/_cip_advect, fs/solver.py:267-332 (fluid cells);
 * dx2 = (float)(dx*dx), dx3 = (float)(dx*dx*dx) folded in double by the caller */
int fs2d_cip_advect(float *fn, float *fxn, float *fyn, const float *fc, const float *fxc, const float *fyc,
                    const float *v, const uint8_t *mask, fs2d_dom d, float dt, float dx, float dx2, float dx3,
                    void *stream);
/* CipMacSolver._set_grad, fs/solver.py:207-211 (all cells) */
int fs2d_set_grad(float *fx, float *fy, const float *f, fs2d_dom d, float dx, void *stream);
/* VorticityConfinement._calc_vorticity, fs/vorticity_confinement.py:27-32 (fluid cells) */
int fs2d_vort_calc(float *w, float *wabs, const float *vc, const uint8_t *mask, fs2d_dom d, float dx, void *stream);
/* VorticityConfinement._add_vorticity, fs/vorticity_confinement.py:34-55; dtw = (float)(dt*weight) */
int fs2d_vort_add(float *vn, const float *vc, const float *w, const float *wabs, const uint8_t *mask, fs2d_dom d,
                  float dx, float dtw, void *stream);
/* VorticityConfinement.apply, fs/vorticity_confinement.py:57-59: fs2d_vort_calc followed by fs2d_vort_add in one pass
 * (the |vorticity| of a fluid neighbour is recomputed from vc instead of being read back; non-fluid neighbours keep
 * their stored value).  Same results in w, wabs and vn as the two calls; vn must not alias vc. */
int fs2d_vort_apply(float *vn, float *w, float *wabs, const float *vc, const uint8_t *mask, fs2d_dom d, float dx, float dtw,
                    void *stream);
/* Velocity source terms of predict_p (fs/pressure_updater.py:23-38) for every cell of rows [r0, r1):
 * src[i][j] = (t2, t3), t2 = (sx.x^2 + sy.y^2 + sy.x*sx.y)/8, t3 = dx*(sx.x + sy.y)/(8*dt) with
 * sx = v(i+1,j) - v(i-1,j), sy = v(i,j+1) - v(i,j-1).  v is constant during one pressure update, so the
 * host computes this once per update; sweeps then evaluate the literal (t1 + t2) - t3. */
int fs2d_pressure_source(float *src, const float *vc, fs2d_dom d, float dt, float dx, void *stream);
/* JacobiPressureUpdater._update, fs/pressure_updater.py:62-66 (not-wall cells), src from
 * fs2d_pressure_source.  inline_bc != 0: neighbour pressures are the post-BC values of `pc`
 * recomputed from `pcode` (pc itself is not modified) == set_pressure_boundary_condition(pc)
 * followed by _update; inline_bc == 0: pc is read as is (caller already applied the BC). */
int fs2d_jacobi_sweep(float *pn, const float *pc, const float *src, const uint8_t *pcode, fs2d_dom d, int inline_bc,
                      void *stream);
/* n_sweeps of {pressure BC, Jacobi} with ping-pong between pa (current) and pb (next); equal to
 * n_sweeps reference iterations (fs/pressure_updater.py:56-60) INCLUDING the final contents of the
 * BC cells of both buffers.  (tgt, src0, src1, kind, n_bc): table as in fs2d_pressure_bc; scratch:
 * >= n_bc floats.  fuse_mask: bit t set (1 <= t <= 12) allows all but the last iterations to run as
 * fused passes of t iterations (fs2d_jacobi_fused) -- the caller must have verified that pass size against
 * the preconditions listed there; 0 = literal iterations only.  orders / n_orders: NULL, or HOST arrays of 13
 * entries indexed by t holding the tile list of fs2d_fused_order(pcode, d, t, 0, 0, ...) (device pointer) and
 * its length for every t of fuse_mask (a NULL entry: that pass size classifies its tiles on the fly).
 * *final_in_b = 1 if the current buffer after the call is pb (n_sweeps odd). */
int fs2d_jacobi_update(float *pa, float *pb, const float *src, const uint8_t *pcode, fs2d_dom d, int n_sweeps,
                       const int32_t *tgt, const int32_t *src0, const int32_t *src1, const uint8_t *kind, float *scratch,
                       int n_bc, int fuse_mask, const int32_t *const *orders, const int *n_orders, int *final_in_b,
                       void *stream);
/* The schedule fs2d_jacobi_update uses: sizes[k] > 0 = one fused pass of that many iterations, 0 = one literal
 * iteration {fs2d_pressure_bc, fs2d_jacobi_sweep}; every entry flips the buffers once.  With a fused pass in the
 * schedule it ends {..., fused pass through fs2d_jacobi_fused_tail, ONE literal iteration} (fs2d_set_tuning(4, 0):
 * {..., two literal iterations}).  (A multi-rank host runs the same schedule with a halo exchange in front of every
 * entry.) */
int fs2d_jacobi_plan(int n_sweeps, int fuse_mask, int *sizes, int cap, int *n_entries);
/* Tile list of one fused pass geometry (rows [r0, r1) of d tiled for T iterations, minus the tile rows
 * [skip_from, skip_from + skip_n), see fs2d_jacobi_fused_part): classifies every tile from pcode -- open fluid / has BC
 * cells, global edges or cells outside the grid / nothing to store -- and writes to `order` (DEVICE array of `cap` >=
 * number-of-tiles ints) the tiles that have work to do, the expensive ones first, as the kernel will deal them to its
 * persistent CTAs.  counts (HOST array of 3): entries written, how many of them are slow tiles, tiles dropped.  The list
 * depends on pcode, d, T and the skipped rows only: build it once per mask and geometry (it synchronises the stream). */
int fs2d_fused_order(const uint8_t *pcode, fs2d_dom d, int T, int skip_from, int skip_n, int32_t *order, int cap, int *counts,
                     void *stream);
/* One fused pass: T reference iterations {BC, sweep} computed on chip, p_in -> relaxed cells of
 * p_out (rows [r0, r1)); bit-identical to T calls of fs2d_pressure_bc + fs2d_jacobi_sweep on the relaxed
 * cells.  BC cells of p_in/p_out are neither read nor written (their values are recomputed from pcode).
 * order / n_order: the tile list of fs2d_fused_order for exactly this (pcode, d, T), or NULL / 0 (the tiles are then
 * classified on the launch stream before the pass: same results, an extra pass over pcode and an unbalanced order).
 * Preconditions (checked by the host layer, fs/_bc_tables.py): Y % 16 == 0 and 16-byte aligned fields;
 * T <= t_max of fs2d_fused_tile; the mask's dependency reach fits the tile halo (fused_reach_ok); no
 * inflow cell reads a wall-BC cell; never-written wall cells hold equal values in p_in and p_out; in a
 * row strip the halo rows [r0-T, r1+T) of p_in and src are up to date. */
int fs2d_jacobi_fused(float *p_out, const float *p_in, const float *src, const uint8_t *pcode, fs2d_dom d, int T,
                      const int32_t *order, int n_order, void *stream);
/* The same pass restricted to part of its tile rows: the tiling of rows [r0, r1) (tile rows of fs2d_fused_tile rows -
 * 2 * halo_rows output rows, anchored at r0) minus the tile rows [skip_from, skip_from + skip_n).  A multi-rank host
 * runs the interior tile rows (a row window through fs2d_jacobi_fused) while the halo SendRecv is in flight and then the
 * first and last tile rows, which read the fresh halo, in ONE launch through this entry point. */
int fs2d_jacobi_fused_part(float *p_out, const float *p_in, const float *src, const uint8_t *pcode, fs2d_dom d, int T,
                           int skip_from, int skip_n, const int32_t *order, int n_order, void *stream);
/* fs2d_jacobi_fused_part whose pass also stores, into the wall-BC cells of p_in (rows [r0, r1) of the launched tile rows;
 * those cells are never read), the BC values of the pass's PENULTIMATE state -- what the reference leaves there
 * (SURVEY T1).  A pass of T iterations ending at iteration n - 1 followed by ONE literal iteration then reproduces both
 * physical buffers of an n-iteration update; fs2d_jacobi_plan returns such a schedule (entry n - 2 fused, entry n - 1
 * literal). */
int fs2d_jacobi_fused_tail(float *p_out, float *p_in, const float *src, const uint8_t *pcode, fs2d_dom d, int T, int skip_from,
                           int skip_n, const int32_t *order, int n_order, void *stream);
/* tile geometry of the fused kernel for T iterations per pass: loaded tile rows x cols, the halo it discards on
 * each side (rows: T; columns: T rounded up to 4 -- TMA box starts must be 16-byte aligned) and the largest T */
int fs2d_fused_tile(int T, int *rows, int *cols, int *halo_rows, int *halo_cols, int *t_max);
/* One colour pass of RedBlackSorPressureUpdater, fs/pressure_updater.py:98-114 (fluid cells of
 * colour `parity`, (i_global + j) % 2); pc may alias pn (even pass, :96); src as above */
int fs2d_rbsor_pass(float *pn, const float *pc, const float *src, const uint8_t *mask, fs2d_dom d, float omega,
                    float one_minus_omega, int parity, void *stream);
/* RedBlackSorPressureUpdater._update, fs/pressure_updater.py:86-96: the odd pass (pn <- pc) followed by the even pass
 * (pn <- pn) in ONE pass over HBM, bit-identical to fs2d_rbsor_pass(pn, pc, .., 1) + fs2d_rbsor_pass(pn, pn, .., 0);
 * pn must not alias pc.  (A rank of a row-strip decomposition keeps the two passes: the neighbour's odd cells travel
 * in between.) */
int fs2d_rbsor_iteration(float *pn, const float *pc, const float *src, const uint8_t *mask, fs2d_dom d, float omega,
                         float one_minus_omega, void *stream);
/* ---- dye transport (3-channel AoS fields; DyeMacSolver / DyeCipMacSolver, fs/solver.py:110-161, :335-401) ---- */
/* DyeBoundaryCondition.set_dye_boundary_condition, fs/boundary_condition.py:94-99: dye[tgt] = bc_dye[tgt] for
 * the inflow cells listed in tgt (linear cell indices into the local array) */
int fs2d_dye_bc(float *dye, const float *bc_dye, const int32_t *tgt, int n, void *stream);
/* clamp_field, fs/solver.py:46-49 (all cells of rows [r0, r1), all `channels` components) */
int fs2d_clamp(float *f, fs2d_dom d, int channels, float low, float high, void *stream);
/* DyeMacSolver._update_dye, fs/solver.py:157-161: dn = dc - dt * advect(vc, dc) (fluid cells) */
int fs2d_dye_mac(float *dn, const float *dc, const float *vc, const uint8_t *mask, fs2d_dom d, float dt, float dx, int scheme,
                 void *stream);
/* DyeCipMacSolver._non_advection_phase_dye, fs/solver.py:378-383 (not-wall cells) */
int fs2d_dye_nonadv(float *dn, const float *dc, const uint8_t *mask, fs2d_dom d, float dt, float dx, float re, void *stream);
/* _non_advection_phase_grad / _advection_phase / _set_grad on the 3-channel dye fields (fs/solver.py:242-332, :207-211);
 * v is the 2-channel advecting velocity */
int fs2d_dye_nonadv_grad(float *fxn, float *fyn, const float *fxc, const float *fyc, const float *fc, const float *fn,
                         const uint8_t *mask, fs2d_dom d, float two_dx, void *stream);
int fs2d_dye_cip_advect(float *fn, float *fxn, float *fyn, const float *fc, const float *fxc, const float *fyc,
                        const float *v, const uint8_t *mask, fs2d_dom d, float dt, float dx, float dx2, float dx3,
                        void *stream);
int fs2d_dye_set_grad(float *fx, float *fy, const float *f, fs2d_dom d, float dx, void *stream);
/* FluidSimulator._to_norm/_to_pressure/_to_vorticity, DyeFluidSimulator._to_dye (fs/fluid_simulator.py:38-58,
 * :121-126; fs/visualization.py:8-22): rgb (rows, Y, 3) <- mode 0 norm + pressure tint, 1 pressure, 2 vorticity,
 * 3 dye; wall cells get the wall colour (0.5, 0.7, 0.5).  Unused inputs may be NULL. */
int fs2d_render(float *rgb, const float *v, const float *p, const float *dye, const uint8_t *mask, fs2d_dom d, float dx,
                int mode, void *stream);
/* limit_field, fs/solver.py:38-43 (all cells, in place) */
int fs2d_limit(float *v, fs2d_dom d, float limit, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FS2D_H */

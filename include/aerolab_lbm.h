/*
 * aerolab_lbm.h -- C ABI of libaerolab_lbm.so, the B200 (sm_100a) D2Q9
 * lattice-Boltzmann wind tunnel.
 *
 * The reference (583phoenix-hue/Airfoil-CFD-Tool) has no FFI for this path:
 * its tunnel is browser JavaScript + GLSL in
 * pages/airfoil_flow_lbm_aerolab.html ("HTML:n" below).  Each entry point
 * names the reference function it stands in for, so a maintainer can replace
 * the iframe component built by pages/Airfoil_Analysis.py:20-42 with calls
 * into this library (see INTEGRATION.md for the ctypes stub).
 *
 * Conventions
 *   - every function returns ALB_OK (0) or a negative ALB_ERR_* code; nothing
 *     throws across the boundary; alb_last_error() gives the message.
 *   - the library owns all device memory; the caller owns every host buffer
 *     (plain C-contiguous arrays, caller-allocated).  No callbacks.
 *   - a handle is one lattice (or one y-slab of a lattice) on one GPU with its
 *     own CUDA stream.  A handle is not thread-safe; different handles are
 *     independent.  alb_step() is asynchronous; getters synchronise.
 *   - fields are row-major [ny][nx], row 0 = bottom of the world window
 *     (world y up, HTML:164,178); populations are SoA [9][ny][nx], fp32, in the
 *     reference's direction order e = (0,0),(1,0),(0,1),(-1,0),(0,-1),(1,1),
 *     (-1,1),(-1,-1),(1,-1) (HTML:238-248).
 *   - there is no CPU fallback: without a CUDA device alb_create() fails with
 *     ALB_ERR_CUDA.
 */
#ifndef AEROLAB_LBM_H
#define AEROLAB_LBM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ALB_VERSION 100            /* 0.1.0 */

#define ALB_OK            0
#define ALB_ERR_INVALID  -1        /* bad argument                              */
#define ALB_ERR_CUDA     -2        /* CUDA runtime error (message has details)  */
#define ALB_ERR_NOMEM    -3        /* device or host allocation failed          */
#define ALB_ERR_STATE    -4        /* call not valid in the handle's state      */
#define ALB_ERR_TIMEOUT  -5        /* a slab waited too long for its neighbour  */

#define ALB_NPANEL       160       /* NP, HTML:131 (161 panel nodes)            */
#define ALB_IPC_BYTES    256       /* size of an alb_ipc_export() blob          */

/* field modes of alb_get_field()/alb_get_rgba(): `fieldMode`, HTML:527, 953 */
#define ALB_FIELD_SPEED  0
#define ALB_FIELD_CP     1
#define ALB_FIELD_VORT   2

typedef struct alb_handle alb_handle;

int         alb_version(void);
const char *alb_error_string(int code);
/* Message of the last failing call on `h`; h == NULL: last failing create on
 * this thread. */
const char *alb_last_error(const alb_handle *h);
int         alb_device_count(int *count);

/* ---- lifetime: replaces the WebGL context/texture/FBO set-up, HTML:86-95,
 *      438-469, and initSim(), HTML:492-500 ---------------------------------- */

/* Whole nx x ny lattice on `device`, initialised like initSim(0.06) with an
 * empty mask, tau = 0.58 (HTML:78), U0 = 0.06 (HTML:472). */
int alb_create(int nx, int ny, int device, alb_handle **out);
/* Rows [y0, y0+ny_local) of an nx x ny_global lattice (multi-GPU y-slabs). */
int alb_create_slab(int nx, int ny_global, int y0, int ny_local, int device,
                    alb_handle **out);
int alb_destroy(alb_handle *h);
int alb_get_dims(const alb_handle *h, int *nx, int *ny_global, int *y0, int *ny_local);

/* ---- control surface ------------------------------------------------------ */

/* U0 slider (HTML:956-959) and TAU (HTML:78, a constant there).  Doubles, like
 * the JS variables; the step uses their fp32 roundings (gl.uniform1f,
 * HTML:521-522), the diagnostics the doubles.  Does not touch the flow. */
int alb_set_params(alb_handle *h, double u0, double tau);
int alb_get_params(const alb_handle *h, double *u0, double *tau);
/* initSim(u0), HTML:492-500: every cell (solids and borders too) := fp32 of
 * the float64 equilibrium at rho = 1, u = (u0, 0).  Also sets U0 = u0.
 * On a slab with connected neighbours the ghost rows are refilled too, while a
 * neighbour that is one step behind may still be pushing halo rows into them:
 * call it only when ALL slabs of the lattice have been synchronised (alb_sync on
 * each, then a barrier between their owners) and let no slab step before all have
 * been reset (DistributedTunnel.reset and LocalMultiTunnel.reset do exactly that). */
int alb_reset(alb_handle *h, double u0);

/* ---- geometry: replaces rotate/panelise/rasterMask/applyGeometry,
 *      HTML:133-182, 559-586.  The flow is NOT re-initialised (HTML:579-584). */

/* xy: npts pairs (x0,y0,x1,y1,...) of chord-normalised coordinates, e.g. the
 * `coords_after` of main.py:607 after the 6-decimal rounding of
 * pages/Airfoil_Analysis.py:34-36.  Rotates by -alpha_deg about (0.25,0),
 * resamples to 161 cosine-spaced nodes, scan-converts this slab's rows on the
 * GPU.  mask_out (nullable): ny_local*nx bytes, 0 or 255. */
int alb_rasterize(alb_handle *h, const double *xy, int npts, double alpha_deg,
                  uint8_t *mask_out);
/* Scan-convert given panel nodes (n >= 2) without rotate/panelise. */
int alb_rasterize_panels(alb_handle *h, const double *xp, const double *yp, int n,
                         uint8_t *mask_out);
/* Upload a mask directly (makeMaskTex, HTML:448-458): ny_global*nx bytes of
 * the WHOLE lattice, nonzero = solid; a slab picks its rows and ghost rows. */
int alb_set_mask(alb_handle *h, const uint8_t *mask_global);
int alb_get_mask(alb_handle *h, uint8_t *mask_out /* ny_local*nx */);
/* Panel nodes of the last alb_rasterize(): sol.xp/sol.yp, HTML:565. */
int alb_get_panels(const alb_handle *h, double *xp, double *yp /* 161 each */);

/* ---- the hot path: simStep(), HTML:510-525 == STEP_FS_SRC, HTML:222-360 --- */

int alb_step(alb_handle *h, int nsteps);
int alb_sync(alb_handle *h);
int alb_step_count(const alb_handle *h, long long *steps);
/* GPU time of the last alb_step() call in milliseconds (CUDA events on the
 * handle's stream); synchronises. */
int alb_last_step_ms(alb_handle *h, float *ms);

/* ---- state: replaces readPixels of the ping-pong set, HTML:547-552 -------- */

int alb_get_populations(alb_handle *h, float *f /* 9*ny_local*nx */);
/* Rows row0 .. row0+nrows-1 of the slab's current populations (0 = first owned row;
 * -1 and ny_local are the ghost rows), dense [9][nrows][nx].  For inspecting bands of
 * lattices too large to copy whole (configs[3] is 19 GB per state).  Synchronises. */
int alb_get_population_rows(alb_handle *h, int row0, int nrows, float *f);
/* Checksum of the current populations: out9[i] = sum over the slab's owned cells of
 * mix64(global cell index, bit pattern of f_i) mod 2^64.  Position dependent and order
 * independent, so the words of all slabs of a lattice add up (mod 2^64) to the words of
 * the same state held by one GPU -- the cheap proof that N GPUs computed what one GPU
 * computes (HTML:283-360 has no decomposition to mirror).  Synchronises. */
int alb_state_hash(alb_handle *h, unsigned long long *out9);
int alb_set_populations(alb_handle *h, const float *f);
/* texC.gba of the current set (HTML:359, 547-552): rho, ux, uy as the last
 * step wrote them (clamped values in the interior, (1,0,0) in solids, (1,U0,0)
 * on the equilibrium border).  Any pointer may be NULL. */
int alb_get_macro(alb_handle *h, float *rho, float *ux, float *uy);
int alb_set_macro(alb_handle *h, const float *rho, const float *ux, const float *uy);
/* Sum of all populations over this slab's rows, in float64. */
int alb_total_mass(alb_handle *h, double *mass);

/* ---- diagnostics ----------------------------------------------------------- */

/* updateFieldsFromMacro(), HTML:596-614: refreshes the sticky autoscale values
 * maxS, cpMin, cpMax (initially 0.6, -1, 1; HTML:593) from the current macro
 * fields.  stats (nullable) receives {maxS, cpMin, cpMax} after the update;
 * U, V, Cp (nullable, ny_local*nx fp32) receive Ufield/Vfield/CpField with NaN
 * in solid cells.  For a slab, stats_partial (see alb_stats_partial) is the
 * building block. */
int alb_update_stats(alb_handle *h, double *stats, float *U, float *V, float *Cp);
/* Raw sweep result of this slab: {mx, cMin, cMax} with mx = 0, cMin = +inf,
 * cMax = -inf when nothing qualified (HTML:598, 608-609). */
int alb_stats_partial(alb_handle *h, double *out3, float *U, float *V, float *Cp);
int alb_set_stats(alb_handle *h, double maxS, double cpMin, double cpMax);
int alb_get_stats(const alb_handle *h, double *stats3);

/* RENDER_FS_SRC.main, HTML:395-420: the scalar t that the palette is indexed
 * with (fp32, NaN in solids), using the sticky stats above and VORT_SCALE =
 * 0.06 (HTML:528).  On a slab the output holds the slab's ny_local rows.  Speed and Cp
 * are cell-local; the vorticity taps of the first / last owned row reach into the
 * neighbouring slab (HTML:411-418; CLAMP_TO_EDGE only at the lattice border), so on a
 * slab mode 2 needs alb_set_macro_ghosts() after the last step, and the sticky stats
 * must have been set to the lattice-wide values (alb_set_stats). */
int alb_get_field(alb_handle *h, int mode, float *t_out /* ny_local*nx */);
/* ux, uy of the first (lo2) and last (hi2) owned row of the current state, 2*nx floats
 * each (nullable): what the neighbouring slabs need for their vorticity taps. */
int alb_get_macro_edges(alb_handle *h, float *lo2, float *hi2);
/* ux, uy of the row below the first owned row (below2 = the lower neighbour's hi2) and of
 * the row above the last one (above2 = the upper neighbour's lo2), 2*nx floats each; NULL
 * where the slab touches the lattice border.  Valid until the next step. */
int alb_set_macro_ghosts(alb_handle *h, const float *below2, const float *above2);
/* Same, passed through the reference palettes (HTML:371-393) to RGBA8. */
int alb_get_rgba(alb_handle *h, int mode, uint8_t *rgba /* ny*nx*4 */);

/* computeForces(), HTML:650-700, on a whole-lattice handle, including both
 * EMAs (0.9/0.1 and 0.85/0.15).  out[10] = {fx, fy, CL_raw, CD_raw, CL_smooth,
 * CD_smooth, sep_frac, surf, rev, any}.  When there are no fluid/solid faces
 * nothing is updated (HTML:672) and any = 0. */
int alb_compute_forces(alb_handle *h, double *out10);
/* This slab's share: {fx, fy, surf, rev} (pressure p = rho/3 over the faces
 * whose FLUID cell lies in this slab). */
int alb_forces_partial(alb_handle *h, double *out4);
int alb_reset_force_emas(alb_handle *h);

/* Momentum-exchange force (not in the reference; BASELINE.json north_star):
 * accumulated inside the step kernel, one value per step, in 2^-40 fixed
 * point (exact, order independent).  n <= ALB_ME_HISTORY most recent steps,
 * oldest first: fxfy[2*k] = Fx, fxfy[2*k+1] = Fy of this slab. */
#define ALB_ME_HISTORY 4095
#define ALB_ME_SCALE   1099511627776.0   /* 2^40 */
int alb_get_me_history(alb_handle *h, int n, long long *fxfy);
/* Last step, whole lattice: out[4] = {Fx, Fy, CL_me, CD_me}. */
int alb_get_me_forces(alb_handle *h, double *out4);
/* Interior cells whose rho or |u| clamp fired (HTML:344-350) since create or
 * the last alb_reset(). */
int alb_clamp_hits(alb_handle *h, long long *hits);

/* U0*CHORD_L/NU_L, HTML:77-79, 865. */
int alb_reynolds(const alb_handle *h, double *re);
/* Stall indicator, HTML:869-884: *state = 0 "Attached" (<5 %), 1 "n% sep"
 * (<25 %), 2 "STALL"; *sep_pct = round(100*sep_frac). */
int alb_stall_state(const alb_handle *h, int *state, int *sep_pct);

/* ---- the page's frame loop, frame(), HTML:902-930, run autonomously ----------
 * nframes frames of steps_per_frame steps each (the page: 4, HTML:80).  After
 * every frame the sticky autoscale values are refreshed (updateFieldsFromMacro)
 * and, on every forces_every-th frame counted since create/reset (the page: 3,
 * HTML:914), computeForces() with its EMAs runs -- all on the device: nothing
 * synchronises with the host inside the loop; one 12-double record per frame is
 * copied to host memory as the frames complete.  controls (nullable): nframes
 * pairs {U0, tau} applied before each frame (the sliders).  series (nullable):
 * nframes x 12 doubles = {CL, CD (EMA; NaN before the first force frame),
 * sep_frac, CL_raw, CD_raw, surf, rev (NaN on frames without forces), maxS,
 * cpMin, cpMax, CL_me, CD_me (momentum exchange of the frame's last step)}.
 * Synchronises at the end.
 * On a slab of a decomposed lattice (alb_create_slab / alb_create_multi) the loop runs
 * the same way, but a slab cannot know the other slabs' extrema and face sums: its
 * records are the frame's RAW PARTIAL reductions, {max s (double), min rho, max rho over
 * the Cp window (+inf / -inf: none), then five 64-bit integers stored bit for bit in the
 * double slots: pressure-face sums fx, fy (2^-40 fixed point, of rho, before the /3),
 * surf, rev, momentum-exchange fx, fy (2^-40), then U0, q = 0.5 U0^2 CHORD_L, and 1.0 on
 * force frames}.  The caller takes max / min / integer sums over the slabs and applies
 * HTML:611-613, 672-699 (aerolab_lbm.distributed.DistributedTunnel.run_frames does);
 * the sticky state of the slab handle itself is left alone. */
#define ALB_FRAME_ROW 12
int alb_run_frames(alb_handle *h, int nframes, int steps_per_frame, int forces_every,
                   const double *controls, double *series);
/* The same in two halves, so that several handles (e.g. the cases of an alpha
 * sweep sharing one GPU) can run their frame loops concurrently: enqueue on each
 * handle, then collect each.  Between the two calls only alb_frames_collect may
 * be used on the handle. */
int alb_frames_enqueue(alb_handle *h, int nframes, int steps_per_frame, int forces_every,
                       const double *controls);
int alb_frames_collect(alb_handle *h, double *series /* nframes x 12, nullable */);

/* ---- tracer particles: initParts/spawn/advect/stepParticles, HTML:721-808 ---
 * Whole-lattice handles only.  Math.random() is replaced by a counter-based
 * generator keyed by (seed, particle, draw), so runs are reproducible. */
#define ALB_MAX_PARTICLES 1000000
/* initParts() with NPART = n (page default 2600, slider 800..5000). */
int alb_particles_init(alb_handle *h, int n, unsigned long long seed);
/* The trail-count slider (HTML:961-967): drop from the end or push spawn(false). */
int alb_particles_resize(alb_handle *h, int n);
/* stepParticles(dt) on the current macroscopic fields; dt in milliseconds
 * (the page passes min(frame time, 40), 16 on the first frame; HTML:903). */
int alb_particles_step(alb_handle *h, double dt_ms);
/* out: n rows of {x, y, life, lane, x0, y0, speed, respawned}: the segment
 * (x0,y0)->(x,y) is what the page strokes, speed is in units of U0. */
int alb_particles_get(alb_handle *h, double *out8, int *n);

/* ---- multi-GPU y-slabs: one-row population halo over NVLink ---------------- */

/* In-process neighbours (several slabs driven by one process). lo = the slab
 * below (smaller y), hi = the slab above; NULL at the lattice edge. */
int alb_connect_local(alb_handle *h, alb_handle *lo, alb_handle *hi);
/* One process driving several GPUs: ndev slabs with (nearly) equal row counts on
 * the given devices, already connected (peer access is enabled as needed);
 * out_slabs receives ndev handles, bottom slab first.  Devices may repeat
 * (several slabs on one GPU).  Geometry calls (alb_rasterize, alb_set_mask) go
 * to every slab; alb_step_multi steps them all; partial diagnostics
 * (alb_forces_partial, alb_stats_partial, alb_get_me_history) add up. */
int alb_create_multi(int nx, int ny, const int *devices, int ndev, alb_handle **out_slabs);
int alb_step_multi(alb_handle **slabs, int nslabs, int nsteps);
/* Destroy connected slabs together: every slab's work is drained first (a neighbour may
 * still be storing halo rows and flags into a slab's memory), then all are freed.
 * Destroying ONE slab of a connected set with alb_destroy is only safe when its
 * neighbours are idle and are never stepped again. */
int alb_destroy_multi(alb_handle **slabs, int nslabs);
/* Cross-process neighbours through CUDA IPC: export a blob, exchange it by any
 * means (torch.distributed all_gather in the Python package), connect. */
int alb_ipc_export(alb_handle *h, void *blob /* ALB_IPC_BYTES */);
int alb_ipc_connect(alb_handle *h, const void *lo_blob, const void *hi_blob);
/* Call on every slab after connect and after any alb_reset/alb_set_populations
 * so the ghost rows hold the neighbours' edge rows before the first step. */
int alb_halo_prime(alb_handle *h);
/* Device pointers for an external halo transport (NCCL send/recv): the three
 * populations that cross each face, each nx floats, in the CURRENT state.
 * send_lo = {f4,f7,f8} of my bottom row, send_hi = {f2,f5,f6} of my top row,
 * recv_lo = ghost row below (f2,f5,f6), recv_hi = ghost row above (f4,f7,f8). */
int alb_halo_ptrs(alb_handle *h, void **send_lo3, void **send_hi3,
                  void **recv_lo3, void **recv_hi3);
/* When 1, alb_step() neither pushes nor waits for halos: the caller moves them
 * between single steps with alb_halo_ptrs(). */
int alb_set_external_halo(alb_handle *h, int on);


/* ---- execution strategy (no effect on results) ------------------------------ */

/* Two LBM steps per pass over HBM (temporal blocking): the
 * deep interior of the lattice is advanced by a fused two-step kernel, everything
 * near borders, the body and slab edges by two list-driven single-step passes.
 * Bit-identical to single steps.  mode: -1 automatic (lattices at least 1024 wide
 * with 1.9 million cells or more),
 * 0 never, 1 whenever a batch has three or more steps left.  All slabs of one
 * lattice must use the same mode.  The environment variable AEROLAB_LBM_DOUBLE
 * (0/1) sets the initial mode of new handles. */
int alb_set_double_steps(alb_handle *h, int mode);
/* How the kernels evaluate the shader's division by tau (HTML:355, `fin - (fin - feq) / tau`).
 * 0: a three-instruction sequence (multiply by RN(1/tau), exact residual, one correction) that the
 *    library has just compared on the device with IEEE division for every fp32 operand of magnitude
 *    [2^-40, 2^8) -- it is used only for a tau that passes with zero mismatches;
 * 1: IEEE division (any other tau; or forced).  The result is bit-identical either way.
 * alb_set_div_mode(h, 1) forces IEEE division, alb_set_div_mode(h, -1) returns to automatic. */
int alb_get_div_mode(const alb_handle *h, int *mode);
int alb_set_div_mode(alb_handle *h, int mode);
/* mode as set; active = 1 when step batches of this handle use double steps. */
int alb_get_double_steps(const alb_handle *h, int *mode, int *active);
/* Host-only (no device needed): the tiling the fused two-step kernel would use for
 * an nx-wide slab of ny_local rows on a GPU with nsm SMs.  out5 = {column segments
 * (over columns [128, pitch-128): the first and last 128-cell task of a row hold the
 * inlet / outlet and are never deep), output columns per segment, rows per segment,
 * units (= column segments x row segments; one warp each), warps per CTA (one CTA per SM)}.
 * A unit reads 128 columns and writes the middle 120. */
int alb_debug_step2_plan(int nx, int ny_local, int nsm, int *out5);
/* Number of CUDA kernels the step batches of this handle have launched so far
 * (alb_step / alb_run_frames; kernels inside replayed CUDA graphs included). */
int alb_launch_count(const alb_handle *h, long long *launches);

/* Self-test of the shared-reciprocal division the fused kernels use for
 * u = j / rho: runs it on about `pairs` generated operand
 * triples (lattice-like values, the whole accepted exponent range and beyond,
 * quotients next to rounding boundaries, zeros/inf/NaN) and compares every
 * accepted result with IEEE division.  out3 = {checked, accepted, wrong};
 * wrong must be 0. */
int alb_selftest_division(alb_handle *h, unsigned long long seed, long long pairs,
                          unsigned long long *out3);

#ifdef __cplusplus
}
#endif
#endif /* AEROLAB_LBM_H */

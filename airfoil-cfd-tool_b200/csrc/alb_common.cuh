// Shared definitions of libaerolab_lbm.so (B200 / sm_100a).
//
// Data layout in HBM:
//   populations  f[2][9][nrows][pitch] fp32, SoA, ping-pong pair
//   nrows = ny_local + 2 (one ghost row below and above: the slab halo; on a
//   whole-lattice handle the ghost rows are never read), pitch = nx rounded up
//   to 128 cells so that a warp task (128 consecutive cells of one row, four
//   per lane, one 128-bit access per lane and population) never straddles rows.
//   mask   u8  [nrows][pitch]   0 / 255, ghost rows included
//   info   u16 [nrows][pitch]   per-cell type + "pull source is solid" bits
//   tclass u8  [nrows][pitch/128] per-warp-task class (all fluid / general /
//                               all solid / all equilibrium)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/aerolab_lbm.h"

// Tuning knob shared by the classifier (alb_geometry.cu) and the step kernel (alb_step.cu):
// 1 = tasks that are all-fluid except the inlet cell x = 0 / the outlet cell x = nx-1 stay in the
// fast kernel.  Measured on B200: +0.5 % at 2048x1024 but -4 % at 32768x16384 (the extra live
// state spills in the hot path) -> off.
#ifndef ALB_EDGE_IN_FAST
#define ALB_EDGE_IN_FAST 0
#endif

namespace alb {

constexpr int TASK_CELLS = 128;          // cells per warp task
constexpr int BLOCK_THREADS = 256;       // 8 warp tasks per CTA
constexpr int TASKS_PER_BLOCK = BLOCK_THREADS / 32;

// world window, HTML:73
constexpr double DX0 = -0.42, DX1 = 1.42, DY0 = -0.46, DY1 = 0.46;

// cell types (info >> 8); priority solid > outlet > equilibrium > interior
// (HTML:287, 301, 314, 324)
enum : int { CT_FLUID = 0, CT_SOLID = 1, CT_OUTLET = 2, CT_EQUIL = 3 };
// warp-task classes
// TC_FLUID_L / TC_FLUID_R: all-fluid except the inlet cell x = 0 (first cell) / the outlet cell
// x = nx-1 (last cell, nx a multiple of 128) -- still handled by the fast kernel
enum : int { TC_FLUID = 0, TC_GENERAL = 1, TC_SOLID = 2, TC_EQUIL = 3, TC_FLUID_L = 4, TC_FLUID_R = 5 };

// info word layout
constexpr unsigned INFO_TYPE_SHIFT = 8, INFO_TYPE_MASK = 3, INFO_PAD = 0x400;

// Accumulators of one fused macroscopic/diagnostics pass (step_kernel<MODE_MACRO,*>): everything is
// a max/min or an integer sum, so the result does not depend on the order of the atomics.
// The device keeps DIAG_SLOTS copies (warps pick one by their index) so that the pre-check loads
// and atomics of millions of warps do not pile up on one L2 line; the copies are merged afterwards.
constexpr int DIAG_SLOTS = 256;
struct alignas(64) DiagAcc {
    unsigned long long smax_bits;    // double bits of max s = hypot(ux/U0, uy/U0) over cells with s < 4 (0: none)
    unsigned long long m2max_bits;   // double bits of ux^2 + uy^2 of that cell (pre-filter for the atomics)
    float rho_min, rho_max;          // over non-solid cells whose Cp lies in (-4, 1.2); +inf / -inf: none
    long long fx, fy;                // sum over fluid/solid faces of rho (2^-40 fixed point) * (solid - fluid)
    unsigned long long surf, rev;    // number of faces, faces whose fluid cell has ux < 0
};

// Momentum-exchange bookkeeping (device).  The step that reads population buffer `parity`
// accumulates into acc[parity]; the NEXT step's first thread commits acc[parity] to
// ring[count % ME_RING] and clears it.  All addresses a kernel touches depend only on the parity,
// so a pair of steps is a static CUDA graph.
constexpr int ME_RING = ALB_ME_HISTORY + 1;
struct MeState {
    long long acc[2][2];
    long long count;          // step number the next commit goes to
    int pending;              // 1: acc[parity of the last step] is not yet in the ring
    int pad;
    long long ring[ME_RING][2];
};

struct StepParams {
    const float *__restrict__ src;
    float *__restrict__ dst;
    const uint16_t *__restrict__ info;
    const uint8_t *__restrict__ tclass;
    const int *__restrict__ gen_list;   // compacted indices of TC_GENERAL tasks (row-major task ids)
    int ngen;
    size_t plane;            // floats per population plane = nrows * pitch
    int pitch;               // cells per row (multiple of 128)
    int tpr;                 // warp tasks per row = pitch / 128
    int ntasks;              // ny_local * tpr
    int nyl;                 // rows owned by this slab
    int nx;
    float tau, inv_tau;      // inv_tau = RN(1/tau)
    int div_mode;            // how x / tau is evaluated: DM_FAST3 (verified for this tau) or DM_IEEE, see div_by_tau()
    float u0;
    float feq0[9];           // feq_i(1, U0, 0) in fp32, source order (HTML:315-317)
    // macro output (macro mode only)
    float *rho, *ux, *uy;
    int write_macro;         // 0: diagnostics only, do not store rho/ux/uy
    DiagAcc *diag;           // nullable: fused autoscale statistics + pressure-face force
    float rho_lo, rho_hi;    // Cp window (-4, 1.2) expressed as a closed rho interval
    double U0d;              // the double U0 of the JS host code
    double m2_lo, m2_hi;     // (4 U0)^2 (1 -/+ 1e-9): below -> s < 4 for sure, above -> s >= 4 for sure
    float m2f_cap;           // fp32 ux^2+uy^2 above this is s >= 4 for sure (m2_hi with fp32 slack)
    // momentum exchange (see MeState); parity = index of the source buffer
    MeState *me;
    int parity;
    unsigned long long *clamp_hits;
    // halo push into the neighbours' ghost rows (nullable)
    float *peer_lo_dst;      // base of the lower neighbour's DESTINATION buffer
    size_t peer_lo_plane;
    size_t peer_lo_row;      // float offset of its upper ghost row
    float *peer_hi_dst;
    size_t peer_hi_plane;
    size_t peer_hi_row;      // float offset of its lower ghost row (row 0)
};

// Task flags of the two-steps-per-pass path (see march2_kernel in alb_march.cu), one byte per warp
// task [nrows][tpr]:
//   TF_DEEP  every cell of the task and every cell within one cell of it is a plain interior fluid
//            cell (type fluid, no solid pull source) and the row is not a slab edge row: the fused
//            kernel may write its state two steps ahead
//   TF_NEED  some task of the 3x3 task neighbourhood is deep: the fused kernel needs the
//            intermediate (one step ahead) state of this task
constexpr unsigned TF_DEEP = 1, TF_NEED = 2;
// list entries of the two-pass path: bit 30 = "clamp hits of this task are counted elsewhere"
constexpr int LIST_NOHIT = 1 << 30, LIST_ID_MASK = LIST_NOHIT - 1;

struct Step2Params {
    const float *__restrict__ src;
    float *__restrict__ dst;
    const uint8_t *__restrict__ tflags;
    size_t plane;
    int pitch, tpr, nyl;
    int wo;                  // output columns per strip (multiple of 4, <= 128*K - 8)
    int hs;                  // output rows per segment
    int nstrips, ntiles;
    // march2_kernel (alb_march.cu): column segments per row, units = nseg * row segments, and the work
    // queue {next unit, warps that have finished}, both zero between launches
    int nseg, nunits;
    int quota;               // units a warp may take before it retires
    int edges, nx;           // the first / last task of a row can be deep (see build_deep_kernel); lattice width
    int *queue;
    float tau, inv_tau;
    float u0, feq0[9];       // inlet state (HTML:314-322), edges only
    int div_mode;            // as in StepParams
    unsigned long long *clamp_hits;
    // fused statistics of the state being written (nullable), same meaning as in StepParams
    DiagAcc *diag;
    float rho_lo, rho_hi;
    double U0d;
    double m2_lo, m2_hi;
    float m2f_cap;
};

struct Handle;

// Device-side mirror of the page's sticky host state (autoscale values HTML:593, force EMAs
// HTML:641), so that whole frames (HTML:902-930) can run without host synchronisation.
struct FrameDev {
    double maxS, cpMin, cpMax;
    double cl_smooth, cd_smooth, sep_frac;
    int ema_valid, pad;
};
constexpr int FRAME_ROW = 12;   // doubles per frame record, see alb_run_frames() in the header
cudaError_t launch_frame_finalize(DiagAcc *d, DiagAcc *published, const MeState *me, int me_parity, FrameDev *st,
                                  int do_forces, double U0, double q, double *row, cudaStream_t s);

// alb_particles.cu -- one tracer (HTML:730-736): position, remaining life, home lane; plus the
// segment of the last step (x0,y0 -> x,y) and its speed for whoever draws the trails
struct ParticleState {
    double x, y, life, lane;
    double x0, y0, speed;
    int respawned, pad;
};
cudaError_t launch_particles_init(ParticleState *ps, unsigned *ctrs, int first, int n, int npart_total,
                                  unsigned long long seed, int slider_push, cudaStream_t s);
cudaError_t launch_particles_step(ParticleState *ps, unsigned *ctrs, int n, unsigned long long seed, double dt,
                                  const uint8_t *mask, const float *ux, const float *uy, int pitch, int nx, int ny,
                                  double U0, cudaStream_t s);

// every file: load its kernels' device code now (called once per device from alb_create_slab)
cudaError_t preload_step_kernels();
cudaError_t preload_march_kernels();
cudaError_t preload_step2_kernels();
cudaError_t preload_diag_kernels();
cudaError_t preload_geometry_kernels();
cudaError_t preload_particle_kernels();

// alb_step.cu -- one step per pass
cudaError_t launch_step_fast(const StepParams &p, cudaStream_t s);
cudaError_t launch_step_general(const StepParams &p, cudaStream_t s);
cudaError_t launch_step_unified(const StepParams &p, cudaStream_t s);
cudaError_t launch_macro(const StepParams &p, cudaStream_t s);
cudaError_t launch_step_fast_list(const StepParams &p, cudaStream_t s);   // p.gen_list/p.ngen = the list
int small_lattice_capacity(int device);
cudaError_t launch_small_lattice(const StepParams &p, float *f0, float *f1, int cur, int nsteps, cudaStream_t s);
int band_lattice_rows(int nx, int nyl, int nsm);            // cells per strip; 0: the lattice does not qualify
size_t band_inbox_bytes(int nx, int nyl, int R);
cudaError_t launch_band_lattice(const StepParams &p, float *f0, float *f1, int cur, int nsteps, int R, void *inbox,
                                long long step_base, int *err, cudaStream_t s);
void host_feq0(float u0, float *out9);

// alb_step2.cu -- helpers of the two-steps-per-pass path
cudaError_t launch_copy_tasks(const StepParams &p, cudaStream_t s);       // dst = src on the listed tasks
// alb_march.cu -- two steps per pass, one independent warp per unit (the default two-step kernel)
void march_plan(Step2Params &p, int nsm);
int march_out_width();     // output columns per warp
int march_warps_per_cta();
cudaError_t launch_march2(const Step2Params &p, int nsm, cudaStream_t s);
cudaError_t launch_div_selftest(unsigned long long seed, int nblocks, int iters, unsigned long long *d_out3, cudaStream_t s);
// mismatches of the three-instruction x / tau against IEEE division over all fp32 x with 2^-40 <= |x| < 2^8
cudaError_t launch_divtau_check(float tau, float rcp, unsigned long long *d_out1, cudaStream_t s);

// alb_geometry.cu
void host_rotate_panelise(const double *xy, int npts, double alpha_deg, double *xp, double *yp);
cudaError_t launch_raster(const double *d_xp, const double *d_yp, int n, uint8_t *mask, int pitch,
                          int nx, int ny_global, int gy_first, int nrows, cudaStream_t s);
cudaError_t launch_build_info(const uint8_t *mask, uint16_t *info, uint8_t *tclass, int *gen_list,
                              int *gen_count, int pitch, int nx, int ny_global, int gy_first, int nrows,
                              cudaStream_t s);

// lists[5] = pass-1 fast, pass-1 general, pass-2 fast, pass-2 general, all-solid (copied, not stepped);
// counts = int[5] (device)
bool march_edges_enabled(int nx, int pitch);
cudaError_t launch_build_lists(const uint16_t *info, const uint8_t *tclass, uint8_t *deep_tmp, uint8_t *tflags,
                               int *const lists[5], int *counts, int pitch, int nrows, int lo_nb, int hi_nb,
                               int edges_nx, cudaStream_t s);

// alb_diag.cu
struct DiagScratch {
    double *d_part = nullptr;     // device partials
    double *h_part = nullptr;     // pinned host mirror
    int nblocks = 0;
};
cudaError_t launch_stats(const uint8_t *mask, const float *rho, const float *ux, const float *uy,
                         int pitch, int nx, int nyl, double u0, float *U, float *V, float *Cp,
                         double *d_part, int nblocks, cudaStream_t s);
cudaError_t launch_forces(const uint8_t *mask, const float *rho, const float *ux, int pitch, int nx,
                          int ny_global, int gy_first, int nyl, double *d_part, int nblocks,
                          cudaStream_t s);
cudaError_t launch_render(const uint8_t *mask, const float *rho, const float *ux, const float *uy,
                          int pitch, int nx, int ny, int lo_ghost, int hi_ghost, int mode, float u0, float maxS,
                          float cpMin, float cpMax, float vortScale, float *t_out, uint8_t *rgba, cudaStream_t s);
cudaError_t launch_mass(const float *f, size_t plane, int pitch, int nx, int nyl, double *d_part,
                        int nblocks, cudaStream_t s);
cudaError_t launch_state_hash(const float *f, size_t plane, int pitch, int nx, int nyl, int gy0,
                              unsigned long long *d_out9, cudaStream_t s);
cudaError_t launch_fill_init(float *f0, float *f1, size_t plane, const float *feq9, float *rho,
                             float *ux, float *uy, float u0, cudaStream_t s);

}  // namespace alb

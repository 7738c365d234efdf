// Two LBM steps per pass over HBM, second generation: march2_kernel.
//
// step2_kernel (round 1) split a CTA into step-1 and step-2 warps that met at a __syncthreads per
// row group; ncu showed a quarter of its issue slots lost at that barrier, and its one shape
// (632-column strips) ruled out lattices narrower than ~8000 columns.  Here every WARP is an
// independent unit of work and there is no barrier of any kind:
//
//   * a warp owns a column segment of 128 cells (4 per lane) and marches up a range of rows;
//   * per row it runs step 1 on the 128 cells.  The nine input vectors were copied into the warp's
//     private shared-memory double buffer two rows ahead with cp.async (every lane copies and later
//     reads its own 16 bytes, so completion is a per-thread cp.async.wait_group: no mbarrier, no
//     __syncwarp).  The part of the intermediate state that the NEXT row's step 2 needs (f0,f1,f3)
//     stays in registers; f2,f5,f6, needed two rows later, go through a private shared-memory slot
//     (again own bytes only);
//   * it then runs step 2 of the row below, whose remaining inputs (f4,f7,f8 of the row just
//     computed) are in registers already, and stores the result to HBM with 128-bit stores;
//   * step 2 of a cell needs step 1 of its x-neighbours, so only lanes 1..30 produce output (120
//     columns per warp, segments overlap by 8 columns: 6.7 % redundant arithmetic instead of a
//     shared intermediate ring and its barriers);
//   * the inlet column x = 0 and the outlet column x = nx-1 belong to the first and the last column
//     segment (widths that are a multiple of 4, march_edges_enabled()): lane 0 overwrites the inlet
//     cell with the equilibrium constants after either collision; the outlet cell copies its left
//     neighbour's previous state (HTML:301-312), which lane 31 has at hand -- the neighbour's
//     intermediate state is in its own registers, and of the source state only f6, f7 of one cell
//     per row are not staged anyway (two 4-byte cp.async in the same group);
//   * units (row segment x column segment) are handed out through an atomic queue to the persistent
//     warps of one CTA per SM, so there is no wave tail, and any lattice at least 256 cells wide can
//     use it (2048- and 4096-wide lattices included).
//
// Measured and rejected (profiles/r2a_*): staging with TMA bulk copies (nine 512-byte
// cp.async.bulk per row and warp on the warp's own mbarrier).  UBLKCP takes its operands from
// uniform registers, so per-lane issue turns into a 9-trip waterfall loop; 28 % of the kernel's
// stall samples sat in that loop and the try_wait spin, and it ran no faster than the round-1
// kernel (134.6 vs 131.6 GLUPS at configs[3]).
//
// Which cells it may write (deep tasks) and which intermediate values it needs (TF_NEED) comes
// from the same task flags as before; everything else is advanced by the list-driven two-pass
// path on the aux stream (alb_api.cu, issue_double).  Same collide_quad() arithmetic as
// everywhere else -> bit-identical to two single steps (HTML:283-360 twice).
#include <stdio.h>
#include <stdlib.h>

#include "alb_lbm.cuh"

// Shape knobs: warps per CTA (one CTA per SM), the tallest row segment the planner may choose, and
// the fixed cost of starting a unit (pipeline fill) in row-steps.  The register file is split over
// the four schedulers (16 K registers each), so a CTA of 16 warps gets 128 registers per thread and
// one of 12 warps 168.  At 128 the plain kernel spills 8 bytes and the DIAG variant 52, whose
// reloads (local memory, long scoreboard) cost it 22 %; at 12 warps neither spills and the plain
// kernel is 3 % faster as well (6.61 vs 6.83 ms per pass at configs[3], profiles/r2f vs r2c).  Short segments win although they recompute two rows each: neighbouring column
// segments re-read each other's edge columns, and the less two warps can drift apart the more of
// those re-reads hit L2 (measured at configs[3]: 24 rows 144.1, 32 rows 144.8, 48 rows 142.6, 64 rows
// 140.8, 96 rows 138.3, 171 rows 135.5 GLUPS).
#ifndef ALB_MARCH_WARPS
#define ALB_MARCH_WARPS 12
#endif
#ifndef ALB_MARCH_WARPS_DIAG
#define ALB_MARCH_WARPS_DIAG 12
#endif
#ifndef ALB_MARCH_HS_MAX
#define ALB_MARCH_HS_MAX 32
#endif
#ifndef ALB_MARCH_UNIT_OVERHEAD
#define ALB_MARCH_UNIT_OVERHEAD 3
#endif
// how many CTAs an SM works through per pass (1: persistent CTAs that hold their SM until the end).
// Every retirement is a chance for the list-driven passes on the aux stream to get an SM.  Measured
// (profiles/r2_scaling_notes.md): configs[3] on one GPU 126.6 / 141.2 / 144.9 / 142.8 / 143.5 GLUPS with
// 1 / 6 / 12 / 24 / 48 generations; two GPUs, 32768 x 2048 rows each (weak): 259.6 with 6, 288.7 with 24
// (one GPU: 145.4) -- with few generations the passes, and through their flags the neighbouring GPU,
// wait 0.2 ms of every 1.0 ms double step for the first CTAs to retire.
#ifndef ALB_MARCH_GENERATIONS
#define ALB_MARCH_GENERATIONS 24
#endif
// L2 eviction hints on the staged loads.  Neighbouring column segments share 8 columns (plus the
// rest of the 64-byte fetch granule), and the two warps reach a given row at different times: the
// shared bytes have to survive in L2 until the second warp arrives or they are fetched from HBM
// twice (ncu, profiles/r2b: 23.1 GB read per pass for 19.3 GB of state).
//   0 no hints   1 first/last two lanes of a warp evict_last, the rest evict_first   2 only the evict_last part
#ifndef ALB_MARCH_L2HINT
#define ALB_MARCH_L2HINT 0
#endif
// registers per thread the kernel is compiled for (0: whatever the launch bounds allow)
#ifndef ALB_MARCH_MAXNREG
#define ALB_MARCH_MAXNREG 0
#endif

namespace alb {

namespace {

constexpr int M_WARPS = ALB_MARCH_WARPS;      // warps per CTA, one CTA per SM
constexpr int M_OUT = 120;                    // output columns per warp (lanes 1..30)
constexpr int M_STAGE = 9 * 128;              // floats of one staged step-1 row (9 planes x 128 columns)
constexpr int M_CARRY = 3 * 128;              // floats of one carried row (f2, f5, f6 of an intermediate row)
constexpr int M_DSLOT = 4 * 2 * 32;             // floats: candidate (ux, uy) of each of the four cells of a quad, per lane
constexpr int M_DCOLD = 4 * 32;                 // floats: per lane the best candidate's exact |u|^2 (double) and fp32 |u|^2
constexpr int M_EDGE = 32;                      // floats: outlet bookkeeping of lane 31 (two staged scalars and nine carried values per row parity)
constexpr int M_WARP_SMEM = 2 * M_STAGE + 2 * M_CARRY + M_DSLOT + M_DCOLD + M_EDGE;   // floats of shared memory per warp

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(unsigned dst, const float *src, [[maybe_unused]] unsigned long long policy) {
#if ALB_MARCH_L2HINT
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(policy) : "memory");
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
#endif
}
__device__ __forceinline__ void cp_async4(unsigned dst, const float *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// x-shifts inside the warp; the outermost cell of lane 0 / lane 31 receives a value that is never used
__device__ __forceinline__ float4 shl(const float4 &v) {   // populations arriving from x-1
    return make_float4(__shfl_up_sync(FULL, v.w, 1), v.x, v.y, v.z);
}
__device__ __forceinline__ float4 shr(const float4 &v) {   // populations arriving from x+1
    return make_float4(v.y, v.z, v.w, __shfl_down_sync(FULL, v.x, 1));
}

// Fused statistics of the state being written (DIAG variant; same result as diag_cell() in
// alb_lbm.cuh).  The per-cell code inside the collision must stay branch-free, call-free and must not
// drag the arg-max bookkeeping through the register allocator (earlier versions: the values kept in
// registers across the collision, then a cold call -- both +26..28 % on the double step that ends a
// frame).  So the hot part is five registers (rho window extrema, the pre-filter level, the current
// candidate's velocity); a cell that may be a new arg-max (rare: within 1e-6 of the level and not the
// candidate itself) only drops its velocity into a private shared-memory slot and sets a bit, and
// the exact fp64 comparison runs after the collision of the quad, outside the hot basic block.
// Warps per CTA and registers per thread of the two variants (see the knobs at the top of the file):
// 3 warps per scheduler at <= 168 registers, no spills.  (14 warps at 144 registers do not launch: a
// scheduler's 16 K registers hold 3 such warps, not 4.)
template <bool DIAG> struct MarchShape {
    static constexpr int warps = DIAG ? ALB_MARCH_WARPS_DIAG : M_WARPS;
    static constexpr int regs = ALB_MARCH_MAXNREG ? ALB_MARCH_MAXNREG : (16384 / (((warps + 3) / 4) * 32)) / 8 * 8;
};

template <bool DIAG, int DM>
__global__ void __maxnreg__(MarchShape<DIAG>::regs)
march2_kernel(const __grid_constant__ Step2Params p) {
    extern __shared__ float4 smem4[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // this lane's own 16 bytes of every row of the warp's private buffers
    float *const stg = reinterpret_cast<float *>(smem4) + (size_t)warp * M_WARP_SMEM + lane * 4;   // [2][9][128]
    float *const car = stg + 2 * M_STAGE;                                                           // [2][3][128]
    // DIAG: this lane's candidate slots [4][2] and cold state {m2 lo, m2 hi, m2f}, stride 32 floats
    [[maybe_unused]] float *const dsl = reinterpret_cast<float *>(smem4) + (size_t)warp * M_WARP_SMEM + 2 * M_STAGE + 2 * M_CARRY + lane;
    [[maybe_unused]] float *const dco = dsl + M_DSLOT;
    // outlet bookkeeping (lane 31 of a unit that holds the last columns of the row): eg[2 k + {0,1}] = f6, f7 of
    // the cell left of the outlet, source state, row of stage k; eg[4 + 9 k + i] = intermediate f_i of that cell
    float *const eg = reinterpret_cast<float *>(smem4) + (size_t)warp * M_WARP_SMEM + 2 * M_STAGE + 2 * M_CARRY + M_DSLOT + M_DCOLD;
    const unsigned stg_u32 = smem_u32(stg);
    const unsigned eg_u32 = smem_u32(eg);
    const size_t plane = p.plane;
    const int pitch = p.pitch, tpr = p.tpr;
    [[maybe_unused]] const float *const src = p.src;
    [[maybe_unused]] float *const dst_base = p.dst;
    unsigned long long policy = 0;
#if ALB_MARCH_L2HINT
    {
        unsigned long long keep, stream;
        asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
#if ALB_MARCH_L2HINT == 1
        asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(stream));
#else
        asm("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(stream));
#endif
        policy = (lane < 2 || lane >= 30) ? keep : stream;
    }
#endif
    const int total_warps = gridDim.x * MarchShape<DIAG>::warps;
    unsigned hits = 0;
    [[maybe_unused]] float d_rmin = INFINITY, d_rmax = -INFINITY, d_thr = -1.0f, d_bux = 0.f, d_buy = 0.f;
    if (DIAG) {
        dco[0] = __int_as_float(__double2loint(-1.0));      // exact |u|^2 of the best candidate so far: none
        dco[32] = __int_as_float(__double2hiint(-1.0));
        dco[64] = -1.0f;                                    // its fp32 |u|^2 (the pre-filter level)
    }

    // units come from the queue; a warp takes at most p.quota of them, so that CTAs retire while the
    // pass is under way and the list-driven passes on the (high-priority) aux stream find SMs
    int unit = 0;
    if (lane == 0) unit = atomicAdd(p.queue, 1);
    unit = __shfl_sync(FULL, unit, 0);
    for (int taken = 1;; taken++) {
        if (unit >= p.nunits) break;
        // fetch the id of the unit after this one now: the atomic's latency hides behind the work
        int next_unit = p.nunits;
        if (lane == 0 && taken < p.quota) next_unit = atomicAdd(p.queue, 1);
        const int rowseg = unit / p.nseg, s = unit - rowseg * p.nseg;
        const int y0 = 2 + rowseg * p.hs;                // owned output rows [y0, y1)
        const int y1 = min(y0 + p.hs, p.nyl);
        // this lane's quad and whether it is one the unit writes.  Without edges: segment s stages columns
        // [124 + 120 s, +128) and writes the middle 120.  With edges (march_plan): segment 0 stages [0, 128) and
        // writes [0, 124) with the inlet cell patched; the last one stages [nx - 128, nx) and writes
        // [nx - 124, nx) with the outlet cell patched; the ones between write 120 columns from 124 + 120 (s-1)
        // on (the last of them clipped where the last segment begins)
        int gx = 124 + M_OUT * s + lane * 4;
        bool own = lane >= 1 && lane <= 30;
        unsigned excl = 0;                               // cells of this quad that are not plain: bit 0 inlet, bit 3 outlet
        if (p.edges) {
            if (s == 0) {
                gx = lane * 4;
                own = lane <= 30;
                excl = lane == 0 ? 1u : 0u;
            } else if (s == p.nseg - 1) {
                gx = p.nx - 128 + lane * 4;
                own = lane >= 1;
                excl = lane == 31 ? 8u : 0u;
            } else {
                gx = M_OUT * s + lane * 4;
                own = own && gx < p.nx - 124;
            }
        }
        const uint8_t *const tfl = p.tflags + (gx >> 7);
        // task flags of rows a, a+1, a+2 (pipelined; rows beyond y1 are of no interest to this unit)
        auto row_flags = [&](int a) -> unsigned { return a <= y1 ? (unsigned)tfl[(size_t)a * tpr] : 0u; };
        // stage k <- the nine input vectors of step 1 of row a (population i comes from row a - e_y(i))
        auto issue = [&](int a, int k) {
            const float *g = p.src + (size_t)a * pitch + gx;
            const unsigned d = stg_u32 + (unsigned)(k * M_STAGE) * 4u;
            const int ey[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
#pragma unroll
            for (int i = 0; i < 9; i++) {
                const float *gi = g + i * plane - (ptrdiff_t)ey[i] * pitch;
                ALB_CHECK_SRC(gi, 4);
                cp_async16(d + i * 512u, gi, policy);
            }
            if (excl & 8u) {
                // the outlet cell's intermediate state is the SOURCE state of its left neighbour (HTML:301-312); the
                // neighbour's step 2 pulls f3, f6, f7 of it: f3 of this row is staged above, f6 and f7 are not
                const float *g6 = p.src + 6 * plane + (size_t)a * pitch + (p.nx - 2);
                cp_async4(eg_u32 + (unsigned)(k * 2 + 0) * 4u, g6);
                cp_async4(eg_u32 + (unsigned)(k * 2 + 1) * 4u, g6 + plane);
            }
        };
        int a = y0 - 1;
        unsigned tf0 = row_flags(a), tf1 = row_flags(a + 1), tf2 = row_flags(a + 2);
        unsigned tfb = 0;                                // flags of the step-2 row a - 1
        bool have = __any_sync(FULL, tf0 & TF_NEED);
        bool have1 = __any_sync(FULL, tf1 & TF_NEED);
        if (have) issue(a, 0);
        cp_async_commit();
        if (have1) issue(a + 1, 1);
        cp_async_commit();
        float4 wn0, wn1, wn3;                            // f0, f1, f3 of intermediate row a-1, already x-shifted
        wn0 = wn1 = wn3 = make_float4(0.f, 0.f, 0.f, 0.f);
        float *d = p.dst + (size_t)(a - 1) * pitch + gx;  // destination of step-2 row a-1
        int k = 0;
#pragma unroll 1
        for (; a <= y1; a++, k ^= 1, d += pitch) {
            const unsigned tf3 = row_flags(a + 3);
            const bool have2 = __any_sync(FULL, tf2 & TF_NEED);
            float4 o[9];
            cp_async_wait<1>();                          // everything but the newest group (row a+1) has landed
            if (have) {
                // ---- step 1 of row a: staged inputs -> intermediate state (registers) ----
                const float *sp = stg + k * M_STAGE;
                const float4 v0 = *reinterpret_cast<const float4 *>(sp + 0 * 128);
                const float4 v1 = *reinterpret_cast<const float4 *>(sp + 1 * 128);
                const float4 v2 = *reinterpret_cast<const float4 *>(sp + 2 * 128);
                const float4 v3 = *reinterpret_cast<const float4 *>(sp + 3 * 128);
                const float4 v4 = *reinterpret_cast<const float4 *>(sp + 4 * 128);
                const float4 v5 = *reinterpret_cast<const float4 *>(sp + 5 * 128);
                const float4 v6 = *reinterpret_cast<const float4 *>(sp + 6 * 128);
                const float4 v7 = *reinterpret_cast<const float4 *>(sp + 7 * 128);
                const float4 v8 = *reinterpret_cast<const float4 *>(sp + 8 * 128);
                o[0] = v0;
                o[1] = shl(v1);
                o[2] = v2;
                o[3] = shr(v3);
                o[4] = v4;
                o[5] = shl(v5);
                o[6] = shr(v6);
                o[7] = shr(v7);
                o[8] = shl(v8);
                const unsigned hm = collide_quad<DM>(o, p.tau, p.inv_tau);
                // clamp hits of step 1: every deep cell is owned by exactly one unit
                if ((tf0 & TF_DEEP) && own && a >= y0 && a < y1) hits += __popc(hm & ~excl);
                if (excl) {
                    if (excl & 1u) {
                        // inlet cell: equilibrium at (1, U0, 0), HTML:314-322
                        o[0].x = p.feq0[0]; o[1].x = p.feq0[1]; o[2].x = p.feq0[2];
                        o[3].x = p.feq0[3]; o[4].x = p.feq0[4]; o[5].x = p.feq0[5];
                        o[6].x = p.feq0[6]; o[7].x = p.feq0[7]; o[8].x = p.feq0[8];
                    } else {
                        // the intermediate state of the cell left of the outlet is what the outlet cell holds after
                        // step 2; its own intermediate state matters only through what that neighbour pulls from it
                        float *const oc = eg + 4 + k * 9;
                        oc[0] = o[0].z; oc[1] = o[1].z; oc[2] = o[2].z;
                        oc[3] = o[3].z; oc[4] = o[4].z; oc[5] = o[5].z;
                        oc[6] = o[6].z; oc[7] = o[7].z; oc[8] = o[8].z;
                        o[3].w = sp[3 * 128 + 2];
                        o[6].w = eg[k * 2 + 0];
                        o[7].w = eg[k * 2 + 1];
                    }
                }
                // what later rows pull from this one, x-shifts applied now
                o[1] = shl(o[1]);
                o[3] = shr(o[3]);
                o[5] = shl(o[5]);
                o[6] = shr(o[6]);
                o[7] = shr(o[7]);
                o[8] = shl(o[8]);
            }
            // every staged value of row a has been consumed by the arithmetic above: refill the buffer
            if (have2) issue(a + 2, k);
            cp_async_commit();                           // one group per row, empty or not
            // ---- step 2 of row a-1: f0,f1,f3 of its own row, f2,f5,f6 of row a-2, f4,f7,f8 of row a ----
            const bool st = (tfb & TF_DEEP) && own;
            const bool any_st = __any_sync(FULL, st);
            float *const cs = car + k * M_CARRY;         // holds f2,f5,f6 of row a-2; receives those of row a
            float4 q[9];
            if (any_st) {
                q[2] = *reinterpret_cast<const float4 *>(cs + 0 * 128);
                q[5] = *reinterpret_cast<const float4 *>(cs + 1 * 128);
                q[6] = *reinterpret_cast<const float4 *>(cs + 2 * 128);
            }
            // this row's f2,f5,f6 go into the slot just read BEFORE step 2 runs: twelve registers fewer
            // are live across the second collision
            if (have) {
                *reinterpret_cast<float4 *>(cs + 0 * 128) = o[2];
                *reinterpret_cast<float4 *>(cs + 1 * 128) = o[5];
                *reinterpret_cast<float4 *>(cs + 2 * 128) = o[6];
            }
            if (any_st) {
                q[0] = wn0; q[1] = wn1; q[3] = wn3;
                q[4] = o[4]; q[7] = o[7]; q[8] = o[8];
                unsigned hm;
                if (DIAG) {
                    // the statistics of the state being written ride along: rho, ux, uy and |u|^2 are consumed
                    // where the collision has them (deep cells have no faces)
                    unsigned cand = 0;
                    hm = collide_quad<DM, ALB_QUAD_GB>(q, p.tau, p.inv_tau, [&](int c, float rho, float ux, float uy, float uu) {
                        const bool stc = st && !((excl >> c) & 1u);
                        const bool in_window = stc && rho >= p.rho_lo && rho <= p.rho_hi;
                        d_rmin = fminf(d_rmin, in_window ? rho : INFINITY);
                        d_rmax = fmaxf(d_rmax, in_window ? rho : -INFINITY);
                        // uu >= level also drops NaN; above the cap s >= 4 for sure; a velocity that IS the
                        // candidate (a uniform free stream is bitwise uniform) cannot change the maximum
                        if (stc && uu >= d_thr && uu <= p.m2f_cap && !(ux == d_bux && uy == d_buy)) {
                            dsl[(c * 2 + 0) * 32] = ux;
                            dsl[(c * 2 + 1) * 32] = uy;
                            cand |= 1u << c;
                        }
                    });
                    if (excl && st) {
                        // the inlet / outlet cell of this quad: its macroscopic values do not come from a collision;
                        // it goes through the exact comparison below unconditionally
                        float rho = 1.0f, ux = p.u0, uy = 0.0f;
                        const int c = (excl & 1u) ? 0 : 3;
                        if (excl & 8u) {
                            const float *const oc = eg + 4 + (k ^ 1) * 9;
                            const float f[9] = {oc[0], oc[1], oc[2], oc[3], oc[4], oc[5], oc[6], oc[7], oc[8]};
                            moments_plain(f, rho, ux, uy);
                        }
                        const bool in_window = rho >= p.rho_lo && rho <= p.rho_hi;
                        d_rmin = fminf(d_rmin, in_window ? rho : INFINITY);
                        d_rmax = fmaxf(d_rmax, in_window ? rho : -INFINITY);
                        dsl[(c * 2 + 0) * 32] = ux;
                        dsl[(c * 2 + 1) * 32] = uy;
                        cand |= 1u << c;
                    }
                    if (cand) {
                        // exact comparison of the flagged cells, in cell order (diag_cell() of alb_lbm.cuh)
                        double best = __hiloint2double(__float_as_int(dco[32]), __float_as_int(dco[0]));
                        float best_f = dco[64];
#pragma unroll
                        for (int c = 0; c < 4; c++) {
                            if (!(cand & (1u << c))) continue;
                            const float ux = dsl[(c * 2 + 0) * 32], uy = dsl[(c * 2 + 1) * 32];
                            const double m2 = __dadd_rn(__dmul_rn((double)ux, (double)ux), __dmul_rn((double)uy, (double)uy));
                            if (m2 > best && m2 < p.m2_hi) {
                                if (m2 >= p.m2_lo && !(speed_ratio(ux, uy, p.U0d) < 4.0)) continue;
                                best_f = fmaxf(best_f, ux * ux + uy * uy);   // only ACCEPTED cells raise the level
                                best = m2;
                                d_bux = ux;
                                d_buy = uy;
                            }
                        }
                        dco[0] = __int_as_float(__double2loint(best));
                        dco[32] = __int_as_float(__double2hiint(best));
                        dco[64] = best_f;
                        d_thr = best_f * (1.0f - 1e-6f);
                    }
                } else {
                    hm = collide_quad<DM, ALB_QUAD_GB>(q, p.tau, p.inv_tau);
                }
                if (st) {
                    hits += __popc(hm & ~excl);
                    if (excl & 1u) {
                        q[0].x = p.feq0[0]; q[1].x = p.feq0[1]; q[2].x = p.feq0[2];
                        q[3].x = p.feq0[3]; q[4].x = p.feq0[4]; q[5].x = p.feq0[5];
                        q[6].x = p.feq0[6]; q[7].x = p.feq0[7]; q[8].x = p.feq0[8];
                    } else if (excl) {
                        const float *const oc = eg + 4 + (k ^ 1) * 9;
                        q[0].w = oc[0]; q[1].w = oc[1]; q[2].w = oc[2];
                        q[3].w = oc[3]; q[4].w = oc[4]; q[5].w = oc[5];
                        q[6].w = oc[6]; q[7].w = oc[7]; q[8].w = oc[8];
                    }
#pragma unroll
                    for (int i = 0; i < 9; i++) ST4(d + i * plane, q[i]);
                }
            }
            // carry: this row's f0,f1,f3 stay in registers for the next iteration
            if (have) {
                wn0 = o[0]; wn1 = o[1]; wn3 = o[3];
            }
            tfb = (a >= y0 && a < y1) ? tf0 : 0u;
            tf0 = tf1; tf1 = tf2; tf2 = tf3;
            have = have1; have1 = have2;
        }
        cp_async_wait<0>();
        unit = __shfl_sync(FULL, next_unit, 0);
    }
    if (DIAG) {
        DiagLocal dl;
        dl.rmin = d_rmin; dl.rmax = d_rmax;
        dl.m2 = __hiloint2double(__float_as_int(dco[32]), __float_as_int(dco[0]));
        dl.m2f = dco[64]; dl.bux = d_bux; dl.buy = d_buy;
        diag_flush<false>(p, dl, lane);
    }
    if (hits && p.clamp_hits) atomicAdd(p.clamp_hits, (unsigned long long)hits);
    // the last warp to leave re-arms the queue for the next launch (stream order does the rest)
    if (lane == 0) {
        __threadfence();
        const int left = atomicAdd(p.queue + 1, 1);
        if (left == total_warps - 1) {
            p.queue[0] = 0;
            p.queue[1] = 0;
        }
    }
}

template <bool DIAG> constexpr size_t MARCH_SMEM = sizeof(float) * (size_t)MarchShape<DIAG>::warps * M_WARP_SMEM;

}  // namespace

// Geometry of the marching kernel for a pitch x nyl slab on nsm SMs: column segments of 120 output
// columns over the columns that can be deep, [128, pitch - 128), and row segments of hs rows over
// rows 2 .. nyl-1.  A unit costs about 2 hs + 2 row-steps (hs + 2 rows of step 1, hs rows of step
// 2) plus a fixed start-up; the persistent warps take units from a queue, so the pass lasts about
// ceil(units / warps) units -- pick the segment height that minimises it.
void march_plan(Step2Params &p, int nsm) {
    if (p.edges)   // [0, 124) | 120 columns each from 124 on | [nx - 124, nx)
        p.nseg = 2 + (p.nx > 248 ? (p.nx - 248 + M_OUT - 1) / M_OUT : 0);
    else
        p.nseg = p.pitch >= 3 * TASK_CELLS ? (p.pitch - 2 * TASK_CELLS + M_OUT - 1) / M_OUT : 0;
    const int rows = p.nyl - 2;
    p.nstrips = p.nseg;
    p.wo = M_OUT;
    if (rows <= 0 || p.nseg == 0) {
        p.hs = 1;
        p.nunits = p.ntiles = 0;
        return;
    }
    static int hs_env = -1;
    if (hs_env < 0) {
        const char *e = getenv("AEROLAB_LBM_S2_HS");
        hs_env = e ? atoi(e) : 0;
    }
    const long long warps = (long long)nsm * M_WARPS;
    int best = rows;
    if (hs_env > 0) {
        best = hs_env < rows ? hs_env : rows;
    } else {
        long long best_cost = -1;
        const int hs_max = rows < ALB_MARCH_HS_MAX ? rows : ALB_MARCH_HS_MAX;
        for (int hs = hs_max; hs >= 1; hs--) {
            const long long units = (long long)p.nseg * ((rows + hs - 1) / hs);
            const long long cost = ((units + warps - 1) / warps) * (2 * hs + 2 + ALB_MARCH_UNIT_OVERHEAD);
            if (best_cost < 0 || cost < best_cost) {
                best_cost = cost;
                best = hs;
            }
        }
    }
    p.hs = best;
    p.nunits = p.ntiles = p.nseg * ((rows + best - 1) / best);
    // CTA generations: about ALB_MARCH_GENERATIONS retirements per SM and pass
    static int gen_env = -1;
    if (gen_env < 0) {
        const char *e = getenv("AEROLAB_LBM_S2_GENERATIONS");
        gen_env = e ? atoi(e) : ALB_MARCH_GENERATIONS;
        if (gen_env < 1) gen_env = 1;
    }
    const long long per_warp = (p.nunits + warps - 1) / warps;
    p.quota = (int)((per_warp + gen_env - 1) / gen_env);
    if (p.quota < 1) p.quota = 1;
}

// The inlet and outlet columns can be part of the fused kernel's domain when the last segment, which
// ends with the outlet cell, starts on a 16-byte boundary (nx a multiple of 4) and does not overlap
// the first one.  All slabs of a lattice decide alike: the rule looks at the global width and the
// environment only.
bool march_edges_enabled(int nx, int pitch) {
    static int env = -1;
    if (env < 0) {
        const char *e = getenv("AEROLAB_LBM_MARCH_EDGES");
        env = e ? atoi(e) : 1;
    }
    (void)pitch;
    return env != 0 && nx % 4 == 0 && nx >= 2 * TASK_CELLS;
}

int march_out_width() { return M_OUT; }
int march_warps_per_cta() { return M_WARPS; }

template <bool DIAG, int DM>
cudaError_t launch_march2_t(const Step2Params &p, cudaStream_t s) {
    // one CTA per SM at a time; every warp takes at most p.quota units from the queue, so the grid
    // must offer at least nunits / quota warps
    const long long need_warps = ((long long)p.nunits + p.quota - 1) / p.quota;
    const int grid = (int)((need_warps + MarchShape<DIAG>::warps - 1) / MarchShape<DIAG>::warps);
    static bool configured[64] = {};       // the attribute is per device (and per instantiation)
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(march2_kernel<DIAG, DM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MARCH_SMEM<DIAG>);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    march2_kernel<DIAG, DM><<<grid, MarchShape<DIAG>::warps * 32, MARCH_SMEM<DIAG>, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_march2(const Step2Params &p, int nsm, cudaStream_t s) {
    if (p.nunits <= 0) return cudaSuccess;
    (void)nsm;
    if (p.div_mode == DM_FAST3) return p.diag ? launch_march2_t<true, DM_FAST3>(p, s) : launch_march2_t<false, DM_FAST3>(p, s);
    return p.diag ? launch_march2_t<true, DM_IEEE>(p, s) : launch_march2_t<false, DM_IEEE>(p, s);
}


// Force the device code of every kernel of this file to be loaded now (see preload_all_kernels in
// alb_api.cu): with CUDA's lazy module loading the FIRST launch of a kernel may have to wait for the
// device to go idle, which never happens while a slab's wait_kernel spins for a neighbour that the
// same host thread was about to step.
#define ALB_PRELOAD(fn)                                                           \
    do {                                                                          \
        cudaFuncAttributes a_;                                                    \
        cudaError_t e_ = cudaFuncGetAttributes(&a_, reinterpret_cast<const void *>(fn)); \
        if (e_ != cudaSuccess) return e_;                                         \
    } while (0)

cudaError_t preload_march_kernels() {
    ALB_PRELOAD((march2_kernel<false, DM_FAST3>));
    ALB_PRELOAD((march2_kernel<true, DM_FAST3>));
    ALB_PRELOAD((march2_kernel<false, DM_IEEE>));
    ALB_PRELOAD((march2_kernel<true, DM_IEEE>));
    return cudaSuccess;
}

}  // namespace alb

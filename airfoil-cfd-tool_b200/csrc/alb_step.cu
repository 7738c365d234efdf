// The hot path: one fused D2Q9 step = pull-stream + half-way bounce-back on the
// mask + inlet/outlet/equilibrium borders + clamp + BGK collision, i.e.
// STEP_FS_SRC.main of the reference (pages/airfoil_flow_lbm_aerolab.html:283-360,
// "HTML:n" below), written for sm_100a.
//
// Arithmetic contract: every fp32 operation is a separately rounded IEEE
// operation in the reference's source order (the file is compiled with
// -fmad=false; divisions and the square root are the IEEE ones), so the result
// is bit-identical to the strict-fp32 CPU oracle.  The only FMAs are inside
// div_by_tau(), which computes the correctly rounded quotient (see there).
//
// Mapping: one warp = one "task" = 128 consecutive cells of one row, four cells
// per lane, so every population is moved with one 128-bit load and one 128-bit
// store per lane.  Populations that stream along x are loaded at the aligned
// own position and shifted by one cell through the neighbouring lane's
// register (shfl); only lane 0 / lane 31 issue one extra scalar load for the
// cell beyond the task.  A per-task class byte selects a branch-free path for
// tasks that are pure interior fluid (the vast majority), pure solid or pure
// equilibrium border; everything else takes the general path that patches the
// pulled populations per cell from a 16-bit info word.
#include <cooperative_groups.h>
#include <stdio.h>
#include <stdlib.h>

#include "alb_lbm.cuh"

namespace cg = cooperative_groups;

namespace alb {

namespace {

// KIND_FAST: every task of the slab, but tasks of class TC_GENERAL are skipped -- 64 registers,
// 4 CTAs per SM.  KIND_GENERAL: only the compacted list of TC_GENERAL tasks (tasks that mix cell
// types or touch the body), with the per-cell patching code.  KIND_UNIFIED: every task, both
// paths in one launch -- for lattices so small that the step is launch-latency bound and
// occupancy is irrelevant.
// KIND_FAST_LIST: the fast path over a compacted task list (pass 1 / pass 2 of the two-pass path that
// accompanies march2_kernel); entries may carry LIST_NOHIT.
constexpr int KIND_FAST = 0, KIND_GENERAL = 1, KIND_UNIFIED = 2, KIND_FAST_LIST = 3;

// DIAG (step mode): also accumulate the autoscale statistics and pressure-face sums of the state
// being WRITTEN (its rho/ux/uy are computed here anyway) -- used for the last step of a batch.
// halo push: a slab's edge rows go straight into the neighbours' ghost rows of the DESTINATION
// buffer (peer memory over NVLink, or the same GPU for in-process slabs); only the populations that
// cross the face: f2,f5,f6 upwards, f4,f7,f8 downwards
__device__ __forceinline__ void push_halo(const StepParams &p, int j, int x0, const float4 (&o)[9]) {
    if (j == p.nyl && p.peer_hi_dst) {
        st4(p.peer_hi_dst + 2 * p.peer_hi_plane + p.peer_hi_row + x0, o[2]);
        st4(p.peer_hi_dst + 5 * p.peer_hi_plane + p.peer_hi_row + x0, o[5]);
        st4(p.peer_hi_dst + 6 * p.peer_hi_plane + p.peer_hi_row + x0, o[6]);
    }
    if (j == 1 && p.peer_lo_dst) {
        st4(p.peer_lo_dst + 4 * p.peer_lo_plane + p.peer_lo_row + x0, o[4]);
        st4(p.peer_lo_dst + 7 * p.peer_lo_plane + p.peer_lo_row + x0, o[7]);
        st4(p.peer_lo_dst + 8 * p.peer_lo_plane + p.peer_lo_row + x0, o[8]);
    }
}

template <int MODE, int KIND, bool DIAG = false, int DM = DM_FAST3>
__global__ void __launch_bounds__(BLOCK_THREADS, (KIND == KIND_FAST || KIND == KIND_FAST_LIST) ? (DIAG ? ALB_DIAG_MINBLOCKS : ALB_FAST_MINBLOCKS) : 2)
step_kernel(const __grid_constant__ StepParams p) {
    const int lane = threadIdx.x & 31;
    int task = blockIdx.x * TASKS_PER_BLOCK + (threadIdx.x >> 5);
    if (KIND != KIND_GENERAL && MODE == MODE_STEP && blockIdx.x == 0 && threadIdx.x == 0 && p.me)
        me_begin_step(p.me, p.parity);   // the previous step of this handle has completed (stream order)
    constexpr bool from_list = KIND == KIND_GENERAL || KIND == KIND_FAST_LIST;
    bool count_hits = true;
    if (from_list) {
        if (task >= p.ngen) return;
        task = p.gen_list[task];
        count_hits = !(task & LIST_NOHIT);
        task &= LIST_ID_MASK;
    } else if (task >= p.ntasks) {
        return;
    }
    const int j = task / p.tpr + 1;          // local row (0 is the lower ghost row)
    const int s = task - (j - 1) * p.tpr;
    const int x0 = s * TASK_CELLS + lane * 4;
    const size_t c = (size_t)j * p.pitch + x0;
    const size_t plane = p.plane;
    const int cls = KIND == KIND_GENERAL ? (int)TC_GENERAL : (int)p.tclass[(size_t)j * p.tpr + s];   // warp-uniform
    if ((KIND == KIND_FAST || KIND == KIND_FAST_LIST) && cls == TC_GENERAL) return;
    // warp-uniform; compile-time false for the fast kinds
    const bool GENERAL = KIND != KIND_FAST && KIND != KIND_FAST_LIST && cls == TC_GENERAL;
    const float *__restrict__ src = p.src;
    [[maybe_unused]] float *const dst_base = p.dst;

    float4 o[9];

    if (KIND != KIND_GENERAL && cls == TC_EQUIL) {
        // HTML:314-322: whole task is inlet/top/bottom equilibrium at (1, U0, 0)
        if (MODE == MODE_STEP) {
#pragma unroll
            for (int i = 0; i < 9; i++) {
                const float v = p.feq0[i];
                o[i] = make_float4(v, v, v, v);
                ST4(p.dst + i * plane + c, o[i]);
            }
            push_halo(p, j, x0, o);   // a bottom / top slab of a single row: its border row is an edge row too
            if (DIAG) {   // 128 identical border cells (1, U0, 0), none of them next to a solid
                DiagLocal d;
                if (lane == 0) diag_cell(p, d, 1.0f, p.u0, 0.0f);
                diag_flush<false>(p, d, lane);
            }
        } else {
            if (p.write_macro) {
                st4(p.rho + c, make_float4(1.0f, 1.0f, 1.0f, 1.0f));
                st4(p.ux + c, make_float4(p.u0, p.u0, p.u0, p.u0));
                st4(p.uy + c, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
            }
            if (p.diag) {   // 128 identical border cells (1, U0, 0), none of them next to a solid
                DiagLocal d;
                if (lane == 0) diag_cell(p, d, 1.0f, p.u0, 0.0f);
                diag_flush<false>(p, d, lane);
            }
        }
        return;
    }
    if (KIND != KIND_GENERAL && cls == TC_SOLID) {
        // HTML:287-294: solid cells swap every population with its opposite
        if (MODE == MODE_STEP) {
            const int opp[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
#pragma unroll
            for (int i = 0; i < 9; i++) o[i] = LD4(src + opp[i] * plane + c);
#pragma unroll
            for (int i = 0; i < 9; i++) ST4(p.dst + i * plane + c, o[i]);
            push_halo(p, j, x0, o);   // nobody pulls from a solid cell, but keep the ghost rows a faithful copy
        } else if (p.write_macro) {
            st4(p.rho + c, make_float4(1.0f, 1.0f, 1.0f, 1.0f));
            st4(p.ux + c, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
            st4(p.uy + c, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
        }
        return;
    }

    // ---- pull (HTML:325-334), all loads issued before first use ------------
    const size_t cm = c - p.pitch;   // row j-1
    const size_t cp = c + p.pitch;   // row j+1
    const float4 v0 = LD4(src + 0 * plane + c);
    const float4 v1 = LD4(src + 1 * plane + c);
    const float4 v2 = LD4(src + 2 * plane + cm);
    const float4 v3 = LD4(src + 3 * plane + c);
    const float4 v4 = LD4(src + 4 * plane + cp);
    const float4 v5 = LD4(src + 5 * plane + cm);
    const float4 v6 = LD4(src + 6 * plane + cm);
    const float4 v7 = LD4(src + 7 * plane + cp);
    const float4 v8 = LD4(src + 8 * plane + cp);
    float l1 = 0.f, l5 = 0.f, l8 = 0.f, r3 = 0.f, r6 = 0.f, r7 = 0.f;
    if (lane == 0 && x0 > 0) {
        l1 = LD1(src + 1 * plane + c - 1);
        l5 = LD1(src + 5 * plane + cm - 1);
        l8 = LD1(src + 8 * plane + cp - 1);
    }
    if (lane == 31 && x0 + 4 < p.pitch) {
        r3 = LD1(src + 3 * plane + c + 4);
        r6 = LD1(src + 6 * plane + cm + 4);
        r7 = LD1(src + 7 * plane + cp + 4);
    }

    float4 own[9];
    uint2 iv = make_uint2(0u, 0u);
    if (GENERAL) {
        // own-cell populations for bounce-back / solid swap (v0, v1, v3 are own already)
        iv = __ldg(reinterpret_cast<const uint2 *>(p.info + c));
        own[0] = v0;
        own[1] = v1;
        own[3] = v3;
        own[2] = LD4(src + 2 * plane + c);
        own[4] = LD4(src + 4 * plane + c);
        own[5] = LD4(src + 5 * plane + c);
        own[6] = LD4(src + 6 * plane + c);
        own[7] = LD4(src + 7 * plane + c);
        own[8] = LD4(src + 8 * plane + c);
    }

    o[0] = v0;
    o[1] = from_left(v1, l1, lane);
    o[2] = v2;
    o[3] = from_right(v3, r3, lane);
    o[4] = v4;
    o[5] = from_left(v5, l5, lane);
    o[6] = from_right(v6, r6, lane);
    o[7] = from_right(v7, r7, lane);
    o[8] = from_left(v8, l8, lane);

    float4 mr, mx, my;   // macro outputs (macro mode)
    long long me_fx = 0, me_fy = 0;
    unsigned hits = 0;
    DiagLocal dl;
    const bool want_diag = (MODE == MODE_MACRO && p.diag != nullptr) || (MODE == MODE_STEP && DIAG);

#pragma unroll
    for (int k = 0; k < 4; k++) {
        float f[9];
#pragma unroll
        for (int i = 0; i < 9; i++) f[i] = comp(o[i], k);

        unsigned info = 0;
        if (GENERAL) {
            info = (k < 2 ? iv.x : iv.y) >> ((k & 1) * 16) & 0xffffu;
            const unsigned links = info & 0xffu;
            if (links && (info >> INFO_TYPE_SHIFT) == CT_FLUID) {
                // HTML:329-330: source cell is solid -> take my own opposite population
                const int opp[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
                const int ex[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
                const int ey[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
#pragma unroll
                for (int i = 1; i < 9; i++) {
                    if (links & (1u << (i - 1))) {
                        const float b = comp(own[opp[i]], k);
                        f[i] = b;
                        if (MODE == MODE_STEP) {
                            // momentum handed to the body, 2*b*e_opp(i), in 2^-40 fixed point
                            const long long q = __double2ll_rn((double)b * 0x1p41);
                            me_fx += -ex[i] * q;
                            me_fy += -ey[i] * q;
                        }
                    }
                }
            }
        }

        const Moments m = moments_clamped(f);
        float rho = m.rho, ux = m.ux, uy = m.uy;
        if (MODE == MODE_STEP) collide<DM>(f, m, p.tau, p.inv_tau);
        bool hit = m.hit;

        if (GENERAL) {
            const int type = (info >> INFO_TYPE_SHIFT) & INFO_TYPE_MASK;
            if (type != CT_FLUID) hit = false;
            if (type == CT_SOLID) {
                const int opp[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
#pragma unroll
                for (int i = 0; i < 9; i++) f[i] = comp(own[opp[i]], k);
                rho = 1.0f; ux = 0.0f; uy = 0.0f;
            } else if (type == CT_EQUIL) {
#pragma unroll
                for (int i = 0; i < 9; i++) f[i] = p.feq0[i];
                rho = 1.0f; ux = p.u0; uy = 0.0f;
            } else if (type == CT_OUTLET) {
                // HTML:301-312: copy all nine populations of (x-1, y), previous state
#pragma unroll
                for (int i = 0; i < 9; i++) f[i] = LD1(src + i * plane + c + k - 1);
                moments_plain(f, rho, ux, uy);
            }
        }
        // inlet / outlet cell of an otherwise all-fluid task (fast kernel): lane 0 cell 0 is the
        // equilibrium inlet (HTML:314-322), lane 31 cell 3 copies x-1 of the previous state
        // (HTML:301-312; nine late scalar loads by that one lane)
        if (ALB_EDGE_IN_FAST && KIND != KIND_GENERAL && cls == TC_FLUID_L && k == 0 && lane == 0) {
#pragma unroll
            for (int i = 0; i < 9; i++) f[i] = p.feq0[i];
            rho = 1.0f; ux = p.u0; uy = 0.0f;
            hit = false;
        }
        if (ALB_EDGE_IN_FAST && KIND != KIND_GENERAL && cls == TC_FLUID_R && k == 3 && lane == 31) {
#pragma unroll
            for (int i = 0; i < 9; i++) f[i] = LD1(src + i * plane + c + 2);   // late loads, one lane per row
            moments_plain(f, rho, ux, uy);
            hit = false;
        }
        if (hit) hits++;
        if (want_diag) {
            if (!GENERAL) {
                diag_cell(p, dl, rho, ux, uy);
            } else if (((info >> INFO_TYPE_SHIFT) & INFO_TYPE_MASK) != CT_SOLID && !(info & INFO_PAD)) {
                diag_cell(p, dl, rho, ux, uy);
                diag_faces(dl, info & 0xffu, rho, ux);
            }
        }

        if (MODE == MODE_STEP) {
#pragma unroll
            for (int i = 0; i < 9; i++) setc(o[i], k, f[i]);
        } else {
            setc(mr, k, rho);
            setc(mx, k, ux);
            setc(my, k, uy);
        }
    }

    if (MODE == MODE_STEP) {
#pragma unroll
        for (int i = 0; i < 9; i++) ST4(p.dst + i * plane + c, o[i]);

        push_halo(p, j, x0, o);

        if (DIAG) {
            if (KIND == KIND_FAST || KIND == KIND_FAST_LIST) diag_flush<false>(p, dl, lane);
            else diag_flush<true>(p, dl, lane);
        }
        if (GENERAL && p.me) {
            // integer sums are exact and order independent: shuffle tree, one atomic per warp
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                me_fx += __shfl_xor_sync(FULL, me_fx, d);
                me_fy += __shfl_xor_sync(FULL, me_fy, d);
            }
            if (lane == 0) {
                long long *acc = p.me->acc[p.parity];
                if (me_fx) atomicAdd(reinterpret_cast<unsigned long long *>(acc), (unsigned long long)me_fx);
                if (me_fy) atomicAdd(reinterpret_cast<unsigned long long *>(acc + 1), (unsigned long long)me_fy);
            }
        }
        if (hits && p.clamp_hits && count_hits) atomicAdd(p.clamp_hits, (unsigned long long)hits);
    } else {
        if (p.write_macro) {
            st4(p.rho + c, mr);
            st4(p.ux + c, mx);
            st4(p.uy + c, my);
        }
        if (want_diag) {
            if (KIND == KIND_FAST) diag_flush<false>(p, dl, lane);
            else diag_flush<true>(p, dl, lane);
        }
    }
}


// ---- small lattices: one persistent cooperative launch for a whole batch of steps -------------
// A 320x160 lattice (the reference's default) moves 3.7 MB per step: it lives in L2 and a step
// is bound by launch latency and by the length of one thread's dependent instruction chain, not
// by HBM.  So: one thread per cell (shortest chain, most warps), all CTAs co-resident, the whole
// batch of steps inside one launch with a grid-wide barrier between steps.  The arithmetic is the
// same moments_clamped()/collide() as the streaming kernels -> bit-identical results.
// The last step of the batch also reduces the statistics / face sums of the final state (p.diag).
template <int DM>
__global__ void __launch_bounds__(BLOCK_THREADS)
small_lattice_kernel(const __grid_constant__ StepParams p, float *f0, float *f1, int cur, int nsteps) {
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31;
    const int tid = blockIdx.x * BLOCK_THREADS + threadIdx.x;
    const int ncell = p.nx * p.nyl;
    const bool active = tid < ncell;
    const int row = active ? tid / p.nx : 0;
    const int x = active ? tid - row * p.nx : 0;
    const size_t c = (size_t)(row + 1) * p.pitch + x;
    const size_t plane = p.plane;
    const unsigned info = active ? p.info[c] : (unsigned)(CT_EQUIL << INFO_TYPE_SHIFT);
    const int type = (info >> INFO_TYPE_SHIFT) & INFO_TYPE_MASK;
    const unsigned links = type == CT_FLUID ? (info & 0xffu) : 0u;
    const int opp[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
    const int ex[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
    const int ey[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};

    for (int s = 0; s < nsteps; s++) {
        const float *src = ((cur + s) & 1) ? f1 : f0;
        float *dst = ((cur + s) & 1) ? f0 : f1;
        [[maybe_unused]] float *const dst_base = dst;
        const int parity = (cur + s) & 1;
        long long *slot = p.me->acc[parity];
        if (tid == 0) me_begin_step(p.me, parity);   // acc[parity] was cleared one step (one barrier) ago
        long long me_fx = 0, me_fy = 0;
        bool hit = false;
        const bool diag_now = p.diag != nullptr && s == nsteps - 1;   // statistics of the final state
        DiagLocal dl;
        if (active) {
            float f[9];
            float rho = 1.0f, ux = p.u0, uy = 0.0f;                    // equilibrium border values
            if (type == CT_FLUID) {
#pragma unroll
                for (int i = 0; i < 9; i++) {
                    if (i > 0 && (links & (1u << (i - 1)))) {
                        const float b = LD1CG(src + opp[i] * plane + c);   // HTML:329-330
                        f[i] = b;
                        const long long q = __double2ll_rn((double)b * 0x1p41);
                        me_fx += -ex[i] * q;
                        me_fy += -ey[i] * q;
                    } else {
                        f[i] = LD1CG(src + i * plane + c - (ptrdiff_t)ey[i] * p.pitch - ex[i]);
                    }
                }
                const Moments m = moments_clamped(f);
                collide<DM>(f, m, p.tau, p.inv_tau);
                hit = m.hit;
                rho = m.rho; ux = m.ux; uy = m.uy;
            } else if (type == CT_SOLID) {
#pragma unroll
                for (int i = 0; i < 9; i++) f[i] = LD1CG(src + opp[i] * plane + c);
            } else if (type == CT_OUTLET) {
#pragma unroll
                for (int i = 0; i < 9; i++) f[i] = LD1CG(src + i * plane + c - 1);
                if (diag_now) moments_plain(f, rho, ux, uy);
            } else {
#pragma unroll
                for (int i = 0; i < 9; i++) f[i] = p.feq0[i];
            }
#pragma unroll
            for (int i = 0; i < 9; i++) { ALB_CHECK_DST(dst + i * plane + c, 1); dst[i * plane + c] = f[i]; }
            if (diag_now && type != CT_SOLID && !(info & INFO_PAD)) {
                diag_cell(p, dl, rho, ux, uy);
                diag_faces(dl, info & 0xffu, rho, ux);
            }
        }
        if (diag_now) diag_flush(p, dl, lane);
        if (__any_sync(FULL, links != 0)) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                me_fx += __shfl_xor_sync(FULL, me_fx, d);
                me_fy += __shfl_xor_sync(FULL, me_fy, d);
            }
            if (lane == 0) {
                if (me_fx) atomicAdd(reinterpret_cast<unsigned long long *>(slot), (unsigned long long)me_fx);
                if (me_fy) atomicAdd(reinterpret_cast<unsigned long long *>(slot + 1), (unsigned long long)me_fy);
            }
        }
        if (hit && p.clamp_hits) atomicAdd(p.clamp_hits, 1ull);
        grid.sync();
    }
}


// ---- small lattices, second generation: the state lives in REGISTERS for the whole batch ----------
// small_lattice_kernel (above) round-trips the 1.84 MB state of the reference's 320x160 lattice
// through L2 on every step and crosses a grid-wide barrier per step: 2.9 us per step, all of it
// latency.  Here one thread owns one cell for the whole batch and the cell's nine populations never
// leave its registers; streaming is MESSAGE PASSING through an L2-resident inbox:
//   * every step a thread publishes its populations as 8-byte words {value, step tag} into
//     inbox[parity][i][cell] and pulls population i from inbox[parity][i][cell - e_i], polling each
//     word until its tag is the current step -- data and flag arrive together (one L2 round trip, no
//     fence), there is no barrier of any kind (no grid barrier, no __syncthreads, no shared memory)
//     and a thread only ever waits for its eight neighbours.  All words of a round are requested
//     together; a word that has not arrived is asked for again;
//   * the inbox is double buffered by step parity.  A word is rewritten two steps later; its writer
//     cannot get there before it has polled -- UNCONDITIONALLY, whatever the cell types -- the word
//     that the reader of its own word publishes one step later, i.e. after the reader has read it:
//     population i of cell c is read by cell c + e_i, and c pulls population opp(i) from that very
//     cell.  (An earlier version polled only where a fluid cell needed the value; all-border threads
//     then ran ahead and overwrote unread words.)  The outlet cell copies all nine populations of
//     its left neighbour, which in turn pulls f3 from the outlet cell: same argument;
//   * the last two states of the batch are written to the two global buffers, so everything else
//     in the library (lazy macroscopic pass, getters, diagnostics) finds what a sequence of single
//     steps would have left.
// Measured on the way (320x160): CTA row bands + shared memory + one poll at a time 2.71 us per step;
// one row-major strip per SM, batched polls 2.35 us (714 instructions per warp and step, of which 125
// are the collision: the shared-memory / inbox case distinctions and the barrier cost the rest).
// Momentum-exchange sums go straight into the history ring slot of their step (threads are not in
// lock step, so the two-accumulator scheme of the streaming kernels does not apply; the host zeroes
// the slots of the batch before the launch and fixes MeState up afterwards).  Same
// moments_clamped()/collide() -> bit-identical.  All CTAs must be co-resident (cooperative launch).
struct BandWord { float v; int tag; };
// GPU-scope relaxed accesses: the 8-byte word is written and read as ONE access, so value and tag
// travel together and no ordering between different words is needed (volatile would compile to
// system-scope accesses)
__device__ __forceinline__ void band_put(BandWord *w, float v, int tag) {
    asm volatile("st.relaxed.gpu.global.v2.b32 [%0], {%1, %2};" ::"l"(w), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
__device__ __forceinline__ void band_poll(const BandWord *w, unsigned &v, unsigned &t) {
    asm volatile("ld.relaxed.gpu.global.v2.b32 {%0, %1}, [%2];" : "=r"(v), "=r"(t) : "l"(w) : "memory");
}

constexpr int BAND_MAX_THREADS = 384;       // 3 warps per scheduler: up to 168 registers per thread
template <int DM>
__global__ void __launch_bounds__(BAND_MAX_THREADS)
band_lattice_kernel(const __grid_constant__ StepParams p, float *f0, float *f1, int cur, int nsteps, int L,
                    BandWord *inbox, long long step_base, int *err) {
    const int nx = p.nx, ncell = p.nx * p.nyl;
    const int tid = threadIdx.x, lane = tid & 31;
    const int cell = blockIdx.x * L + tid;               // row-major index over the owned rows
    const bool active = tid < L && cell < ncell;
    const int y = active ? cell / nx : 0, x = active ? cell - y * nx : 0;
    const size_t plane = p.plane;
    const size_t c = (size_t)(y + 1) * p.pitch + x;
    const unsigned info = active ? p.info[c] : (unsigned)(CT_EQUIL << INFO_TYPE_SHIFT);
    const int type = (info >> INFO_TYPE_SHIFT) & INFO_TYPE_MASK;
    const unsigned links = type == CT_FLUID ? (info & 0xffu) : 0u;
    const int opp[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
    const int ex[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
    const int ey[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
    // word index (within one parity's [9][ncell] block) this thread publishes population i to, and the
    // one it pulls population i from; bit i of `rem`: that source cell exists (row-major arithmetic:
    // at the ends of a row the "neighbour" is a cell of the adjacent row -- a border cell never uses
    // the value, but it polls it all the same, which is what keeps the pairing symmetric)
    unsigned rem = 0;
    int pidx[9], ridx[9];
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const int sc = cell - ey[i] * nx - ex[i];
        pidx[i] = i * ncell + cell;
        ridx[i] = i * ncell + sc;
        if (active && i > 0 && sc >= 0 && sc < ncell) rem |= 1u << i;
    }
    const bool outlet = active && type == CT_OUTLET && cell > 0;
    float f[9];
    {
        const float *src = cur ? f1 : f0;
#pragma unroll
        for (int i = 0; i < 9; i++) f[i] = active ? __ldcg(src + i * plane + c) : 0.0f;
    }
    for (int s = 0; s < nsteps; s++) {
        const int tag = (int)(step_base + s + 1);
        BandWord *box = inbox + (size_t)(s & 1) * 9 * ncell;      // [9][ncell], indexed by SOURCE cell
        if (s == nsteps - 1 && active) {
            // the state before the last step goes to the buffer that will hold the previous state
            float *prev = ((cur + nsteps - 1) & 1) ? f1 : f0;
            [[maybe_unused]] float *const dst_base = prev;
#pragma unroll
            for (int i = 0; i < 9; i++) { ALB_CHECK_DST(prev + i * plane + c, 1); prev[i * plane + c] = f[i]; }
        }
        long long me_fx = 0, me_fy = 0;
        bool hit = false;
        const bool diag_now = p.diag != nullptr && s == nsteps - 1;   // statistics of the final state
        DiagLocal dl;
        if (active) {
#pragma unroll
            for (int i = 0; i < 9; i++) band_put(box + pidx[i], f[i], tag);
            // pull, whatever the cell type (see above): all words of a round are in flight together
            float g[9];
            g[0] = f[0];
#pragma unroll
            for (int i = 1; i < 9; i++) g[i] = 0.0f;                   // sources outside the lattice: border cells, unused
            unsigned pend = rem;
            for (int spin = 0; pend; spin++) {
                unsigned v[9], tg[9];
#pragma unroll
                for (int i = 1; i < 9; i++)
                    if (pend & (1u << i)) band_poll(box + ridx[i], v[i], tg[i]);
#pragma unroll
                for (int i = 1; i < 9; i++)
                    if ((pend & (1u << i)) && (int)tg[i] == tag) {
                        g[i] = __uint_as_float(v[i]);
                        pend &= ~(1u << i);
                    }
                if (spin > (1 << 22)) {      // seconds: a neighbour is gone -- give up instead of hanging the GPU
                    *err = 1;
                    break;
                }
            }
            float rho = 1.0f, ux = p.u0, uy = 0.0f;                    // equilibrium border values
            if (type == CT_FLUID) {
                if (links) {
#pragma unroll
                    for (int i = 1; i < 9; i++) {
                        if (links & (1u << (i - 1))) {
                            const float b = f[opp[i]];                 // HTML:329-330: my own opposite population
                            g[i] = b;
                            const long long q = __double2ll_rn((double)b * 0x1p41);
                            me_fx += -ex[i] * q;
                            me_fy += -ey[i] * q;
                        }
                    }
                }
                const Moments m = moments_clamped(g);
                collide<DM>(g, m, p.tau, p.inv_tau);
                hit = m.hit;
                rho = m.rho; ux = m.ux; uy = m.uy;
            } else if (type == CT_SOLID) {
#pragma unroll
                for (int i = 0; i < 9; i++) g[i] = f[opp[i]];
            } else if (outlet) {
                // HTML:301-312: all nine populations of (x-1, y), previous state
                unsigned pend9 = 0x1ffu;
                for (int spin = 0; pend9; spin++) {
                    unsigned v[9], tg[9];
#pragma unroll
                    for (int i = 0; i < 9; i++)
                        if (pend9 & (1u << i)) band_poll(box + (pidx[i] - 1), v[i], tg[i]);
#pragma unroll
                    for (int i = 0; i < 9; i++)
                        if ((pend9 & (1u << i)) && (int)tg[i] == tag) {
                            g[i] = __uint_as_float(v[i]);
                            pend9 &= ~(1u << i);
                        }
                    if (spin > (1 << 22)) {
                        *err = 1;
                        break;
                    }
                }
                if (diag_now) moments_plain(g, rho, ux, uy);
            } else {
#pragma unroll
                for (int i = 0; i < 9; i++) g[i] = p.feq0[i];
            }
#pragma unroll
            for (int i = 0; i < 9; i++) f[i] = g[i];
            if (diag_now && type != CT_SOLID && !(info & INFO_PAD)) {
                diag_cell(p, dl, rho, ux, uy);
                diag_faces(dl, info & 0xffu, rho, ux);
            }
        }
        if (diag_now) diag_flush(p, dl, lane);
        if (__any_sync(FULL, links != 0)) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                me_fx += __shfl_xor_sync(FULL, me_fx, d);
                me_fy += __shfl_xor_sync(FULL, me_fy, d);
            }
            if (lane == 0) {
                long long *slot = p.me->ring[(step_base + s) % ME_RING];
                if (me_fx) atomicAdd(reinterpret_cast<unsigned long long *>(slot), (unsigned long long)me_fx);
                if (me_fy) atomicAdd(reinterpret_cast<unsigned long long *>(slot + 1), (unsigned long long)me_fy);
            }
        }
        if (hit && p.clamp_hits) atomicAdd(p.clamp_hits, 1ull);
    }
    if (active) {
        float *dst = ((cur + nsteps) & 1) ? f1 : f0;
        [[maybe_unused]] float *const dst_base = dst;
#pragma unroll
        for (int i = 0; i < 9; i++) { ALB_CHECK_DST(dst + i * plane + c, 1); dst[i * plane + c] = f[i]; }
    }
}

// momentum-exchange bookkeeping around a band_lattice_kernel batch (see MeState): before -- commit the
// pending sums of the previous step to the ring and clear both accumulators; after -- the batch's
// sums sit in the ring already: advance the counter and mirror the last step into the accumulator a
// streaming step would have left it in (frame_finalize_kernel reads it there)
__global__ void me_before_band_kernel(MeState *m, int parity, long long step_base, int nsteps) {
    if (threadIdx.x == 0) {
        if (m->pending) {
            const long long c = m->count;          // == step_base - 1: not among the slots zeroed below
            m->ring[c % ME_RING][0] = m->acc[parity ^ 1][0];
            m->ring[c % ME_RING][1] = m->acc[parity ^ 1][1];
            m->count = c + 1;
        }
        m->acc[0][0] = m->acc[0][1] = m->acc[1][0] = m->acc[1][1] = 0;
        m->pending = 0;
    }
    // the ring slots of this batch collect atomic sums: zero them
    for (int s = threadIdx.x; s < nsteps; s += blockDim.x) {
        m->ring[(step_base + s) % ME_RING][0] = 0;
        m->ring[(step_base + s) % ME_RING][1] = 0;
    }
}
__global__ void me_after_band_kernel(MeState *m, long long count, int last_parity) {
    m->count = count;
    m->acc[last_parity][0] = m->ring[(count - 1) % ME_RING][0];
    m->acc[last_parity][1] = m->ring[(count - 1) % ME_RING][1];
    m->pending = 0;
}

}  // namespace

// The fast kernel and the general kernel of one step read the same source state and write
// disjoint cells, so the caller may run them concurrently on two streams.
template <int KIND>
cudaError_t launch_step_kind(const StepParams &p, int nblocks, cudaStream_t s) {
    if (p.div_mode == DM_FAST3) {
        if (p.diag) step_kernel<MODE_STEP, KIND, true, DM_FAST3><<<nblocks, BLOCK_THREADS, 0, s>>>(p);
        else step_kernel<MODE_STEP, KIND, false, DM_FAST3><<<nblocks, BLOCK_THREADS, 0, s>>>(p);
    } else {
        if (p.diag) step_kernel<MODE_STEP, KIND, true, DM_IEEE><<<nblocks, BLOCK_THREADS, 0, s>>>(p);
        else step_kernel<MODE_STEP, KIND, false, DM_IEEE><<<nblocks, BLOCK_THREADS, 0, s>>>(p);
    }
    return cudaGetLastError();
}

cudaError_t launch_step_fast(const StepParams &p, cudaStream_t s) {
    return launch_step_kind<KIND_FAST>(p, (p.ntasks + TASKS_PER_BLOCK - 1) / TASKS_PER_BLOCK, s);
}

cudaError_t launch_step_general(const StepParams &p, cudaStream_t s) {
    if (p.ngen == 0) return cudaSuccess;
    return launch_step_kind<KIND_GENERAL>(p, (p.ngen + TASKS_PER_BLOCK - 1) / TASKS_PER_BLOCK, s);
}

cudaError_t launch_step_unified(const StepParams &p, cudaStream_t s) {
    return launch_step_kind<KIND_UNIFIED>(p, (p.ntasks + TASKS_PER_BLOCK - 1) / TASKS_PER_BLOCK, s);
}

cudaError_t launch_step_fast_list(const StepParams &p, cudaStream_t s) {
    // at least one CTA even for an empty list: its first thread does the momentum-exchange bookkeeping
    return launch_step_kind<KIND_FAST_LIST>(p, p.ngen > 0 ? (p.ngen + TASKS_PER_BLOCK - 1) / TASKS_PER_BLOCK : 1, s);
}

// Largest number of cells the persistent small-lattice kernel can own on this device (all CTAs
// must be co-resident for the grid barrier); 0 when cooperative launches are unsupported.
int small_lattice_capacity(int device) {
    int coop = 0, sms = 0, per_sm = 0;
    if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device) != cudaSuccess || !coop) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 0;
    int per_sm2 = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, small_lattice_kernel<DM_FAST3>, BLOCK_THREADS, 0) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, small_lattice_kernel<DM_IEEE>, BLOCK_THREADS, 0) != cudaSuccess)
        return 0;
    return sms * (per_sm < per_sm2 ? per_sm : per_sm2) * BLOCK_THREADS;
}

cudaError_t launch_small_lattice(const StepParams &p, float *f0, float *f1, int cur, int nsteps, cudaStream_t s) {
    const int ncell = p.nx * p.nyl;
    const int nblocks = (ncell + BLOCK_THREADS - 1) / BLOCK_THREADS;
    StepParams pp = p;
    void *args[] = {&pp, &f0, &f1, &cur, &nsteps};
    const void *fn = p.div_mode == DM_FAST3 ? (const void *)small_lattice_kernel<DM_FAST3> : (const void *)small_lattice_kernel<DM_IEEE>;
    return cudaLaunchCooperativeKernel(fn, dim3(nblocks), dim3(BLOCK_THREADS), args, 0, s);
}

// Cells per CTA (= threads) of band_lattice_kernel for this lattice on a device with nsm SMs, or 0
// when the lattice does not qualify: one thread per cell, one CTA per SM at most (every thread spins
// on its neighbours, so all must be resident), the cells spread evenly over the SMs.
int band_lattice_rows(int nx, int nyl, int nsm) {
    if (nx < 3 || nyl < 3 || nsm < 1) return 0;
    const long long ncell = (long long)nx * nyl;
    long long L = (ncell + nsm - 1) / nsm;
    if (L < 32) L = 32;
    if (L > BAND_MAX_THREADS) return 0;
    return (int)L;
}
size_t band_inbox_bytes(int nx, int nyl, int /*L*/) {
    return 2ull * 9 * (size_t)nx * nyl * sizeof(BandWord);
}

// nsteps <= ME_RING / 2 steps of the whole lattice in one launch; step_base = number of steps taken so
// far (== MeState::count once the pending step is committed); parity = momentum-exchange slot of the
// NEXT streaming step (== cur on a handle that never ran a double step)
cudaError_t launch_band_lattice(const StepParams &p, float *f0, float *f1, int cur, int nsteps, int R, void *inbox,
                                long long step_base, int *err, cudaStream_t s) {
    me_before_band_kernel<<<1, 256, 0, s>>>(p.me, p.parity, step_base, nsteps);
    const int L = R;                                           // cells per CTA
    const int nbands = (p.nx * p.nyl + L - 1) / L;
    const int threads = (L + 31) / 32 * 32;
    const size_t smem = 0;
    const void *fn = p.div_mode == DM_FAST3 ? (const void *)band_lattice_kernel<DM_FAST3> : (const void *)band_lattice_kernel<DM_IEEE>;
    cudaError_t e = cudaSuccess;
    StepParams pp = p;
    BandWord *ib = reinterpret_cast<BandWord *>(inbox);
    void *args[] = {&pp, &f0, &f1, &cur, &nsteps, &R, &ib, &step_base, &err};
    e = cudaLaunchCooperativeKernel(fn, dim3(nbands), dim3(threads), args, smem, s);
    if (e != cudaSuccess) return e;
    me_after_band_kernel<<<1, 1, 0, s>>>(p.me, step_base + nsteps, (cur + nsteps - 1) & 1);
    return cudaGetLastError();
}

cudaError_t launch_macro(const StepParams &p, cudaStream_t s) {
    const int nblocks = (p.ntasks + TASKS_PER_BLOCK - 1) / TASKS_PER_BLOCK;
    step_kernel<MODE_MACRO, KIND_FAST><<<nblocks, BLOCK_THREADS, 0, s>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || p.ngen == 0) return e;
    const int gblocks = (p.ngen + TASKS_PER_BLOCK - 1) / TASKS_PER_BLOCK;
    step_kernel<MODE_MACRO, KIND_GENERAL><<<gblocks, BLOCK_THREADS, 0, s>>>(p);
    return cudaGetLastError();
}

// feq_i(rho = 1, ux = U0, uy = 0) exactly as the shader evaluates it
// (HTML:276-281, 315-317).  Host code; built with -ffp-contract=off.
void host_feq0(float u0, float *out9) {
    const float w0 = 4.0f / 9.0f, ws = 1.0f / 9.0f, wd = 1.0f / 36.0f;
    const float W[9] = {w0, ws, ws, ws, ws, wd, wd, wd, wd};
    const float EXf[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
    const float EYf[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
    volatile float rho = 1.0f, ux = u0, uy = 0.0f;   // volatile: keep the operations separate
    for (int i = 0; i < 9; i++) {
        volatile float a = EXf[i] * ux;
        volatile float b = EYf[i] * uy;
        volatile float eu = a + b;
        volatile float uxx = ux * ux;
        volatile float uyy = uy * uy;
        volatile float uu = uxx + uyy;
        volatile float wr = W[i] * rho;
        volatile float t1 = 3.0f * eu;
        volatile float s1 = 1.0f + t1;
        volatile float t2 = 4.5f * eu;
        volatile float t3 = t2 * eu;
        volatile float s2 = s1 + t3;
        volatile float t4 = 1.5f * uu;
        volatile float s3 = s2 - t4;
        out9[i] = wr * s3;
    }
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

cudaError_t preload_step_kernels() {
    ALB_PRELOAD((step_kernel<MODE_STEP, KIND_FAST, false, DM_FAST3>));
    ALB_PRELOAD((step_kernel<MODE_STEP, KIND_FAST, true, DM_FAST3>));
    ALB_PRELOAD((step_kernel<MODE_STEP, KIND_FAST, false, DM_IEEE>));
    ALB_PRELOAD((step_kernel<MODE_STEP, KIND_FAST, true, DM_IEEE>));
    ALB_PRELOAD((step_kernel<MODE_STEP, KIND_GENERAL, false, DM_FAST3>));
    ALB_PRELOAD((step_kernel<MODE_STEP, KIND_GENERAL, true, DM_FAST3>));
    ALB_PRELOAD((step_kernel<MODE_STEP, KIND_GENERAL, false, DM_IEEE>));
    ALB_PRELOAD((step_kernel<MODE_STEP, KIND_GENERAL, true, DM_IEEE>));
    ALB_PRELOAD((step_kernel<MODE_STEP, KIND_UNIFIED, false, DM_FAST3>));
    ALB_PRELOAD((step_kernel<MODE_STEP, KIND_UNIFIED, true, DM_FAST3>));
    ALB_PRELOAD((step_kernel<MODE_STEP, KIND_UNIFIED, false, DM_IEEE>));
    ALB_PRELOAD((step_kernel<MODE_STEP, KIND_UNIFIED, true, DM_IEEE>));
    ALB_PRELOAD((step_kernel<MODE_STEP, KIND_FAST_LIST, false, DM_FAST3>));
    ALB_PRELOAD((step_kernel<MODE_STEP, KIND_FAST_LIST, true, DM_FAST3>));
    ALB_PRELOAD((step_kernel<MODE_STEP, KIND_FAST_LIST, false, DM_IEEE>));
    ALB_PRELOAD((step_kernel<MODE_STEP, KIND_FAST_LIST, true, DM_IEEE>));
    ALB_PRELOAD((step_kernel<MODE_MACRO, KIND_FAST>));
    ALB_PRELOAD((step_kernel<MODE_MACRO, KIND_GENERAL>));
    ALB_PRELOAD((small_lattice_kernel<DM_FAST3>));
    ALB_PRELOAD((small_lattice_kernel<DM_IEEE>));
    ALB_PRELOAD((band_lattice_kernel<DM_FAST3>));
    ALB_PRELOAD((band_lattice_kernel<DM_IEEE>));
    ALB_PRELOAD(me_before_band_kernel);
    ALB_PRELOAD(me_after_band_kernel);
    return cudaSuccess;
}

}  // namespace alb

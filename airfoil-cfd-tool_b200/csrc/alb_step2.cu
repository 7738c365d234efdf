// Two LBM steps per pass over HBM: the fused step2_kernel (and its measured-and-rejected staging
// variants), the copy kernel for all-solid tasks, the tiling plan, and the self-test of the
// shared-reciprocal division.  Same arithmetic as alb_step.cu (alb_lbm.cuh) -> bit-identical to
// two single steps.  DESIGN.md section 4.2.
#include <stdio.h>
#include <stdlib.h>

#include "alb_lbm.cuh"

namespace alb {

namespace {

// Shape of the fused two-step kernel: RB rows per group, K 128-cell segments per strip.
#ifndef ALB_S2_RB
#define ALB_S2_RB 2
#endif
#ifndef ALB_S2_K
#define ALB_S2_K 5
#endif
#ifndef ALB_S2_HS_MAX
#define ALB_S2_HS_MAX 256
#endif
#ifndef ALB_S2_HS_MIN
#define ALB_S2_HS_MIN 64
#endif
// (A variant that staged whole row groups through a shared-memory ring filled by a producer warp with
// TMA bulk copies, 2-3 stages deep, was measured at 96-112 GLUPS and removed: the staging ring costs
// the shared memory of 4-8 compute warps.  See DESIGN.md section 4.2 and the git history.)
// step2_kernel: 1 = the A warps prefetch their next task with TMA bulk copies into private shared-memory
// staging instead of loading it into registers just before the barrier.  Measured on 32768x16384:
// long_scoreboard stalls drop from 15 % to 3 % of the samples, but the step is SLOWER (116 vs 125 GLUPS;
// a cp.async version: 116-123) -- see DESIGN.md section 4.2.
#ifndef ALB_S2_ASYNC
#define ALB_S2_ASYNC 0
#endif
// step2_kernel, experimental: re-split the register file between the roles after launch (setmaxnreg)
#ifndef ALB_S2_SETMAXNREG
#define ALB_S2_SETMAXNREG 0
#endif
#ifndef ALB_S2_REGS_A
#define ALB_S2_REGS_A 96
#endif
#ifndef ALB_S2_REGS_B
#define ALB_S2_REGS_B 64
#endif


// ---- two steps per pass over HBM (temporal blocking) ---------------------------------------------
// The single-step kernel already moves exactly the algorithmic 72 B per cell update and runs at the
// DRAM ceiling, so the only way up is to touch HBM less often: step2_kernel advances the deep
// interior of the lattice by TWO steps while reading the state once and writing it once (36 B per
// cell update).  A CTA owns a column strip of WI = 128*K cells and marches up the rows.  Half of
// its warps ("A") run step 1 exactly like the fast kernel (aligned 128-bit loads from HBM, shuffle
// shifts) but store the result into a ring of 3*RB rows in shared memory; the other half ("B")
// run step 2 out of that ring, one row group behind, and store to HBM.  While the A warps wait
// for their loads the B warps compute, so one __syncthreads per row group is all the coordination
// needed.  Step 2 of a cell needs step 1 of its 8 neighbours, hence the strip's outermost 4
// columns and the rows above/below a segment are computed redundantly (A only) and never stored.
// Same moments_clamped()/collide() as everywhere else -> bit-identical to two single steps.
// Everything that is not "deep" (border cells, the body and its surroundings, slab edge rows) is
// advanced by two passes of the list-driven single-step kernels through a third buffer.
// dst = src on the listed tasks (all-solid tasks over a double step, see build_lists_kernel)
__global__ void __launch_bounds__(BLOCK_THREADS)
copy_tasks_kernel(const __grid_constant__ StepParams p) {
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * TASKS_PER_BLOCK + (threadIdx.x >> 5);
    if (t >= p.ngen) return;
    const int task = p.gen_list[t] & LIST_ID_MASK;
    const int j = task / p.tpr + 1;
    const size_t c = (size_t)j * p.pitch + (task - (j - 1) * p.tpr) * TASK_CELLS + lane * 4;
    const size_t plane = p.plane;
    [[maybe_unused]] const float *const src = p.src;
    [[maybe_unused]] float *const dst_base = p.dst;
    float4 v[9];
#pragma unroll
    for (int i = 0; i < 9; i++) v[i] = LD4(p.src + i * plane + c);
#pragma unroll
    for (int i = 0; i < 9; i++) ST4(p.dst + i * plane + c, v[i]);
}

#if ALB_S2_ASYNC
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

#endif

// loads that must stay where they are written (issued BEFORE the barrier that ends a super-step)
__device__ __forceinline__ float4 ld4_pinned(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float ld1_pinned(const float *p) {
    float r;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
[[maybe_unused]] constexpr int S2_STG = 4 + 128 + 4;    // floats per population in an A warp's private staging buffer
#define LD4P(ptr) (ALB_CHECK_SRC((ptr), 4), ld4_pinned(ptr))
#define LD1P(ptr) (ALB_CHECK_SRC((ptr), 1), ld1_pinned(ptr))

#ifndef ALB_S2_MINB
#define ALB_S2_MINB ((16 / (ALB_S2_RB * ALB_S2_K)) >= 1 ? 16 / (ALB_S2_RB * ALB_S2_K) : 1)
#endif
// DIAG: step 2 also reduces the autoscale statistics of the state it writes (deep cells have no
// faces), once per tile -- for batches that END with a double step.
template <int RB, int K, bool DIAG = false>
__global__ void __launch_bounds__(2 * RB * K * 32, ALB_S2_MINB)
step2_kernel(const __grid_constant__ Step2Params p) {
    extern __shared__ float4 ring4[];
    float *ring = reinterpret_cast<float *>(ring4);   // [RS][9][WI]
    constexpr int WI = 128 * K, NW = RB * K, RS = 2 * RB + 2;
    [[maybe_unused]] float *stage_all = ring + (size_t)RS * 9 * WI;   // ALB_S2_ASYNC: [NW][9][S2_STG]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool role_b = warp >= NW;
    const int w = role_b ? warp - NW : warp;
    const int r = w / K, seg = w - r * K;
    const int strip = blockIdx.x % p.nstrips, sgm = blockIdx.x / p.nstrips;
    const int y0 = 2 + sgm * p.hs;                 // owned output rows [y0, y1)
    const int y1 = min(y0 + p.hs, p.nyl);
    const int a0 = y0 - 1;                         // first intermediate row
    const int nga = (y1 - y0 + 2 + RB - 1) / RB;   // row groups of step 1
    const int col = seg * 128 + lane * 4;          // column of this lane's quad inside the strip
    const int gx = strip * p.wo - 4 + col;         // and on the lattice
    const bool inx = gx >= 0 && gx < p.pitch;
    const bool ownx = col >= 4 && col < min(p.wo, p.pitch - strip * p.wo) + 4;
    const size_t plane = p.plane;
    const float *__restrict__ src = p.src;
    [[maybe_unused]] float *const dst_base = p.dst;
    const uint8_t *tfl = p.tflags + (inx ? (gx >> 7) : 0);
    unsigned hits = 0;
#if ALB_S2_SETMAXNREG
    // Experimental: the step-1 warps need ~100 registers (36 of them hold the next task's loads across
    // the barrier), the step-2 warps far fewer; with warpgroup-aligned roles the register file can be
    // re-split after launch, so that 24 warps fit without spills (compile with RB*K a multiple of 4).
    static_assert(!ALB_S2_SETMAXNREG || (RB * K) % 4 == 0, "roles must be whole warpgroups");
    if (!role_b) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(ALB_S2_REGS_A));
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ALB_S2_REGS_B));
#endif

    if (!role_b) {
        // ---- A warps: step 1, HBM -> ring.  The loads of the NEXT row group are issued before the
        // barrier that ends the current one (their 36 registers are dead by then), so HBM latency
        // overlaps the barrier wait and the B warps' work instead of heading every task; the task
        // flags are fetched two groups ahead for the same reason.
        // flags of this warp's tasks in row-group order (0 once past the last intermediate row):
        // a running row / pointer pair instead of index arithmetic per group
        int jf = a0 + r, of = jf * p.tpr;       // rows * tasks per row < 2^30 (checked in alb_create)
        auto next_flags = [&]() -> unsigned {
            const unsigned v = (jf <= y1 && inx) ? tfl[of] : 0u;
            jf += RB;
            of += RB * p.tpr;
            return v;
        };
#if ALB_S2_ASYNC
        // Each A warp owns a private staging buffer of one task, 9 rows of 4 + 128 + 4 floats (the
        // task's 128 cells plus the quad to its left and right, so the x-neighbours come along).
        // As soon as the populations of the current task are in registers, one lane starts nine
        // 1-D bulk copies (TMA, completion on the warp's own mbarrier) of the NEXT task into the
        // same buffer: HBM latency is covered by a whole task of arithmetic plus the barrier, no
        // registers are held, and the per-lane address arithmetic of nine LDG.128 disappears.
        float *const stg_row = stage_all + (size_t)w * 9 * S2_STG;
        unsigned long long *const bar = reinterpret_cast<unsigned long long *>(stage_all + (size_t)NW * 9 * S2_STG) + w;
        if (lane == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        // lattice columns [c_lo, c_hi) of this warp's segment incl. the two extra quads, clipped to the row
        const int seg_x = strip * p.wo - 4 + seg * 128;
        const int c_lo = max(seg_x - 4, 0), c_hi = min(seg_x + 132, p.pitch);
        const bool seg_in = c_hi > c_lo;
        const float *stg = stg_row + 4 + lane * 4;
        auto issue_loads = [&](int g) {
            if (lane == 0 && seg_in) {
                const int j = a0 + g * RB + r;
                const unsigned bytes = (unsigned)(c_hi - c_lo) * 4u;
                mbar_arrive_expect_tx(bar, 9u * bytes);
                const float *g0 = src + (size_t)j * p.pitch + c_lo;
                float *s0 = stg_row + (c_lo - (seg_x - 4));
                const int ey[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
#pragma unroll
                for (int i = 0; i < 9; i++) {
                    const float *gs = g0 + i * plane - (ptrdiff_t)ey[i] * p.pitch;
                    ALB_CHECK_SRC(gs, c_hi - c_lo);
                    tma_load_1d(s0 + i * S2_STG, gs, bytes, bar);
                }
            }
        };
        unsigned tf = next_flags(), tf1 = next_flags(), phase = 0;
        bool have = __any_sync(FULL, tf & TF_NEED);
        if (have) issue_loads(0);
        for (int g = 0; g <= nga; g++) {
            const unsigned tf2 = next_flags();
            const bool have_next = __any_sync(FULL, tf1 & TF_NEED);
            float4 o[9];
            if (have) {
                if (seg_in) mbar_wait(bar, phase);
                phase ^= 1u;
                const float4 v0 = *reinterpret_cast<const float4 *>(stg + 0 * S2_STG);
                const float4 v1 = *reinterpret_cast<const float4 *>(stg + 1 * S2_STG);
                const float4 v2 = *reinterpret_cast<const float4 *>(stg + 2 * S2_STG);
                const float4 v3 = *reinterpret_cast<const float4 *>(stg + 3 * S2_STG);
                const float4 v4 = *reinterpret_cast<const float4 *>(stg + 4 * S2_STG);
                const float4 v5 = *reinterpret_cast<const float4 *>(stg + 5 * S2_STG);
                const float4 v6 = *reinterpret_cast<const float4 *>(stg + 6 * S2_STG);
                const float4 v7 = *reinterpret_cast<const float4 *>(stg + 7 * S2_STG);
                const float4 v8 = *reinterpret_cast<const float4 *>(stg + 8 * S2_STG);
                float l1 = 0.f, l5 = 0.f, l8 = 0.f, r3 = 0.f, r6 = 0.f, r7 = 0.f;
                if (lane == 0) {
                    l1 = stg[1 * S2_STG - 1];
                    l5 = stg[5 * S2_STG - 1];
                    l8 = stg[8 * S2_STG - 1];
                }
                if (lane == 31) {
                    r3 = stg[3 * S2_STG + 4];
                    r6 = stg[6 * S2_STG + 4];
                    r7 = stg[7 * S2_STG + 4];
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
            }
            // the staged values are in registers (the shuffles consumed them): refill the buffer
            if (have_next) {
                __syncwarp();
                issue_loads(g + 1);
            }
            if (have) {
                const int j = a0 + g * RB + r;
                const unsigned hm = collide_quad(o, p.tau, p.inv_tau, p.inv_tau_lo);
                if ((tf & TF_DEEP) && ownx && j >= y0 && j < y1) hits += __popc(hm);
                float *slot = ring + ((size_t)((j - a0) % RS) * 9) * WI + col;
#pragma unroll
                for (int i = 0; i < 9; i++) *reinterpret_cast<float4 *>(slot + i * WI) = o[i];
            }
            tf = tf1;
            tf1 = tf2;
            have = have_next;
            __syncthreads();
        }
#else
        float4 v0, v1, v2, v3, v4, v5, v6, v7, v8;
        float l1 = 0.f, l5 = 0.f, l8 = 0.f, r3 = 0.f, r6 = 0.f, r7 = 0.f;
        v0 = v1 = v2 = v3 = v4 = v5 = v6 = v7 = v8 = make_float4(0.f, 0.f, 0.f, 0.f);
        auto issue_loads = [&](int g) {
            const int j = a0 + g * RB + r;
            const size_t c = (size_t)j * p.pitch + gx;
            const size_t cm = c - p.pitch, cp = c + p.pitch;
            if (inx) {
                v0 = LD4P(src + 0 * plane + c);
                v1 = LD4P(src + 1 * plane + c);
                v2 = LD4P(src + 2 * plane + cm);
                v3 = LD4P(src + 3 * plane + c);
                v4 = LD4P(src + 4 * plane + cp);
                v5 = LD4P(src + 5 * plane + cm);
                v6 = LD4P(src + 6 * plane + cm);
                v7 = LD4P(src + 7 * plane + cp);
                v8 = LD4P(src + 8 * plane + cp);
            }
            if (lane == 0 && gx > 0) {
                l1 = LD1P(src + 1 * plane + c - 1);
                l5 = LD1P(src + 5 * plane + cm - 1);
                l8 = LD1P(src + 8 * plane + cp - 1);
            }
            if (lane == 31 && gx + 4 < p.pitch) {
                r3 = LD1P(src + 3 * plane + c + 4);
                r6 = LD1P(src + 6 * plane + cm + 4);
                r7 = LD1P(src + 7 * plane + cp + 4);
            }
        };
        unsigned tf = next_flags(), tf1 = next_flags();
        bool have = __any_sync(FULL, tf & TF_NEED);
        if (have) issue_loads(0);
        for (int g = 0; g <= nga; g++) {
            const unsigned tf2 = next_flags();
            if (have) {
                const int j = a0 + g * RB + r;
                float4 o[9];
                o[0] = v0;
                o[1] = from_left(v1, l1, lane);
                o[2] = v2;
                o[3] = from_right(v3, r3, lane);
                o[4] = v4;
                o[5] = from_left(v5, l5, lane);
                o[6] = from_right(v6, r6, lane);
                o[7] = from_right(v7, r7, lane);
                o[8] = from_left(v8, l8, lane);
                const unsigned hm = collide_quad(o, p.tau, p.inv_tau, p.inv_tau_lo);
                // clamp hits of step 1: every deep cell is owned by exactly one tile
                if ((tf & TF_DEEP) && ownx && j >= y0 && j < y1) hits += __popc(hm);
                float *slot = ring + ((size_t)((j - a0) % RS) * 9) * WI + col;
#pragma unroll
                for (int i = 0; i < 9; i++) *reinterpret_cast<float4 *>(slot + i * WI) = o[i];
            }
            tf = tf1;
            tf1 = tf2;
            have = __any_sync(FULL, tf & TF_NEED);
            if (have) issue_loads(g + 1);
            __syncthreads();
        }
#endif  // ALB_S2_ASYNC
    } else {
        // ---- B warps: step 2, ring -> HBM, one row group behind ----
        // flags of this warp's output rows, group 1 first (0 for rows outside [y0, y1) and for lanes
        // outside the strip's own columns)
        int jf = y0 - 2 + r, of = jf * p.tpr;
        auto next_flags = [&]() -> unsigned {
            const unsigned v = ((unsigned)(jf - y0) < (unsigned)(y1 - y0) && ownx) ? tfl[of] : 0u;
            jf += RB;
            of += RB * p.tpr;
            return v;
        };
        unsigned tf = 0u, tf1 = next_flags();
        [[maybe_unused]] DiagLocal dl;
        for (int g = 0; g <= nga; g++) {
            const unsigned tf2 = next_flags();
            const bool st = (tf & TF_DEEP) != 0;
            if (__any_sync(FULL, st)) {
                const int j = y0 - 2 + (g - 1) * RB + r;
                const int q = j - a0;
                const float *s0 = ring + ((size_t)(q % RS) * 9) * WI + col;
                const float *sm = ring + ((size_t)((q - 1) % RS) * 9) * WI + col;
                const float *sp = ring + ((size_t)((q + 1) % RS) * 9) * WI + col;
                float4 o[9];
                const float4 v0 = *reinterpret_cast<const float4 *>(s0 + 0 * WI);
                const float4 v1 = *reinterpret_cast<const float4 *>(s0 + 1 * WI);
                const float4 v2 = *reinterpret_cast<const float4 *>(sm + 2 * WI);
                const float4 v3 = *reinterpret_cast<const float4 *>(s0 + 3 * WI);
                const float4 v4 = *reinterpret_cast<const float4 *>(sp + 4 * WI);
                const float4 v5 = *reinterpret_cast<const float4 *>(sm + 5 * WI);
                const float4 v6 = *reinterpret_cast<const float4 *>(sm + 6 * WI);
                const float4 v7 = *reinterpret_cast<const float4 *>(sp + 7 * WI);
                const float4 v8 = *reinterpret_cast<const float4 *>(sp + 8 * WI);
                float l1 = 0.f, l5 = 0.f, l8 = 0.f, r3 = 0.f, r6 = 0.f, r7 = 0.f;
                if (lane == 0 && col > 0) {
                    l1 = s0[1 * WI - 1];
                    l5 = sm[5 * WI - 1];
                    l8 = sp[8 * WI - 1];
                }
                if (lane == 31 && col + 4 < WI) {
                    r3 = s0[3 * WI + 4];
                    r6 = sm[6 * WI + 4];
                    r7 = sp[7 * WI + 4];
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
                float mac[4][3];
                const unsigned hm = collide_quad<ALB_QUAD_GB>(o, p.tau, p.inv_tau, p.inv_tau_lo, DIAG ? mac : nullptr);
                if (st) {
                    hits += __popc(hm);
                    float *d = p.dst + (size_t)j * p.pitch + gx;
#pragma unroll
                    for (int i = 0; i < 9; i++) ST4(d + i * plane, o[i]);
                    if (DIAG) {
#pragma unroll
                        for (int k = 0; k < 4; k++) diag_cell(p, dl, mac[k][0], mac[k][1], mac[k][2]);
                    }
                }
            }
            tf = tf1;
            tf1 = tf2;
            __syncthreads();
        }
        if (DIAG) diag_flush<false>(p, dl, lane);
    }
    if (hits && p.clamp_hits) atomicAdd(p.clamp_hits, (unsigned long long)hits);
}


// ---- self-test of div_pair() against true division (tests/test_gpu_div.py) -----------------------
// Counter-based generator; operand regimes: lattice-like values, wide exponent ranges, numerators
// placed within a few ulps of a rounding boundary of the quotient, zeros, and out-of-range values
// (which must be rejected, never silently wrong).
__device__ __forceinline__ unsigned long long splitmix(unsigned long long &x) {
    unsigned long long z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void div_selftest_kernel(unsigned long long seed, int iters, unsigned long long *out3) {
    unsigned long long st = seed + 0x632BE59BD9B4E019ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
    unsigned long long checked = 0, accepted = 0, wrong = 0;
    for (int it = 0; it < iters; it++) {
        const unsigned long long a = splitmix(st), b = splitmix(st);
        const int regime = (int)(a >> 61);
        float r, jx, jy;
        const unsigned mant_r = (unsigned)b & 0x7fffffu, mant_x = (unsigned)(b >> 23) & 0x7fffffu, mant_y = (unsigned)(a >> 8) & 0x7fffffu;
        const unsigned sx = ((unsigned)(a >> 40) & 1u) << 31, sy = ((unsigned)(a >> 41) & 1u) << 31;
        if (regime <= 2) {            // lattice-like: rho in [0.25, 4), |j| in [2^-32, 1)
            r = __uint_as_float(((125u + (unsigned)(b >> 50) % 4u) << 23) | mant_r);
            jx = __uint_as_float(sx | ((95u + (unsigned)(a >> 32) % 32u) << 23) | mant_x);
            jy = __uint_as_float(sy | ((95u + (unsigned)(a >> 48) % 32u) << 23) | mant_y);
        } else if (regime <= 4) {     // the whole accepted range and beyond it on both sides
            r = __uint_as_float(((80u + (unsigned)(b >> 50) % 96u) << 23) | mant_r);
            jx = __uint_as_float(sx | ((60u + (unsigned)(a >> 32) % 116u) << 23) | mant_x);
            jy = __uint_as_float(sy | ((60u + (unsigned)(a >> 48) % 116u) << 23) | mant_y);
        } else if (regime <= 6) {     // numerator = q*r moved by -2..2 ulps: quotients next to rounding boundaries
            r = __uint_as_float(((126u + (unsigned)(b >> 50) % 2u) << 23) | mant_r);
            const float q1 = __uint_as_float(sx | ((100u + (unsigned)(a >> 32) % 28u) << 23) | mant_x);
            const float q2 = __uint_as_float(sy | ((100u + (unsigned)(a >> 48) % 28u) << 23) | (mant_y | 1u));
            jx = __uint_as_float(__float_as_uint(__fmul_rn(q1, r)) + (unsigned)(a >> 20) % 5u - 2u);
            jy = __uint_as_float(__float_as_uint(__fmul_rz(q2, r)) + (unsigned)(a >> 24) % 5u - 2u);
        } else {                      // special values
            const float sp[8] = {0.0f, -0.0f, 1.0f, INFINITY, NAN, 1e-45f, 3e38f, -1.0f};
            r = (b >> 60) & 1 ? sp[(b >> 40) & 7] : __uint_as_float((127u << 23) | mant_r);
            jx = sp[(a >> 32) & 7];
            jy = (a >> 36) & 1 ? sp[(a >> 44) & 7] : __uint_as_float(sy | (120u << 23) | mant_y);
        }
        float vx, vy;
        const bool nums_ok = div_pair(jx, jy, r, vx, vy);
        const bool ok = quad_accept(r, vx * vx + vy * vy, nums_ok);
        checked++;
        if (ok) {
            accepted++;
            const float tx = __fdiv_rn(jx, r), ty = __fdiv_rn(jy, r);
            if (__float_as_uint(tx) != __float_as_uint(vx) || __float_as_uint(ty) != __float_as_uint(vy)) wrong++;
        }
    }
    atomicAdd(out3 + 0, checked);
    atomicAdd(out3 + 1, accepted);
    atomicAdd(out3 + 2, wrong);
}

}  // namespace

cudaError_t launch_div_selftest(unsigned long long seed, int nblocks, int iters, unsigned long long *d_out3, cudaStream_t s) {
    div_selftest_kernel<<<nblocks, 256, 0, s>>>(seed, iters, d_out3);
    return cudaGetLastError();
}

cudaError_t launch_copy_tasks(const StepParams &p, cudaStream_t s) {
    if (p.ngen <= 0) return cudaSuccess;
    copy_tasks_kernel<<<(p.ngen + TASKS_PER_BLOCK - 1) / TASKS_PER_BLOCK, BLOCK_THREADS, 0, s>>>(p);
    return cudaGetLastError();
}

int step2_strip_width() { return 128 * ALB_S2_K; }

void step2_plan(Step2Params &p, int nsm) {
    constexpr int WI = 128 * ALB_S2_K;
    const int wo_max = WI - 8;
    p.nstrips = (p.pitch + wo_max - 1) / wo_max;
    int wo = (p.pitch + p.nstrips - 1) / p.nstrips;
    wo = (wo + 7) / 8 * 8;                 // full 32-byte sectors per strip where possible
    if (wo > wo_max) wo = wo_max;
    while ((long long)(p.nstrips - 1) * wo >= p.pitch) p.nstrips--;   // rounding up may have emptied the last strip
    p.wo = wo;
    const int rows = p.nyl - 2;            // rows 2 .. nyl-1 can be deep
    static int hs_env = -1;
    if (hs_env < 0) {
        const char *e = getenv("AEROLAB_LBM_S2_HS");
        hs_env = e ? atoi(e) : 0;
    }
    if (rows <= 0) {
        p.hs = 1;
        p.ntiles = 0;
        return;
    }
    if (hs_env > 0) {
        p.hs = hs_env;
    } else {
        // Every tile costs about (rows + 6) row-group times (two recomputed rows, pipeline fill and
        // drain) and one CTA runs per SM, so the step takes ceil(tiles / SMs) * (hs + 6): pick the
        // segment height in [ALB_S2_HS_MIN, ALB_S2_HS_MAX] that minimises it.  Taller segments are not
        // better per se: the list-driven passes on the aux stream only get SMs when a tile retires
        // (measured on 32768x16384: 128..256 rows 125 GLUPS, 443 rows 123, 1024 rows 115).
        long long best_cost = -1;
        int best = rows < ALB_S2_HS_MAX ? rows : ALB_S2_HS_MAX;
        for (int nsegs = (rows + ALB_S2_HS_MAX - 1) / ALB_S2_HS_MAX; nsegs <= rows; nsegs++) {
            const int hs = (rows + nsegs - 1) / nsegs;
            if (hs < ALB_S2_HS_MIN && best_cost >= 0) break;
            const long long tiles = (long long)p.nstrips * ((rows + hs - 1) / hs);
            const long long cost = ((tiles + nsm - 1) / nsm) * (hs + 6);
            if (best_cost < 0 || cost < best_cost) {
                best_cost = cost;
                best = hs;
            }
        }
        p.hs = best;
    }
    const int nsegs = (rows + p.hs - 1) / p.hs;
    p.ntiles = p.nstrips * nsegs;
}

cudaError_t launch_step2(const Step2Params &p, cudaStream_t s) {
    if (p.ntiles <= 0) return cudaSuccess;
    constexpr int RB = ALB_S2_RB, K = ALB_S2_K;
    constexpr size_t smem = sizeof(float) * ((size_t)(2 * RB + 2) * 9 * 128 * K + (ALB_S2_ASYNC ? (size_t)RB * K * (9 * (4 + 128 + 4) + 2) : 0));
    static bool configured[64] = {};       // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(step2_kernel<RB, K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(step2_kernel<RB, K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    if (p.diag) step2_kernel<RB, K, true><<<p.ntiles, 2 * RB * K * 32, smem, s>>>(p);
    else step2_kernel<RB, K, false><<<p.ntiles, 2 * RB * K * 32, smem, s>>>(p);
    return cudaGetLastError();
}

}  // namespace alb

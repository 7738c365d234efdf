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

#include "alb_common.cuh"

namespace cg = cooperative_groups;

namespace alb {

namespace {

constexpr unsigned FULL = 0xffffffffu;

// Tuning knobs (defaults chosen from B200 measurements, see DESIGN.md / profiles/):
//   ALB_LD_HINT  0 ld.global.nc   1 ld.global.cs (evict first)   2 ld.global.nc.L1::no_allocate
//   ALB_ST_HINT  0 st.global      1 st.global.cs (evict first)
//   ALB_FAST_MINBLOCKS  resident CTAs per SM the fast kernel is compiled for
#ifndef ALB_LD_HINT
#define ALB_LD_HINT 0
#endif
#ifndef ALB_ST_HINT
#define ALB_ST_HINT 0
#endif
//   ALB_EDGE_IN_FAST  (alb_common.cuh) inlet/outlet cells of otherwise all-fluid tasks patched in the fast kernel
//   ALB_DIAG_MINBLOCKS  resident CTAs per SM the DIAG variant of the fast kernel is compiled for
#ifndef ALB_DIAG_MINBLOCKS
#define ALB_DIAG_MINBLOCKS 4
#endif
#ifndef ALB_FAST_MINBLOCKS
#define ALB_FAST_MINBLOCKS 4
#endif

// ALB_DEBUG_BOUNDS=1 (compute-sanitizer is not available on the pool): every population load and
// store of the step kernels is checked against the source / destination allocation; a violation
// prints the address and traps, which the C ABI reports as a CUDA error.
#ifndef ALB_DEBUG_BOUNDS
#define ALB_DEBUG_BOUNDS 0
#endif
#if ALB_DEBUG_BOUNDS
// the kernels keep the bases in locals named src / dst_base / plane
#define ALB_CHECK_SRC(ptr, n) alb_check((ptr), (n), src, 9 * plane, "load")
#define ALB_CHECK_DST(ptr, n) alb_check((ptr), (n), dst_base, 9 * plane, "store")
__device__ __noinline__ void alb_check(const float *ptr, int n, const float *base, size_t len, const char *what) {
    if (ptr < base || ptr + n > base + len || (n == 4 && (reinterpret_cast<uintptr_t>(ptr) & 15))) {
        printf("alb bounds violation: %s of %d floats at offset %lld (allocation %llu floats)\n", what, n,
               (long long)(ptr - base), (unsigned long long)len);
        __trap();
    }
}
#else
#define ALB_CHECK_SRC(ptr, n) ((void)0)
#define ALB_CHECK_DST(ptr, n) ((void)0)
#endif
#define LD4(ptr) (ALB_CHECK_SRC((ptr), 4), ld4(ptr))
#define LD1(ptr) (ALB_CHECK_SRC((ptr), 1), __ldg(ptr))
#define LD1CG(ptr) (ALB_CHECK_SRC((ptr), 1), __ldcg(ptr))
#define ST4(ptr, v) (ALB_CHECK_DST((ptr), 4), st4((ptr), (v)))

__device__ __forceinline__ float4 ld4(const float *p) {
#if ALB_LD_HINT == 1
    return __ldcs(reinterpret_cast<const float4 *>(p));
#elif ALB_LD_HINT == 2
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
#else
    return __ldg(reinterpret_cast<const float4 *>(p));
#endif
}
__device__ __forceinline__ void st4(float *p, const float4 &v) {
#if ALB_ST_HINT == 1
    __stcs(reinterpret_cast<float4 *>(p), v);
#else
    *reinterpret_cast<float4 *>(p) = v;
#endif
}

// x / tau, correctly rounded (== IEEE division), for the uniform divisor tau.
// rcp = RN(1/tau) and rcp_lo = RN(1/tau - rcp) are computed once on the host.  x*(rcp + rcp_lo),
// rounded once by the FMA, is a faithful estimate of the quotient; by Markstein's theorem one
// correction with the exact residual (FMA) and the correctly rounded reciprocal then yields
// RN(x/tau).  4 instructions instead of the ~10 + slow path of the generic division.  (Operands
// here are differences of populations: 0 or >= 2^-30 in magnitude, far from underflow.)  Checked
// exhaustively against true division in tests/test_div_by_tau.py.
__device__ __forceinline__ float div_by_tau(float x, float tau, float rcp, float rcp_lo) {
    const float t = __fmul_rn(x, rcp_lo);
    float q = __fmaf_rn(x, rcp, t);
    const float r = __fmaf_rn(-tau, q, x);
    q = __fmaf_rn(r, rcp, q);
    return q;
}

struct Moments {
    float rho, ux, uy;
    bool hit;
};

// HTML:335-350: moments of the streamed populations, then the stability clamps.
__device__ __forceinline__ Moments moments_clamped(const float (&f)[9]) {
    Moments m;
    float rho = f[0];
    rho = rho + f[1];
    rho = rho + f[2];
    rho = rho + f[3];
    rho = rho + f[4];
    rho = rho + f[5];
    rho = rho + f[6];
    rho = rho + f[7];
    rho = rho + f[8];
    float ux = (f[1] + f[5] + f[8] - f[3] - f[6] - f[7]) / rho;
    float uy = (f[2] + f[5] + f[6] - f[4] - f[7] - f[8]) / rho;
    const float uMax = 0.35f, rhoMin = 0.5f, rhoMax = 2.0f;
    float rc = fminf(fmaxf(rho, rhoMin), rhoMax);
    m.hit = (rc != rho);
    float spd2 = ux * ux + uy * uy;
    if (spd2 > uMax * uMax) {
        float k = uMax / sqrtf(spd2);
        ux *= k;
        uy *= k;
        m.hit = true;
    }
    m.rho = rc;
    m.ux = ux;
    m.uy = uy;
    return m;
}

// plain moments of the outlet rule (HTML:305-307): no clamp
__device__ __forceinline__ void moments_plain(const float (&f)[9], float &rho, float &ux, float &uy) {
    rho = f[0] + f[1] + f[2] + f[3] + f[4] + f[5] + f[6] + f[7] + f[8];
    ux = (f[1] + f[5] + f[8] - f[3] - f[6] - f[7]) / rho;
    uy = (f[2] + f[5] + f[6] - f[4] - f[7] - f[8]) / rho;
}

// HTML:276-281 and 352-356.  feq_i = wt(i)*rho*(1+3eu+4.5eu*eu-1.5uu), left to
// right; opposite directions share 3*eu and 4.5*eu*eu (negating eu negates the
// first exactly and leaves the second unchanged, so sharing is bit-neutral).
__device__ __forceinline__ void collide(float (&f)[9], const Moments &m, float tau, float rcp, float rcp_lo) {
    const float w0 = 4.0f / 9.0f, ws = 1.0f / 9.0f, wd = 1.0f / 36.0f;
    const float rho = m.rho, ux = m.ux, uy = m.uy;
    const float uu = ux * ux + uy * uy;
    const float c15 = 1.5f * uu;
    const float wr0 = w0 * rho, wrs = ws * rho, wrd = wd * rho;
    {
        float eq = wr0 * (1.0f - c15);
        f[0] = f[0] - div_by_tau(f[0] - eq, tau, rcp, rcp_lo);
    }
#define ALB_PAIR(A, B, EU, WR)                                      \
    {                                                               \
        const float eu = (EU);                                      \
        const float t1 = 3.0f * eu;                                 \
        const float t2 = (4.5f * eu) * eu;                          \
        const float ea = (WR) * (((1.0f + t1) + t2) - c15);         \
        const float eb = (WR) * (((1.0f - t1) + t2) - c15);         \
        f[A] = f[A] - div_by_tau(f[A] - ea, tau, rcp, rcp_lo);              \
        f[B] = f[B] - div_by_tau(f[B] - eb, tau, rcp, rcp_lo);              \
    }
    ALB_PAIR(1, 3, ux, wrs)
    ALB_PAIR(2, 4, uy, wrs)
    ALB_PAIR(5, 7, ux + uy, wrd)
    ALB_PAIR(6, 8, uy - ux, wrd)
#undef ALB_PAIR
}

__device__ __forceinline__ float comp(const float4 &v, int k) {
    return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w));
}
__device__ __forceinline__ void setc(float4 &v, int k, float a) {
    if (k == 0) v.x = a;
    else if (k == 1) v.y = a;
    else if (k == 2) v.z = a;
    else v.w = a;
}

// populations arriving from x-1: own aligned vector shifted right by one cell
__device__ __forceinline__ float4 from_left(const float4 &v, float edge, int lane) {
    float t = __shfl_up_sync(FULL, v.w, 1);
    if (lane == 0) t = edge;
    return make_float4(t, v.x, v.y, v.z);
}
// populations arriving from x+1
__device__ __forceinline__ float4 from_right(const float4 &v, float edge, int lane) {
    float t = __shfl_down_sync(FULL, v.x, 1);
    if (lane == 31) t = edge;
    return make_float4(v.y, v.z, v.w, t);
}

constexpr int MODE_STEP = 0, MODE_MACRO = 1;

// first thread of a step: commit the previous step's momentum-exchange sums, clear that accumulator
__device__ __forceinline__ void me_begin_step(MeState *m, int parity) {
    const int prev = parity ^ 1;
    if (m->pending) {
        const long long c = m->count;
        m->ring[c % ME_RING][0] = m->acc[prev][0];
        m->ring[c % ME_RING][1] = m->acc[prev][1];
        m->count = c + 1;
    }
    m->acc[prev][0] = 0;
    m->acc[prev][1] = 0;
    m->pending = 1;
}

// ---- fused diagnostics of the macro pass (HTML:596-614 statistics, HTML:649-700 faces) ----------
struct DiagLocal {
    float rmin = INFINITY, rmax = -INFINITY;
    float m2f = -1.0f;       // fp32 pre-filter: largest fp32 ux^2+uy^2 among the cells accepted so far
    double m2 = -1.0;        // largest ux^2+uy^2 among cells with s < 4
    float bux = 0.f, buy = 0.f;
    long long fx = 0, fy = 0;
    unsigned surf = 0, rev = 0;
};

__device__ __forceinline__ double speed_ratio(float ux, float uy, double U0) {
    return hypot(__ddiv_rn((double)ux, U0), __ddiv_rn((double)uy, U0));   // Math.hypot(ux/U0, uy/U0)
}

// One non-solid lattice cell.  s is monotone in ux^2+uy^2 (exact in double), so only the arg-max
// candidate ever needs the hypot; cells within 1e-9 of the s < 4 cut are decided exactly.
__device__ __forceinline__ void diag_cell(const StepParams &p, DiagLocal &d, float rho, float ux, float uy) {
    if (rho >= p.rho_lo && rho <= p.rho_hi) {
        d.rmin = fminf(d.rmin, rho);
        d.rmax = fmaxf(d.rmax, rho);
    }
    // fp32 pre-filter (relative error of m2f < 2e-7): a cell can only be the arg-max if its fp32
    // value is within 1e-6 of the largest fp32 value seen so far; everything else skips the fp64 part
    const float m2f = ux * ux + uy * uy;
    if (!(m2f >= d.m2f * (1.0f - 1e-6f)) || m2f > p.m2f_cap) return;   // also drops NaN and s >= 4 for sure
    const double m2 = __dadd_rn(__dmul_rn((double)ux, (double)ux), __dmul_rn((double)uy, (double)uy));
    if (m2 > d.m2 && m2 < p.m2_hi) {
        if (m2 >= p.m2_lo && !(speed_ratio(ux, uy, p.U0d) < 4.0)) return;
        d.m2f = fmaxf(d.m2f, m2f);   // only ACCEPTED cells (s < 4) may raise the pre-filter level
        d.m2 = m2;
        d.bux = ux;
        d.buy = uy;
    }
}

// faces of a non-solid cell: bit i-1 of `links` (i = 1..4) says the cell at x - e_i is solid
__device__ __forceinline__ void diag_faces(DiagLocal &d, unsigned links, float rho, float ux) {
    const unsigned faces = links & 0xfu;
    if (!faces) return;
    const long long q = __double2ll_rn((double)rho * 0x1p40);
    const int n = __popc(faces);
    if (faces & 1u) d.fx -= q;   // solid at x-1: force on the body points to -x
    if (faces & 4u) d.fx += q;   // solid at x+1
    if (faces & 2u) d.fy -= q;   // solid at y-1
    if (faces & 8u) d.fy += q;   // solid at y+1
    d.surf += n;
    if (ux < 0.0f) d.rev += n;
}

__device__ __forceinline__ void atomic_min_float(float *a, float v) {
    int *ai = reinterpret_cast<int *>(a);
    int old = *ai;
    while (v < __int_as_float(old)) {
        const int assumed = old;
        old = atomicCAS(ai, assumed, __float_as_int(v));
        if (old == assumed) break;
    }
}
__device__ __forceinline__ void atomic_max_float(float *a, float v) {
    int *ai = reinterpret_cast<int *>(a);
    int old = *ai;
    while (v > __int_as_float(old)) {
        const int assumed = old;
        old = atomicCAS(ai, assumed, __float_as_int(v));
        if (old == assumed) break;
    }
}

// warp tree, then at most a handful of atomics per warp -- and none at all once the global
// extrema have settled (plain-load pre-check)
// FACES = false for tasks that cannot have fluid/solid faces (all-fluid, all-equilibrium): the
// four face sums are known to be zero and are left out of the shuffle tree.
template <bool FACES = true>
__device__ __forceinline__ void diag_flush(const StepParams &p, DiagLocal &d, int lane) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        d.rmin = fminf(d.rmin, __shfl_xor_sync(FULL, d.rmin, s));
        d.rmax = fmaxf(d.rmax, __shfl_xor_sync(FULL, d.rmax, s));
        const double om = __shfl_xor_sync(FULL, d.m2, s);
        const float ox = __shfl_xor_sync(FULL, d.bux, s), oy = __shfl_xor_sync(FULL, d.buy, s);
        if (om > d.m2) { d.m2 = om; d.bux = ox; d.buy = oy; }
        if (FACES) {
            d.fx += __shfl_xor_sync(FULL, d.fx, s);
            d.fy += __shfl_xor_sync(FULL, d.fy, s);
            d.surf += __shfl_xor_sync(FULL, d.surf, s);
            d.rev += __shfl_xor_sync(FULL, d.rev, s);
        }
    }
    if (lane != 0) return;
    DiagAcc *g = p.diag + ((blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) & (DIAG_SLOTS - 1));
    // pre-checks through L1 (ld.global.ca): a stale value only makes the filter less tight, the
    // atomics below re-check against the true value
    if (d.rmin < __ldca(&g->rho_min)) atomic_min_float(&g->rho_min, d.rmin);
    if (d.rmax > __ldca(&g->rho_max)) atomic_max_float(&g->rho_max, d.rmax);
    if (d.m2 >= 0.0) {
        const double cur = __longlong_as_double((long long)__ldca(&g->m2max_bits));
        if (d.m2 >= cur * (1.0 - 1e-12)) {
            const double sr = speed_ratio(d.bux, d.buy, p.U0d);
            if (sr < 4.0) {
                atomicMax(&g->smax_bits, (unsigned long long)__double_as_longlong(sr));
                atomicMax(&g->m2max_bits, (unsigned long long)__double_as_longlong(d.m2));
            }
        }
    }
    if (FACES && d.surf) {
        atomicAdd(reinterpret_cast<unsigned long long *>(&g->fx), (unsigned long long)d.fx);
        atomicAdd(reinterpret_cast<unsigned long long *>(&g->fy), (unsigned long long)d.fy);
        atomicAdd(&g->surf, (unsigned long long)d.surf);
        atomicAdd(&g->rev, (unsigned long long)d.rev);
    }
}

// KIND_FAST: every task of the slab, but tasks of class TC_GENERAL are skipped -- 64 registers,
// 4 CTAs per SM.  KIND_GENERAL: only the compacted list of TC_GENERAL tasks (tasks that mix cell
// types or touch the body), with the per-cell patching code.  KIND_UNIFIED: every task, both
// paths in one launch -- for lattices so small that the step is launch-latency bound and
// occupancy is irrelevant.
constexpr int KIND_FAST = 0, KIND_GENERAL = 1, KIND_UNIFIED = 2;

// DIAG (step mode): also accumulate the autoscale statistics and pressure-face sums of the state
// being WRITTEN (its rho/ux/uy are computed here anyway) -- used for the last step of a batch.
template <int MODE, int KIND, bool DIAG = false>
__global__ void __launch_bounds__(BLOCK_THREADS, KIND == KIND_FAST ? (DIAG ? ALB_DIAG_MINBLOCKS : ALB_FAST_MINBLOCKS) : 2)
step_kernel(const __grid_constant__ StepParams p) {
    const int lane = threadIdx.x & 31;
    int task = blockIdx.x * TASKS_PER_BLOCK + (threadIdx.x >> 5);
    if (KIND != KIND_GENERAL && MODE == MODE_STEP && blockIdx.x == 0 && threadIdx.x == 0 && p.me)
        me_begin_step(p.me, p.parity);   // the previous step of this handle has completed (stream order)
    constexpr bool from_list = KIND == KIND_GENERAL;
    if (from_list) {
        if (task >= p.ngen) return;
        task = p.gen_list[task];
    } else if (task >= p.ntasks) {
        return;
    }
    const int j = task / p.tpr + 1;          // local row (0 is the lower ghost row)
    const int s = task - (j - 1) * p.tpr;
    const int x0 = s * TASK_CELLS + lane * 4;
    const size_t c = (size_t)j * p.pitch + x0;
    const size_t plane = p.plane;
    const int cls = from_list ? (int)TC_GENERAL : (int)p.tclass[(size_t)j * p.tpr + s];   // warp-uniform
    if (KIND == KIND_FAST && cls == TC_GENERAL) return;
    const bool GENERAL = KIND != KIND_FAST && cls == TC_GENERAL;   // warp-uniform; compile-time false for KIND_FAST
    const float *__restrict__ src = p.src;
    [[maybe_unused]] float *const dst_base = p.dst;

    float4 o[9];

    if (!from_list && cls == TC_EQUIL) {
        // HTML:314-322: whole task is inlet/top/bottom equilibrium at (1, U0, 0)
        if (MODE == MODE_STEP) {
#pragma unroll
            for (int i = 0; i < 9; i++) {
                const float v = p.feq0[i];
                ST4(p.dst + i * plane + c, make_float4(v, v, v, v));
            }
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
    if (!from_list && cls == TC_SOLID) {
        // HTML:287-294: solid cells swap every population with its opposite
        if (MODE == MODE_STEP) {
            const int opp[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
#pragma unroll
            for (int i = 0; i < 9; i++) o[i] = LD4(src + opp[i] * plane + c);
#pragma unroll
            for (int i = 0; i < 9; i++) ST4(p.dst + i * plane + c, o[i]);
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
        if (MODE == MODE_STEP) collide(f, m, p.tau, p.inv_tau, p.inv_tau_lo);
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

        // halo push: my edge rows go straight into the neighbours' ghost rows
        // (peer memory over NVLink, or the same GPU for in-process slabs)
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

        if (DIAG) {
            if (KIND == KIND_FAST) diag_flush<false>(p, dl, lane);
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
        if (hits && p.clamp_hits) atomicAdd(p.clamp_hits, (unsigned long long)hits);
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
                collide(f, m, p.tau, p.inv_tau, p.inv_tau_lo);
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

}  // namespace

// The fast kernel and the general kernel of one step read the same source state and write
// disjoint cells, so the caller may run them concurrently on two streams.
cudaError_t launch_step_fast(const StepParams &p, cudaStream_t s) {
    const int nblocks = (p.ntasks + TASKS_PER_BLOCK - 1) / TASKS_PER_BLOCK;
    if (p.diag) step_kernel<MODE_STEP, KIND_FAST, true><<<nblocks, BLOCK_THREADS, 0, s>>>(p);
    else step_kernel<MODE_STEP, KIND_FAST><<<nblocks, BLOCK_THREADS, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_step_general(const StepParams &p, cudaStream_t s) {
    if (p.ngen == 0) return cudaSuccess;
    const int gblocks = (p.ngen + TASKS_PER_BLOCK - 1) / TASKS_PER_BLOCK;
    if (p.diag) step_kernel<MODE_STEP, KIND_GENERAL, true><<<gblocks, BLOCK_THREADS, 0, s>>>(p);
    else step_kernel<MODE_STEP, KIND_GENERAL><<<gblocks, BLOCK_THREADS, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_step_unified(const StepParams &p, cudaStream_t s) {
    const int nblocks = (p.ntasks + TASKS_PER_BLOCK - 1) / TASKS_PER_BLOCK;
    if (p.diag) step_kernel<MODE_STEP, KIND_UNIFIED, true><<<nblocks, BLOCK_THREADS, 0, s>>>(p);
    else step_kernel<MODE_STEP, KIND_UNIFIED><<<nblocks, BLOCK_THREADS, 0, s>>>(p);
    return cudaGetLastError();
}

// Largest number of cells the persistent small-lattice kernel can own on this device (all CTAs
// must be co-resident for the grid barrier); 0 when cooperative launches are unsupported.
int small_lattice_capacity(int device) {
    int coop = 0, sms = 0, per_sm = 0;
    if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device) != cudaSuccess || !coop) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, small_lattice_kernel, BLOCK_THREADS, 0) != cudaSuccess)
        return 0;
    return sms * per_sm * BLOCK_THREADS;
}

cudaError_t launch_small_lattice(const StepParams &p, float *f0, float *f1, int cur, int nsteps, cudaStream_t s) {
    const int ncell = p.nx * p.nyl;
    const int nblocks = (ncell + BLOCK_THREADS - 1) / BLOCK_THREADS;
    StepParams pp = p;
    void *args[] = {&pp, &f0, &f1, &cur, &nsteps};
    return cudaLaunchCooperativeKernel((const void *)small_lattice_kernel, dim3(nblocks), dim3(BLOCK_THREADS), args, 0, s);
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

}  // namespace alb

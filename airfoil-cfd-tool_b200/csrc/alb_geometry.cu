// Subsystem (a): airfoil coordinates -> solid mask on the lattice.
//
//   host_rotate_panelise : rotate() + panelise()       HTML:133-157 (float64, libm)
//   raster_kernel        : rasterMask()                HTML:160-182 (float64 on the GPU)
//   build_info_kernel    : per-cell type + link bits   (derived; HTML:287-334 branch tests)
//   build_tclass_kernel  : per-warp-task class
//
// "HTML:n" = pages/airfoil_flow_lbm_aerolab.html of the reference.  The mask
// must be bit-exact: the scan conversion reproduces the reference's quirks --
// y sampled at cell centres but x at integer node positions, the polygon is
// not closed, crossings are sorted numerically and an unpaired last crossing
// is dropped.  All float64 operations are individually rounded (-fmad=false
// plus explicit _rn intrinsics), in the reference's operand order.
#include <math.h>

#include "alb_common.cuh"

namespace alb {

// HTML:133-157.  xy: npts (x, y) pairs.  xp/yp: ALB_NPANEL + 1 nodes.
void host_rotate_panelise(const double *xy, int npts, double alpha_deg, double *xp, double *yp) {
    const int NP = ALB_NPANEL;
    // rotate(): a = -aDeg*PI/180 about (0.25, 0)
    volatile double a = -alpha_deg * M_PI / 180;
    const double ca = cos(a), sa = sin(a);
    const double px = 0.25, py = 0;
    std::vector<double> xs(npts), ys(npts), arc(npts);
    for (int i = 0; i < npts; i++) {
        const double dx = xy[2 * i] - px, dy = xy[2 * i + 1] - py;
        volatile double t1 = dx * ca, t2 = dy * sa, t3 = dx * sa, t4 = dy * ca;
        volatile double rx = px + t1;
        volatile double ry = py + t3;
        xs[i] = rx - t2;
        ys[i] = ry + t4;
    }
    // panelise(): cumulative arc length, cosine spacing, linear search
    arc[0] = 0;
    for (int i = 1; i < npts; i++) arc[i] = arc[i - 1] + hypot(xs[i] - xs[i - 1], ys[i] - ys[i - 1]);
    const double L = arc[npts - 1];
    for (int i = 0; i <= NP; i++) {
        volatile double ang = M_PI * i / NP;
        volatile double half = L * 0.5;
        volatile double om = 1 - cos(ang);
        const double s = half * om;
        int j = 0;
        while (j < npts - 2 && arc[j + 1] < s) j++;
        volatile double den = arc[j + 1] - arc[j];
        den = den + 1e-12;
        const double t = (s - arc[j]) / den;
        volatile double mx = (xs[j + 1] - xs[j]) * t;
        volatile double my = (ys[j + 1] - ys[j]) * t;
        xp[i] = xs[j] + mx;
        yp[i] = ys[j] + my;
    }
}

namespace {

constexpr int RASTER_THREADS = 128;
constexpr int MAX_CROSS = 1024;   // crossings kept per row (a 161-node outline has <= 160)

// One CTA per lattice row.  HTML:163-179.
__global__ void __launch_bounds__(RASTER_THREADS)
raster_kernel(const double *__restrict__ xp, const double *__restrict__ yp, int n, uint8_t *mask,
              int pitch, int nx, int ny_global, int gy_first, int nrows) {
    __shared__ double xs_raw[MAX_CROSS];
    __shared__ double xs[MAX_CROSS];
    __shared__ int span0[MAX_CROSS / 2], span1[MAX_CROSS / 2];
    __shared__ int count;
    const int j = blockIdx.x;
    if (j >= nrows) return;
    const int iy = gy_first + j;
    uint8_t *row = mask + (size_t)j * pitch;
    if (iy < 0 || iy >= ny_global) {   // ghost row outside the lattice: fluid, never read
        for (int x = threadIdx.x; x < pitch; x += RASTER_THREADS) row[x] = 0;
        return;
    }
    if (threadIdx.x == 0) count = 0;
    __syncthreads();
    // wy = DY0 + (iy+0.5)/NY*(DY1-DY0)
    const double wy = __dadd_rn(DY0, __dmul_rn(__ddiv_rn((double)iy + 0.5, (double)ny_global), DY1 - DY0));
    for (int i = threadIdx.x; i < n - 1; i += RASTER_THREADS) {
        const double y1 = yp[i], y2 = yp[i + 1];
        if ((y1 > wy) != (y2 > wy)) {
            const double x1 = xp[i], x2 = xp[i + 1];
            // x1 + (x2-x1)*(wy-y1)/(y2-y1)
            const double num = __dmul_rn(__dsub_rn(x2, x1), __dsub_rn(wy, y1));
            const double xc = __dadd_rn(x1, __ddiv_rn(num, __dsub_rn(y2, y1)));
            const int k = atomicAdd(&count, 1);
            if (k < MAX_CROSS) xs_raw[k] = xc;
        }
    }
    __syncthreads();
    const int m = min(count, MAX_CROSS);
    // numeric ascending sort by ranking (equal values are interchangeable)
    for (int k = threadIdx.x; k < m; k += RASTER_THREADS) {
        const double v = xs_raw[k];
        int rank = 0;
        for (int q = 0; q < m; q++) {
            const double w = xs_raw[q];
            rank += (w < v) || (w == v && q < k);
        }
        xs[rank] = v;
    }
    __syncthreads();
    const int npairs = m / 2;   // unpaired last crossing dropped (HTML:174)
    for (int k = threadIdx.x; k < npairs; k += RASTER_THREADS) {
        // ix0 = ceil((xs[k]-DX0)/(DX1-DX0)*NX), ix1 = floor(...), clamped
        const double a = ceil(__dmul_rn(__ddiv_rn(__dsub_rn(xs[2 * k], DX0), DX1 - DX0), (double)nx));
        const double b = floor(__dmul_rn(__ddiv_rn(__dsub_rn(xs[2 * k + 1], DX0), DX1 - DX0), (double)nx));
        const double a2 = fmax(0.0, a);
        const double b2 = fmin((double)(nx - 1), b);
        // empty span when b2 < a2; values are within [0, nx-1] or the span is empty
        span0[k] = (a2 <= (double)(nx - 1)) ? (int)a2 : nx;
        span1[k] = (b2 >= 0.0) ? (int)b2 : -1;
    }
    __syncthreads();
    for (int x = threadIdx.x; x < pitch; x += RASTER_THREADS) {
        uint8_t v = 0;
        if (x < nx)
            for (int k = 0; k < npairs; k++)
                if (x >= span0[k] && x <= span1[k]) { v = 255; break; }
        row[x] = v;
    }
}

// Per-cell info word: bits 0..7 = "pull source x - e_i is solid" for i = 1..8
// (interior fluid cells only), bits 8..9 = cell type.  Padding cells (x >= nx)
// are typed equilibrium so that they only ever receive constants.
__global__ void build_info_kernel(const uint8_t *__restrict__ mask, uint16_t *__restrict__ info,
                                  int pitch, int nx, int ny_global, int gy_first, int nrows) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (x >= pitch || j >= nrows) return;
    const int gy = gy_first + j;
    const size_t c = (size_t)j * pitch + x;
    unsigned type, links = 0, pad = 0;
    if (x >= nx || gy < 0 || gy >= ny_global) {
        type = CT_EQUIL;
        pad = INFO_PAD;       // not a lattice cell: only ever receives constants, never counted
    } else if (mask[c]) {
        type = CT_SOLID;
    } else if (x == nx - 1) {
        type = CT_OUTLET;
    } else if (x == 0 || gy == ny_global - 1 || gy == 0) {
        type = CT_EQUIL;
    } else {
        type = CT_FLUID;
    }
    if (!pad && type != CT_SOLID && j >= 1 && j <= nrows - 2) {
        // "the cell at x - e_i is a solid lattice cell".  The step uses the bits of interior
        // fluid cells (bounce-back); the pressure-face force uses bits 0..3 of every non-solid cell.
        const int ex[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
        const int ey[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
#pragma unroll
        for (int i = 1; i < 9; i++) {
            const int xs = x - ex[i], gys = gy - ey[i];
            if (xs < 0 || xs >= nx || gys < 0 || gys >= ny_global) continue;
            if (mask[(size_t)(j - ey[i]) * pitch + xs]) links |= 1u << (i - 1);
        }
    }
    info[c] = (uint16_t)((type << INFO_TYPE_SHIFT) | links | pad);
}

// Also appends the TC_GENERAL tasks of the owned rows (1 .. nrows-2) to gen_list as
// (row - 1) * tpr + segment, the task numbering of the step kernel.
__global__ void build_tclass_kernel(const uint16_t *__restrict__ info, uint8_t *__restrict__ tclass,
                                    int *gen_list, int *gen_count, int pitch, int nrows) {
    const int lane = threadIdx.x & 31;
    const int tpr = pitch / TASK_CELLS;
    const int task = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (task >= tpr * nrows) return;
    const int j = task / tpr, s = task - j * tpr;
    const uint16_t *p = info + (size_t)j * pitch + s * TASK_CELLS + lane * 4;
    bool all_fluid = true, all_solid = true, all_equil = true;
    bool fluid_l = ALB_EDGE_IN_FAST && (s == 0), fluid_r = ALB_EDGE_IN_FAST && (s == tpr - 1);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const unsigned v = p[k];
        const bool is_fluid = (v == (CT_FLUID << 8));
        all_fluid &= is_fluid;
        all_solid &= (((v >> 8) & 3) == CT_SOLID);
        all_equil &= (v == (CT_EQUIL << 8));   // no solid neighbour, not padding
        // inlet cell = first cell of the row, outlet cell = last cell of the row (no padding)
        fluid_l &= (lane == 0 && k == 0) ? (v == (CT_EQUIL << 8)) : is_fluid;
        fluid_r &= (lane == 31 && k == 3) ? (v == (CT_OUTLET << 8)) : is_fluid;
    }
    all_fluid = __all_sync(0xffffffffu, all_fluid);
    all_solid = __all_sync(0xffffffffu, all_solid);
    all_equil = __all_sync(0xffffffffu, all_equil);
    fluid_l = __all_sync(0xffffffffu, fluid_l);
    fluid_r = __all_sync(0xffffffffu, fluid_r);
    if (lane == 0) {
        const int cls = all_fluid ? TC_FLUID : all_solid ? TC_SOLID : all_equil ? TC_EQUIL
                        : fluid_l ? TC_FLUID_L : fluid_r ? TC_FLUID_R : TC_GENERAL;
        tclass[task] = (uint8_t)cls;
        if (cls == TC_GENERAL && j >= 1 && j <= nrows - 2) gen_list[atomicAdd(gen_count, 1)] = (j - 1) * tpr + s;
    }
}


// ---- task sets of the two-steps-per-pass path (march2_kernel, alb_march.cu) ----------------------
// A cell is "plain" when its info word is 0: interior fluid, no solid pull source, not padding.
// A task is deep when every cell within one cell of it (its own 128 cells, the last cell of the
// task to its left, the first cell of the task to its right, in rows j-1, j, j+1) is plain, and
// the row is neither the first nor the last owned row (those need the neighbouring slab's
// intermediate state, or are the equilibrium border rows of a whole lattice).
// With a neighbouring slab below / above (lo_nb / hi_nb) the first / last TWO rows stay shallow, so
// that the fused kernel never reads a ghost row and needs no synchronisation with the neighbours.
// edges_nx > 0 (the lattice width, when march_edges_enabled() says so): the first and the last task of a
// row qualify as well when their one special cell is exactly the inlet cell x = 0 (equilibrium, no
// solid next to it) or the outlet cell x = nx-1 (ditto), everything before the outlet is plain and
// everything after it padding; the fused kernel patches that one cell (alb_march.cu) and never
// writes padding (which holds the equilibrium constants in every buffer anyway).  Without them a
// 2048-wide lattice leaves an eighth of its cells to the list-driven passes.
__device__ __forceinline__ bool edge_task_plain(const uint16_t *row, int special, unsigned special_info) {
    bool ok = true;
    for (int c = 0; c < special; c++) ok = ok && row[c] == 0u;
    ok = ok && row[special] == special_info;
    const unsigned after = special == 0 ? 0u : ((CT_EQUIL << INFO_TYPE_SHIFT) | INFO_PAD);
    for (int c = special + 1; c < TASK_CELLS; c++) ok = ok && row[c] == after;
    return ok;
}

__global__ void build_deep_kernel(const uint16_t *__restrict__ info, const uint8_t *__restrict__ tclass,
                                  uint8_t *__restrict__ deep, int pitch, int nrows, int lo_nb, int hi_nb, int edges_nx) {
    const int tpr = pitch / TASK_CELLS;
    const int task = blockIdx.x * blockDim.x + threadIdx.x;
    if (task >= tpr * nrows) return;
    const int j = task / tpr, s = task - j * tpr;
    const bool edges = edges_nx > 0;
    bool d = j >= 2 + lo_nb && j <= nrows - 3 - hi_nb && (edges || (s >= 1 && s <= tpr - 2));
    if (d) {
        for (int jj = j - 1; jj <= j + 1; jj++) {
            const uint16_t *row = info + (size_t)jj * pitch + s * TASK_CELLS;
            if (s == 0) d = d && edge_task_plain(row, 0, CT_EQUIL << INFO_TYPE_SHIFT);
            else if (s == tpr - 1) d = d && edge_task_plain(row, edges_nx - 1 - s * TASK_CELLS, CT_OUTLET << INFO_TYPE_SHIFT);
            else d = d && tclass[jj * tpr + s] == TC_FLUID;
            if (s > 0) d = d && row[-1] == 0;
            if (s < tpr - 1) d = d && row[TASK_CELLS] == 0;
        }
    }
    deep[task] = d ? 1 : 0;
}

// warp-aggregated append of (id | extra) for the lanes with pred set
__device__ __forceinline__ void list_append(int *list, int *count, bool pred, int value) {
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (pred) list[base + __popc(m & ((1u << lane) - 1))] = value;
}

// Pass 1 of the two-pass path produces the intermediate state of every task within one task of a
// shallow (= not deep) task; pass 2 advances the shallow tasks themselves.  Deep tasks that are
// in pass 1 only as neighbours carry LIST_NOHIT: the fused kernel counts their clamp hits.
// All-solid tasks are left out of both passes: a solid cell only swaps its own populations with
// their opposites (HTML:287-294) and nobody ever pulls from it, so after TWO steps it is back
// where it was -- such tasks (list 4) are just copied when the destination buffer does not hold
// their values already.  (Not the last two tasks of a row: the outlet cell copies the populations
// of its left neighbour whatever that is, HTML:301-312.)
__global__ void build_lists_kernel(const uint8_t *__restrict__ tclass, const uint8_t *__restrict__ deep,
                                   uint8_t *__restrict__ tflags, int *l1f, int *l1g, int *l2f, int *l2g, int *lsol,
                                   int *counts, int tpr, int nrows) {
    const int task = blockIdx.x * blockDim.x + threadIdx.x;   // blockDim is a multiple of 32: whole warps
    const bool valid = task < tpr * nrows;
    const int j = valid ? task / tpr : 0, s = valid ? task - j * tpr : 0;
    const bool owned = valid && j >= 1 && j <= nrows - 2;
    bool any_shallow = false, any_deep = false, is_deep = false, gen = false, solid = false;
    if (owned) {
        is_deep = deep[task] != 0;
        gen = tclass[task] == TC_GENERAL;
        solid = tclass[task] == TC_SOLID && s < tpr - 2;
        for (int jj = max(1, j - 1); jj <= min(nrows - 2, j + 1); jj++)
            for (int ss = max(0, s - 1); ss <= min(tpr - 1, s + 1); ss++) {
                const bool d = deep[jj * tpr + ss] != 0;
                any_deep |= d;
                any_shallow |= !d;
            }
    }
    if (valid) tflags[task] = (uint8_t)((is_deep ? TF_DEEP : 0u) | (any_deep ? TF_NEED : 0u));
    const int id = (j - 1) * tpr + s;
    list_append(l1f, counts + 0, owned && any_shallow && !gen && !solid, id | (is_deep ? LIST_NOHIT : 0));
    list_append(l1g, counts + 1, owned && any_shallow && gen, id | (is_deep ? LIST_NOHIT : 0));   // deep edge tasks
    list_append(l2f, counts + 2, owned && !is_deep && !gen && !solid, id);
    list_append(l2g, counts + 3, owned && !is_deep && gen, id);
    list_append(lsol, counts + 4, owned && solid, id);
}

}  // namespace

cudaError_t launch_raster(const double *d_xp, const double *d_yp, int n, uint8_t *mask, int pitch,
                          int nx, int ny_global, int gy_first, int nrows, cudaStream_t s) {
    raster_kernel<<<nrows, RASTER_THREADS, 0, s>>>(d_xp, d_yp, n, mask, pitch, nx, ny_global, gy_first, nrows);
    return cudaGetLastError();
}

cudaError_t launch_build_info(const uint8_t *mask, uint16_t *info, uint8_t *tclass, int *gen_list,
                              int *gen_count, int pitch, int nx, int ny_global, int gy_first, int nrows,
                              cudaStream_t s) {
    dim3 grid((pitch + 255) / 256, nrows);
    build_info_kernel<<<grid, 256, 0, s>>>(mask, info, pitch, nx, ny_global, gy_first, nrows);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(gen_count, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    const int ntask = (pitch / TASK_CELLS) * nrows;
    build_tclass_kernel<<<(ntask + 7) / 8, 256, 0, s>>>(info, tclass, gen_list, gen_count, pitch, nrows);
    return cudaGetLastError();
}

cudaError_t launch_build_lists(const uint16_t *info, const uint8_t *tclass, uint8_t *deep_tmp, uint8_t *tflags,
                               int *const lists[5], int *counts, int pitch, int nrows, int lo_nb, int hi_nb,
                               int edges_nx, cudaStream_t s) {
    const int tpr = pitch / TASK_CELLS, ntask = tpr * nrows;
    build_deep_kernel<<<(ntask + 255) / 256, 256, 0, s>>>(info, tclass, deep_tmp, pitch, nrows, lo_nb, hi_nb, edges_nx);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(counts, 0, 5 * sizeof(int), s);
    if (e != cudaSuccess) return e;
    build_lists_kernel<<<(ntask + 255) / 256, 256, 0, s>>>(tclass, deep_tmp, tflags, lists[0], lists[1], lists[2],
                                                           lists[3], lists[4], counts, tpr, nrows);
    return cudaGetLastError();
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

cudaError_t preload_geometry_kernels() {
    ALB_PRELOAD(raster_kernel);
    ALB_PRELOAD(build_info_kernel);
    ALB_PRELOAD(build_tclass_kernel);
    ALB_PRELOAD(build_deep_kernel);
    ALB_PRELOAD(build_lists_kernel);
    return cudaSuccess;
}

}  // namespace alb

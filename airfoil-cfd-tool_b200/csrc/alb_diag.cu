// Subsystem (c): diagnostics on the macroscopic fields.
//
//   stats_kernel   updateFieldsFromMacro()        HTML:596-614
//   forces_kernel  computeForces(): pressure faces + separation counts  HTML:649-700
//   render_kernel  RENDER_FS_SRC.main (+palettes) HTML:371-422
//   mass_kernel    total population sum (float64)
//   fill_kernel    equilibriumInitData()/initSim  HTML:474-500
//
// "HTML:n" = pages/airfoil_flow_lbm_aerolab.html.  Reductions are made
// deterministic by construction: a fixed grid of CTAs, each reducing a fixed
// set of cells in a fixed order (warp shuffle tree, then one thread sums the
// warps), one partial per CTA, summed on the host in CTA order.
#include <math.h>

#include "alb_common.cuh"

namespace alb {

namespace {

constexpr int DIAG_THREADS = 256;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, d));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, d));
    return v;
}

// HTML:596-614.  part[3*b + {0,1,2}] = {mx, cMin, cMax} of CTA b.
__global__ void __launch_bounds__(DIAG_THREADS)
stats_kernel(const uint8_t *__restrict__ mask, const float *__restrict__ rho,
             const float *__restrict__ ux, const float *__restrict__ uy, int pitch, int nx, int nyl,
             double U0, float *U, float *V, float *Cp, double *part) {
    __shared__ double sm[3][DIAG_THREADS / 32];
    double mx = 0.0, cmin = INFINITY, cmax = -INFINITY;
    const double cden = __dmul_rn(__dmul_rn(1.5, U0), U0);   // 1.5*U0*U0
    const size_t ncell = (size_t)nx * nyl;
    for (size_t t = (size_t)blockIdx.x * DIAG_THREADS + threadIdx.x; t < ncell;
         t += (size_t)gridDim.x * DIAG_THREADS) {
        const int y = (int)(t / nx), x = (int)(t - (size_t)y * nx);
        const size_t c = (size_t)(y + 1) * pitch + x;   // padded storage, ghost row below
        const size_t o = t;                             // dense output index
        if (mask[c]) {
            if (U) U[o] = NAN;
            if (V) V[o] = NAN;
            if (Cp) Cp[o] = NAN;
            continue;
        }
        const double u = __ddiv_rn((double)ux[c], U0), v = __ddiv_rn((double)uy[c], U0);
        const double cp = __ddiv_rn(__dsub_rn((double)rho[c], 1.0), cden);
        if (U) U[o] = (float)u;
        if (V) V[o] = (float)v;
        if (Cp) Cp[o] = (float)cp;
        const double s = hypot(u, v);
        if (s > mx && s < 4.0) mx = s;
        if (cp > -4.0 && cp < 1.2) {
            if (cp < cmin) cmin = cp;
            if (cp > cmax) cmax = cp;
        }
    }
    mx = warp_max(mx);
    cmin = warp_min(cmin);
    cmax = warp_max(cmax);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sm[0][w] = mx; sm[1][w] = cmin; sm[2][w] = cmax; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < DIAG_THREADS / 32; k++) {
            mx = fmax(mx, sm[0][k]);
            cmin = fmin(cmin, sm[1][k]);
            cmax = fmax(cmax, sm[2][k]);
        }
        part[3 * blockIdx.x + 0] = mx;
        part[3 * blockIdx.x + 1] = cmin;
        part[3 * blockIdx.x + 2] = cmax;
    }
}

// HTML:649-700, enumerated from the fluid side so that a slab only needs its
// own rho/ux rows plus the (static) mask ghost rows: for every non-solid cell
// and each of its four neighbours that is in the lattice and solid, the face
// contributes p = rho/3 along (solid - fluid).  part[4*b + {0,1,2,3}] =
// {fx, fy, surf, rev}.
__global__ void __launch_bounds__(DIAG_THREADS)
forces_kernel(const uint8_t *__restrict__ mask, const float *__restrict__ rho,
              const float *__restrict__ ux, int pitch, int nx, int ny_global, int gy_first, int nyl,
              double *part) {
    __shared__ double sm[4][DIAG_THREADS / 32];
    double fx = 0, fy = 0, surf = 0, rev = 0;
    const size_t ncell = (size_t)nx * nyl;
    for (size_t t = (size_t)blockIdx.x * DIAG_THREADS + threadIdx.x; t < ncell;
         t += (size_t)gridDim.x * DIAG_THREADS) {
        const int y = (int)(t / nx), x = (int)(t - (size_t)y * nx);
        const int j = y + 1, gy = gy_first + j;
        const size_t c = (size_t)j * pitch + x;
        if (mask[c]) continue;
        // neighbour offsets from the fluid cell to the candidate solid cell
        const int dx[4] = {-1, 0, 1, 0}, dy[4] = {0, -1, 0, 1};
        double p = 0;
        bool have = false;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int xs = x + dx[k], gys = gy + dy[k];
            if (xs < 0 || xs >= nx || gys < 0 || gys >= ny_global) continue;
            if (!mask[(size_t)(j + dy[k]) * pitch + xs]) continue;
            if (!have) { p = __ddiv_rn((double)rho[c], 3.0); have = true; }
            // JS: face (FACE_DX,FACE_DY) = fluid - solid = (-dx,-dy); f += p*(-FACE)
            fx += p * (double)dx[k];
            fy += p * (double)dy[k];
            surf += 1.0;
            if (ux[c] < 0.0f) rev += 1.0;
        }
    }
    fx = warp_sum(fx); fy = warp_sum(fy); surf = warp_sum(surf); rev = warp_sum(rev);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sm[0][w] = fx; sm[1][w] = fy; sm[2][w] = surf; sm[3][w] = rev; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < DIAG_THREADS / 32; k++) {
            fx += sm[0][k]; fy += sm[1][k]; surf += sm[2][k]; rev += sm[3][k];
        }
        part[4 * blockIdx.x + 0] = fx;
        part[4 * blockIdx.x + 1] = fy;
        part[4 * blockIdx.x + 2] = surf;
        part[4 * blockIdx.x + 3] = rev;
    }
}

__global__ void __launch_bounds__(DIAG_THREADS)
mass_kernel(const float *__restrict__ f, size_t plane, int pitch, int nx, int nyl, double *part) {
    __shared__ double sm[DIAG_THREADS / 32];
    double m = 0;
    const size_t ncell = (size_t)nx * nyl;
    for (size_t t = (size_t)blockIdx.x * DIAG_THREADS + threadIdx.x; t < ncell;
         t += (size_t)gridDim.x * DIAG_THREADS) {
        const int y = (int)(t / nx), x = (int)(t - (size_t)y * nx);
        const size_t c = (size_t)(y + 1) * pitch + x;
#pragma unroll
        for (int i = 0; i < 9; i++) m += (double)f[i * plane + c];
    }
    m = warp_sum(m);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sm[w] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < DIAG_THREADS / 32; k++) m += sm[k];
        part[blockIdx.x] = m;
    }
}

// Checksum of the population bit patterns, one 64-bit word per plane: the sum over the slab's
// owned cells of mix64(global cell index, bits) modulo 2^64.  Every term depends on WHERE the
// value sits on the global lattice, and integer sums are order independent, so the per-slab
// words of any decomposition add up to the word of the whole lattice (bench.py `check`,
// tests/test_gpu_large.py; numpy twin: aerolab_lbm.tunnel.state_hash_numpy).
__device__ __forceinline__ unsigned long long mix64(unsigned long long g, unsigned v) {
    unsigned long long z = g * 0x9E3779B97F4A7C15ull + (unsigned long long)v * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void __launch_bounds__(DIAG_THREADS)
hash_kernel(const float *__restrict__ f, size_t plane, int pitch, int nx, int nyl, int gy0, unsigned long long *out9) {
    unsigned long long h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    const size_t ncell = (size_t)nx * nyl;
    for (size_t t = (size_t)blockIdx.x * DIAG_THREADS + threadIdx.x; t < ncell; t += (size_t)gridDim.x * DIAG_THREADS) {
        const int y = (int)(t / nx), x = (int)(t - (size_t)y * nx);
        const size_t c = (size_t)(y + 1) * pitch + x;
        const unsigned long long g = (unsigned long long)(gy0 + y) * nx + x;
#pragma unroll
        for (int i = 0; i < 9; i++) h[i] += mix64(g, __float_as_uint(f[i * plane + c]));
    }
#pragma unroll
    for (int i = 0; i < 9; i++) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) h[i] += __shfl_xor_sync(FULL, h[i], d);
        if ((threadIdx.x & 31) == 0 && h[i]) atomicAdd(out9 + i, h[i]);
    }
}

__device__ __forceinline__ float mixf(float a, float b, float u) { return a * (1.0f - u) + b * u; }
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ uint8_t unorm8(float v) {
    return (uint8_t)(int)floorf(clampf(v, 0.0f, 1.0f) * 255.0f + 0.5f);
}

__constant__ float SPEED_PAL[10][3] = {{5, 5, 20}, {0, 20, 120}, {0, 60, 200}, {0, 140, 220}, {0, 220, 220},
                                       {0, 210, 140}, {80, 200, 0}, {220, 210, 0}, {255, 120, 0}, {220, 20, 0}};
__constant__ float CP_PAL[8][3] = {{20, 50, 160}, {40, 110, 210}, {100, 175, 235}, {190, 220, 245},
                                   {248, 248, 248}, {248, 214, 140}, {240, 150, 60}, {205, 50, 25}};

__device__ __forceinline__ void palette(const float (*C)[3], int nseg, float t, float *rgb) {
    t = clampf(t, 0.0f, 1.0f);
    const float f = t * (float)nseg;
    int i = (int)floorf(f);
    if (i > nseg - 1) i = nseg - 1;
    if (i < 0) i = 0;
    const float u = f - (float)i;
#pragma unroll
    for (int k = 0; k < 3; k++) rgb[k] = mixf(C[i][k] / 255.0f, C[i + 1][k] / 255.0f, u);
}

// HTML:395-422.  ny = rows of this slab.  lo_ghost / hi_ghost: the row below the first / above the
// last owned row belongs to a neighbouring slab and its ux, uy sit in the ghost rows of the arrays
// (alb_set_macro_ghosts); otherwise the vorticity taps clamp to the edge like the page's textures.
__global__ void render_kernel(const uint8_t *__restrict__ mask, const float *__restrict__ rho,
                              const float *__restrict__ ux, const float *__restrict__ uy, int pitch,
                              int nx, int ny, int lo_ghost, int hi_ghost, int mode, float U0, float maxS, float cpMin,
                              float cpMax, float vortScale, float *t_out, uint8_t *rgba) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= nx || y >= ny) return;
    const size_t c = (size_t)(y + 1) * pitch + x;
    const size_t o = (size_t)y * nx + x;
    float t = NAN;
    const bool solid = mask[c] != 0;
    if (!solid) {
        if (mode == ALB_FIELD_SPEED) {
            const float s = sqrtf(ux[c] * ux[c] + uy[c] * uy[c]) / U0;
            t = s / fmaxf(maxS * 0.92f, 1e-6f);
        } else if (mode == ALB_FIELD_CP) {
            const float cp = (rho[c] - 1.0f) / (1.5f * U0 * U0);
            const float range = fmaxf(cpMax - cpMin, 1e-6f);
            t = (cp - cpMin) / range;
        } else {
            // CLAMP_TO_EDGE neighbour taps (HTML:443-444, 411-414)
            const int xr = min(x + 1, nx - 1), xl = max(x - 1, 0);
            const int yu = (y == ny - 1 && hi_ghost) ? ny : min(y + 1, ny - 1);
            const int yd = (y == 0 && lo_ghost) ? -1 : max(y - 1, 0);
            const float dvydx = (uy[(size_t)(y + 1) * pitch + xr] - uy[(size_t)(y + 1) * pitch + xl]) * 0.5f;
            const float duxdy = (ux[(size_t)(yu + 1) * pitch + x] - ux[(size_t)(yd + 1) * pitch + x]) * 0.5f;
            const float vort = dvydx - duxdy;
            t = vort / fmaxf(U0 * vortScale, 1e-6f);
        }
    }
    if (t_out) t_out[o] = t;
    if (rgba) {
        float col[3];
        if (solid) {
            col[0] = 0.039f; col[1] = 0.043f; col[2] = 0.078f;
        } else if (mode == ALB_FIELD_SPEED) {
            palette(SPEED_PAL, 9, t, col);
        } else if (mode == ALB_FIELD_CP) {
            palette(CP_PAL, 7, t, col);
        } else {
            const float tc = clampf(t, -1.0f, 1.0f);
            const float base[3] = {0.06f, 0.07f, 0.11f};
            const float neg[3] = {0.15f, 0.5f, 0.98f}, pos[3] = {0.98f, 0.28f, 0.18f};
#pragma unroll
            for (int k = 0; k < 3; k++)
                col[k] = tc < 0.0f ? mixf(base[k], neg[k], -tc) : mixf(base[k], pos[k], tc);
        }
        rgba[4 * o + 0] = unorm8(col[0]);
        rgba[4 * o + 1] = unorm8(col[1]);
        rgba[4 * o + 2] = unorm8(col[2]);
        rgba[4 * o + 3] = 255;
    }
}

// HTML:474-500: every cell of both ping-pong sets := the same nine fp32 values.
__global__ void fill_kernel(float *f0, float *f1, size_t plane, float e0, float e1, float e2, float e3,
                            float e4, float e5, float e6, float e7, float e8, float *rho, float *ux,
                            float *uy, float u0) {
    const float e[9] = {e0, e1, e2, e3, e4, e5, e6, e7, e8};
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < plane;
         t += (size_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int i = 0; i < 9; i++) {
            f0[i * plane + t] = e[i];
            f1[i * plane + t] = e[i];
        }
        rho[t] = 1.0f;
        ux[t] = u0;
        uy[t] = 0.0f;
    }
}

// End of a frame, one warp: merge the DIAG_SLOTS accumulator copies, then lane 0 performs
// updateFieldsFromMacro()'s sticky update (HTML:611-613) and, on force frames, computeForces()'s
// normalisation and EMAs (HTML:672-679, 699) -- the same float64 operations the host path
// performs.  `row` may be mapped host memory.  Finally the merged reductions are published for
// the host and all copies re-armed for the next frame (saves a host->device copy per frame).
__global__ void frame_finalize_kernel(DiagAcc *d, DiagAcc *published, const MeState *me, int me_parity,
                                      FrameDev *st, int do_forces, double U0, double q, double *row) {
    const int lane = threadIdx.x;
    DiagAcc m;
    m.smax_bits = 0; m.m2max_bits = 0; m.rho_min = INFINITY; m.rho_max = -INFINITY;
    m.fx = 0; m.fy = 0; m.surf = 0; m.rev = 0;
    DiagAcc z = m;
    for (int k = lane; k < DIAG_SLOTS; k += 32) {
        const DiagAcc a = d[k];
        m.smax_bits = max(m.smax_bits, a.smax_bits);
        m.m2max_bits = max(m.m2max_bits, a.m2max_bits);
        m.rho_min = fminf(m.rho_min, a.rho_min);
        m.rho_max = fmaxf(m.rho_max, a.rho_max);
        m.fx += a.fx; m.fy += a.fy; m.surf += a.surf; m.rev += a.rev;
        d[k] = z;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        m.smax_bits = max(m.smax_bits, __shfl_xor_sync(FULL, m.smax_bits, s));
        m.m2max_bits = max(m.m2max_bits, __shfl_xor_sync(FULL, m.m2max_bits, s));
        m.rho_min = fminf(m.rho_min, __shfl_xor_sync(FULL, m.rho_min, s));
        m.rho_max = fmaxf(m.rho_max, __shfl_xor_sync(FULL, m.rho_max, s));
        m.fx += __shfl_xor_sync(FULL, m.fx, s);
        m.fy += __shfl_xor_sync(FULL, m.fy, s);
        m.surf += __shfl_xor_sync(FULL, m.surf, s);
        m.rev += __shfl_xor_sync(FULL, m.rev, s);
    }
    if (lane != 0) return;
    *published = m;
    if (st == nullptr) {
        // a slab of a decomposed lattice: publish the raw partial reductions of this frame; the host
        // combines the slabs (max / min / exact integer sums) and applies the page's logic afterwards
        row[0] = __longlong_as_double((long long)m.smax_bits);
        row[1] = (double)m.rho_min;
        row[2] = (double)m.rho_max;
        row[3] = __longlong_as_double(m.fx);
        row[4] = __longlong_as_double(m.fy);
        row[5] = __longlong_as_double((long long)m.surf);
        row[6] = __longlong_as_double((long long)m.rev);
        row[7] = __longlong_as_double(me->acc[me_parity][0]);
        row[8] = __longlong_as_double(me->acc[me_parity][1]);
        row[9] = U0;
        row[10] = q;
        row[11] = (double)do_forces;
        return;
    }
    const double cden = __dmul_rn(__dmul_rn(1.5, U0), U0);
    const double smax = __longlong_as_double((long long)m.smax_bits);
    if (smax > 0) st->maxS = smax;
    if (m.rho_min <= m.rho_max) {
        const double cmin = __ddiv_rn(__dsub_rn((double)m.rho_min, 1.0), cden);
        const double cmax = __ddiv_rn(__dsub_rn((double)m.rho_max, 1.0), cden);
        if (isfinite(cmin)) st->cpMin = cmin;
        if (isfinite(cmax)) st->cpMax = cmax;
    }
    double cl_raw = NAN, cd_raw = NAN, surf = NAN, rev = NAN;
    if (do_forces) {
        surf = (double)m.surf;
        rev = (double)m.rev;
        if (m.surf > 0) {
            const double fx = __ddiv_rn(__ddiv_rn((double)m.fx, ALB_ME_SCALE), 3.0);
            const double fy = __ddiv_rn(__ddiv_rn((double)m.fy, ALB_ME_SCALE), 3.0);
            cl_raw = __ddiv_rn(fy, q);
            cd_raw = __ddiv_rn(fx, q);
            if (!st->ema_valid) {
                st->cl_smooth = cl_raw;
                st->cd_smooth = cd_raw;
                st->ema_valid = 1;
            } else {
                st->cl_smooth = __dadd_rn(__dmul_rn(st->cl_smooth, 0.9), __dmul_rn(cl_raw, 0.1));
                st->cd_smooth = __dadd_rn(__dmul_rn(st->cd_smooth, 0.9), __dmul_rn(cd_raw, 0.1));
            }
            st->sep_frac = __dadd_rn(__dmul_rn(st->sep_frac, 0.85), __dmul_rn(__ddiv_rn(rev, surf), 0.15));
        }
    }
    const double mfx = __ddiv_rn((double)me->acc[me_parity][0], ALB_ME_SCALE);
    const double mfy = __ddiv_rn((double)me->acc[me_parity][1], ALB_ME_SCALE);
    row[0] = st->ema_valid ? st->cl_smooth : NAN;
    row[1] = st->ema_valid ? st->cd_smooth : NAN;
    row[2] = st->sep_frac;
    row[3] = cl_raw;
    row[4] = cd_raw;
    row[5] = surf;
    row[6] = rev;
    row[7] = st->maxS;
    row[8] = st->cpMin;
    row[9] = st->cpMax;
    row[10] = __ddiv_rn(mfy, q);
    row[11] = __ddiv_rn(mfx, q);
}

}  // namespace

cudaError_t launch_frame_finalize(DiagAcc *d, DiagAcc *published, const MeState *me, int me_parity, FrameDev *st,
                                  int do_forces, double U0, double q, double *row, cudaStream_t s) {
    frame_finalize_kernel<<<1, 32, 0, s>>>(d, published, me, me_parity, st, do_forces, U0, q, row);
    return cudaGetLastError();
}

cudaError_t launch_stats(const uint8_t *mask, const float *rho, const float *ux, const float *uy,
                         int pitch, int nx, int nyl, double u0, float *U, float *V, float *Cp,
                         double *d_part, int nblocks, cudaStream_t s) {
    stats_kernel<<<nblocks, DIAG_THREADS, 0, s>>>(mask, rho, ux, uy, pitch, nx, nyl, u0, U, V, Cp, d_part);
    return cudaGetLastError();
}

cudaError_t launch_forces(const uint8_t *mask, const float *rho, const float *ux, int pitch, int nx,
                          int ny_global, int gy_first, int nyl, double *d_part, int nblocks,
                          cudaStream_t s) {
    forces_kernel<<<nblocks, DIAG_THREADS, 0, s>>>(mask, rho, ux, pitch, nx, ny_global, gy_first, nyl, d_part);
    return cudaGetLastError();
}

cudaError_t launch_render(const uint8_t *mask, const float *rho, const float *ux, const float *uy,
                          int pitch, int nx, int ny, int lo_ghost, int hi_ghost, int mode, float u0, float maxS,
                          float cpMin, float cpMax, float vortScale, float *t_out, uint8_t *rgba, cudaStream_t s) {
    dim3 grid((nx + 255) / 256, ny);
    render_kernel<<<grid, 256, 0, s>>>(mask, rho, ux, uy, pitch, nx, ny, lo_ghost, hi_ghost, mode, u0, maxS, cpMin,
                                       cpMax, vortScale, t_out, rgba);
    return cudaGetLastError();
}

cudaError_t launch_mass(const float *f, size_t plane, int pitch, int nx, int nyl, double *d_part,
                        int nblocks, cudaStream_t s) {
    mass_kernel<<<nblocks, DIAG_THREADS, 0, s>>>(f, plane, pitch, nx, nyl, d_part);
    return cudaGetLastError();
}

cudaError_t launch_state_hash(const float *f, size_t plane, int pitch, int nx, int nyl, int gy0,
                              unsigned long long *d_out9, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_out9, 0, 9 * sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    hash_kernel<<<148 * 8, DIAG_THREADS, 0, s>>>(f, plane, pitch, nx, nyl, gy0, d_out9);
    return cudaGetLastError();
}

cudaError_t launch_fill_init(float *f0, float *f1, size_t plane, const float *e, float *rho, float *ux,
                             float *uy, float u0, cudaStream_t s) {
    fill_kernel<<<148 * 8, 256, 0, s>>>(f0, f1, plane, e[0], e[1], e[2], e[3], e[4], e[5], e[6], e[7], e[8],
                                        rho, ux, uy, u0);
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

cudaError_t preload_diag_kernels() {
    ALB_PRELOAD(stats_kernel);
    ALB_PRELOAD(forces_kernel);
    ALB_PRELOAD(mass_kernel);
    ALB_PRELOAD(hash_kernel);
    ALB_PRELOAD(render_kernel);
    ALB_PRELOAD(fill_kernel);
    ALB_PRELOAD(frame_finalize_kernel);
    return cudaSuccess;
}

}  // namespace alb

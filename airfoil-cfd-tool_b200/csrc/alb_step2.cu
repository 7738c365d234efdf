// Helpers of the two-steps-per-pass path (march2_kernel lives in alb_march.cu): the copy kernel for
// all-solid tasks and the self-tests of the two division shortcuts (div_pair, div_by_tau) against
// IEEE division.  DESIGN.md section 4.2.
#include <stdio.h>
#include <stdlib.h>

#include "alb_lbm.cuh"

namespace alb {

namespace {

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

// ---- exhaustive check of div_by_tau<DM_FAST3> for one tau (alb_api.cu: refresh_params) -----------
// One thread per (sign, exponent, 2^11 mantissas): all fp32 x with 2^-40 <= |x| < 2^8.
constexpr int DIVTAU_E0 = 127 - 40, DIVTAU_E1 = 127 + 8, DIVTAU_CHUNK = 11;
__global__ void divtau_check_kernel(float tau, float rcp, unsigned long long *out) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned per_exp = 1u << (23 - DIVTAU_CHUNK);
    const unsigned nexp = DIVTAU_E1 - DIVTAU_E0;
    if (t >= 2u * nexp * per_exp) return;
    const unsigned sign = t / (nexp * per_exp), rest = t - sign * nexp * per_exp;
    const unsigned e = DIVTAU_E0 + rest / per_exp, m0 = (rest % per_exp) << DIVTAU_CHUNK;
    unsigned bad = 0;
    for (unsigned m = m0; m < m0 + (1u << DIVTAU_CHUNK); m++) {
        const float x = __uint_as_float((sign << 31) | (e << 23) | m);
        bad += __float_as_uint(div_by_tau<DM_FAST3>(x, tau, rcp)) != __float_as_uint(__fdiv_rn(x, tau));
    }
    if (bad) atomicAdd(out, (unsigned long long)bad);
}

}  // namespace

cudaError_t launch_divtau_check(float tau, float rcp, unsigned long long *d_out1, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_out1, 0, sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    const unsigned nthreads = 2u * (DIVTAU_E1 - DIVTAU_E0) * (1u << (23 - DIVTAU_CHUNK));
    divtau_check_kernel<<<(nthreads + 255) / 256, 256, 0, s>>>(tau, rcp, d_out1);
    return cudaGetLastError();
}

cudaError_t launch_div_selftest(unsigned long long seed, int nblocks, int iters, unsigned long long *d_out3, cudaStream_t s) {
    div_selftest_kernel<<<nblocks, 256, 0, s>>>(seed, iters, d_out3);
    return cudaGetLastError();
}

cudaError_t launch_copy_tasks(const StepParams &p, cudaStream_t s) {
    if (p.ngen <= 0) return cudaSuccess;
    copy_tasks_kernel<<<(p.ngen + TASKS_PER_BLOCK - 1) / TASKS_PER_BLOCK, BLOCK_THREADS, 0, s>>>(p);
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

cudaError_t preload_step2_kernels() {
    ALB_PRELOAD(copy_tasks_kernel);
    ALB_PRELOAD(div_selftest_kernel);
    ALB_PRELOAD(divtau_check_kernel);
    return cudaSuccess;
}

}  // namespace alb

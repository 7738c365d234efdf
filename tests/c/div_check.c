/* Exhaustive check of the step kernel's div_by_tau() sequence against IEEE division.
 * Mirrors airfoil-cfd-tool_b200/csrc/alb_lbm.cuh:div_by_tau with C99 fmaf (exact FMA).
 * Build: gcc -O2 -fopenmp -ffp-contract=off [-mfma]  (see tests/test_div_by_tau.py) */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* DM_FAST3 of alb_lbm.cuh: q0 = RN(x*rcp) with rcp = RN(1/tau); exact residual r = x - tau*q0 (FMA);
 * q = RN(q0 + r*rcp) (FMA).  Whether it equals IEEE division for every x depends on tau -- the
 * library checks each tau on the device before using it; this is the same check on the CPU. */
static inline float div_by_tau(float x, float tau, float rcp)
{
    float q = x * rcp;
    float r = fmaf(-tau, q, x);
    return fmaf(r, rcp, q);
}

/* Returns the number of x in [lo_bits, hi_bits) (stepping by `stride` in bit space, both signs)
 * for which div_by_tau(x) != x / tau bitwise. first_bad receives one offending bit pattern. */
long div_check(float tau, uint32_t lo_bits, uint32_t hi_bits, uint32_t stride, uint32_t *first_bad)
{
    const float rcp = 1.0f / tau;
    long bad = 0;
    uint32_t fb = 0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
    for (int64_t b = lo_bits; b < (int64_t)hi_bits; b += stride) {
        uint32_t u = (uint32_t)b;
        float x, q1, q2;
        memcpy(&x, &u, 4);
        for (int sgn = 0; sgn < 2; sgn++) {
            float xs = sgn ? -x : x;
            q1 = div_by_tau(xs, tau, rcp);
            q2 = xs / tau;
            uint32_t a, c;
            memcpy(&a, &q1, 4);
            memcpy(&c, &q2, 4);
            if (a != c) {
                bad++;
#pragma omp critical
                fb = u;
            }
        }
    }
    if (first_bad) *first_bad = fb;
    return bad;
}

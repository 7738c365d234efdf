/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 * Pinned against the reference's own shader / JavaScript text executed by the
 * minimal interpreters in tests/refexec (tests/test_reference_pins.py): bitwise
 * equal populations, macro fields, statistics, force EMAs and render output.
 *
 * Strict IEEE-754 binary32 restatement, in source operation order, of the
 * reference's D2Q9 step and its host-side diagnostics.  "HTML:n" cites
 * pages/airfoil_flow_lbm_aerolab.html of the reference.
 *
 *   orc_init            equilibriumInitData / initSim      HTML:474-500
 *   orc_step            STEP_FS_SRC.main (+dir/wt/opp/feq)  HTML:234-360
 *   orc_field_stats     updateFieldsFromMacro              HTML:596-614
 *   orc_forces          computeForces (pressure faces+sep) HTML:649-700
 *   orc_render_scalar   RENDER_FS_SRC.main, scalar t only  HTML:395-420
 *   orc_rgba            palettes + mix -> RGBA8            HTML:371-393,421
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -fno-fast-math (oracle/build.py).
 * -ffp-contract=off matters: an FMA anywhere changes the last bit and, after
 * 1,000 steps, the fields by ~1e-5 relative.
 *
 * Layout: populations SoA, F[q][row][x] (q = 0..8), row 0 = bottom of the
 * world window.  A "slab" view is supported so the multi-GPU decomposition can
 * be checked on the CPU: the arrays hold `nrows` rows, local row j is global
 * row gy0 + j, and rows j0 <= j < j1 are updated.  Whole domain: gy0 = 0,
 * nrows = ny, j0 = 0, j1 = ny.
 *
 * Not in the reference (asked for by BASELINE.json north_star): a
 * momentum-exchange force.  For every interior fluid cell x and direction i
 * whose pull source x - e_i is solid, the half-way bounce-back returns the
 * population f_opp(i)(x); the body gains momentum 2 * f_opp(i)(x) * e_opp(i).
 * The sum is taken in 2^-40 fixed point so that it is exact and independent of
 * summation order (the CUDA kernel accumulates the same integers).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>

static const int EX[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};   /* HTML:238-248 */
static const int EY[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
static const int OPP[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};     /* HTML:254-264 */

static inline float wt(int i)                               /* HTML:234-236, 249-253 */
{
    const float w0 = 4.0f / 9.0f;
    const float ws = 1.0f / 9.0f;
    const float wd = 1.0f / 36.0f;
    if (i == 0) return w0;
    if (i >= 1 && i <= 4) return ws;
    return wd;
}

static inline float feq(int i, float rho, float ux, float uy) /* HTML:276-281 */
{
    float ex = (float)EX[i], ey = (float)EY[i];
    float eu = ex * ux + ey * uy;
    float uu = ux * ux + uy * uy;
    return wt(i) * rho * (1.0f + 3.0f * eu + 4.5f * eu * eu - 1.5f * uu);
}

static inline float clampf(float x, float lo, float hi)     /* GLSL clamp = min(max(x,lo),hi) */
{
    float m = x > lo ? x : lo;
    return m < hi ? m : hi;
}

void orc_weights(float *w9)
{
    for (int i = 0; i < 9; i++) w9[i] = wt(i);
}

/* HTML:474-490: feq(rho=1, u=(u0,0)) in float64, rounded to fp32 on store. */
void orc_init(int nx, int nrows, double u0, float *F, float *rho, float *ux, float *uy)
{
    const double w0 = 4.0 / 9.0, ws = 1.0 / 9.0, wd = 1.0 / 36.0;
    const double W[9] = {w0, ws, ws, ws, ws, wd, wd, wd, wd};
    float f[9];
    for (int i = 0; i < 9; i++) {
        double eu = EX[i] * u0, uu = u0 * u0;
        f[i] = (float)(W[i] * (1 + 3 * eu + 4.5 * eu * eu - 1.5 * uu));
    }
    size_t n = (size_t)nx * nrows;
    for (int i = 0; i < 9; i++)
        for (size_t c = 0; c < n; c++) F[(size_t)i * n + c] = f[i];
    if (rho) for (size_t c = 0; c < n; c++) rho[c] = 1.0f;
    if (ux)  for (size_t c = 0; c < n; c++) ux[c] = (float)u0;
    if (uy)  for (size_t c = 0; c < n; c++) uy[c] = 0.0f;
}

static inline int64_t me_fixed(float f)
{
    /* 2*f in 2^-40 fixed point, round-to-nearest-even (exact for |f| >= 2^-18) */
    return (int64_t)llrint((double)f * 0x1p41);
}

/*
 * One step, HTML:283-360.  mask: nonzero = solid (R8 texel 255 -> 1.0 > 0.5).
 * rho/ux/uy may be NULL.  me_fx/me_fy (may be NULL) receive the fixed-point
 * momentum-exchange sums of THIS step (they are added to, not cleared).
 * clamp_hits (may be NULL) counts interior cells where a clamp changed rho or u.
 */
void orc_step(int nx, int ny_global, int gy0, int nrows, int j0, int j1,
              const uint8_t *mask, const float *src, float *dst,
              float *rho_o, float *ux_o, float *uy_o, float tau, float U0,
              int64_t *me_fx, int64_t *me_fy, int64_t *clamp_hits)
{
    const size_t n = (size_t)nx * nrows;
    int64_t sfx = 0, sfy = 0, shits = 0;
    (void)nrows;
#pragma omp parallel for schedule(static) reduction(+ : sfx, sfy, shits)
    for (int j = j0; j < j1; j++) {
        const int gy = gy0 + j;
        for (int x = 0; x < nx; x++) {
            const size_t c = (size_t)j * nx + x;
            float o[9], r, vx, vy;
            if (mask[c]) {                                   /* HTML:287-294 */
                for (int i = 0; i < 9; i++) o[i] = src[(size_t)OPP[i] * n + c];
                r = 1.0f; vx = 0.0f; vy = 0.0f;
            } else if (x == nx - 1) {                        /* HTML:301-312 */
                const size_t s = c - 1;
                for (int i = 0; i < 9; i++) o[i] = src[(size_t)i * n + s];
                r = o[0] + o[1] + o[2] + o[3] + o[4] + o[5] + o[6] + o[7] + o[8];
                vx = (o[1] + o[5] + o[8] - o[3] - o[6] - o[7]) / r;
                vy = (o[2] + o[5] + o[6] - o[4] - o[7] - o[8]) / r;
            } else if (x == 0 || gy == ny_global - 1 || gy == 0) { /* HTML:314-322 */
                r = 1.0f; vx = U0; vy = 0.0f;
                for (int i = 0; i < 9; i++) o[i] = feq(i, r, vx, vy);
            } else {                                         /* HTML:324-359 */
                float fin[9];
                for (int i = 0; i < 9; i++) {
                    const size_t s = (size_t)(j - EY[i]) * nx + (x - EX[i]);
                    if (mask[s]) {
                        fin[i] = src[(size_t)OPP[i] * n + c];
                        /* momentum handed to the body along e_opp(i) = -e_i */
                        int64_t q = me_fixed(fin[i]);
                        sfx += -EX[i] * q;
                        sfy += -EY[i] * q;
                    } else {
                        fin[i] = src[(size_t)i * n + s];
                    }
                }
                r = 0.0f;
                for (int i = 0; i < 9; i++) r += fin[i];
                vx = (fin[1] + fin[5] + fin[8] - fin[3] - fin[6] - fin[7]) / r;
                vy = (fin[2] + fin[5] + fin[6] - fin[4] - fin[7] - fin[8]) / r;
                const float uMax = 0.35f, rhoMin = 0.5f, rhoMax = 2.0f;
                int hit = 0;
                float rc = clampf(r, rhoMin, rhoMax);
                if (rc != r) hit = 1;
                r = rc;
                float spd2 = vx * vx + vy * vy;
                if (spd2 > uMax * uMax) {
                    float k = uMax / sqrtf(spd2);
                    vx *= k; vy *= k;
                    hit = 1;
                }
                shits += hit;
                for (int i = 0; i < 9; i++) {
                    float eq = feq(i, r, vx, vy);
                    o[i] = fin[i] - (fin[i] - eq) / tau;
                }
            }
            for (int i = 0; i < 9; i++) dst[(size_t)i * n + c] = o[i];
            if (rho_o) rho_o[c] = r;
            if (ux_o) ux_o[c] = vx;
            if (uy_o) uy_o[c] = vy;
        }
    }
    if (me_fx) *me_fx += sfx;
    if (me_fy) *me_fy += sfy;
    if (clamp_hits) *clamp_hits += shits;
}

/*
 * nsteps whole-domain steps, ping-ponging between A and B (HTML:510-525).
 * Returns 0 if the final state is in A, 1 if it is in B.  me_hist (may be NULL)
 * receives 2 int64 per step (fixed-point Fx, Fy).
 */
int orc_run(int nx, int ny, const uint8_t *mask, float *A, float *B,
            float *rho, float *ux, float *uy, float tau, float U0, int nsteps,
            int64_t *me_hist, int64_t *clamp_hits)
{
    int cur = 0;
    for (int s = 0; s < nsteps; s++) {
        float *srcp = cur ? B : A, *dstp = cur ? A : B;
        int64_t fx = 0, fy = 0;
        orc_step(nx, ny, 0, ny, 0, ny, mask, srcp, dstp, rho, ux, uy, tau, U0,
                 &fx, &fy, clamp_hits);
        if (me_hist) { me_hist[2 * s] = fx; me_hist[2 * s + 1] = fy; }
        cur = 1 - cur;
    }
    return cur;
}

/*
 * HTML:596-614.  U, V, Cp: fp32 arrays (NaN in solids), any may be NULL.
 * out[0..2] = raw mx, cMin, cMax of the sweep (mx starts at 0, cMin at +inf,
 * cMax at -inf); the caller applies "if(mx>0) maxS=mx" etc. (HTML:611-613).
 * U0 is the JS double, not the fp32 uniform.
 */
void orc_field_stats(int nx, int ny, const uint8_t *mask, const float *rho,
                     const float *ux, const float *uy, double U0,
                     float *U, float *V, float *Cp, double *out)
{
    double mx = 0, cMin = INFINITY, cMax = -INFINITY;
    size_t n = (size_t)nx * ny;
    for (size_t idx = 0; idx < n; idx++) {
        if (mask[idx]) {
            if (U) U[idx] = NAN;
            if (V) V[idx] = NAN;
            if (Cp) Cp[idx] = NAN;
            continue;
        }
        double r = rho[idx], x = ux[idx], y = uy[idx];
        double u = x / U0, v = y / U0;
        if (U) U[idx] = (float)u;
        if (V) V[idx] = (float)v;
        double cp = (r - 1) / (1.5 * U0 * U0);
        if (Cp) Cp[idx] = (float)cp;
        double s = hypot(u, v);
        if (s > mx && s < 4) mx = s;
        if (cp > -4 && cp < 1.2) { if (cp < cMin) cMin = cp; if (cp > cMax) cMax = cp; }
    }
    out[0] = mx; out[1] = cMin; out[2] = cMax;
}

/*
 * HTML:649-700.  out[0]=fx, out[1]=fy (lattice pressure units), out[2]=any,
 * out[3]=surf, out[4]=rev.  Normalisation by q and the EMAs are the caller's
 * (HTML:676-679, 699).  Sequential double accumulation in the JS loop order.
 */
void orc_forces(int nx, int ny, const uint8_t *mask, const float *rho,
                const float *ux, double *out)
{
    static const int FDX[4] = {1, 0, -1, 0}, FDY[4] = {0, 1, 0, -1};
    double fx = 0, fy = 0;
    int any = 0;
    long surf = 0, rev = 0;
    for (int y = 0; y < ny; y++)
        for (int x = 0; x < nx; x++) {
            size_t c = (size_t)y * nx + x;
            if (!mask[c]) continue;
            for (int j = 0; j < 4; j++) {
                int xn = x + FDX[j], yn = y + FDY[j];
                if (xn < 0 || xn >= nx || yn < 0 || yn >= ny) continue;
                size_t nc = (size_t)yn * nx + xn;
                if (mask[nc]) continue;
                any = 1;
                double r = rho[nc];
                double p = r / 3;
                fx += p * (-FDX[j]);
                fy += p * (-FDY[j]);
                surf++;
                if (ux[nc] < 0) rev++;
            }
        }
    out[0] = fx; out[1] = fy; out[2] = any; out[3] = (double)surf; out[4] = (double)rev;
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/*
 * Scalar part of RENDER_FS_SRC.main (HTML:395-420): the fp32 value t that is
 * fed to the palette.  mode 0 speed, 1 Cp, 2 vorticity.  Solid cells get NaN.
 * U0, maxS, cpMin, cpMax, vortScale are the fp32 uniforms.  Neighbour taps
 * use CLAMP_TO_EDGE (HTML:443-444).
 */
void orc_render_scalar(int nx, int ny, const uint8_t *mask, const float *rho,
                       const float *ux, const float *uy, int mode, float U0,
                       float maxS, float cpMin, float cpMax, float vortScale,
                       float *t_out)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < ny; y++)
        for (int x = 0; x < nx; x++) {
            size_t c = (size_t)y * nx + x;
            float t;
            if (mask[c]) { t_out[c] = NAN; continue; }
            if (mode == 0) {
                float s = sqrtf(ux[c] * ux[c] + uy[c] * uy[c]) / U0;
                t = s / fmaxf(maxS * 0.92f, 1e-6f);
            } else if (mode == 1) {
                float cp = (rho[c] - 1.0f) / (1.5f * U0 * U0);
                float range = fmaxf(cpMax - cpMin, 1e-6f);
                t = (cp - cpMin) / range;
            } else {
                size_t cR = (size_t)y * nx + clampi(x + 1, 0, nx - 1);
                size_t cL = (size_t)y * nx + clampi(x - 1, 0, nx - 1);
                size_t cU = (size_t)clampi(y + 1, 0, ny - 1) * nx + x;
                size_t cD = (size_t)clampi(y - 1, 0, ny - 1) * nx + x;
                float dvydx = (uy[cR] - uy[cL]) * 0.5f;
                float duxdy = (ux[cU] - ux[cD]) * 0.5f;
                float vort = dvydx - duxdy;
                t = vort / fmaxf(U0 * vortScale, 1e-6f);
            }
            t_out[c] = t;
        }
}

/* GLSL mix(a,b,u) = a*(1-u) + b*u */
static inline float mixf(float a, float b, float u) { return a * (1.0f - u) + b * u; }

static void palette(const float (*C)[3], int nseg, float t, float *rgb)
{
    /* HTML:375-379 / 384-387: t clamped to [0,1], f=t*nseg, i=floor(f) clamped */
    t = clampf(t, 0.0f, 1.0f);
    float f = t * (float)nseg;
    int i = (int)floorf(f);
    if (i > nseg - 1) i = nseg - 1;
    if (i < 0) i = 0;
    float u = f - (float)i;
    for (int k = 0; k < 3; k++) rgb[k] = mixf(C[i][k] / 255.0f, C[i + 1][k] / 255.0f, u);
}

static inline uint8_t unorm8(float v)
{
    /* RGBA8 framebuffer write: clamp to [0,1], scale by 255, round to nearest */
    v = clampf(v, 0.0f, 1.0f);
    return (uint8_t)(int)floorf(v * 255.0f + 0.5f);
}

/* HTML:371-393, 397, 421: palette lookup of t -> RGBA8 (solid = fixed colour). */
void orc_rgba(int nx, int ny, const uint8_t *mask, const float *t_in, int mode,
              uint8_t *rgba)
{
    static const float SP[10][3] = {{5, 5, 20}, {0, 20, 120}, {0, 60, 200}, {0, 140, 220},
                                    {0, 220, 220}, {0, 210, 140}, {80, 200, 0}, {220, 210, 0},
                                    {255, 120, 0}, {220, 20, 0}};
    static const float CP[8][3] = {{20, 50, 160}, {40, 110, 210}, {100, 175, 235}, {190, 220, 245},
                                   {248, 248, 248}, {248, 214, 140}, {240, 150, 60}, {205, 50, 25}};
    size_t n = (size_t)nx * ny;
    for (size_t c = 0; c < n; c++) {
        float col[3];
        if (mask[c]) { col[0] = 0.039f; col[1] = 0.043f; col[2] = 0.078f; }
        else if (mode == 0) palette(SP, 9, t_in[c], col);
        else if (mode == 1) palette(CP, 7, t_in[c], col);
        else {
            float t = clampf(t_in[c], -1.0f, 1.0f);
            const float base[3] = {0.06f, 0.07f, 0.11f};
            const float neg[3] = {0.15f, 0.5f, 0.98f}, pos[3] = {0.98f, 0.28f, 0.18f};
            for (int k = 0; k < 3; k++)
                col[k] = t < 0.0f ? mixf(base[k], neg[k], -t) : mixf(base[k], pos[k], t);
        }
        rgba[4 * c + 0] = unorm8(col[0]);
        rgba[4 * c + 1] = unorm8(col[1]);
        rgba[4 * c + 2] = unorm8(col[2]);
        rgba[4 * c + 3] = 255;
    }
}

/* Sum of all nine populations over rows [j0,j1), in float64 (mass diagnostics). */
double orc_total_mass(int nx, int nrows, int j0, int j1, const float *F)
{
    size_t n = (size_t)nx * nrows;
    double m = 0;
    for (int i = 0; i < 9; i++)
        for (int j = j0; j < j1; j++)
            for (int x = 0; x < nx; x++) m += F[(size_t)i * n + (size_t)j * nx + x];
    return m;
}

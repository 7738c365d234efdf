// Tracer particles of the tunnel: spawn/initParts/advect/stepParticles and
// the bilinear sampler sampleScalar/sampleUV of pages/airfoil_flow_lbm_aerolab.html:616-639,
// 727-808 ("HTML:n").  The reference draws the trails on a 2-D canvas and seeds them with
// Math.random(); here every particle owns a counter-based random stream (splitmix64 of seed,
// particle index and draw counter), so a run is reproducible and the CPU oracle can replay it.
// All arithmetic is float64 in the reference's operand order (the file is built with -fmad=false).
#include <math.h>

#include "alb_common.cuh"

namespace alb {

namespace {

__host__ __device__ inline unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// uniform double in [0, 1): the stand-in for Math.random()
__device__ inline double prand(unsigned long long seed, int pid, unsigned &ctr) {
    const unsigned long long k = splitmix64(seed ^ splitmix64(((unsigned long long)(unsigned)pid << 32) | ctr));
    ctr++;
    return (double)(k >> 11) * 0x1.0p-53;
}

struct Sampler {
    const uint8_t *mask;
    const float *ux, *uy;
    int pitch, nx, ny;
    double U0;
};

// sampleScalar() for U and V at once (HTML:616-639); false = null
__device__ bool sample_uv(const Sampler &f, double wx, double wy, double &u, double &v) {
    if (wx < DX0 || wx > DX1 || wy < DY0 || wy > DY1) return false;
    const double fx = (wx - DX0) / (DX1 - DX0) * f.nx - 0.5;
    const double fy = (wy - DY0) / (DY1 - DY0) * f.ny - 0.5;
    const int ix = max(0, min((int)floor(fx), f.nx - 2));
    const int iy = max(0, min((int)floor(fy), f.ny - 2));
    const double tx = fx - ix, ty = fy - iy;
    const double ws[4] = {(1 - tx) * (1 - ty), tx * (1 - ty), (1 - tx) * ty, tx * ty};
    const int cx[4] = {ix, ix + 1, ix, ix + 1}, cy[4] = {iy, iy, iy + 1, iy + 1};
    double su = 0, wu = 0, sv = 0, wv = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const size_t c = (size_t)(cy[k] + 1) * f.pitch + cx[k];
        if (f.mask[c]) continue;
        // Ufield/Vfield are Float32Arrays of ux/U0, uy/U0 (HTML:603-604)
        const double uu = (double)(float)((double)f.ux[c] / f.U0);
        const double vv = (double)(float)((double)f.uy[c] / f.U0);
        if (isfinite(uu)) { su += uu * ws[k]; wu += ws[k]; }
        if (isfinite(vv)) { sv += vv * ws[k]; wv += ws[k]; }
    }
    if (!(wu > 0) || !(wv > 0)) return false;
    u = su / wu;
    v = sv / wv;
    return true;
}

// spawn(edge, lane), HTML:730-736
__device__ void spawn(ParticleState &p, bool edge, bool have_lane, double lane, unsigned long long seed, int pid,
                      unsigned &ctr) {
    if (!have_lane) lane = DY0 + prand(seed, pid, ctr) * (DY1 - DY0);
    if (edge || prand(seed, pid, ctr) < 0.82) {
        p.x = DX0 + 0.001;
        p.y = lane;
        p.life = 220 + prand(seed, pid, ctr) * 300;
    } else {
        p.x = DX0 + prand(seed, pid, ctr) * (DX1 - DX0);
        p.y = DY0 + prand(seed, pid, ctr) * (DY1 - DY0);
        p.life = 150 + prand(seed, pid, ctr) * 250;
    }
    p.lane = lane;
}

// initParts() for particles [first, n); first > 0 = the slider's push(spawn(false)), HTML:737-753, 961-967
__global__ void particles_init_kernel(ParticleState *ps, unsigned *ctrs, int first, int n, int npart_total,
                                      unsigned long long seed, int slider_push) {
    const int i = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned ctr = 0;
    ParticleState p;
    if (slider_push) {
        spawn(p, false, false, 0.0, seed, i, ctr);
    } else {
        double lane;
        if (prand(seed, i, ctr) < 0.35) {
            const double c = (DY0 + DY1) / 2, half = (DY1 - DY0) / 6;
            lane = c + (prand(seed, i, ctr) - 0.5) * 2 * half;
        } else {
            lane = DY0 + ((i + 0.5) / npart_total) * (DY1 - DY0) + (prand(seed, i, ctr) - 0.5) * 0.003;
        }
        spawn(p, true, true, lane, seed, i, ctr);
        p.life *= prand(seed, i, ctr);
        p.x = DX0 + prand(seed, i, ctr) * (DX1 - DX0) * 0.95;
    }
    p.x0 = p.x;
    p.y0 = p.y;
    p.speed = 0;
    p.respawned = 1;
    ps[i] = p;
    ctrs[i] = ctr;
}

// stepParticles(dt) without the canvas strokes, HTML:754-808
__global__ void particles_step_kernel(ParticleState *ps, unsigned *ctrs, int n, unsigned long long seed, double dt,
                                      Sampler f) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ParticleState p = ps[i];
    unsigned ctr = ctrs[i];
    const double STALL_SPEED2 = 3e-6, STALL_DRAIN = 0.18;
    // advect()
    bool ok = false;
    double nx = 0, ny = 0, speed = 0;
    double u1, v1;
    if (sample_uv(f, p.x, p.y, u1, v1)) {
        const double kBase = 0.00105 * dt;
        const double speed1 = hypot(u1, v1);
        double dtEff = kBase;
        const double maxDisp = 0.05;
        if (speed1 * dtEff > maxDisp) dtEff = maxDisp / fmax(speed1, 1e-6);
        const double midx = p.x + u1 * dtEff * 0.5, midy = p.y + v1 * dtEff * 0.5;
        double u2, v2;
        if (!sample_uv(f, midx, midy, u2, v2)) { u2 = u1; v2 = v1; }
        nx = p.x + u2 * dtEff;
        ny = p.y + v2 * dtEff;
        speed = hypot(u2, v2);
        ok = true;
    }
    const bool stalled = ok && (speed * speed < STALL_SPEED2);
    p.life -= dt * (stalled ? STALL_DRAIN : 0.06);
    if (!ok || p.life <= 0) {
        const double lane = p.lane;
        spawn(p, true, true, lane, seed, i, ctr);
        p.x0 = p.x;
        p.y0 = p.y;
        p.speed = 0;
        p.respawned = 1;
    } else {
        p.x0 = p.x;
        p.y0 = p.y;
        p.x = nx;
        p.y = ny;
        p.speed = speed;
        p.respawned = 0;
    }
    ps[i] = p;
    ctrs[i] = ctr;
}

}  // namespace

cudaError_t launch_particles_init(ParticleState *ps, unsigned *ctrs, int first, int n, int npart_total,
                                  unsigned long long seed, int slider_push, cudaStream_t s) {
    if (n <= first) return cudaSuccess;
    particles_init_kernel<<<(n - first + 127) / 128, 128, 0, s>>>(ps, ctrs, first, n, npart_total, seed, slider_push);
    return cudaGetLastError();
}

cudaError_t launch_particles_step(ParticleState *ps, unsigned *ctrs, int n, unsigned long long seed, double dt,
                                  const uint8_t *mask, const float *ux, const float *uy, int pitch, int nx, int ny,
                                  double U0, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    Sampler f{mask, ux, uy, pitch, nx, ny, U0};
    particles_step_kernel<<<(n + 127) / 128, 128, 0, s>>>(ps, ctrs, n, seed, dt, f);
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

cudaError_t preload_particle_kernels() {
    ALB_PRELOAD(particles_init_kernel);
    ALB_PRELOAD(particles_step_kernel);
    return cudaSuccess;
}

}  // namespace alb

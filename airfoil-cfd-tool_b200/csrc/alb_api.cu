// C ABI of libaerolab_lbm.so (declared in include/aerolab_lbm.h).
//
// Host-side driver around the kernels: the CUDA equivalent of the reference
// page's WebGL plumbing, initSim/simStep/readMacro and the stats/forces host
// code (pages/airfoil_flow_lbm_aerolab.html:424-552, 579-614, 641-700,
// 862-885).  No CPU fallback exists: every entry point that computes needs a
// CUDA device and reports ALB_ERR_CUDA otherwise.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <chrono>
#include <string.h>

#include <map>
#include <mutex>
#include <new>

#include "alb_common.cuh"

using namespace alb;

namespace {

constexpr int GRAPH_STEPS = 8;                // steps replayed per CUDA-graph launch (even)
constexpr int DIAG_BLOCKS = 148 * 4;          // fixed reduction grid -> deterministic partials
constexpr int UNIFIED_MAX_TASKS = 148 * 2 * 8;   // one wave of the unified kernel (2 CTAs/SM x 8 tasks)
constexpr long long WAIT_TIMEOUT_NS = 20LL * 1000 * 1000 * 1000;
// Lattices at least this wide and this large advance two steps per pass (see step_batch).  Measured
// with march2_kernel (tools/sweep_double.py, single vs double steps, GLUPS): 1536x768 83 / 79, 2048x512
// 84 / 73, 2000x1000 73 / 79, 2048x1024 75 / 98, 1024x4096 74 / 110, 4096x2048 85 / 123, 32768x16384
// 94 / 147-153.  Below two million cells the state lives in L2 and the launches of the list-driven passes
// cost more than the saved traffic.
constexpr int DOUBLE_MIN_NX = 1024;
constexpr long long DOUBLE_MIN_CELLS = 1900000;

thread_local std::string g_create_error;

struct Peer {
    float *base = nullptr;      // the neighbour's population block (f[k] at base + k*9*plane, k = 0..2)
    int *flags = nullptr;       // the neighbour's flag words (device memory, peer mapped)
    size_t plane = 0;
    int nyl = 0;
    void *ipc_base = nullptr;   // non-null when opened through CUDA IPC (to close on destroy)
};

struct IpcBlob {
    cudaIpcMemHandle_t mem;     // 64 bytes
    int nx, nyl, pitch, device;
    unsigned long long plane;
    unsigned long long flags_offset_bytes;
    int magic;
};
static_assert(sizeof(IpcBlob) <= ALB_IPC_BYTES, "blob too large");

}  // namespace

struct alb_handle {
    int nx = 0, ny_global = 0, y0 = 0, nyl = 0, nrows = 0, pitch = 0, tpr = 0, device = 0;
    size_t plane = 0;
    char *block = nullptr;        // one allocation: f[0], f[1], f[2], flag words (exported through IPC)
    float *f[3] = {nullptr, nullptr, nullptr};   // ping-pong pair + the intermediate state of the two-pass path
    int *flags = nullptr;         // [0] steps completed by the lower neighbour, [1] by the upper one
    int cur = 0;
    int parity = 0;               // parity of the NEXT step (momentum-exchange slot); == cur until a double step ran
    float *rho = nullptr, *ux = nullptr, *uy = nullptr;
    bool macro_valid = true;
    bool ghost_macro_valid = false;   // ux/uy ghost rows hold the neighbours' edge rows of the current state
    bool diag_valid = false;      // h_diag holds the fused statistics/forces of the current state
    DiagAcc *d_diag = nullptr;
    DiagAcc *d_diag_pub = nullptr;    // copy published by the frame-finalize kernel
    int diag_slots_host = 0;          // how many slots of h_diag the last copy filled (merged on read)
    bool diag_prearmed = false;       // d_diag was re-armed on the device: skip the next init copy
    DiagAcc *h_diag = nullptr;    // pinned: results
    DiagAcc *h_diag_init = nullptr;   // pinned: the constant initial value (zero sums, +/-inf extrema)
    double thr_u0 = -1;           // U0 the cached thresholds below were derived for
    float rho_lo = 0, rho_hi = 0;
    double m2_lo = -1, m2_hi = -1;
    uint8_t *mask = nullptr;
    uint16_t *info = nullptr;
    uint8_t *tclass = nullptr;
    int *gen_list = nullptr;      // TC_GENERAL tasks of this slab, [0] of gen_count = how many
    int *gen_count = nullptr;
    int ngen = 0;
    // two steps per pass (march2_kernel): task flags, the four task lists of the two-pass path
    uint8_t *tflags = nullptr, *deep_tmp = nullptr;
    int *lists[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // see launch_build_lists
    int *list_counts = nullptr;
    int nlist[5] = {0, 0, 0, 0, 0};
    int *s2_queue = nullptr;      // work queue of march2_kernel: {next unit, finished warps}, zero between launches
    bool solid_synced = false;    // both ping-pong buffers hold the same values on the all-solid tasks (lists[4])
    int double_mode = -1;         // -1 automatic, 0 never, 1 whenever possible (AEROLAB_LBM_DOUBLE / alb_set_option)
    int graph_parity = 0;         // h->parity the graph was captured at
    // AEROLAB_LBM_TRACE=<step>: CUDA events around the parts of the double step that starts at that
    // step count, printed by the next alb_sync()/alb_last_step_ms() (a measurement aid, see tools/)
    long long trace_step = -1;
    bool trace_armed = false;
    cudaEvent_t tev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    long long launches = 0;       // kernels launched by step batches so far (graph replays included)
    long long graph_launches = 0; // kernels inside one replay of the captured graph
    int prev_idx = 1;             // buffer that holds the PREVIOUS state (what the lazy macro pass reads);
                                  // -1 after a batch that ended with a double step: not materialised
    int small_capacity = 0;       // cells the persistent small-lattice kernel can hold on this GPU
    int band_rows = 0;            // cells per CTA (strip) of band_lattice_kernel (0: this lattice does not qualify)
    void *band_inbox = nullptr;   // its L2-resident message words between neighbouring bands
    int nsm = 148;                // SMs of the device
    double u0 = 0.06, tau = 0.58;
    float u0f = 0, tauf = 0, inv_tau = 0;
    int div_mode = 1;             // DM_IEEE until the three-instruction x / tau has been verified for tauf
    bool div_forced = false;      // alb_set_div_mode(h, 1): stay on IEEE division whatever tau is
    float feq0[9];
    cudaStream_t stream = nullptr;
    cudaStream_t aux = nullptr;   // runs the general-task kernel concurrently with the fast kernel
    cudaStream_t aux2 = nullptr;  // double steps: the general-task kernel of a pass beside its fast-list kernel
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_fork = nullptr, ev_join = nullptr, ev_fork2 = nullptr, ev_join2 = nullptr;
    bool timed = false;
    long long steps = 0;          // user-visible step count
    long long sync_steps = 0;     // monotonic, drives the halo flags and the ME ring
    // Slab protocol: the signal that ends a batch is held back until the next batch starts.  Until
    // then the neighbours cannot begin the step that overwrites the ghost rows of my PREVIOUS state,
    // which the lazy macroscopic pass (run_macro_pass / ensure_prev) still reads.  -1: nothing held.
    long long pending_signal = -1;
    MeState *me = nullptr;
    cudaGraphExec_t graph = nullptr;   // GRAPH_STEPS steps starting at cur == 0 (see alb_step)
    unsigned long long *clamp_hits = nullptr;
    double *d_xp = nullptr, *d_yp = nullptr;
    double xp[ALB_NPANEL + 1], yp[ALB_NPANEL + 1];
    bool have_panels = false;
    double *d_part = nullptr, *h_part = nullptr;
    float *d_tmp[3] = {nullptr, nullptr, nullptr};   // dense staging for U/V/Cp, field, rgba
    double maxS = 0.6, cpMin = -1.0, cpMax = 1.0;    // HTML:593
    bool ema_valid = false;
    double cl_smooth = 0, cd_smooth = 0, sep_frac = 0;
    Peer lo, hi;
    FrameDev *d_frame = nullptr, *h_frame = nullptr;   // device state of the frame loop + pinned staging
    double *d_rows = nullptr, *h_rows = nullptr;      // per-frame records (device, pinned host)
    int rows_cap = 0;
    int frames_pending = 0;                           // frames enqueued by alb_frames_enqueue, not yet collected
    bool frames_partial = false;                      // ... by a slab: the records are raw partial reductions
    long long frame_counter = 0;                      // statCounter, HTML:594, 913
    ParticleState *parts = nullptr;
    unsigned *part_ctr = nullptr;
    int nparts = 0, parts_cap = 0;
    unsigned long long part_seed = 0;
    bool external_halo = false;
    bool use_graph = true;        // replay step batches as a CUDA graph when nothing per-step is dynamic
    int *h_err = nullptr;         // mapped pinned: set by a wait kernel that timed out
    int *d_err = nullptr;
    std::string err;

    int fail(int code, const char *what, cudaError_t e = cudaSuccess) {
        char buf[512];
        if (e != cudaSuccess)
            snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
        else
            snprintf(buf, sizeof buf, "%s", what);
        err = buf;
        return code;
    }
    bool whole() const { return y0 == 0 && nyl == ny_global; }
    double chord_l() const { return nx / (DX1 - DX0); }   // HTML:77
    double qdyn() const { return 0.5 * u0 * u0 * chord_l(); }   // HTML:676
};

#define CK(expr)                                                          \
    do {                                                                  \
        cudaError_t e_ = (expr);                                          \
        if (e_ != cudaSuccess) return h->fail(ALB_ERR_CUDA, #expr, e_);   \
    } while (0)
#define NEED(h) \
    if (!(h)) return ALB_ERR_INVALID; \
    if (cudaSetDevice((h)->device) != cudaSuccess) return (h)->fail(ALB_ERR_CUDA, "cudaSetDevice")
#define ARG(cond, msg) \
    if (!(cond)) return h->fail(ALB_ERR_INVALID, msg)
// state-changing calls are not allowed between alb_frames_enqueue and alb_frames_collect
#define NO_PENDING_FRAMES(h) \
    if ((h)->frames_pending) return (h)->fail(ALB_ERR_STATE, "frames are in flight: call alb_frames_collect first")

namespace {

__global__ void wait_kernel(volatile int *flag_a, volatile int *flag_b, int target, int *err,
                            long long timeout_ns) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((flag_a && (int)(*flag_a - target) < 0) || (flag_b && (int)(*flag_b - target) < 0)) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if ((long long)(t1 - t0) > timeout_ns) {
            *err = 1;
            break;
        }
        __nanosleep(100);
    }
    __threadfence_system();
}

__global__ void signal_kernel(int *peer_a, int *peer_b, int value) {
    __threadfence_system();
    if (peer_a) *(volatile int *)peer_a = value;
    if (peer_b) *(volatile int *)peer_b = value;
    __threadfence_system();
}

// With CUDA's lazy module loading (the default since 12.2) the FIRST launch of a kernel may need the
// device to go idle.  A slab's wait_kernel spins until a neighbour signals; if the same host thread
// that was going to step that neighbour is meanwhile stuck in the first launch of some kernel on
// the waiting device, nothing ever moves (found with in-process slabs on two GPUs: 20 s timeouts
// whenever a kernel variant had not run on the device before).  So every kernel is loaded when the
// first handle on a device is created, before anything can spin.
cudaError_t preload_all_kernels(int device) {
    static std::mutex mu;
    static bool done[64] = {};
    std::lock_guard<std::mutex> lock(mu);
    if (device >= 0 && device < 64 && done[device]) return cudaSuccess;
    cudaFuncAttributes attr;
    cudaError_t e = cudaFuncGetAttributes(&attr, reinterpret_cast<const void *>(wait_kernel));
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&attr, reinterpret_cast<const void *>(signal_kernel));
    if (e == cudaSuccess) e = preload_step_kernels();
    if (e == cudaSuccess) e = preload_march_kernels();
    if (e == cudaSuccess) e = preload_step2_kernels();
    if (e == cudaSuccess) e = preload_diag_kernels();
    if (e == cudaSuccess) e = preload_geometry_kernels();
    if (e == cudaSuccess) e = preload_particle_kernels();
    if (e == cudaSuccess && device >= 0 && device < 64) done[device] = true;
    return e;
}

void drop_graph(alb_handle *h) {
    if (h->graph) {
        cudaGraphExecDestroy(h->graph);
        h->graph = nullptr;
    }
}

// How the kernels may divide by this tau (div_by_tau in alb_lbm.cuh).  The three-instruction
// sequence is used only for a tau for which the device has just compared it with IEEE division
// over every operand the kernels can see (805 M values, well under a millisecond); the verdicts are
// remembered per process.  AEROLAB_LBM_DIV=ieee forces true division (A/B measurements).
int div_mode_for(alb_handle *h, float tau) {
    static std::mutex mu;
    static std::map<uint32_t, int> verdicts;
    static const bool force_ieee = getenv("AEROLAB_LBM_DIV") && strcmp(getenv("AEROLAB_LBM_DIV"), "ieee") == 0;
    if (force_ieee) return 1;
    uint32_t key;
    memcpy(&key, &tau, 4);
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = verdicts.find(key);
        if (it != verdicts.end()) return it->second;
    }
    int mode = 1;                                  // DM_IEEE unless proven otherwise
    const float rcp = 1.0f / tau;
    if (isfinite(rcp) && tau >= 0x1p-20f && tau <= 0x1p20f) {
        unsigned long long *d = reinterpret_cast<unsigned long long *>(h->d_part), bad = 1;
        if (cudaStreamSynchronize(h->stream) == cudaSuccess &&          // d_part is reduction scratch
            launch_divtau_check(tau, rcp, d, h->stream) == cudaSuccess &&
            cudaMemcpyAsync(&bad, d, sizeof bad, cudaMemcpyDeviceToHost, h->stream) == cudaSuccess &&
            cudaStreamSynchronize(h->stream) == cudaSuccess && bad == 0)
            mode = 0;                              // DM_FAST3
        else
            cudaGetLastError();
    }
    std::lock_guard<std::mutex> lock(mu);
    verdicts[key] = mode;
    return mode;
}

void refresh_params(alb_handle *h) {
    h->u0f = (float)h->u0;
    const float tauf = (float)h->tau;
    if (tauf != h->tauf || h->inv_tau == 0) {
        h->tauf = tauf;
        h->inv_tau = 1.0f / tauf;
        h->div_mode = h->div_forced ? 1 : div_mode_for(h, tauf);
    }
    host_feq0(h->u0f, h->feq0);
}

StepParams make_params(alb_handle *h, int src_idx, int dst_idx = -1, int parity = -1) {
    StepParams p;
    memset(&p, 0, sizeof p);
    p.src = h->f[src_idx];
    p.dst = h->f[dst_idx < 0 ? 1 - src_idx : dst_idx];
    p.info = h->info;
    p.tclass = h->tclass;
    p.gen_list = h->gen_list;
    p.ngen = h->ngen;
    p.plane = h->plane;
    p.pitch = h->pitch;
    p.tpr = h->tpr;
    p.ntasks = h->nyl * h->tpr;
    p.nyl = h->nyl;
    p.nx = h->nx;
    p.tau = h->tauf;
    p.inv_tau = h->inv_tau;
    p.div_mode = h->div_mode;
    p.u0 = h->u0f;
    memcpy(p.feq0, h->feq0, sizeof p.feq0);
    p.rho = h->rho;
    p.ux = h->ux;
    p.uy = h->uy;
    p.clamp_hits = h->clamp_hits;
    p.me = h->me;
    p.parity = parity < 0 ? src_idx : parity;
    return p;
}

// ---- thresholds of the fused diagnostics -------------------------------------------------------
// Cp = (rho-1)/(1.5*U0*U0) is a monotone function of the fp32 rho, so the window -4 < Cp < 1.2 of
// HTML:609 is a closed interval [rho_lo, rho_hi] of floats; it is found by bisection over the
// ordered bit patterns with the very expression the reference evaluates (float64, HTML:605).
double cp_of(float rho, double U0) {
    volatile double c = 1.5 * U0;
    c = c * U0;
    volatile double n = (double)rho - 1;
    return n / c;
}
uint32_t fkey(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
float fromkey(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
void refresh_thresholds(alb_handle *h) {
    if (h->thr_u0 == h->u0) return;
    const double U0 = h->u0;
    h->thr_u0 = U0;
    const double cden = 1.5 * U0 * U0;
    if (!(cden > 0) || !isfinite(cden)) {          // U0 == 0: Cp and u/U0 are inf/NaN, nothing qualifies
        h->rho_lo = INFINITY; h->rho_hi = -INFINITY; h->m2_lo = h->m2_hi = -1;
        return;
    }
    const uint32_t kmin = fkey(-3.0e38f), kmax = fkey(3.0e38f);
    // smallest float with cp > -4
    uint32_t lo = kmin, hi = kmax;
    if (!(cp_of(fromkey(hi), U0) > -4)) { h->rho_lo = INFINITY; }
    else {
        while (lo < hi) { uint32_t mid = lo + (hi - lo) / 2; if (cp_of(fromkey(mid), U0) > -4) hi = mid; else lo = mid + 1; }
        h->rho_lo = fromkey(lo);
    }
    // largest float with cp < 1.2
    lo = kmin; hi = kmax;
    if (!(cp_of(fromkey(lo), U0) < 1.2)) { h->rho_hi = -INFINITY; }
    else {
        while (lo < hi) { uint32_t mid = lo + (hi - lo + 1) / 2; if (cp_of(fromkey(mid), U0) < 1.2) lo = mid; else hi = mid - 1; }
        h->rho_hi = fromkey(lo);
    }
    const double cut = (4 * U0) * (4 * U0);
    h->m2_lo = cut * (1 - 1e-9);
    h->m2_hi = cut * (1 + 1e-9);
}

void arm_diag(alb_handle *h, StepParams &p) {
    refresh_thresholds(h);
    p.diag = h->d_diag;
    p.rho_lo = h->rho_lo;
    p.rho_hi = h->rho_hi;
    p.U0d = h->u0;
    p.m2_lo = h->m2_lo;
    p.m2_hi = h->m2_hi;
    p.m2f_cap = (float)(h->m2_hi * (1 + 1e-5));
}

void arm_diag2(alb_handle *h, Step2Params &q) {
    refresh_thresholds(h);
    q.diag = h->d_diag;
    q.rho_lo = h->rho_lo;
    q.rho_hi = h->rho_hi;
    q.U0d = h->u0;
    q.m2_lo = h->m2_lo;
    q.m2_hi = h->m2_hi;
    q.m2f_cap = (float)(h->m2_hi * (1 + 1e-5));
}

int ensure_prev(alb_handle *h);

// One pass over the previous state that yields the autoscale statistics (HTML:596-614) and the
// pressure-face sums (HTML:649-700) of the current state, optionally also storing rho/ux/uy.
int run_macro_pass(alb_handle *h, bool write_macro) {
    CK(cudaMemcpyAsync(h->d_diag, h->h_diag_init, sizeof(DiagAcc) * DIAG_SLOTS, cudaMemcpyHostToDevice, h->stream));
    h->diag_prearmed = false;
    int r = ensure_prev(h);
    if (r) return r;
    StepParams p = make_params(h, h->prev_idx, h->cur);
    p.write_macro = write_macro ? 1 : 0;
    arm_diag(h, p);
    CK(launch_macro(p, h->stream));
    CK(cudaMemcpyAsync(h->h_diag, h->d_diag, sizeof(DiagAcc) * DIAG_SLOTS, cudaMemcpyDeviceToHost, h->stream));
    h->diag_slots_host = DIAG_SLOTS;
    if (write_macro) h->macro_valid = true;
    h->diag_valid = true;     // h_diag is readable after the next stream synchronisation
    return ALB_OK;
}

// Materialise rho/ux/uy of the current state: they are a function of the
// PREVIOUS state (still intact in the other ping-pong buffer), the mask and the
// parameters the last step ran with -- so this is called before any of those
// change.  Costs one read of the populations instead of 12 B/cell on every step.
int ensure_macro(alb_handle *h) {
    if (h->macro_valid) return ALB_OK;
    return run_macro_pass(h, true);
}

int rebuild_info(alb_handle *h) {
    CK(launch_build_info(h->mask, h->info, h->tclass, h->gen_list, h->gen_count, h->pitch, h->nx,
                         h->ny_global, h->y0 - 1, h->nrows, h->stream));
    // the host needs the number of general tasks to size that kernel's grid (mask changes are rare)
    CK(cudaMemcpyAsync(&h->ngen, h->gen_count, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(launch_build_lists(h->info, h->tclass, h->deep_tmp, h->tflags, h->lists, h->list_counts, h->pitch, h->nrows,
                          h->y0 > 0 ? 1 : 0, h->y0 + h->nyl < h->ny_global ? 1 : 0,
                          march_edges_enabled(h->nx, h->pitch) ? h->nx : 0, h->stream));
    CK(cudaMemcpyAsync(h->nlist, h->list_counts, 5 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    h->solid_synced = false;
    CK(cudaStreamSynchronize(h->stream));
    h->diag_valid = false;      // faces and the solid set changed
    drop_graph(h);              // grid sizes depend on the number of general tasks
    return ALB_OK;
}

int do_reset(alb_handle *h, double u0) {
    // HTML:474-490: float64 equilibrium at rho = 1, u = (u0, 0), rounded to fp32 on store
    const double w0 = 4.0 / 9.0, ws = 1.0 / 9.0, wd = 1.0 / 36.0;
    const double W[9] = {w0, ws, ws, ws, ws, wd, wd, wd, wd};
    const int EX[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
    float e[9];
    for (int i = 0; i < 9; i++) {
        volatile double eu = EX[i] * u0, uu = u0 * u0;
        volatile double t1 = 3 * eu, s1 = 1 + t1, t2 = 4.5 * eu, t3 = t2 * eu, s2 = s1 + t3;
        volatile double t4 = 1.5 * uu, s3 = s2 - t4;
        e[i] = (float)(W[i] * s3);
    }
    h->u0 = u0;
    refresh_params(h);
    CK(launch_fill_init(h->f[0], h->f[1], h->plane, e, h->rho, h->ux, h->uy, (float)u0, h->stream));
    CK(cudaMemsetAsync(h->me, 0, sizeof(MeState), h->stream));
    CK(cudaMemcpyAsync(&h->me->count, &h->sync_steps, sizeof(long long), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));   // sync_steps is host memory
    CK(cudaMemsetAsync(h->clamp_hits, 0, sizeof(unsigned long long), h->stream));
    h->cur = 0;
    h->parity = 0;
    h->prev_idx = 1;
    h->solid_synced = true;     // both buffers were filled alike
    drop_graph(h);
    h->steps = 0;
    h->frame_counter = 0;
    h->macro_valid = true;
    h->ghost_macro_valid = false;
    h->diag_valid = false;
    return ALB_OK;
}

int copy_out_rows(alb_handle *h, void *dst_host, const void *src_dev_row1, size_t elem) {
    // dense [nyl][nx] <- padded rows 1..nyl
    CK(cudaMemcpy2DAsync(dst_host, (size_t)h->nx * elem, src_dev_row1, (size_t)h->pitch * elem,
                         (size_t)h->nx * elem, h->nyl, cudaMemcpyDeviceToHost, h->stream));
    return ALB_OK;
}

void print_trace(alb_handle *h) {
    if (!h->trace_armed) return;
    h->trace_armed = false;
    if (cudaEventSynchronize(h->tev[6]) != cudaSuccess) return;
    const char *names[7] = {"fork", "pass1 flags in", "pass1 done", "pass2 flags in", "pass2 done", "march2_kernel done", "joined"};
    fprintf(stderr, "[alb trace] device %d rows %d..%d, double step at step %lld:", h->device, h->y0, h->y0 + h->nyl - 1,
            h->trace_step);
    for (int k = 1; k < 7; k++) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, h->tev[0], h->tev[k]) == cudaSuccess) fprintf(stderr, "  %s %.3f ms", names[k], ms);
    }
    fprintf(stderr, "\n");
}

int check_wait_error(alb_handle *h) {
    print_trace(h);
    if (h->h_err && *h->h_err) {
        *h->h_err = 0;
        return h->fail(ALB_ERR_TIMEOUT, "timed out waiting for a neighbouring slab's halo");
    }
    return ALB_OK;
}

void free_handle(alb_handle *h) {
    if (!h) return;
    // AEROLAB_LBM_TRACE_DESTROY=1 (measurement aid): where the time of alb_destroy goes
    static const bool trace = getenv("AEROLAB_LBM_TRACE_DESTROY") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto t_prev = now();
    auto lap = [&](const char *what) {
        if (!trace) return;
        const auto t = now();
        fprintf(stderr, "[alb destroy] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t_prev).count());
        t_prev = t;
    };
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    lap("stream sync");
    if (h->lo.ipc_base) cudaIpcCloseMemHandle(h->lo.ipc_base);
    if (h->hi.ipc_base) cudaIpcCloseMemHandle(h->hi.ipc_base);
    cudaFree(h->block);
    lap("populations");
    cudaFree(h->rho);
    cudaFree(h->ux);
    cudaFree(h->uy);
    cudaFree(h->mask);
    cudaFree(h->info);
    cudaFree(h->tclass);
    cudaFree(h->gen_list);
    cudaFree(h->gen_count);
    cudaFree(h->tflags);
    cudaFree(h->deep_tmp);
    for (auto &l : h->lists) cudaFree(l);
    cudaFree(h->list_counts);
    cudaFree(h->s2_queue);
    cudaFree(h->band_inbox);
    cudaFree(h->me);
    cudaFree(h->parts);
    cudaFree(h->d_frame);
    lap("device arrays");
    if (h->h_frame) cudaFreeHost(h->h_frame);
    if (h->h_rows) cudaFreeHost(h->h_rows);
    lap("pinned frame buffers");
    cudaFree(h->part_ctr);
    if (h->graph) cudaGraphExecDestroy(h->graph);
    lap("graph");
    cudaFree(h->clamp_hits);
    cudaFree(h->d_xp);
    cudaFree(h->d_yp);
    cudaFree(h->d_part);
    for (auto &t : h->d_tmp) cudaFree(t);
    for (auto &ev : h->tev)
        if (ev) cudaEventDestroy(ev);
    lap("more device arrays, events");
    if (h->h_part) cudaFreeHost(h->h_part);
    if (h->h_err) cudaFreeHost(h->h_err);
    if (h->h_diag) cudaFreeHost(h->h_diag);
    if (h->h_diag_init) cudaFreeHost(h->h_diag_init);
    lap("pinned small buffers");
    cudaFree(h->d_diag);
    cudaFree(h->d_diag_pub);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->ev_fork2) cudaEventDestroy(h->ev_fork2);
    if (h->ev_join2) cudaEventDestroy(h->ev_join2);
    if (h->aux) cudaStreamDestroy(h->aux);
    if (h->aux2) cudaStreamDestroy(h->aux2);
    if (h->stream) cudaStreamDestroy(h->stream);
    lap("events, streams");
    delete h;
}

int ensure_tmp(alb_handle *h, int k) {
    if (!h->d_tmp[k]) CK(cudaMalloc(&h->d_tmp[k], sizeof(float) * (size_t)h->nx * h->nyl));
    return ALB_OK;
}

}  // namespace

extern "C" {

int alb_version(void) { return ALB_VERSION; }

const char *alb_error_string(int code) {
    switch (code) {
        case ALB_OK: return "ok";
        case ALB_ERR_INVALID: return "invalid argument";
        case ALB_ERR_CUDA: return "CUDA error";
        case ALB_ERR_NOMEM: return "out of memory";
        case ALB_ERR_STATE: return "invalid state";
        case ALB_ERR_TIMEOUT: return "halo wait timed out";
        default: return "unknown error";
    }
}

const char *alb_last_error(const alb_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int alb_device_count(int *count) {
    if (!count) return ALB_ERR_INVALID;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        g_create_error = std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e);
        return ALB_ERR_CUDA;
    }
    return ALB_OK;
}

int alb_create_slab(int nx, int ny_global, int y0, int ny_local, int device, alb_handle **out) {
    if (!out) return ALB_ERR_INVALID;
    *out = nullptr;
    if (nx < 3 || ny_global < 3 || ny_local < 1 || y0 < 0 || y0 + ny_local > ny_global ||
        ny_global > 65000 || nx > (1 << 24) ||
        ((long long)((nx + TASK_CELLS - 1) / TASK_CELLS) * (ny_local + 2)) > 0x7fffffffLL / 2) {
        g_create_error = "alb_create: need 3 <= nx <= 2^24, 3 <= ny <= 65000, 0 <= y0, y0 + ny_local <= ny";
        return ALB_ERR_INVALID;
    }
    cudaGetLastError();   // do not inherit a stale error from unrelated earlier CUDA calls of the process
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device (this library has no CPU fallback): ") +
                         (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return ALB_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) {
        g_create_error = "alb_create: device index out of range";
        return ALB_ERR_INVALID;
    }
    alb_handle *h = new (std::nothrow) alb_handle;
    if (!h) return ALB_ERR_NOMEM;
    h->nx = nx;
    h->ny_global = ny_global;
    h->y0 = y0;
    h->nyl = ny_local;
    h->nrows = ny_local + 2;
    h->pitch = (nx + TASK_CELLS - 1) / TASK_CELLS * TASK_CELLS;
    h->tpr = h->pitch / TASK_CELLS;
    h->plane = (size_t)h->nrows * h->pitch;
    h->device = device;
    int rc = ALB_OK;
    auto body = [&]() -> int {
        CK(cudaSetDevice(device));
        CK(preload_all_kernels(device));
        CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        int prio_lo = 0, prio_hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        // The aux stream runs the general-task kernel of single steps and the list-driven passes of
        // double steps.  HIGH priority: march2_kernel fills every SM (one CTA owns the whole register
        // file), so the short passes only get SMs when one of its CTAs retires (every warp takes a
        // bounded number of units from the queue, see march_plan), and at that moment they must win
        // against the next CTA of the fused kernel -- otherwise pass 2, and with it the neighbouring
        // slabs, would wait for the end of the whole pass.  The units come from a queue, so an SM lent
        // to a pass costs no wave.  (Round 1's strip kernel had whole waves of tiles; there a
        // high-priority aux stream cost 10 %.)  AEROLAB_LBM_AUX_PRIO=0 selects normal priority.
        const char *aux_prio = getenv("AEROLAB_LBM_AUX_PRIO");
        CK(cudaStreamCreateWithPriority(&h->aux, cudaStreamNonBlocking,
                                        aux_prio && atoi(aux_prio) == 0 ? prio_lo : prio_hi));
        if (!(getenv("AEROLAB_LBM_AUX2") && atoi(getenv("AEROLAB_LBM_AUX2")) == 0)) {   // A/B switch for measurements
            CK(cudaStreamCreateWithPriority(&h->aux2, cudaStreamNonBlocking,
                                            aux_prio && atoi(aux_prio) == 0 ? prio_lo : prio_hi));
            CK(cudaEventCreateWithFlags(&h->ev_fork2, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&h->ev_join2, cudaEventDisableTiming));
        }
        CK(cudaEventCreate(&h->ev0));
        CK(cudaEventCreate(&h->ev1));
        CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
        const size_t pop_bytes = sizeof(float) * 27 * h->plane;
        CK(cudaMalloc(&h->block, pop_bytes + 256));
        h->f[0] = reinterpret_cast<float *>(h->block);
        h->f[1] = h->f[0] + 9 * h->plane;
        h->f[2] = h->f[0] + 18 * h->plane;
        h->flags = reinterpret_cast<int *>(h->block + pop_bytes);
        CK(cudaMemsetAsync(h->flags, 0, 256, h->stream));
        CK(cudaMemsetAsync(h->f[2], 0, sizeof(float) * 9 * h->plane, h->stream));   // ghost rows are read before any push
        CK(cudaMalloc(&h->rho, sizeof(float) * h->plane));
        CK(cudaMalloc(&h->ux, sizeof(float) * h->plane));
        CK(cudaMalloc(&h->uy, sizeof(float) * h->plane));
        CK(cudaMalloc(&h->mask, h->plane));
        CK(cudaMalloc(&h->info, sizeof(uint16_t) * h->plane));
        CK(cudaMalloc(&h->tclass, (size_t)h->nrows * h->tpr));
        CK(cudaMalloc(&h->gen_list, sizeof(int) * (size_t)h->nrows * h->tpr));
        CK(cudaMalloc(&h->gen_count, sizeof(int)));
        CK(cudaMalloc(&h->tflags, (size_t)h->nrows * h->tpr));
        CK(cudaMalloc(&h->deep_tmp, (size_t)h->nrows * h->tpr));
        for (auto &l : h->lists) CK(cudaMalloc(&l, sizeof(int) * (size_t)h->nrows * h->tpr));
        CK(cudaMalloc(&h->list_counts, 5 * sizeof(int)));
        CK(cudaMalloc(&h->s2_queue, 2 * sizeof(int)));
        CK(cudaMemsetAsync(h->s2_queue, 0, 2 * sizeof(int), h->stream));
        CK(cudaMalloc(&h->me, sizeof(MeState)));
        CK(cudaMalloc(&h->clamp_hits, sizeof(unsigned long long)));
        CK(cudaMalloc(&h->d_xp, sizeof(double) * 1024));
        CK(cudaMalloc(&h->d_yp, sizeof(double) * 1024));
        CK(cudaMalloc(&h->d_part, sizeof(double) * 4 * DIAG_BLOCKS));
        CK(cudaHostAlloc(&h->h_part, sizeof(double) * 4 * DIAG_BLOCKS, cudaHostAllocDefault));
        CK(cudaMalloc(&h->d_diag, sizeof(DiagAcc) * DIAG_SLOTS));
        CK(cudaMalloc(&h->d_diag_pub, sizeof(DiagAcc)));
        CK(cudaHostAlloc(&h->h_diag, sizeof(DiagAcc) * DIAG_SLOTS, cudaHostAllocDefault));
        CK(cudaHostAlloc(&h->h_diag_init, sizeof(DiagAcc) * DIAG_SLOTS, cudaHostAllocDefault));
        memset(h->h_diag_init, 0, sizeof(DiagAcc) * DIAG_SLOTS);
        for (int k = 0; k < DIAG_SLOTS; k++) {
            h->h_diag_init[k].rho_min = INFINITY;
            h->h_diag_init[k].rho_max = -INFINITY;
        }
        CK(cudaMalloc(&h->d_frame, sizeof(FrameDev)));
        CK(cudaHostAlloc(&h->h_frame, sizeof(FrameDev), cudaHostAllocDefault));
        CK(cudaHostAlloc(&h->h_err, sizeof(int), cudaHostAllocMapped));
        *h->h_err = 0;
        CK(cudaHostGetDevicePointer(&h->d_err, h->h_err, 0));
        CK(cudaMemsetAsync(h->mask, 0, h->plane, h->stream));
        int r = rebuild_info(h);
        if (r) return r;
        r = do_reset(h, 0.06);   // HTML:472, 503
        if (r) return r;
        CK(cudaStreamSynchronize(h->stream));
        h->small_capacity = small_lattice_capacity(device);
        CK(cudaDeviceGetAttribute(&h->nsm, cudaDevAttrMultiProcessorCount, device));
        if (h->whole() && h->small_capacity > 0 && !(getenv("AEROLAB_LBM_BAND") && atoi(getenv("AEROLAB_LBM_BAND")) == 0)) {
            h->band_rows = band_lattice_rows(h->nx, h->nyl, h->nsm);
            if (h->band_rows > 0) {
                const size_t bytes = band_inbox_bytes(h->nx, h->nyl, h->band_rows);
                CK(cudaMalloc(&h->band_inbox, bytes));
                CK(cudaMemset(h->band_inbox, 0, bytes));        // tag 0 is never a step tag
            }
        }
        h->use_graph = getenv("AEROLAB_LBM_NO_GRAPH") == nullptr;   // A/B switch for measurements
        if (const char *e = getenv("AEROLAB_LBM_DOUBLE")) h->double_mode = atoi(e) != 0 ? 1 : 0;
        if (const char *e = getenv("AEROLAB_LBM_TRACE")) {
            h->trace_step = atoll(e);
            for (auto &ev : h->tev) CK(cudaEventCreate(&ev));
        }
        return ALB_OK;
    };
    rc = body();
    if (rc != ALB_OK) {
        g_create_error = h->err;
        free_handle(h);
        return rc;
    }
    *out = h;
    return ALB_OK;
}

int alb_create(int nx, int ny, int device, alb_handle **out) {
    return alb_create_slab(nx, ny, 0, ny, device, out);
}

int alb_destroy(alb_handle *h) {
    if (!h) return ALB_ERR_INVALID;
    free_handle(h);
    return ALB_OK;
}

int alb_get_dims(const alb_handle *h, int *nx, int *ny_global, int *y0, int *ny_local) {
    if (!h) return ALB_ERR_INVALID;
    if (nx) *nx = h->nx;
    if (ny_global) *ny_global = h->ny_global;
    if (y0) *y0 = h->y0;
    if (ny_local) *ny_local = h->nyl;
    return ALB_OK;
}

int alb_set_params(alb_handle *h, double u0, double tau) {
    NEED(h);
    NO_PENDING_FRAMES(h);
    ARG(isfinite(u0) && isfinite(tau) && (float)tau != 0.0f, "alb_set_params: u0 and tau must be finite, tau != 0");
    if (u0 == h->u0 && tau == h->tau) return ALB_OK;
    int r = ensure_macro(h);
    if (r) return r;
    h->u0 = u0;
    h->tau = tau;
    h->diag_valid = false;      // statistics and force normalisation depend on U0
    drop_graph(h);              // tau, U0 and the inlet constants are kernel arguments
    refresh_params(h);
    return ALB_OK;
}

int alb_get_params(const alb_handle *h, double *u0, double *tau) {
    if (!h) return ALB_ERR_INVALID;
    if (u0) *u0 = h->u0;
    if (tau) *tau = h->tau;
    return ALB_OK;
}

int alb_reset(alb_handle *h, double u0) {
    NEED(h);
    NO_PENDING_FRAMES(h);
    ARG(isfinite(u0), "alb_reset: u0 must be finite");
    return do_reset(h, u0);
}

int alb_rasterize_panels(alb_handle *h, const double *xp, const double *yp, int n, uint8_t *mask_out) {
    NEED(h);
    NO_PENDING_FRAMES(h);
    ARG(xp && yp && n >= 2 && n <= 1024, "alb_rasterize_panels: need 2 <= n <= 1024 panel nodes");
    int r = ensure_macro(h);
    if (r) return r;
    CK(cudaMemcpyAsync(h->d_xp, xp, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_yp, yp, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));   // xp/yp are caller memory
    CK(launch_raster(h->d_xp, h->d_yp, n, h->mask, h->pitch, h->nx, h->ny_global, h->y0 - 1, h->nrows,
                     h->stream));
    r = rebuild_info(h);
    if (r) return r;
    if (mask_out) {
        r = copy_out_rows(h, mask_out, h->mask + h->pitch, 1);
        if (r) return r;
        CK(cudaStreamSynchronize(h->stream));
    }
    return ALB_OK;
}

int alb_rasterize(alb_handle *h, const double *xy, int npts, double alpha_deg, uint8_t *mask_out) {
    NEED(h);
    ARG(xy && npts >= 2 && npts <= 100000, "alb_rasterize: need 2 <= npts <= 100000 coordinate pairs");
    ARG(isfinite(alpha_deg), "alb_rasterize: alpha must be finite");
    for (int i = 0; i < 2 * npts; i++) ARG(isfinite(xy[i]), "alb_rasterize: coordinates must be finite");
    try {
        host_rotate_panelise(xy, npts, alpha_deg, h->xp, h->yp);
    } catch (const std::bad_alloc &) {
        return h->fail(ALB_ERR_NOMEM, "alb_rasterize: host allocation failed");
    }
    h->have_panels = true;
    return alb_rasterize_panels(h, h->xp, h->yp, ALB_NPANEL + 1, mask_out);
}

int alb_set_mask(alb_handle *h, const uint8_t *mask_global) {
    NEED(h);
    NO_PENDING_FRAMES(h);
    ARG(mask_global, "alb_set_mask: mask is NULL");
    int r = ensure_macro(h);
    if (r) return r;
    CK(cudaMemsetAsync(h->mask, 0, h->plane, h->stream));
    const int gy_lo = h->y0 - 1 < 0 ? 0 : h->y0 - 1;
    const int gy_hi = h->y0 + h->nyl + 1 > h->ny_global ? h->ny_global : h->y0 + h->nyl + 1;
    const int j_lo = gy_lo - (h->y0 - 1);
    CK(cudaMemcpy2DAsync(h->mask + (size_t)j_lo * h->pitch, h->pitch, mask_global + (size_t)gy_lo * h->nx,
                         h->nx, h->nx, gy_hi - gy_lo, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return rebuild_info(h);
}

int alb_get_mask(alb_handle *h, uint8_t *mask_out) {
    NEED(h);
    ARG(mask_out, "alb_get_mask: output is NULL");
    int r = copy_out_rows(h, mask_out, h->mask + h->pitch, 1);
    if (r) return r;
    CK(cudaStreamSynchronize(h->stream));
    return ALB_OK;
}

int alb_get_panels(const alb_handle *h, double *xp, double *yp) {
    if (!h) return ALB_ERR_INVALID;
    if (!h->have_panels) return const_cast<alb_handle *>(h)->fail(ALB_ERR_STATE, "alb_get_panels: no alb_rasterize() yet");
    if (xp) memcpy(xp, h->xp, sizeof h->xp);
    if (yp) memcpy(yp, h->yp, sizeof h->yp);
    return ALB_OK;
}

}  // extern "C" (reopened below)

namespace {

void set_peers(alb_handle *h, StepParams &p, int dst_idx) {
    if (h->lo.base) {
        p.peer_lo_dst = h->lo.base + (size_t)dst_idx * 9 * h->lo.plane;
        p.peer_lo_plane = h->lo.plane;
        p.peer_lo_row = (size_t)(h->lo.nyl + 1) * h->pitch;
    }
    if (h->hi.base) {
        p.peer_hi_dst = h->hi.base + (size_t)dst_idx * 9 * h->hi.plane;
        p.peer_hi_plane = h->hi.plane;
        p.peer_hi_row = 0;
    }
}

// my step k needs the neighbours' k completed steps: their edge rows of state k are in my ghost
// rows, and they no longer read the ghost rows I am about to overwrite.
void halo_wait(alb_handle *h, long long sync_step, cudaStream_t st) {
    wait_kernel<<<1, 1, 0, st>>>(h->lo.base ? h->flags + 0 : nullptr, h->hi.base ? h->flags + 1 : nullptr,
                                 (int)sync_step, h->d_err, WAIT_TIMEOUT_NS);
}
// I am the UPPER neighbour of lo (its flags[1]) and the LOWER neighbour of hi (its flags[0])
void halo_signal(alb_handle *h, long long steps_done, cudaStream_t st) {
    signal_kernel<<<1, 1, 0, st>>>(h->lo.flags ? h->lo.flags + 1 : nullptr, h->hi.flags ? h->hi.flags + 0 : nullptr,
                                   (int)steps_done);
}

// AEROLAB_LBM_DEFER_SIGNAL=0 (diagnosis only): send the closing signal of a batch at once, as round 1 did
bool defer_closing_signal() {
    static const bool on = !(getenv("AEROLAB_LBM_DEFER_SIGNAL") && atoi(getenv("AEROLAB_LBM_DEFER_SIGNAL")) == 0);
    return on;
}

// The previous batch's closing signal: from here on the neighbours may overwrite the ghost rows of
// what was this slab's previous state.
void flush_pending_signal(alb_handle *h) {
    if (h->pending_signal < 0) return;
    halo_signal(h, h->pending_signal, h->stream);
    h->launches++;
    h->pending_signal = -1;
}

// Enqueue one step that reads buffer src_idx.  Used directly and under stream capture.
int issue_step(alb_handle *h, int src_idx, int parity, bool halo, long long sync_step, bool diag, bool last_of_batch) {
    StepParams p = make_params(h, src_idx, 1 - src_idx, parity);
    h->solid_synced = false;    // solid cells swap their populations: the two buffers differ there now
    if (diag) {
        // the last step of a batch also reduces the statistics / face sums of the state it writes
        if (!h->diag_prearmed)
            CK(cudaMemcpyAsync(h->d_diag, h->h_diag_init, sizeof(DiagAcc) * DIAG_SLOTS, cudaMemcpyHostToDevice, h->stream));
        h->diag_prearmed = false;
        arm_diag(h, p);
    }
    if (halo) {
        halo_wait(h, sync_step, h->stream);
        set_peers(h, p, 1 - src_idx);
        h->launches += 2;       // wait + signal
    }
    h->launches += (p.ntasks <= UNIFIED_MAX_TASKS || p.ngen == 0) ? 1 : 2;
    if (p.ntasks <= UNIFIED_MAX_TASKS) {
        // small lattice: launch-latency bound, one launch for both paths
        CK(launch_step_unified(p, h->stream));
    } else {
        if (p.ngen > 0) {
            // fork: the general-task kernel runs on the aux stream beside the fast kernel
            CK(cudaEventRecord(h->ev_fork, h->stream));
            CK(cudaStreamWaitEvent(h->aux, h->ev_fork, 0));
            CK(launch_step_general(p, h->aux));
            CK(cudaEventRecord(h->ev_join, h->aux));
        }
        CK(launch_step_fast(p, h->stream));
        if (p.ngen > 0) CK(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    }
    if (halo) {
        if (last_of_batch && defer_closing_signal()) h->pending_signal = sync_step + 1;
        else halo_signal(h, sync_step + 1, h->stream);
    }
    return ALB_OK;
}

// Enqueue TWO steps that read buffer src_idx and leave the result in buffer 1 - src_idx.
//   main stream: march2_kernel -- every deep task, two steps per pass over HBM (36 B per cell update)
//   aux stream:  the two-pass path for everything else: pass 1 writes the intermediate state of the
//                shallow tasks and their neighbours into f[2], pass 2 advances the shallow tasks from
//                f[2] into the destination.  The slab halo (edge rows are always shallow) is pushed
//                by these passes exactly as by single steps, so march2_kernel needs no flags at all:
//                next to a neighbouring slab TWO edge rows are shallow and it never reads a ghost row.
int issue_double(alb_handle *h, int src_idx, int parity, bool halo, long long sync_step, bool copy_solid,
                 bool diag = false, bool last_of_batch = false) {
    const int dst_idx = 1 - src_idx;
    if (diag) {
        // a batch that ends with a double step: its second step reduces the statistics / face sums
        if (!h->diag_prearmed)
            CK(cudaMemcpyAsync(h->d_diag, h->h_diag_init, sizeof(DiagAcc) * DIAG_SLOTS, cudaMemcpyHostToDevice, h->stream));
        h->diag_prearmed = false;
    }
    const bool trace = h->trace_step >= 0 && sync_step == h->trace_step && h->tev[0];
    if (trace) CK(cudaEventRecord(h->tev[0], h->stream));
    CK(cudaEventRecord(h->ev_fork, h->stream));
    CK(cudaStreamWaitEvent(h->aux, h->ev_fork, 0));
    for (int pass = 0; pass < 2; pass++) {
        StepParams p = pass == 0 ? make_params(h, src_idx, 2, parity) : make_params(h, 2, dst_idx, parity ^ 1);
        if (halo) {
            halo_wait(h, sync_step + pass, h->aux);
            set_peers(h, p, pass == 0 ? 2 : dst_idx);
        }
        if (diag && pass == 1) arm_diag(h, p);
        if (trace) CK(cudaEventRecord(h->tev[1 + 2 * pass], h->aux));      // the neighbours' flags have arrived
        // the two kernels of a pass write disjoint tasks (and the momentum-exchange sums of a step go to an
        // accumulator that its own bookkeeping does not touch, as in single steps): side by side
        const bool side = h->aux2 && h->nlist[2 * pass + 1] > 0;
        if (side) {
            CK(cudaEventRecord(h->ev_fork2, h->aux));
            CK(cudaStreamWaitEvent(h->aux2, h->ev_fork2, 0));
        }
        p.gen_list = h->lists[2 * pass];
        p.ngen = h->nlist[2 * pass];
        CK(launch_step_fast_list(p, h->aux));
        p.gen_list = h->lists[2 * pass + 1];
        p.ngen = h->nlist[2 * pass + 1];
        CK(launch_step_general(p, side ? h->aux2 : h->aux));
        if (side) {
            CK(cudaEventRecord(h->ev_join2, h->aux2));
            CK(cudaStreamWaitEvent(h->aux, h->ev_join2, 0));
        }
        h->launches += 1 + (p.ngen > 0 ? 1 : 0) + (halo ? 2 : 0);
        if (pass == 1 && copy_solid && h->nlist[4] > 0) {
            // all-solid tasks return to their state after two steps: copy, unless the destination
            // still holds the same values from the previous double step
            StepParams c = make_params(h, src_idx, dst_idx, parity);
            c.gen_list = h->lists[4];
            c.ngen = h->nlist[4];
            CK(launch_copy_tasks(c, h->aux));
            h->launches++;
        }
        if (halo) {
            if (pass == 1 && last_of_batch && defer_closing_signal()) h->pending_signal = sync_step + 2;
            else halo_signal(h, sync_step + pass + 1, h->aux);
        }
        if (trace) CK(cudaEventRecord(h->tev[2 + 2 * pass], h->aux));      // pass finished and signalled
    }
    CK(cudaEventRecord(h->ev_join, h->aux));
    Step2Params q;
    memset(&q, 0, sizeof q);
    q.src = h->f[src_idx];
    q.dst = h->f[dst_idx];
    q.tflags = h->tflags;
    q.plane = h->plane;
    q.pitch = h->pitch;
    q.tpr = h->tpr;
    q.nyl = h->nyl;
    q.tau = h->tauf;
    q.inv_tau = h->inv_tau;
    q.div_mode = h->div_mode;
    q.clamp_hits = h->clamp_hits;
    q.queue = h->s2_queue;
    q.edges = march_edges_enabled(h->nx, h->pitch) ? 1 : 0;
    q.nx = h->nx;
    q.u0 = h->u0f;
    memcpy(q.feq0, h->feq0, sizeof q.feq0);
    if (diag) arm_diag2(h, q);
    // no flags here: next to a neighbouring slab two edge rows are shallow, the fused kernel never
    // reads a ghost row, and the GPUs only meet in the short list-driven passes
    march_plan(q, h->nsm);
    CK(launch_march2(q, h->nsm, h->stream));
    h->launches += q.ntiles > 0 ? 1 : 0;
    if (trace) CK(cudaEventRecord(h->tev[5], h->stream));                  // fused kernel finished
    CK(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    if (trace) {
        CK(cudaEventRecord(h->tev[6], h->stream));                         // joined
        h->trace_armed = true;
    }
    return ALB_OK;
}

}  // namespace

// The macroscopic fields of the current state are a function of the state one step earlier.  After
// a batch that ended with a double step that state was never written: advance the state of two
// steps ago (still intact in the other ping-pong buffer) by one step into the scratch buffer,
// without any side effect (no momentum-exchange bookkeeping, no clamp counting, no halo push --
// the ghost rows of the scratch buffer still hold the neighbours' rows of exactly that state).
namespace {
int ensure_prev(alb_handle *h) {
    if (h->prev_idx >= 0) return ALB_OK;
    StepParams p = make_params(h, 1 - h->cur, 2, 0);
    p.me = nullptr;
    p.clamp_hits = nullptr;
    if (p.ntasks <= UNIFIED_MAX_TASKS) {
        CK(launch_step_unified(p, h->stream));
    } else {
        CK(launch_step_fast(p, h->stream));
        CK(launch_step_general(p, h->stream));
    }
    h->prev_idx = 2;
    return ALB_OK;
}

// Two steps per pass pay off once the fused kernel has enough tiles to fill the GPU; all slabs of
// a lattice must decide alike (they address each other's buffers by index), so the rule only
// looks at the global lattice and the environment.
bool double_steps_enabled(const alb_handle *h) {
    if (h->external_halo) return false;
    if (h->double_mode >= 0) return h->double_mode != 0;
    return h->nx >= DOUBLE_MIN_NX && (long long)h->nx * h->ny_global >= DOUBLE_MIN_CELLS;
}

// Capture GRAPH_STEPS steps starting from buffer 0.  Every address the kernels touch depends only
// on the step's parity (MeState), so the instantiated graph can be replayed for the whole run; it
// is dropped whenever a kernel argument changes (mask, parameters, neighbours).
int capture_graph(alb_handle *h, bool doubles) {
    cudaGraph_t g = nullptr;
    CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    int r = ALB_OK;
    int cur = 0, parity = h->parity;
    const long long launches_before = h->launches;
    for (int s = 0; s < GRAPH_STEPS && r == ALB_OK;) {
        if (doubles) {
            r = issue_double(h, cur, parity, false, 0, s == 0);   // replayed after anything: the first one always copies
            s += 2;
        } else {
            r = issue_step(h, cur, parity, false, 0, false, false);
            parity ^= 1;
            s += 1;
        }
        cur ^= 1;
    }
    cudaError_t e = cudaStreamEndCapture(h->stream, &g);
    if (r != ALB_OK) {
        if (g) cudaGraphDestroy(g);
        return r;
    }
    if (e != cudaSuccess) return h->fail(ALB_ERR_CUDA, "cudaStreamEndCapture", e);
    e = cudaGraphInstantiate(&h->graph, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) {
        h->graph = nullptr;
        return h->fail(ALB_ERR_CUDA, "cudaGraphInstantiate", e);
    }
    h->graph_parity = h->parity;
    h->graph_launches = h->launches - launches_before;   // captured, not executed
    h->launches = launches_before;
    return ALB_OK;
}

}  // namespace

extern "C" {

// nsteps >= 1 steps; the last one also reduces the new state's statistics / face sums
static int step_batch(alb_handle *h, int nsteps) {
    // AEROLAB_LBM_FAKE_HALO=1 (measurement aid): issue the wait/signal kernels of the slab protocol even
    // without neighbours (they return at once), to see what the protocol itself costs on one GPU
    static const bool fake_halo = getenv("AEROLAB_LBM_FAKE_HALO") != nullptr;
    const bool halo = !h->external_halo && (h->lo.base || h->hi.base || fake_halo);
    int left = nsteps;
    flush_pending_signal(h);
    const bool doubles = double_steps_enabled(h);
    const bool persistent = !doubles && h->whole() && !halo && !h->external_halo && nsteps >= 2 &&
                            (long long)h->nx * h->nyl <= h->small_capacity;
    if (persistent && h->band_rows > 0 && h->parity == h->cur) {
        // small lattice, state in registers for the whole batch (band_lattice_kernel): chunks of at
        // most ME_RING / 2 steps, the last one also reduces the statistics / face sums of the final state
        if (!h->diag_prearmed)
            CK(cudaMemcpyAsync(h->d_diag, h->h_diag_init, sizeof(DiagAcc) * DIAG_SLOTS, cudaMemcpyHostToDevice, h->stream));
        h->diag_prearmed = false;
        int cur = h->cur;
        for (int done_b = 0; done_b < nsteps;) {
            int n = nsteps - done_b;
            if (n > ME_RING / 2) n = ME_RING / 2;
            if (nsteps - done_b - n == 1) n--;             // never leave a single step for the last chunk
            StepParams p = make_params(h, cur);
            if (done_b + n == nsteps) arm_diag(h, p);
            CK(launch_band_lattice(p, h->f[0], h->f[1], cur, n, h->band_rows, h->band_inbox, h->sync_steps + done_b,
                                   h->d_err, h->stream));
            h->launches += 3;
            cur = (cur + n) & 1;
            done_b += n;
        }
        h->cur = (h->cur + nsteps) & 1;
        h->parity = h->cur;
        h->prev_idx = 1 - h->cur;
        h->solid_synced = false;
        left = 0;
    } else if (persistent) {
        // small lattice: the whole batch of steps in ONE cooperative launch (grid barrier per
        // step); its last iteration also reduces the statistics / face sums of the final state
        // (measured: as fast as a kernel without that code, and no extra launch)
        StepParams p = make_params(h, h->cur);        // parity == cur: no double step ever ran on this handle
        if (!h->diag_prearmed)
            CK(cudaMemcpyAsync(h->d_diag, h->h_diag_init, sizeof(DiagAcc) * DIAG_SLOTS, cudaMemcpyHostToDevice, h->stream));
        h->diag_prearmed = false;
        arm_diag(h, p);
        CK(launch_small_lattice(p, h->f[0], h->f[1], h->cur, nsteps, h->stream));
        h->launches++;
        h->cur = (h->cur + nsteps) & 1;
        h->parity = h->cur;
        h->prev_idx = 1 - h->cur;
        h->solid_synced = false;
        left = 0;
    }
    const bool use_graph = !persistent && !halo && !h->external_halo && h->use_graph && nsteps >= GRAPH_STEPS + 1;
    int done = nsteps - left;
    while (left > 0) {
        if (use_graph && h->cur == 0 && left > GRAPH_STEPS) {   // ">": the last step is issued below, with diagnostics
            if (h->graph && h->graph_parity != h->parity) drop_graph(h);
            if (!h->graph) {
                int r = capture_graph(h, doubles);
                if (r) return r;
            }
            CK(cudaGraphLaunch(h->graph, h->stream));
            h->launches += h->graph_launches;
            h->solid_synced = doubles;
            h->prev_idx = doubles ? -1 : 1;      // the graph starts at cur == 0 and has an even number of steps
            left -= GRAPH_STEPS;           // an even number of steps: cur and parity are unchanged
            done += GRAPH_STEPS;
            continue;
        }
        if (doubles && left >= 2 && left != 3) {
            // a batch may END with a double step (which then carries the diagnostics); the state in
            // between is not materialised, ensure_prev() recomputes it if the fields are asked for
            int r = issue_double(h, h->cur, h->parity, halo, h->sync_steps + done, !h->solid_synced, left == 2, left == 2);
            if (r) return r;
            h->solid_synced = true;
            h->prev_idx = -1;
            h->cur = 1 - h->cur;           // two steps: parity unchanged
            left -= 2;
            done += 2;
            continue;
        }
        int r = issue_step(h, h->cur, h->parity, halo, h->sync_steps + done, left == 1, left == 1);
        if (r) return r;
        h->prev_idx = h->cur;
        h->cur = 1 - h->cur;
        h->parity ^= 1;
        left--;
        done++;
    }
    h->steps += nsteps;
    h->sync_steps += nsteps;
    CK(cudaGetLastError());
    h->macro_valid = false;
    h->ghost_macro_valid = false;
    h->diag_valid = true;       // the last step reduced the new state's statistics and face sums
    return ALB_OK;
}

int alb_step(alb_handle *h, int nsteps) {
    NEED(h);
    NO_PENDING_FRAMES(h);
    ARG(nsteps >= 0, "alb_step: nsteps must be >= 0");
    if (nsteps == 0) return ALB_OK;
    CK(cudaEventRecord(h->ev0, h->stream));
    int r = step_batch(h, nsteps);
    if (r) return r;
    CK(cudaEventRecord(h->ev1, h->stream));
    CK(cudaMemcpyAsync(h->h_diag, h->d_diag, sizeof(DiagAcc) * DIAG_SLOTS, cudaMemcpyDeviceToHost, h->stream));
    h->diag_slots_host = DIAG_SLOTS;
    h->timed = true;
    return ALB_OK;
}

int alb_frames_enqueue(alb_handle *h, int nframes, int steps_per_frame, int forces_every, const double *controls) {
    NEED(h);
    ARG(nframes >= 0 && steps_per_frame >= 1, "alb_run_frames: need nframes >= 0 and steps_per_frame >= 1");
    // a slab of a decomposed lattice records the raw partial reductions of every frame (see
    // alb_frames_collect); the caller combines the slabs
    const bool partial = !h->whole() || h->lo.base || h->hi.base || h->external_halo;
    if (h->frames_pending) return h->fail(ALB_ERR_STATE, "alb_frames_enqueue: collect the previous batch first");
    if (nframes == 0) return ALB_OK;
    if (controls)
        for (int f = 0; f < 2 * nframes; f++) ARG(isfinite(controls[f]), "alb_run_frames: controls must be finite");
    if (nframes > h->rows_cap) {
        if (h->h_rows) cudaFreeHost(h->h_rows);
        h->d_rows = nullptr; h->h_rows = nullptr; h->rows_cap = 0;
        // mapped pinned memory: the finalize kernel stores each frame's record straight to the host
        CK(cudaHostAlloc(&h->h_rows, sizeof(double) * FRAME_ROW * nframes, cudaHostAllocMapped));
        CK(cudaHostGetDevicePointer(&h->d_rows, h->h_rows, 0));
        h->rows_cap = nframes;
    }
    // sticky host state -> device
    FrameDev st;
    st.maxS = h->maxS; st.cpMin = h->cpMin; st.cpMax = h->cpMax;
    st.cl_smooth = h->cl_smooth; st.cd_smooth = h->cd_smooth; st.sep_frac = h->sep_frac;
    st.ema_valid = h->ema_valid ? 1 : 0; st.pad = 0;
    CK(cudaStreamSynchronize(h->stream));          // h_frame may still be the target of an earlier copy
    *h->h_frame = st;
    CK(cudaMemcpyAsync(h->d_frame, h->h_frame, sizeof(FrameDev), cudaMemcpyHostToDevice, h->stream));
    CK(cudaEventRecord(h->ev0, h->stream));
    for (int f = 0; f < nframes; f++) {
        if (controls) {
            const double u0 = controls[2 * f], tau = controls[2 * f + 1];
            ARG((float)tau != 0.0f, "alb_run_frames: tau must be non-zero");
            if (u0 != h->u0 || tau != h->tau) {
                // slider change between frames (HTML:956-959): only the next steps see it
                h->u0 = u0;
                h->tau = tau;
                drop_graph(h);
                refresh_params(h);
            }
        }
        int r = step_batch(h, steps_per_frame);
        if (r) return r;
        h->frame_counter++;
        const int do_forces = forces_every > 0 && (h->frame_counter % forces_every == 0);
        CK(launch_frame_finalize(h->d_diag, h->d_diag_pub, h->me, h->parity ^ 1, partial ? nullptr : h->d_frame, do_forces,
                                 h->u0, h->qdyn(), h->d_rows + (size_t)FRAME_ROW * f, h->stream));
        h->diag_prearmed = true;
    }
    CK(cudaEventRecord(h->ev1, h->stream));
    h->timed = true;
    CK(cudaMemcpyAsync(h->h_diag, h->d_diag_pub, sizeof(DiagAcc), cudaMemcpyDeviceToHost, h->stream));
    h->diag_slots_host = 1;
    CK(cudaMemcpyAsync(h->h_frame, h->d_frame, sizeof(FrameDev), cudaMemcpyDeviceToHost, h->stream));
    h->frames_pending = nframes;
    h->frames_partial = partial;
    return ALB_OK;
}

int alb_frames_collect(alb_handle *h, double *series) {
    NEED(h);
    const int nframes = h->frames_pending;
    if (nframes == 0) return ALB_OK;
    CK(cudaStreamSynchronize(h->stream));
    h->frames_pending = 0;
    if (!h->frames_partial) {
        const FrameDev &o = *h->h_frame;
        h->maxS = o.maxS; h->cpMin = o.cpMin; h->cpMax = o.cpMax;
        h->cl_smooth = o.cl_smooth; h->cd_smooth = o.cd_smooth; h->sep_frac = o.sep_frac;
        h->ema_valid = o.ema_valid != 0;
    }
    if (series) memcpy(series, h->h_rows, sizeof(double) * FRAME_ROW * nframes);
    return check_wait_error(h);
}

int alb_run_frames(alb_handle *h, int nframes, int steps_per_frame, int forces_every, const double *controls,
                   double *series) {
    int r = alb_frames_enqueue(h, nframes, steps_per_frame, forces_every, controls);
    if (r) return r;
    return alb_frames_collect(h, series);
}

int alb_sync(alb_handle *h) {
    NEED(h);
    CK(cudaStreamSynchronize(h->stream));
    return check_wait_error(h);
}

int alb_step_count(const alb_handle *h, long long *steps) {
    if (!h || !steps) return ALB_ERR_INVALID;
    *steps = h->steps;
    return ALB_OK;
}

int alb_last_step_ms(alb_handle *h, float *ms) {
    NEED(h);
    ARG(ms, "alb_last_step_ms: output is NULL");
    if (!h->timed) return h->fail(ALB_ERR_STATE, "alb_last_step_ms: no alb_step() yet");
    CK(cudaEventSynchronize(h->ev1));
    CK(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return check_wait_error(h);
}

int alb_get_populations(alb_handle *h, float *f) {
    NEED(h);
    ARG(f, "alb_get_populations: output is NULL");
    const size_t dense = (size_t)h->nx * h->nyl;
    for (int i = 0; i < 9; i++) {
        int r = copy_out_rows(h, f + i * dense, h->f[h->cur] + i * h->plane + h->pitch, sizeof(float));
        if (r) return r;
    }
    CK(cudaStreamSynchronize(h->stream));
    return check_wait_error(h);
}

int alb_get_population_rows(alb_handle *h, int row0, int nrows, float *f) {
    NEED(h);
    ARG(f && nrows >= 1 && row0 >= -1 && row0 + nrows <= h->nyl + 1,
        "alb_get_population_rows: need -1 <= row0, nrows >= 1, row0 + nrows <= ny_local + 1");
    const size_t dense = (size_t)h->nx * nrows;
    for (int i = 0; i < 9; i++)
        CK(cudaMemcpy2DAsync(f + i * dense, (size_t)h->nx * sizeof(float),
                             h->f[h->cur] + i * h->plane + (size_t)(row0 + 1) * h->pitch, (size_t)h->pitch * sizeof(float),
                             (size_t)h->nx * sizeof(float), nrows, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return check_wait_error(h);
}

int alb_state_hash(alb_handle *h, unsigned long long *out9) {
    NEED(h);
    ARG(out9, "alb_state_hash: output is NULL");
    unsigned long long *d = reinterpret_cast<unsigned long long *>(h->d_part);   // 9 words of the reduction scratch
    CK(launch_state_hash(h->f[h->cur], h->plane, h->pitch, h->nx, h->nyl, h->y0, d, h->stream));
    CK(cudaMemcpyAsync(out9, d, 9 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return check_wait_error(h);
}

int alb_set_populations(alb_handle *h, const float *f) {
    NEED(h);
    NO_PENDING_FRAMES(h);
    ARG(f, "alb_set_populations: input is NULL");
    int r = ensure_macro(h);
    if (r) return r;
    const size_t dense = (size_t)h->nx * h->nyl;
    for (int i = 0; i < 9; i++)
        CK(cudaMemcpy2DAsync(h->f[h->cur] + i * h->plane + h->pitch, sizeof(float) * h->pitch, f + i * dense,
                             sizeof(float) * h->nx, sizeof(float) * h->nx, h->nyl, cudaMemcpyHostToDevice,
                             h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->diag_valid = false;
    h->solid_synced = false;
    return ALB_OK;
}

int alb_get_macro(alb_handle *h, float *rho, float *ux, float *uy) {
    NEED(h);
    int r = ensure_macro(h);
    if (r) return r;
    if (rho && (r = copy_out_rows(h, rho, h->rho + h->pitch, sizeof(float)))) return r;
    if (ux && (r = copy_out_rows(h, ux, h->ux + h->pitch, sizeof(float)))) return r;
    if (uy && (r = copy_out_rows(h, uy, h->uy + h->pitch, sizeof(float)))) return r;
    CK(cudaStreamSynchronize(h->stream));
    return check_wait_error(h);
}

int alb_set_macro(alb_handle *h, const float *rho, const float *ux, const float *uy) {
    NEED(h);
    NO_PENDING_FRAMES(h);
    const float *srcs[3] = {rho, ux, uy};
    float *dsts[3] = {h->rho, h->ux, h->uy};
    for (int k = 0; k < 3; k++)
        if (srcs[k])
            CK(cudaMemcpy2DAsync(dsts[k] + h->pitch, sizeof(float) * h->pitch, srcs[k], sizeof(float) * h->nx,
                                 sizeof(float) * h->nx, h->nyl, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->macro_valid = true;
    h->ghost_macro_valid = false;
    h->diag_valid = false;
    return ALB_OK;
}

int alb_total_mass(alb_handle *h, double *mass) {
    NEED(h);
    ARG(mass, "alb_total_mass: output is NULL");
    CK(launch_mass(h->f[h->cur], h->plane, h->pitch, h->nx, h->nyl, h->d_part, DIAG_BLOCKS, h->stream));
    CK(cudaMemcpyAsync(h->h_part, h->d_part, sizeof(double) * DIAG_BLOCKS, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    double m = 0;
    for (int b = 0; b < DIAG_BLOCKS; b++) m += h->h_part[b];
    *mass = m;
    return ALB_OK;
}

// merge the accumulator copies that the last device->host copy delivered (call after a sync)
static DiagAcc merged_diag(const alb_handle *h) {
    DiagAcc m = h->h_diag[0];
    for (int k = 1; k < h->diag_slots_host; k++) {
        const DiagAcc &a = h->h_diag[k];
        if (a.smax_bits > m.smax_bits) m.smax_bits = a.smax_bits;
        if (a.m2max_bits > m.m2max_bits) m.m2max_bits = a.m2max_bits;
        m.rho_min = fminf(m.rho_min, a.rho_min);
        m.rho_max = fmaxf(m.rho_max, a.rho_max);
        m.fx += a.fx; m.fy += a.fy; m.surf += a.surf; m.rev += a.rev;
    }
    return m;
}

// true when the fused pass can serve the request: the previous state is still the source of the
// current macroscopic fields (i.e. they were not injected by reset/set_macro)
static int fused_diag(alb_handle *h) {
    if (!h->diag_valid) {
        if (h->macro_valid) return 1;            // fields exist only as arrays: use the array kernels
        int r = run_macro_pass(h, false);
        if (r) return r;
    }
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return h->fail(ALB_ERR_CUDA, "cudaStreamSynchronize", e);
    return ALB_OK;
}

int alb_stats_partial(alb_handle *h, double *out3, float *U, float *V, float *Cp) {
    NEED(h);
    int r;
    if (!U && !V && !Cp) {
        r = fused_diag(h);
        if (r < 0) return r;
        if (r == ALB_OK) {
            const DiagAcc d = merged_diag(h);
            double smax;
            memcpy(&smax, &d.smax_bits, 8);
            if (out3) {
                out3[0] = smax;                                                    // 0 when nothing qualified
                out3[1] = d.rho_min <= d.rho_max ? cp_of(d.rho_min, h->u0) : INFINITY;
                out3[2] = d.rho_min <= d.rho_max ? cp_of(d.rho_max, h->u0) : -INFINITY;
            }
            return check_wait_error(h);
        }
    }
    r = ensure_macro(h);
    if (r) return r;
    float *dU = nullptr, *dV = nullptr, *dC = nullptr;
    if (U) { if ((r = ensure_tmp(h, 0))) return r; dU = h->d_tmp[0]; }
    if (V) { if ((r = ensure_tmp(h, 1))) return r; dV = h->d_tmp[1]; }
    if (Cp) { if ((r = ensure_tmp(h, 2))) return r; dC = h->d_tmp[2]; }
    CK(launch_stats(h->mask, h->rho, h->ux, h->uy, h->pitch, h->nx, h->nyl, h->u0, dU, dV, dC, h->d_part,
                    DIAG_BLOCKS, h->stream));
    CK(cudaMemcpyAsync(h->h_part, h->d_part, sizeof(double) * 3 * DIAG_BLOCKS, cudaMemcpyDeviceToHost, h->stream));
    const size_t bytes = sizeof(float) * (size_t)h->nx * h->nyl;
    if (U) CK(cudaMemcpyAsync(U, dU, bytes, cudaMemcpyDeviceToHost, h->stream));
    if (V) CK(cudaMemcpyAsync(V, dV, bytes, cudaMemcpyDeviceToHost, h->stream));
    if (Cp) CK(cudaMemcpyAsync(Cp, dC, bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    double mx = 0, cmin = INFINITY, cmax = -INFINITY;
    for (int b = 0; b < DIAG_BLOCKS; b++) {
        mx = fmax(mx, h->h_part[3 * b]);
        cmin = fmin(cmin, h->h_part[3 * b + 1]);
        cmax = fmax(cmax, h->h_part[3 * b + 2]);
    }
    if (out3) { out3[0] = mx; out3[1] = cmin; out3[2] = cmax; }
    return check_wait_error(h);
}

int alb_update_stats(alb_handle *h, double *stats, float *U, float *V, float *Cp) {
    double raw[3];
    int r = alb_stats_partial(h, raw, U, V, Cp);
    if (r) return r;
    if (raw[0] > 0) h->maxS = raw[0];            // HTML:611-613
    if (isfinite(raw[1])) h->cpMin = raw[1];
    if (isfinite(raw[2])) h->cpMax = raw[2];
    if (stats) { stats[0] = h->maxS; stats[1] = h->cpMin; stats[2] = h->cpMax; }
    return ALB_OK;
}

int alb_set_stats(alb_handle *h, double maxS, double cpMin, double cpMax) {
    if (!h) return ALB_ERR_INVALID;
    h->maxS = maxS; h->cpMin = cpMin; h->cpMax = cpMax;
    return ALB_OK;
}

int alb_get_stats(const alb_handle *h, double *stats3) {
    if (!h || !stats3) return ALB_ERR_INVALID;
    stats3[0] = h->maxS; stats3[1] = h->cpMin; stats3[2] = h->cpMax;
    return ALB_OK;
}

static int render_common(alb_handle *h, int mode, float *t_out, uint8_t *rgba) {
    NEED(h);
    ARG(mode >= 0 && mode <= 2, "field mode must be 0 (speed), 1 (Cp) or 2 (vorticity)");
    const int lo_ghost = h->y0 > 0 ? 1 : 0, hi_ghost = h->y0 + h->nyl < h->ny_global ? 1 : 0;
    if (mode == ALB_FIELD_VORT && (lo_ghost || hi_ghost) && !h->ghost_macro_valid)
        return h->fail(ALB_ERR_STATE, "vorticity on a slab needs the neighbours' edge rows: alb_get_macro_edges on "
                                      "every slab, then alb_set_macro_ghosts, after the last step");
    int r = ensure_macro(h);
    if (r) return r;
    if ((r = ensure_tmp(h, 0))) return r;
    if ((r = ensure_tmp(h, 1))) return r;
    float *dt = h->d_tmp[0];
    uint8_t *drgba = reinterpret_cast<uint8_t *>(h->d_tmp[1]);
    CK(launch_render(h->mask, h->rho, h->ux, h->uy, h->pitch, h->nx, h->nyl, lo_ghost, hi_ghost, mode, h->u0f,
                     (float)h->maxS, (float)h->cpMin, (float)h->cpMax, 0.06f, dt, rgba ? drgba : nullptr, h->stream));
    const size_t n = (size_t)h->nx * h->nyl;
    if (t_out) CK(cudaMemcpyAsync(t_out, dt, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
    if (rgba) CK(cudaMemcpyAsync(rgba, drgba, 4 * n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return check_wait_error(h);
}

int alb_get_field(alb_handle *h, int mode, float *t_out) {
    if (h && !t_out) return h->fail(ALB_ERR_INVALID, "alb_get_field: output is NULL");
    return render_common(h, mode, t_out, nullptr);
}

int alb_get_rgba(alb_handle *h, int mode, uint8_t *rgba) {
    if (h && !rgba) return h->fail(ALB_ERR_INVALID, "alb_get_rgba: output is NULL");
    return render_common(h, mode, nullptr, rgba);
}

int alb_get_macro_edges(alb_handle *h, float *lo2, float *hi2) {
    NEED(h);
    ARG(lo2 || hi2, "alb_get_macro_edges: both outputs are NULL");
    int r = ensure_macro(h);
    if (r) return r;
    const size_t rowb = sizeof(float) * h->nx;
    const float *srcs[2] = {h->ux, h->uy};
    for (int k = 0; k < 2; k++) {
        if (lo2) CK(cudaMemcpyAsync(lo2 + (size_t)k * h->nx, srcs[k] + (size_t)1 * h->pitch, rowb, cudaMemcpyDeviceToHost, h->stream));
        if (hi2) CK(cudaMemcpyAsync(hi2 + (size_t)k * h->nx, srcs[k] + (size_t)h->nyl * h->pitch, rowb, cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    return check_wait_error(h);
}

int alb_set_macro_ghosts(alb_handle *h, const float *below2, const float *above2) {
    NEED(h);
    int r = ensure_macro(h);       // the ghost rows belong to the macroscopic fields of the CURRENT state
    if (r) return r;
    const size_t rowb = sizeof(float) * h->nx;
    float *dsts[2] = {h->ux, h->uy};
    for (int k = 0; k < 2; k++) {
        if (below2) CK(cudaMemcpyAsync(dsts[k], below2 + (size_t)k * h->nx, rowb, cudaMemcpyHostToDevice, h->stream));
        if (above2) CK(cudaMemcpyAsync(dsts[k] + (size_t)(h->nyl + 1) * h->pitch, above2 + (size_t)k * h->nx, rowb,
                                       cudaMemcpyHostToDevice, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));    // caller memory
    const bool need_lo = h->y0 > 0, need_hi = h->y0 + h->nyl < h->ny_global;
    h->ghost_macro_valid = (!need_lo || below2) && (!need_hi || above2);
    return ALB_OK;
}

int alb_forces_partial(alb_handle *h, double *out4) {
    NEED(h);
    ARG(out4, "alb_forces_partial: output is NULL");
    int r = fused_diag(h);
    if (r < 0) return r;
    if (r == ALB_OK) {
        // HTML:663-668: p = rho/3 per face; here (sum of rho)/3, exact integer sum (differs from the
        // sequential float64 sum by rounding only, ~1e-16 relative)
        const DiagAcc d = merged_diag(h);
        out4[0] = (double)d.fx / ALB_ME_SCALE / 3;
        out4[1] = (double)d.fy / ALB_ME_SCALE / 3;
        out4[2] = (double)d.surf;
        out4[3] = (double)d.rev;
        return check_wait_error(h);
    }
    r = ensure_macro(h);
    if (r) return r;
    CK(launch_forces(h->mask, h->rho, h->ux, h->pitch, h->nx, h->ny_global, h->y0 - 1, h->nyl, h->d_part,
                     DIAG_BLOCKS, h->stream));
    CK(cudaMemcpyAsync(h->h_part, h->d_part, sizeof(double) * 4 * DIAG_BLOCKS, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    double s[4] = {0, 0, 0, 0};
    for (int b = 0; b < DIAG_BLOCKS; b++)
        for (int k = 0; k < 4; k++) s[k] += h->h_part[4 * b + k];
    memcpy(out4, s, sizeof s);
    return check_wait_error(h);
}

int alb_compute_forces(alb_handle *h, double *out10) {
    NEED(h);
    ARG(out10, "alb_compute_forces: output is NULL");
    if (!h->whole()) return h->fail(ALB_ERR_STATE, "alb_compute_forces needs a whole-lattice handle; use alb_forces_partial");
    double s[4];
    int r = alb_forces_partial(h, s);
    if (r) return r;
    const bool any = s[2] > 0;
    double cl_raw = 0, cd_raw = 0;
    if (any) {                                   // HTML:672-679, 699
        const double q = h->qdyn();
        cl_raw = s[1] / q;
        cd_raw = s[0] / q;
        if (!h->ema_valid) {
            h->cl_smooth = cl_raw;
            h->cd_smooth = cd_raw;
            h->ema_valid = true;
        } else {
            h->cl_smooth = h->cl_smooth * 0.9 + cl_raw * 0.1;
            h->cd_smooth = h->cd_smooth * 0.9 + cd_raw * 0.1;
        }
        h->sep_frac = h->sep_frac * 0.85 + (s[3] / s[2]) * 0.15;
    }
    out10[0] = s[0]; out10[1] = s[1]; out10[2] = cl_raw; out10[3] = cd_raw;
    out10[4] = h->ema_valid ? h->cl_smooth : NAN;
    out10[5] = h->ema_valid ? h->cd_smooth : NAN;
    out10[6] = h->sep_frac; out10[7] = s[2]; out10[8] = s[3]; out10[9] = any ? 1.0 : 0.0;
    return ALB_OK;
}

int alb_reset_force_emas(alb_handle *h) {
    if (!h) return ALB_ERR_INVALID;
    h->ema_valid = false;
    h->cl_smooth = h->cd_smooth = h->sep_frac = 0;
    return ALB_OK;
}

int alb_get_me_history(alb_handle *h, int n, long long *fxfy) {
    NEED(h);
    ARG(fxfy && n >= 0 && n <= ALB_ME_HISTORY, "alb_get_me_history: need 0 <= n <= ALB_ME_HISTORY");
    if (n > h->steps) return h->fail(ALB_ERR_STATE, "alb_get_me_history: fewer steps taken than requested");
    CK(cudaStreamSynchronize(h->stream));
    struct { long long acc[2][2]; long long count; int pending; int pad; } head;
    CK(cudaMemcpy(&head, h->me, sizeof head, cudaMemcpyDeviceToHost));
    for (int k = 0; k < n;) {
        const long long step = h->sync_steps - n + k;
        if (step < head.count) {
            // committed steps sit in the ring: one copy per contiguous run (the ring wraps at most once)
            const long long slot = step % ME_RING;
            long long run = head.count - step;
            if (run > n - k) run = n - k;
            if (run > ME_RING - slot) run = ME_RING - slot;
            CK(cudaMemcpy(fxfy + 2 * k, &h->me->ring[slot][0], sizeof(long long) * 2 * (size_t)run, cudaMemcpyDeviceToHost));
            k += (int)run;
        } else {
            // the last step: still in the accumulator of its parity (= the buffer it read from)
            const int parity = h->parity ^ 1;
            fxfy[2 * k] = head.acc[parity][0];
            fxfy[2 * k + 1] = head.acc[parity][1];
            k++;
        }
    }
    return check_wait_error(h);
}

int alb_get_me_forces(alb_handle *h, double *out4) {
    NEED(h);
    ARG(out4, "alb_get_me_forces: output is NULL");
    long long v[2];
    int r = alb_get_me_history(h, 1, v);
    if (r) return r;
    const double fx = (double)v[0] / ALB_ME_SCALE, fy = (double)v[1] / ALB_ME_SCALE;
    const double q = h->qdyn();
    out4[0] = fx; out4[1] = fy; out4[2] = fy / q; out4[3] = fx / q;
    return ALB_OK;
}

int alb_clamp_hits(alb_handle *h, long long *hits) {
    NEED(h);
    ARG(hits, "alb_clamp_hits: output is NULL");
    unsigned long long v = 0;
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(&v, h->clamp_hits, sizeof v, cudaMemcpyDeviceToHost));
    *hits = (long long)v;
    return ALB_OK;
}

int alb_reynolds(const alb_handle *h, double *re) {
    if (!h || !re) return ALB_ERR_INVALID;
    const double nu_l = (h->tau - 0.5) / 3;       // HTML:79
    *re = h->u0 * h->chord_l() / nu_l;            // HTML:865
    return ALB_OK;
}

int alb_stall_state(const alb_handle *h, int *state, int *sep_pct) {
    if (!h) return ALB_ERR_INVALID;
    const double x100 = h->sep_frac * 100, fl = floor(x100);
    const int pct = (int)(fl + (x100 - fl >= 0.5 ? 1.0 : 0.0));   // Math.round (ties up; x - floor(x) is exact), HTML:869
    if (sep_pct) *sep_pct = pct;
    if (state) *state = pct < 5 ? 0 : (pct < 25 ? 1 : 2);  // HTML:872-884
    return ALB_OK;
}

/* ---- tracer particles -------------------------------------------------------- */

static int particles_reserve(alb_handle *h, int n) {
    if (n <= h->parts_cap) return ALB_OK;
    ParticleState *np = nullptr;
    unsigned *nc = nullptr;
    CK(cudaMalloc(&np, sizeof(ParticleState) * n));
    CK(cudaMalloc(&nc, sizeof(unsigned) * n));
    if (h->nparts > 0) {
        CK(cudaMemcpyAsync(np, h->parts, sizeof(ParticleState) * h->nparts, cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaMemcpyAsync(nc, h->part_ctr, sizeof(unsigned) * h->nparts, cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    cudaFree(h->parts);
    cudaFree(h->part_ctr);
    h->parts = np;
    h->part_ctr = nc;
    h->parts_cap = n;
    return ALB_OK;
}

int alb_particles_init(alb_handle *h, int n, unsigned long long seed) {
    NEED(h);
    ARG(n >= 0 && n <= ALB_MAX_PARTICLES, "alb_particles_init: need 0 <= n <= ALB_MAX_PARTICLES");
    if (!h->whole()) return h->fail(ALB_ERR_STATE, "particles need a whole-lattice handle");
    h->nparts = 0;
    int r = particles_reserve(h, n);
    if (r) return r;
    h->part_seed = seed;
    CK(launch_particles_init(h->parts, h->part_ctr, 0, n, n, seed, 0, h->stream));
    h->nparts = n;
    return ALB_OK;
}

int alb_particles_resize(alb_handle *h, int n) {
    NEED(h);
    ARG(n >= 0 && n <= ALB_MAX_PARTICLES, "alb_particles_resize: need 0 <= n <= ALB_MAX_PARTICLES");
    if (!h->whole()) return h->fail(ALB_ERR_STATE, "particles need a whole-lattice handle");
    if (n > h->nparts) {
        int r = particles_reserve(h, n);
        if (r) return r;
        CK(launch_particles_init(h->parts, h->part_ctr, h->nparts, n, n, h->part_seed, 1, h->stream));
    }
    h->nparts = n;
    return ALB_OK;
}

int alb_particles_step(alb_handle *h, double dt_ms) {
    NEED(h);
    NO_PENDING_FRAMES(h);
    ARG(isfinite(dt_ms), "alb_particles_step: dt must be finite");
    if (!h->whole()) return h->fail(ALB_ERR_STATE, "particles need a whole-lattice handle");
    int r = ensure_macro(h);
    if (r) return r;
    CK(launch_particles_step(h->parts, h->part_ctr, h->nparts, h->part_seed, dt_ms, h->mask, h->ux, h->uy, h->pitch,
                             h->nx, h->nyl, h->u0, h->stream));
    return ALB_OK;
}

int alb_particles_get(alb_handle *h, double *out8, int *n) {
    NEED(h);
    if (n) *n = h->nparts;
    if (out8 && h->nparts > 0) {
        static_assert(sizeof(ParticleState) == 8 * sizeof(double), "ParticleState must be 8 doubles wide");
        std::vector<ParticleState> tmp;
        try {
            tmp.resize(h->nparts);
        } catch (const std::bad_alloc &) {
            return h->fail(ALB_ERR_NOMEM, "alb_particles_get: host allocation failed");
        }
        CK(cudaMemcpyAsync(tmp.data(), h->parts, sizeof(ParticleState) * h->nparts, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        for (int i = 0; i < h->nparts; i++) {
            const ParticleState &p = tmp[i];
            double *o = out8 + 8 * (size_t)i;
            o[0] = p.x; o[1] = p.y; o[2] = p.life; o[3] = p.lane; o[4] = p.x0; o[5] = p.y0; o[6] = p.speed;
            o[7] = (double)p.respawned;
        }
    }
    return check_wait_error(h);
}

/* ---- multi-GPU y-slabs ------------------------------------------------------ */

int alb_connect_local(alb_handle *h, alb_handle *lo, alb_handle *hi) {
    NEED(h);
    alb_handle *peers[2] = {lo, hi};
    for (alb_handle *p : peers) {
        if (!p) continue;
        ARG(p->nx == h->nx && p->ny_global == h->ny_global, "alb_connect_local: neighbour has a different lattice");
        if (p->device != h->device) {
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, h->device, p->device));
            if (!can) return h->fail(ALB_ERR_CUDA, "alb_connect_local: no peer access between the two devices");
            cudaError_t e = cudaDeviceEnablePeerAccess(p->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return h->fail(ALB_ERR_CUDA, "cudaDeviceEnablePeerAccess", e);
            cudaGetLastError();
        }
    }
    if (lo) ARG(lo->y0 + lo->nyl == h->y0, "alb_connect_local: lower neighbour does not end where this slab starts");
    if (hi) ARG(h->y0 + h->nyl == hi->y0, "alb_connect_local: upper neighbour does not start where this slab ends");
    h->lo = Peer();
    h->hi = Peer();
    if (lo) { h->lo.base = lo->f[0]; h->lo.flags = lo->flags; h->lo.plane = lo->plane; h->lo.nyl = lo->nyl; }
    if (hi) { h->hi.base = hi->f[0]; h->hi.flags = hi->flags; h->hi.plane = hi->plane; h->hi.nyl = hi->nyl; }
    return ALB_OK;
}

int alb_create_multi(int nx, int ny, const int *devices, int ndev, alb_handle **out_slabs) {
    if (!out_slabs || !devices || ndev < 1 || ndev > ny) {
        g_create_error = "alb_create_multi: need 1 <= ndev <= ny, devices and out_slabs non-NULL";
        return ALB_ERR_INVALID;
    }
    for (int k = 0; k < ndev; k++) out_slabs[k] = nullptr;
    const int base = ny / ndev, rem = ny % ndev;
    int y0 = 0, rc = ALB_OK;
    for (int k = 0; k < ndev && rc == ALB_OK; k++) {
        const int n = base + (k < rem ? 1 : 0);
        rc = alb_create_slab(nx, ny, y0, n, devices[k], &out_slabs[k]);
        y0 += n;
    }
    for (int k = 0; k < ndev && rc == ALB_OK; k++) {
        rc = alb_connect_local(out_slabs[k], k > 0 ? out_slabs[k - 1] : nullptr, k + 1 < ndev ? out_slabs[k + 1] : nullptr);
        if (rc != ALB_OK) g_create_error = out_slabs[k]->err;
    }
    if (rc != ALB_OK) {
        for (int k = 0; k < ndev; k++) {
            if (out_slabs[k]) free_handle(out_slabs[k]);
            out_slabs[k] = nullptr;
        }
    }
    return rc;
}

int alb_step_multi(alb_handle **slabs, int nslabs, int nsteps) {
    if (!slabs || nslabs < 1 || nsteps < 0) return ALB_ERR_INVALID;
    // Chunks keep every slab's stream fed while bounding how far one slab is enqueued ahead.  Slabs
    // that share a GPU are stepped one step at a time: then every flag a wait kernel needs was
    // signalled by work enqueued EARLIER, so hardware queues shared between streams cannot deadlock.
    int chunk = 16;
    for (int a = 0; a < nslabs; a++)
        for (int b = a + 1; b < nslabs; b++)
            if (slabs[a] && slabs[b] && slabs[a]->device == slabs[b]->device) chunk = 1;
    for (int done = 0; done < nsteps; done += chunk) {
        const int n = nsteps - done < chunk ? nsteps - done : chunk;
        // all closing signals of the previous round first, so that no wait kernel below depends on
        // work that is enqueued after it
        for (int k = 0; k < nslabs; k++) {
            if (!slabs[k]) return ALB_ERR_INVALID;
            if (cudaSetDevice(slabs[k]->device) != cudaSuccess) return slabs[k]->fail(ALB_ERR_CUDA, "cudaSetDevice");
            flush_pending_signal(slabs[k]);
        }
        for (int k = 0; k < nslabs; k++) {
            int rc = alb_step(slabs[k], n);
            if (rc != ALB_OK) return rc;
        }
    }
    return ALB_OK;
}

int alb_destroy_multi(alb_handle **slabs, int nslabs) {
    if (!slabs || nslabs < 0) return ALB_ERR_INVALID;
    for (int k = 0; k < nslabs; k++)
        if (slabs[k] && cudaSetDevice(slabs[k]->device) == cudaSuccess && slabs[k]->stream) {
            cudaStreamSynchronize(slabs[k]->stream);
            if (slabs[k]->aux) cudaStreamSynchronize(slabs[k]->aux);
        }
    for (int k = 0; k < nslabs; k++) {
        if (slabs[k]) free_handle(slabs[k]);
        slabs[k] = nullptr;
    }
    return ALB_OK;
}

int alb_ipc_export(alb_handle *h, void *blob) {
    NEED(h);
    ARG(blob, "alb_ipc_export: blob is NULL");
    IpcBlob b;
    memset(&b, 0, sizeof b);
    CK(cudaIpcGetMemHandle(&b.mem, h->block));
    b.nx = h->nx; b.nyl = h->nyl; b.pitch = h->pitch; b.device = h->device;
    b.plane = h->plane;
    b.flags_offset_bytes = sizeof(float) * 27 * h->plane;
    b.magic = 0x414c4231;
    memset(blob, 0, ALB_IPC_BYTES);
    memcpy(blob, &b, sizeof b);
    return ALB_OK;
}

static int open_peer(alb_handle *h, const void *blob, Peer *out) {
    IpcBlob b;
    memcpy(&b, blob, sizeof b);
    ARG(b.magic == 0x414c4231, "alb_ipc_connect: not an alb_ipc_export() blob");
    ARG(b.nx == h->nx && b.pitch == h->pitch, "alb_ipc_connect: neighbour has a different lattice width");
    void *base = nullptr;
    CK(cudaIpcOpenMemHandle(&base, b.mem, cudaIpcMemLazyEnablePeerAccess));
    out->ipc_base = base;
    out->base = reinterpret_cast<float *>(base);
    out->flags = reinterpret_cast<int *>(reinterpret_cast<char *>(base) + b.flags_offset_bytes);
    out->plane = b.plane;
    out->nyl = b.nyl;
    return ALB_OK;
}

int alb_ipc_connect(alb_handle *h, const void *lo_blob, const void *hi_blob) {
    NEED(h);
    if (h->lo.ipc_base) { cudaIpcCloseMemHandle(h->lo.ipc_base); }
    if (h->hi.ipc_base) { cudaIpcCloseMemHandle(h->hi.ipc_base); }
    h->lo = Peer();
    h->hi = Peer();
    int r;
    if (lo_blob && (r = open_peer(h, lo_blob, &h->lo))) return r;
    if (hi_blob && (r = open_peer(h, hi_blob, &h->hi))) return r;
    return ALB_OK;
}

int alb_halo_prime(alb_handle *h) {
    NEED(h);
    // push my edge rows of the CURRENT state into the neighbours' ghost rows, then publish my step count
    const size_t rowb = sizeof(float) * h->pitch;
    const int lo_pops[3] = {4, 7, 8}, hi_pops[3] = {2, 5, 6};
    if (h->lo.base)
        for (int k = 0; k < 3; k++) {
            const int i = lo_pops[k];
            CK(cudaMemcpyAsync(h->lo.base + (size_t)h->cur * 9 * h->lo.plane + i * h->lo.plane +
                                   (size_t)(h->lo.nyl + 1) * h->pitch,
                               h->f[h->cur] + i * h->plane + (size_t)1 * h->pitch, rowb, cudaMemcpyDefault,
                               h->stream));
        }
    if (h->hi.base)
        for (int k = 0; k < 3; k++) {
            const int i = hi_pops[k];
            CK(cudaMemcpyAsync(h->hi.base + (size_t)h->cur * 9 * h->hi.plane + i * h->hi.plane,
                               h->f[h->cur] + i * h->plane + (size_t)h->nyl * h->pitch, rowb, cudaMemcpyDefault,
                               h->stream));
        }
    signal_kernel<<<1, 1, 0, h->stream>>>(h->lo.flags ? h->lo.flags + 1 : nullptr,
                                          h->hi.flags ? h->hi.flags + 0 : nullptr, (int)h->sync_steps);
    CK(cudaGetLastError());
    h->pending_signal = -1;     // the neighbours now know this slab's step count
    return ALB_OK;
}

int alb_halo_ptrs(alb_handle *h, void **send_lo3, void **send_hi3, void **recv_lo3, void **recv_hi3) {
    NEED(h);
    float *cur = h->f[h->cur];
    const int lo_pops[3] = {4, 7, 8}, hi_pops[3] = {2, 5, 6};
    for (int k = 0; k < 3; k++) {
        if (send_lo3) send_lo3[k] = cur + lo_pops[k] * h->plane + (size_t)1 * h->pitch;
        if (send_hi3) send_hi3[k] = cur + hi_pops[k] * h->plane + (size_t)h->nyl * h->pitch;
        if (recv_lo3) recv_lo3[k] = cur + hi_pops[k] * h->plane;   // ghost row 0 receives the lower slab's f2,f5,f6
        if (recv_hi3) recv_hi3[k] = cur + lo_pops[k] * h->plane + (size_t)(h->nyl + 1) * h->pitch;
    }
    return ALB_OK;
}

int alb_set_external_halo(alb_handle *h, int on) {
    if (!h) return ALB_ERR_INVALID;
    h->external_halo = on != 0;
    return ALB_OK;
}

int alb_selftest_division(alb_handle *h, unsigned long long seed, long long pairs, unsigned long long *out3) {
    NEED(h);
    ARG(out3 && pairs > 0, "alb_selftest_division: need pairs > 0 and an output array");
    unsigned long long *d = nullptr;
    CK(cudaMalloc(&d, 3 * sizeof(unsigned long long)));
    const int nblocks = 148 * 8, iters = (int)((pairs + (long long)nblocks * 256 - 1) / ((long long)nblocks * 256));
    cudaError_t e = cudaMemsetAsync(d, 0, 3 * sizeof(unsigned long long), h->stream);
    if (e == cudaSuccess) e = launch_div_selftest(seed, nblocks, iters, d, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out3, d, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d);
    if (e != cudaSuccess) return h->fail(ALB_ERR_CUDA, "alb_selftest_division", e);
    return ALB_OK;
}

int alb_debug_step2_plan(int nx, int ny_local, int nsm, int *out5) {
    if (!out5 || nx < 1 || ny_local < 1 || nsm < 1) return ALB_ERR_INVALID;
    Step2Params q;
    memset(&q, 0, sizeof q);
    q.pitch = (nx + TASK_CELLS - 1) / TASK_CELLS * TASK_CELLS;
    q.nyl = ny_local;
    q.edges = march_edges_enabled(nx, q.pitch) ? 1 : 0;
    q.nx = nx;
    march_plan(q, nsm);
    out5[0] = q.nseg;
    out5[1] = q.wo;
    out5[2] = q.hs;
    out5[3] = q.nunits;
    out5[4] = march_warps_per_cta();
    return ALB_OK;
}

int alb_launch_count(const alb_handle *h, long long *launches) {
    if (!h || !launches) return ALB_ERR_INVALID;
    *launches = h->launches;
    return ALB_OK;
}

int alb_get_double_steps(const alb_handle *h, int *mode, int *active) {
    if (!h) return ALB_ERR_INVALID;
    if (mode) *mode = h->double_mode;
    if (active) *active = double_steps_enabled(h) ? 1 : 0;
    return ALB_OK;
}

int alb_get_div_mode(const alb_handle *h, int *mode) {
    if (!h || !mode) return ALB_ERR_INVALID;
    *mode = h->div_mode;
    return ALB_OK;
}

int alb_set_div_mode(alb_handle *h, int mode) {
    NEED(h);
    NO_PENDING_FRAMES(h);
    ARG(mode == -1 || mode == 1, "alb_set_div_mode: mode must be -1 (verified shortcut when possible) or 1 (IEEE division)");
    const int m = mode == 1 ? 1 : div_mode_for(h, h->tauf);
    if (m != h->div_mode) drop_graph(h);
    h->div_mode = m;
    h->div_forced = mode == 1;
    return ALB_OK;
}

int alb_set_double_steps(alb_handle *h, int mode) {
    if (!h) return ALB_ERR_INVALID;
    NO_PENDING_FRAMES(h);
    ARG(mode >= -1 && mode <= 1, "alb_set_double_steps: mode must be -1, 0 or 1");
    if (mode != h->double_mode) drop_graph(h);
    h->double_mode = mode;
    return ALB_OK;
}

}  // extern "C"

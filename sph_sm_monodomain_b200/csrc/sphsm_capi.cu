// sphsm_capi.cu — libsphsm_b200.so: handle, step orchestration and the extern "C" entry points declared in
// include/sphsm_b200.h.  All device work is in the hand-written sm_100a kernels of sphsm_{sort,sm,pass}.cuh;
// there is no CPU fallback anywhere in this file.
#include "../../include/sphsm_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <string>
#include <unordered_set>
#include <vector>

#include "sphsm_comm.cuh"
#include "sphsm_pass.cuh"
#include "sphsm_pass2.cuh"
#include "sphsm_pass3.cuh"
#include "sphsm_pass4.cuh"
#include "sphsm_pass4w.cuh"
#include "sphsm_pass5.cuh"
#include "sphsm_sm.cuh"
#include "sphsm_sort.cuh"
#include "sphsm_types.cuh"

using namespace sphsm;

// byte offsets inside the reference's Particle (Particle.h:10-29)
enum : int {
    OFF_POS = 0, OFF_VEL = 12, OFF_PVEL = 24, OFF_IVEL = 36, OFF_CVEL = 48, OFF_ACC = 60, OFF_MASS = 72, OFF_ORIG = 76,
    OFF_GOAL = 88, OFF_FIXED = 100, OFF_DENS = 104, OFF_PRES = 108, OFF_VM = 112, OFF_IVM = 116, OFF_IION = 120,
    OFF_STIM = 124, OFF_W = 128
};

static std::string g_create_error;
static int (*g_nccl_destroy)(void *) = nullptr;  // set once libnccl is loaded (sphsm_destroy runs before its definition)

enum KernelGroup { KG_HASH = 0, KG_SORT, KG_GRID, KG_MOMENTS, KG_GOAL, KG_PASS_A, KG_PASS_B, KG_OTHER };
static const char *kGroupNames[SPHSM_NUM_KERNEL_GROUPS] = {"hash", "radix_sort", "cell_bounds+reorder", "sm_moments+solve",
                                                           "goal+corrected_vel", "pass_a(density+xsph)",
                                                           "pass_b(cell+force+laplacian+integrate)", "other"};

struct sphsm_handle {
    sphsm_params prm;
    DevParams dp;
    DevParams *d_dp = nullptr;  // global-memory copy of dp for the fast passes (sphsm_pass2.cuh)
    DevParams dp_uploaded{};
    int n = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t launch_stream = nullptr;  // where LAUNCH enqueues: `stream`, except while the side chain below is built
    cudaStream_t side_stream = nullptr;    // slab step: moment sums + allreduce + solve run here, beside the sort
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_meta = nullptr, ev_bnd = nullptr, ev_int = nullptr;
    bool moments_forked = false;
    bool allreduce_pending = false;        // slab step: the forked sums still need their allreduce (issued after exchange 1 on a shared communicator)
    cudaEvent_t pev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // SPHSM_HOST_PROF: device-side brackets
    double pacc[4] = {0, 0, 0, 0};
    double meta_wait_us = 0.0;             // SPHSM_HOST_PROF: host time spent waiting for the plane boundaries
    // NCCL mode: a rank-local step error (halo overflow, a particle crossing two planes) must not leave the peers waiting in
    // a collective.  It is recorded, rides as one extra element on the NEXT step's moment allreduce, and every rank returns the
    // error at the end of that step (so all ranks stop after the same step, at most one step late).
    int local_error = 0;
    std::string local_error_msg;
    double *h_flag = nullptr;              // pinned: the summed error flag of the latest moment allreduce
    cudaEvent_t ev_flag = nullptr;
    bool flag_pending = false, peer_error = false, failed = false;
    bool reordered = false;                // slab step: the gather was queued before the plane boundaries reached the host
    bool split = false;                    // slab step: exchange 2 in flight on the side stream beside the interior planes
    Arrays cur{}, alt{};
    uint32_t *keys[2] = {nullptr, nullptr}, *vals[2] = {nullptr, nullptr};
    uint32_t *ghist = nullptr, *tile_state = nullptr, *tile_counter = nullptr;
    int sorted_buf = 0;  // which keys[] / vals[] hold the sorted result
    uint32_t *cell_count = nullptr, *tile_sums = nullptr;  // counting sort: per-cell counts (kept zero between steps), scan scratch
    bool bounds_ready = false;                             // grid_sort already produced the cell_start table
    // CUDA graphs for small single-GPU steps (launch-latency bound): see graph_step
    bool dry_run = false;  // replaying a captured step: the host-side state transitions run, launches and stream calls do not
    struct StepGraph {
        std::string sig;
        cudaGraphExec_t exec;
    };
    std::vector<StepGraph> graphs;
    std::vector<std::string> seen_sigs;
    std::unordered_set<long long> host_cells;  // cells occupied by the positions the host handed in (small sets only): see warp_path
    int *big_cells = nullptr, *big_count = nullptr;  // counting sort: worklist of cells too full for the one-thread in-cell sort
    bool counts_ready = false;  // pass B already filed keys / ranks / per-cell counts of the CURRENT positions (single-GPU fast step)
    int *cell_start = nullptr, *slot_of = nullptr;
    SmState *sm = nullptr;
    double *partial = nullptr, *totals = nullptr;
    float *scratch = nullptr;
    uint8_t *d_aos = nullptr;
    size_t aos_cap_bytes = 0;
    // asynchronous I/O (sphsm_*_async): copy streams beside the compute stream, their own device staging
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t ev_in_ready = nullptr, ev_in_free = nullptr, ev_out_ready = nullptr, ev_out_done = nullptr;
    float *io_in_f = nullptr, *io_out_f = nullptr;
    uint8_t *io_in_b = nullptr;
    int *io_out_i = nullptr;
    size_t io_in_cap = 0, io_out_cap = 0;
    float *d_tmp = nullptr;  // staging for position lists (stim_mesh / stim_cube / init_fluid)
    size_t tmp_cap = 0;
    int *d_itmp = nullptr;
    size_t itmp_cap = 0;
    bool grid_valid = false, rest_dirty = true, inter_live = true, slot_of_valid = false;
    int sort_passes = 1, max_tiles = 1, red_blocks = 1;
    long long launches = 0;
    int total_steps = 0;
    bool stage_timing = false;
    double stage_time[7] = {0, 0, 0, 0, 0, 0, 0};
    cudaEvent_t ev[SPHSM_NUM_KERNEL_GROUPS + 2] = {};
    cudaEvent_t ev_step0 = nullptr, ev_step1 = nullptr;
    float last_step_ms = 0.f;
    bool profiling = false;
    float group_ms[SPHSM_NUM_KERNEL_GROUPS] = {};
    int group_launches[SPHSM_NUM_KERNEL_GROUPS] = {};
    std::string err;
    // ---- multi-GPU slab layer (sphsm_comm.cuh) ----
    int comm_mode = 0;        // 0 none, 1 NCCL (one process per GPU), 2 local group (virtual ranks on one device, for tests)
    int nranks = 1, rank = 0;
    void *nccl_comm = nullptr;
    void *nccl_comm_red = nullptr;  // communicator of the moment allreduce: nccl_comm, or a split of it (SPHSM_SPLIT_COMM)
    bool slab_applied = false;
    int send_cap = 0;         // particles per exchange-1 message
    int alloc_n = 0;          // slots allocated per array (capacity + room for two halo messages in slab mode)
    uint8_t *msg_send[2] = {nullptr, nullptr}, *msg_recv[2] = {nullptr, nullptr};  // [0] left neighbour, [1] right neighbour
    int *d_err = nullptr, *d_meta = nullptr, *h_meta = nullptr;
    int b2 = 0, b3 = 0;       // start of the 2nd / of the last owned plane (exchange-2 ranges)
    int n_global = 0;         // particles uploaded before sphsm_comm_set_slab filtered them (ids are global)
    int mom_n = 0;            // slab step: extent of the PRE-reorder arrays (old slots + both message regions) the REST-state sums scan
    int mom_begin = 0, mom_end = 0;  // slab step: the slots this rank integrated last step; its share of the per-step moment sums
    struct GroupTimer *gt = nullptr;
};

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char b_[512];                                                                          \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            if (h) h->err = b_; else g_create_error = b_;                                          \
            return SPHSM_ERR_CUDA;                                                                 \
        }                                                                                          \
    } while (0)

// SPHSM_SYNC_DEBUG=1 in the environment synchronises after every launch and names the kernel that faulted
static const bool g_sync_debug = getenv("SPHSM_SYNC_DEBUG") != nullptr;
// SPHSM_PASS selects the generation of the fast-path neighbour passes: 4 = sphsm_pass4.cuh (production), 2 = sphsm_pass2.cuh
// (its predecessor), 3 = the experimental warp-staged passes (sphsm_pass3.cuh; correct, but measured 2x slower at 8M:
// profiles/r01_v5_staged_*.json)
static const int g_pass_gen = getenv("SPHSM_PASS") ? atoi(getenv("SPHSM_PASS")) : 4;
#define LAUNCH(kern, grid, block, ...)                                                                  \
    do {                                                                                                \
        if (!h->dry_run) kern<<<(grid), (block), 0, h->launch_stream>>>(__VA_ARGS__);                   \
        h->launches++;                                                                                  \
        if (g_sync_debug) {                                                                             \
            cudaError_t e_ = cudaStreamSynchronize(h->launch_stream);                                   \
            if (e_ != cudaSuccess) {                                                                    \
                h->err = std::string("kernel ") + #kern + " failed: " + cudaGetErrorString(e_);         \
                fprintf(stderr, "[sphsm] %s\n", h->err.c_str());                                        \
                return SPHSM_ERR_CUDA;                                                                  \
            }                                                                                           \
        }                                                                                               \
    } while (0)

static int fail(sphsm_handle *h, int code, const char *msg) {
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}
static inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }
static inline void set_n(sphsm_handle *h, int n) {  // single-GPU meaning: every slot is computed
    h->n = n;
    h->dp.n = n;
    h->dp.own_begin = 0;
    h->dp.own_end = n;
}

// ---------------------------------------------------------------------------------------------------
// defaults: the reference ctor, cpp:13-69, with its float/double promotions
extern "C" int sphsm_abi_version(void) { return SPHSM_ABI_VERSION; }

extern "C" int sphsm_default_params(sphsm_params *p) {
    if (!p) return SPHSM_ERR_INVALID;
    memset(p, 0, sizeof(*p));
    p->struct_size = (uint32_t)sizeof(*p);
    p->device = 0;
    p->capacity = 50000;
    p->world[0] = p->world[1] = p->world[2] = 1.5f;
    p->kernel_h = 0.04f;
    p->gravity[0] = 0.0f; p->gravity[1] = -9.8f; p->gravity[2] = 0.0f;
    p->K = 0.5f;
    p->stand_density = 1112.0f;
    const float max_vel2 = 3.0f * 3.0f + 3.0f * 3.0f + 3.0f * 3.0f;                          // max_vel.magnitudeSquared()
    p->time_delta = (float)(0.4 * (double)p->kernel_h / (double)sqrtf(max_vel2));            // cpp:47
    p->wall_hit = -1.0f;
    p->mu = 100.0f;
    p->velocity_mixing = 1.0f;
    const float pi = 3.1415926535897932f;                                                    // m3Pi, m3Real.h:9
    p->poly6_constant = (float)((double)315.0f / ((double)(64.0f * pi) * pow((double)p->kernel_h, 9.0)));  // cpp:54
    p->spiky_constant = (float)((double)45.0f / ((double)pi * pow((double)p->kernel_h, 6.0)));             // cpp:55
    p->bspline_constant = 1.0f / (pi * p->kernel_h * p->kernel_h * p->kernel_h);                           // cpp:57
    p->alpha = 0.3f; p->beta = 0.4f;
    p->quadratic_match = 0; p->volume_conservation = 1; p->allow_flip = 0;
    p->Cm = 1.f; p->Beta = 50;
    {
        float sigma_i = 0.893, sigma_e = 0.67;  // cpp:15 (double literals narrowed to float)
        p->sigma = sigma_i * sigma_e / (sigma_i + sigma_e);
    }
    p->stim_strength = 300.0f;
    p->FH_Vt = -75.0; p->FH_Vp = 15.0; p->FH_Vr = -85.0;
    p->C1 = 0.175; p->C2 = 0.03; p->C3 = 0.011; p->C4 = 0.55;
    p->voltage_constant = 1; p->max_pressure = 15000; p->max_voltage = 200;
    p->particle_mass = 0.2f;
    p->diagnostics = 1;
    p->strict = 0;
    p->slab_axis = -1;
    return SPHSM_OK;
}

// largest float x (searching around `guess`) with pred(x) true, pred monotone (true below, false above)
template <class Pred>
static float max_float_where(Pred pred, float guess) {
    float x = guess;
    while (!pred(x)) x = nextafterf(x, 0.0f);
    while (pred(nextafterf(x, INFINITY))) x = nextafterf(x, INFINITY);
    return x;
}

static void derive_dev_params(sphsm_handle *h) {
    const sphsm_params &q = h->prm;
    DevParams &d = h->dp;
    memset(&d, 0, sizeof d);
    d.n = h->n;
    d.cell_size = q.kernel_h;  // Cell_Size == kernel, cpp:17,31
    d.h = q.kernel_h;
    d.h2 = q.kernel_h * q.kernel_h;
    for (int a = 0; a < 3; a++) {
        d.world[a] = q.world[a];
        d.gravity[a] = q.gravity[a];
        d.g[a] = (int)ceilf(q.world[a] / d.cell_size);  // cpp:32-35
    }
    if (q.slab_axis >= 0 && q.slab_axis <= 2) {
        d.perm[2] = q.slab_axis;
        d.perm[0] = (q.slab_axis + 1) % 3;
        d.perm[1] = (q.slab_axis + 2) % 3;
        if (d.perm[0] > d.perm[1]) std::swap(d.perm[0], d.perm[1]);
    } else {
        d.perm[0] = 0; d.perm[1] = 1; d.perm[2] = 2;
    }
    d.ga = d.g[d.perm[0]] + 2; d.gb = d.g[d.perm[1]] + 2; d.gc = d.g[d.perm[2]];  // ga, gb: one empty border cell per side
    d.c_off = 0; d.gcl = d.gc; d.slab_lo = 0; d.slab_hi = d.gc;
    d.slab_on = 0; d.own_begin = 0; d.own_end = h->n; d.hole_begin = 0; d.hole_len = 0;
    d.num_cells = d.ga * d.gb * d.gcl;
    d.K = q.K; d.rho0 = q.stand_density; d.dt = q.time_delta; d.inv_dt = 1.0f / q.time_delta;  // cpp:661
    d.wall_hit = q.wall_hit; d.mu = q.mu; d.mix = q.velocity_mixing;
    d.c_poly6 = q.poly6_constant; d.c_spiky = q.spiky_constant; d.c_bspline = q.bspline_constant;
    d.alpha = q.alpha; d.beta = q.beta;
    d.quadratic = q.quadratic_match; d.volume = q.volume_conservation; d.allow_flip = q.allow_flip;
    d.Cm = q.Cm; d.Beta = q.Beta; d.sigma = q.sigma;
    d.diff_coef = q.sigma / (q.Beta * q.Cm);  // cpp:571
    d.Vr = q.FH_Vr;
    d.fh_denom = q.FH_Vp - q.FH_Vr;                 // cpp:579
    d.fh_asd = (q.FH_Vt - q.FH_Vr) / d.fh_denom;    // cpp:580
    d.C1 = q.C1; d.C2 = q.C2; d.C3 = q.C3; d.C4 = q.C4;
    d.voltage_constant = q.voltage_constant; d.max_pressure = q.max_pressure; d.max_voltage = q.max_voltage;
    const float hh = d.h;
    d.r2_spiky = max_float_where([hh](float x) { return sqrtf(x) <= hh; }, hh * hh);
    d.r2_q1 = max_float_where([hh](float x) { return (sqrtf(x) / hh) < 1.0f; }, hh * hh);
    d.r2_q2 = max_float_where([hh](float x) { return (sqrtf(x) / hh) < 2.0f; }, 4.0f * hh * hh);
    d.bs_a1 = (float)(4.5 * (double)d.c_bspline / (double)hh);
    d.bs_b1 = (float)(-3.0 * (double)d.c_bspline);
    d.bs_a2 = (float)(-1.5 * (double)d.c_bspline / (double)hh);
    d.bs_b2 = (float)(3.0 * (double)d.c_bspline);
    d.poly6_self = (float)((double)d.c_poly6 * pow((double)(d.h2 - 0.0f), 3.0));  // Poly6(0.0f), cpp:151,483
}

// ---------------------------------------------------------------------------------------------------
static int alloc_arrays(sphsm_handle *h, Arrays &a, int cap, bool with_cold) {
    size_t n4 = ((size_t)cap + 8) * sizeof(float4);  // tail: the pair loops read slot j+1 (masked) up to j+1 == n
    CU(cudaMalloc(&a.P, n4)); CU(cudaMalloc(&a.VEL, n4)); CU(cudaMalloc(&a.O, n4)); CU(cudaMalloc(&a.E, n4));
    CU(cudaMalloc(&a.ID, (size_t)cap * sizeof(int)));
    CU(cudaMalloc(&a.C, n4)); CU(cudaMalloc(&a.V, n4)); CU(cudaMalloc(&a.S, (size_t)cap * sizeof(float2)));
    CU(cudaMalloc(&a.ACC, n4)); CU(cudaMalloc(&a.GOAL, n4)); CU(cudaMalloc(&a.PV, n4)); CU(cudaMalloc(&a.PB, n4));
    CU(cudaMemset(a.PB, 0, n4));
    CU(cudaMalloc(&a.VN, ((size_t)cap + 8) * sizeof(float)));
    CU(cudaMemset(a.VN, 0, ((size_t)cap + 8) * sizeof(float)));
    CU(cudaMemset(a.C, 0, n4)); CU(cudaMemset(a.V, 0, n4)); CU(cudaMemset(a.S, 0, (size_t)cap * sizeof(float2)));
    CU(cudaMemset(a.ACC, 0, n4)); CU(cudaMemset(a.GOAL, 0, n4)); CU(cudaMemset(a.PV, 0, n4));
    if (with_cold) {
        CU(cudaMalloc(&a.COLD_GOAL, n4)); CU(cudaMalloc(&a.COLD_PV, n4));
        CU(cudaMemset(a.COLD_GOAL, 0, n4)); CU(cudaMemset(a.COLD_PV, 0, n4));
    }
    return SPHSM_OK;
}
static void free_arrays(Arrays &a, bool with_cold) {
    cudaFree(a.P); cudaFree(a.VEL); cudaFree(a.O); cudaFree(a.E); cudaFree(a.ID); cudaFree(a.C); cudaFree(a.V);
    cudaFree(a.S); cudaFree(a.ACC); cudaFree(a.GOAL); cudaFree(a.PV); cudaFree(a.PB); cudaFree(a.VN);
    if (with_cold) { cudaFree(a.COLD_GOAL); cudaFree(a.COLD_PV); }
}

static int setup_grid_buffers(sphsm_handle *h) {
    // (re)allocates what depends on the number of cells
    if (h->cell_start) cudaFree(h->cell_start);
    h->cell_start = nullptr;
    CU(cudaMalloc(&h->cell_start, ((size_t)h->dp.num_cells + 2) * sizeof(int)));
    if (h->cell_count) cudaFree(h->cell_count);
    if (h->tile_sums) cudaFree(h->tile_sums);
    h->cell_count = h->tile_sums = nullptr;
    CU(cudaMalloc(&h->cell_count, ((size_t)h->dp.num_cells + 2 + SCAN_IPT) * sizeof(uint32_t)));
    CU(cudaMemset(h->cell_count, 0, ((size_t)h->dp.num_cells + 2 + SCAN_IPT) * sizeof(uint32_t)));
    h->counts_ready = false;
    CU(cudaMalloc(&h->tile_sums, ((size_t)(h->dp.num_cells + 2) / SCAN_TILE + 2) * sizeof(uint32_t)));
    int bits = 1;
    while ((1ll << bits) < (long long)h->dp.num_cells + 1) bits++;
    h->sort_passes = (bits + RADIX_BITS - 1) / RADIX_BITS;
    if (h->sort_passes > MAX_SORT_PASSES) return fail(h, SPHSM_ERR_INVALID, "grid too large for the radix sort key");
    return SPHSM_OK;
}

extern "C" int sphsm_create(const sphsm_params *p, sphsm_handle **out) {
    sphsm_handle *h = nullptr;
    if (!p || !out) return fail(nullptr, SPHSM_ERR_INVALID, "null argument");
    if (p->struct_size != sizeof(sphsm_params)) return fail(nullptr, SPHSM_ERR_INVALID, "sphsm_params.struct_size mismatch (ABI)");
    if (p->capacity <= 0 || p->kernel_h <= 0.f || p->world[0] <= 0.f || p->world[1] <= 0.f || p->world[2] <= 0.f)
        return fail(nullptr, SPHSM_ERR_INVALID, "capacity, kernel_h and world must be positive");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (p->device < 0 || p->device >= ndev) return fail(nullptr, SPHSM_ERR_CUDA, "no such CUDA device (no CPU fallback exists)");
    CU(cudaSetDevice(p->device));
    sphsm_handle *nh = new sphsm_handle();
    nh->prm = *p;
    nh->n = 0;
    derive_dev_params(nh);
    {
        long long cells = (long long)nh->dp.ga * nh->dp.gb * nh->dp.gcl;
        if (cells <= 0 || cells >= (1ll << 30)) {
            delete nh;
            return fail(nullptr, SPHSM_ERR_INVALID, "grid has too many cells");
        }
    }
    h = nh;
    // slab mode appends up to two halo messages behind the local particles before every sort: room for them
    if (p->slab_axis >= 0) {
        int hc = p->reserved[0];  // halo capacity override (particles per message)
        if (hc <= 0) {
            // default: 2.5 times the average population of a cell plane across the slab axis at full capacity (a cell
            // plane of a regular lattice holds one OR two lattice planes), plus slack for migrants
            const double planes = std::max(1.0, ceil((double)p->world[p->slab_axis] / (double)p->kernel_h));
            hc = (int)std::min((double)p->capacity, 2.5 * (double)p->capacity / planes) + 4096;
        }
        h->send_cap = hc;
    }
    h->alloc_n = p->capacity + 2 * h->send_cap;
    const int cap = h->alloc_n;
    int rc;
    CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    {
        int prio_lo = 0, prio_hi = 0;  // the side stream carries the short chains everything else waits for
        CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CU(cudaStreamCreateWithPriority(&h->side_stream, cudaStreamNonBlocking, prio_hi));
    }
    CU(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_meta, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_flag, cudaEventDisableTiming));
    CU(cudaMallocHost(&h->h_flag, sizeof(double)));
    *h->h_flag = 0.0;
    CU(cudaStreamCreateWithFlags(&h->h2d_stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
    for (cudaEvent_t *e : {&h->ev_in_ready, &h->ev_in_free, &h->ev_out_ready, &h->ev_out_done}) CU(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_bnd, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_int, cudaEventDisableTiming));
    h->launch_stream = h->stream;
    if ((rc = alloc_arrays(h, h->cur, cap, true)) != 0) return rc;
    if ((rc = alloc_arrays(h, h->alt, cap, false)) != 0) return rc;
    h->alt.COLD_GOAL = h->cur.COLD_GOAL;
    h->alt.COLD_PV = h->cur.COLD_PV;
    for (int k = 0; k < 2; k++) {
        CU(cudaMalloc(&h->keys[k], (size_t)cap * sizeof(uint32_t)));
        CU(cudaMalloc(&h->vals[k], (size_t)cap * sizeof(uint32_t)));
    }
    h->max_tiles = cdiv(cap, SORT_TILE);
    CU(cudaMalloc(&h->ghist, MAX_SORT_PASSES * RADIX * sizeof(uint32_t)));
    CU(cudaMalloc(&h->tile_state, (size_t)MAX_SORT_PASSES * h->max_tiles * RADIX * sizeof(uint32_t)));
    CU(cudaMalloc(&h->tile_counter, MAX_SORT_PASSES * sizeof(uint32_t)));
    CU(cudaMalloc(&h->slot_of, (size_t)cap * sizeof(int)));
    CU(cudaMalloc(&h->big_cells, ((size_t)cap / BIG_CELL + 2) * sizeof(int)));
    CU(cudaMalloc(&h->big_count, sizeof(int)));
    CU(cudaMemset(h->big_count, 0, sizeof(int)));
    CU(cudaMalloc(&h->d_dp, sizeof(DevParams)));
    CU(cudaMalloc(&h->sm, sizeof(SmState)));
    CU(cudaMemset(h->sm, 0, sizeof(SmState)));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, p->device));
    h->red_blocks = std::max(1, std::min(8 * prop.multiProcessorCount, cdiv(cap, 256)));  // enough loads in flight to cover HBM latency
    CU(cudaMalloc(&h->partial, (size_t)h->red_blocks * 10 * 9 * sizeof(double)));
    CU(cudaMalloc(&h->totals, 128 * sizeof(double)));
    CU(cudaMalloc(&h->scratch, 162 * sizeof(float)));
    if ((rc = setup_grid_buffers(h)) != 0) return rc;
    for (auto &e : h->ev) CU(cudaEventCreate(&e));
    CU(cudaEventCreate(&h->ev_step0));
    CU(cudaEventCreate(&h->ev_step1));
    *out = h;
    return SPHSM_OK;
}

extern "C" int sphsm_destroy(sphsm_handle *h) {
    if (!h) return SPHSM_OK;
    cudaSetDevice(h->prm.device);
    cudaStreamSynchronize(h->stream);
    free_arrays(h->cur, true);
    free_arrays(h->alt, false);
    for (int k = 0; k < 2; k++) { cudaFree(h->keys[k]); cudaFree(h->vals[k]); }
    cudaFree(h->cell_count); cudaFree(h->tile_sums); cudaFree(h->big_cells); cudaFree(h->big_count);
    cudaFree(h->ghist); cudaFree(h->tile_state); cudaFree(h->tile_counter); cudaFree(h->cell_start); cudaFree(h->slot_of);
    cudaFree(h->d_dp); cudaFree(h->sm); cudaFree(h->partial); cudaFree(h->totals); cudaFree(h->scratch); cudaFree(h->d_aos); cudaFree(h->d_tmp);
    cudaFree(h->d_itmp);
    for (int k = 0; k < 2; k++) { cudaFree(h->msg_send[k]); cudaFree(h->msg_recv[k]); }
    cudaFree(h->d_err); cudaFree(h->d_meta);
    if (h->h_meta) cudaFreeHost(h->h_meta);
    if (h->nccl_comm_red && h->nccl_comm_red != h->nccl_comm && g_nccl_destroy) g_nccl_destroy(h->nccl_comm_red);
    if (h->nccl_comm && g_nccl_destroy) g_nccl_destroy(h->nccl_comm);
    for (auto &e : h->ev) if (e) cudaEventDestroy(e);
    if (h->ev_step0) cudaEventDestroy(h->ev_step0);
    if (h->ev_step1) cudaEventDestroy(h->ev_step1);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    for (auto &gx : h->graphs) cudaGraphExecDestroy(gx.exec);
    h->graphs.clear();
    if (h->ev_meta) cudaEventDestroy(h->ev_meta);
    if (h->ev_flag) cudaEventDestroy(h->ev_flag);
    if (h->h_flag) cudaFreeHost(h->h_flag);
    for (cudaEvent_t e : {h->ev_in_ready, h->ev_in_free, h->ev_out_ready, h->ev_out_done})
        if (e) cudaEventDestroy(e);
    if (h->h2d_stream) cudaStreamDestroy(h->h2d_stream);
    if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
    cudaFree(h->io_in_f); cudaFree(h->io_in_b); cudaFree(h->io_out_f); cudaFree(h->io_out_i);
    if (h->ev_bnd) cudaEventDestroy(h->ev_bnd);
    if (h->ev_int) cudaEventDestroy(h->ev_int);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return SPHSM_OK;
}

extern "C" const char *sphsm_last_error(sphsm_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" int sphsm_get_params(sphsm_handle *h, sphsm_params *out) {
    if (!h || !out) return SPHSM_ERR_INVALID;
    *out = h->prm;
    return SPHSM_OK;
}

extern "C" int sphsm_set_params(sphsm_handle *h, const sphsm_params *p) {
    if (!h || !p) return SPHSM_ERR_INVALID;
    if (p->struct_size != sizeof(sphsm_params)) return fail(h, SPHSM_ERR_INVALID, "sphsm_params.struct_size mismatch (ABI)");
    if (p->capacity != h->prm.capacity || p->device != h->prm.device || p->kernel_h != h->prm.kernel_h ||
        memcmp(p->world, h->prm.world, sizeof p->world) != 0 || p->slab_axis != h->prm.slab_axis)
        return fail(h, SPHSM_ERR_INVALID, "capacity, device, kernel_h, world and slab_axis are fixed at create");
    const DevParams old = h->dp;
    h->prm = *p;
    derive_dev_params(h);
    h->dp.c_off = old.c_off; h->dp.gcl = old.gcl; h->dp.slab_lo = old.slab_lo; h->dp.slab_hi = old.slab_hi; h->dp.num_cells = old.num_cells;
    h->dp.slab_on = old.slab_on; h->dp.own_begin = old.own_begin; h->dp.own_end = old.own_end;
    return SPHSM_OK;
}

// ---------------------------------------------------------------------------------------------------
// AoS <-> SoA
__device__ __forceinline__ float ldf(const uint8_t *rec, int off) { return *reinterpret_cast<const float *>(rec + off); }
__device__ __forceinline__ void stf(uint8_t *rec, int off, float v) { *reinterpret_cast<float *>(rec + off) = v; }

__global__ void k_aos_to_soa(int first, int count, const uint8_t *__restrict__ aos, int stride, Arrays a) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int i = first + k;
    const uint8_t *r = aos + (size_t)k * stride;
    const float mass = ldf(r, OFF_MASS), dens = ldf(r, OFF_DENS), Vm = ldf(r, OFF_VM);
    const int fixed = r[OFF_FIXED] != 0;
    a.P[i] = make_float4(ldf(r, OFF_POS), ldf(r, OFF_POS + 4), ldf(r, OFF_POS + 8), mass);
    a.VEL[i] = make_float4(ldf(r, OFF_VEL), ldf(r, OFF_VEL + 4), ldf(r, OFF_VEL + 8), dens);
    a.O[i] = make_float4(ldf(r, OFF_ORIG), ldf(r, OFF_ORIG + 4), ldf(r, OFF_ORIG + 8), __int_as_float(fixed ? i + 1 : 0));
    a.E[i] = make_float4(Vm, ldf(r, OFF_IION), ldf(r, OFF_W), ldf(r, OFF_STIM));
    a.ID[i] = i;
    a.C[i] = make_float4(ldf(r, OFF_CVEL), ldf(r, OFF_CVEL + 4), ldf(r, OFF_CVEL + 8), __fdiv_rn(mass, dens));
    a.V[i] = make_float4(ldf(r, OFF_IVEL), ldf(r, OFF_IVEL + 4), ldf(r, OFF_IVEL + 8), __fdiv_rn(mass, dens));
    a.S[i] = make_float2(ldf(r, OFF_PRES), Vm);
    a.ACC[i] = make_float4(ldf(r, OFF_ACC), ldf(r, OFF_ACC + 4), ldf(r, OFF_ACC + 8), ldf(r, OFF_IVM));
    const float4 goal = make_float4(ldf(r, OFF_GOAL), ldf(r, OFF_GOAL + 4), ldf(r, OFF_GOAL + 8), 0.f);
    const float4 pv = make_float4(ldf(r, OFF_PVEL), ldf(r, OFF_PVEL + 4), ldf(r, OFF_PVEL + 8), 0.f);
    a.GOAL[i] = goal; a.PV[i] = pv;
    a.COLD_GOAL[i] = goal; a.COLD_PV[i] = pv;
}

__global__ void k_soa_to_aos(int first, int count, Arrays a, uint8_t *__restrict__ aos, int stride) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= count) return;
    s += first;
    const int id = a.ID[s];
    uint8_t *r = aos + (size_t)id * stride;
    const float4 p = a.P[s], v = a.VEL[s], o = a.O[s], e = a.E[s], c = a.C[s], iv = a.V[s], acc = a.ACC[s];
    const float2 sp = a.S[s];
    const int flags = __float_as_int(o.w);
    const float4 goal = flags ? a.COLD_GOAL[flags - 1] : a.GOAL[s];
    const float4 pv = flags ? a.COLD_PV[flags - 1] : a.PV[s];
    stf(r, OFF_POS, p.x); stf(r, OFF_POS + 4, p.y); stf(r, OFF_POS + 8, p.z);
    stf(r, OFF_VEL, v.x); stf(r, OFF_VEL + 4, v.y); stf(r, OFF_VEL + 8, v.z);
    stf(r, OFF_PVEL, pv.x); stf(r, OFF_PVEL + 4, pv.y); stf(r, OFF_PVEL + 8, pv.z);
    stf(r, OFF_IVEL, iv.x); stf(r, OFF_IVEL + 4, iv.y); stf(r, OFF_IVEL + 8, iv.z);
    stf(r, OFF_CVEL, c.x); stf(r, OFF_CVEL + 4, c.y); stf(r, OFF_CVEL + 8, c.z);
    stf(r, OFF_ACC, acc.x); stf(r, OFF_ACC + 4, acc.y); stf(r, OFF_ACC + 8, acc.z);
    stf(r, OFF_MASS, p.w);
    stf(r, OFF_ORIG, o.x); stf(r, OFF_ORIG + 4, o.y); stf(r, OFF_ORIG + 8, o.z);
    stf(r, OFF_GOAL, goal.x); stf(r, OFF_GOAL + 4, goal.y); stf(r, OFF_GOAL + 8, goal.z);
    r[OFF_FIXED] = flags ? 1 : 0; r[OFF_FIXED + 1] = 0; r[OFF_FIXED + 2] = 0; r[OFF_FIXED + 3] = 0;
    stf(r, OFF_DENS, v.w); stf(r, OFF_PRES, sp.x); stf(r, OFF_VM, e.x); stf(r, OFF_IVM, acc.w);
    stf(r, OFF_IION, e.y); stf(r, OFF_STIM, e.w); stf(r, OFF_W, e.z);
}

__global__ void k_positions_out(int first, int count, Arrays a, float *__restrict__ xyz) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= count) return;
    s += first;
    const int id = a.ID[s];
    const float4 p = a.P[s];
    xyz[3 * (size_t)id] = p.x; xyz[3 * (size_t)id + 1] = p.y; xyz[3 * (size_t)id + 2] = p.z;
}

// Init_Particle, cpp:101-125
__global__ void k_init_particles(int first, int count, const float *__restrict__ xyz, Arrays a, float mass, float rho0) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int i = first + k;
    const float x = xyz[3 * (size_t)k], y = xyz[3 * (size_t)k + 1], z = xyz[3 * (size_t)k + 2];
    a.P[i] = make_float4(x, y, z, mass);
    a.VEL[i] = make_float4(0.f, 0.f, 0.f, rho0);
    a.O[i] = make_float4(x, y, z, __int_as_float(0));
    a.E[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    a.ID[i] = i;
    a.C[i] = make_float4(0.f, 0.f, 0.f, __fdiv_rn(mass, rho0));
    a.V[i] = make_float4(0.f, 0.f, 0.f, __fdiv_rn(mass, rho0));
    a.S[i] = make_float2(0.f, 0.f);
    a.ACC[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    a.GOAL[i] = make_float4(x, y, z, 0.f);
    a.PV[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    a.COLD_GOAL[i] = make_float4(x, y, z, 0.f);
    a.COLD_PV[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void k_slot_of(int n, const int *__restrict__ id, int *__restrict__ slot_of) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) slot_of[id[s]] = s;
}

// ---- stimulation / fixation ------------------------------------------------------------------------------
// set_stim (cpp:704-717) for a LIST of centres: stim = strength where the squared distance to any centre <= radius
__global__ void __launch_bounds__(256) k_set_stim_list(int n, Arrays a, const float *__restrict__ centres, int m, float radius, float strength) {
    __shared__ float sc[256 * 3];
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    float4 p = make_float4(0, 0, 0, 0);
    if (s < n) p = a.P[s];
    bool hit = false;
    for (int base = 0; base < m; base += 256) {
        const int cnt = min(256, m - base);
        __syncthreads();
        for (int k = threadIdx.x; k < cnt * 3; k += blockDim.x) sc[k] = centres[(size_t)base * 3 + k];
        __syncthreads();
        if (!hit && s < n) {
            for (int k = 0; k < cnt; k++) {
                const float dx = __fsub_rn(p.x, sc[3 * k]), dy = __fsub_rn(p.y, sc[3 * k + 1]), dz = __fsub_rn(p.z, sc[3 * k + 2]);
                if (dist2_exact(dx, dy, dz) <= radius) { hit = true; break; }
            }
        }
        if (__syncthreads_and(hit || s >= n)) break;
    }
    if (hit && s < n) a.E[s].w = strength;
}

// mode 0: turnOnStim_Mesh's fixation rule (cpp:759), 1: turnOnStim_Cube's (cpp:738); comparisons against double
// literals are done in double exactly as the reference's promotions make them.
__global__ void k_fix_rule(int n, Arrays a, int mode, int diag) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const float4 p = a.P[s];
    bool fix;
    if (mode == 0) {
        const double x = p.x, y = p.y;
        fix = (x >= 0.0 && x <= 0.07) || (x >= 0.90 && y >= 0.80);
    } else {
        fix = (p.y == 0.0f && p.x <= 0.48f) || (p.y == 0.0f && (double)p.x >= 1.0);
    }
    if (fix && __float_as_int(a.O[s].w) == 0) {
        const int id = a.ID[s];
        if (diag) { a.COLD_GOAL[id] = a.GOAL[s]; a.COLD_PV[id] = a.PV[s]; }
        a.O[s].w = __int_as_float(id + 1);
    }
}

__global__ void k_set_masks(int n, Arrays a, const uint8_t *__restrict__ fixed, const float *__restrict__ stim, int diag) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int id = a.ID[s];
    if (fixed) {
        const int was = __float_as_int(a.O[s].w);
        if (fixed[id] && !was) {
            if (diag) { a.COLD_GOAL[id] = a.GOAL[s]; a.COLD_PV[id] = a.PV[s]; }
            a.O[s].w = __int_as_float(id + 1);
        } else if (!fixed[id] && was) {
            a.GOAL[s] = a.COLD_GOAL[was - 1];
            a.PV[s] = a.COLD_PV[was - 1];
            a.O[s].w = __int_as_float(0);
        }
    }
    if (stim) a.E[s].w = stim[id];
}

// turnOffStim, cpp:764-783
__global__ void k_stim_off(int n, Arrays a) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    a.E[s] = make_float4(0.0f, 0.0f, 0.0f, -10000.0f);  // Vm, Iion, w = 0; stim = -10000
    a.S[s] = make_float2(-10000.0f, 0.0f);              // pres = -10000
    a.ACC[s].w = 0.0f;                                  // Inter_Vm = 0
}

// ---------------------------------------------------------------------------------------------------
static int ensure_tmp(sphsm_handle *h, size_t floats) {
    if (h->tmp_cap >= floats) return SPHSM_OK;
    if (h->d_tmp) cudaFree(h->d_tmp);
    h->d_tmp = nullptr; h->tmp_cap = 0;
    CU(cudaMalloc(&h->d_tmp, floats * sizeof(float)));
    h->tmp_cap = floats;
    return SPHSM_OK;
}
static int ensure_itmp(sphsm_handle *h, size_t ints) {
    if (h->itmp_cap >= ints) return SPHSM_OK;
    if (h->d_itmp) cudaFree(h->d_itmp);
    h->d_itmp = nullptr; h->itmp_cap = 0;
    CU(cudaMalloc(&h->d_itmp, ints * sizeof(int)));
    h->itmp_cap = ints;
    return SPHSM_OK;
}
static int ensure_aos(sphsm_handle *h, size_t bytes) {
    if (h->aos_cap_bytes >= bytes) return SPHSM_OK;
    if (h->d_aos) cudaFree(h->d_aos);
    h->d_aos = nullptr; h->aos_cap_bytes = 0;
    CU(cudaMalloc(&h->d_aos, bytes));
    h->aos_cap_bytes = bytes;
    return SPHSM_OK;
}
// positions (or the grid) are about to change behind pass B's back: its pre-filed counts for the next sort are void
static void drop_counts(sphsm_handle *h) {
    if (!h->counts_ready) return;
    h->counts_ready = false;
    if (h->cell_count) cudaMemsetAsync(h->cell_count, 0, ((size_t)h->dp.num_cells + 2 + SCAN_IPT) * sizeof(uint32_t), h->stream);
}
static void state_changed(sphsm_handle *h, bool rest) {
    drop_counts(h);
    h->grid_valid = false;
    h->slot_of_valid = false;
    h->inter_live = true;
    if (rest) {
        h->rest_dirty = true;
        h->slab_applied = false;  // the particle set was replaced: sphsm_comm_set_slab must be applied again
    }
}

// Host-side estimate of the mean cell occupancy of a small particle set, from the positions as the caller hands them in
// (Init_Fluid / upload): what decides between the thread-per-particle and the warp-per-particle passes is candidates per
// stencil row, i.e. particles per occupied cell, and the host never sees the cell table.  Only tracked up to the warp-path limit.
static void note_host_positions(sphsm_handle *h, const float *x, size_t stride_floats, int count, bool reset) {
    if (reset) h->host_cells.clear();
    if (h->n + count > 4 * WARP_PATH_MAX || (int)h->host_cells.size() > 4 * WARP_PATH_MAX) return;
    const float cs = h->prm.kernel_h;
    for (int k = 0; k < count; k++) {
        const float *q = x + (size_t)k * stride_floats;
        if (!(fabsf(q[0]) < 1e9f && fabsf(q[1]) < 1e9f && fabsf(q[2]) < 1e9f)) continue;  // NaN / absurd: no cell
        const long long cx = (long long)(q[0] / cs), cy = (long long)(q[1] / cs), cz = (long long)(q[2] / cs);
        h->host_cells.insert((cx & 0x1fffff) | ((cy & 0x1fffff) << 21) | ((cz & 0x1fffff) << 42));
    }
}

extern "C" int sphsm_init_fluid(sphsm_handle *h, const float *xyz, int n) {
    if (!h || (!xyz && n > 0) || n < 0) return SPHSM_ERR_INVALID;
    CU(cudaSetDevice(h->prm.device));
    const int take = std::min(n, h->prm.capacity - h->n);  // cpp:103: the rest is silently dropped
    if (take <= 0) return SPHSM_OK;
    int rc;
    if ((rc = ensure_tmp(h, (size_t)take * 3)) != 0) return rc;
    CU(cudaMemcpyAsync(h->d_tmp, xyz, (size_t)take * 3 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    LAUNCH(k_init_particles, cdiv(take, 256), 256, h->n, take, h->d_tmp, h->cur, h->prm.particle_mass, h->prm.stand_density);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->stream));  // xyz may be pageable / freed by the caller
    note_host_positions(h, xyz, 3, take, h->n == 0);
    set_n(h, h->n + take);
    state_changed(h, true);
    return SPHSM_OK;
}

extern "C" int sphsm_upload_aos(sphsm_handle *h, const void *particles, int n, int stride) {
    if (!h || !particles || n < 0 || stride < SPHSM_PARTICLE_STRIDE) return SPHSM_ERR_INVALID;
    if (n > h->prm.capacity) return fail(h, SPHSM_ERR_CAPACITY, "more particles than capacity");
    CU(cudaSetDevice(h->prm.device));
    int rc;
    if ((rc = ensure_aos(h, (size_t)std::max(n, 1) * stride)) != 0) return rc;
    CU(cudaMemcpyAsync(h->d_aos, particles, (size_t)n * stride, cudaMemcpyHostToDevice, h->stream));
    if (n > 0) LAUNCH(k_aos_to_soa, cdiv(n, 256), 256, 0, n, h->d_aos, stride, h->cur);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->stream));
    if (stride % 4 == 0) {
        h->n = 0;
        note_host_positions(h, reinterpret_cast<const float *>(static_cast<const uint8_t *>(particles) + OFF_POS), (size_t)stride / 4, n, true);
    } else h->host_cells.clear();
    set_n(h, n);
    state_changed(h, true);
    return SPHSM_OK;
}

extern "C" int sphsm_download_aos(sphsm_handle *h, void *particles, int n, int stride) {
    if (!h || !particles || n < 0 || stride < SPHSM_PARTICLE_STRIDE) return SPHSM_ERR_INVALID;
    // slab mode: n counts GLOBAL particles (the caller's array is indexed by original id); only the particles this rank
    // owns are written, everything else in the caller's array keeps its bytes
    const bool slab = h->dp.slab_on != 0;
    if (n > (slab ? h->prm.capacity : h->n)) return fail(h, SPHSM_ERR_INVALID, "n exceeds the number of particles");
    CU(cudaSetDevice(h->prm.device));
    int rc;
    if ((rc = ensure_aos(h, (size_t)std::max(std::max(h->n, n), 1) * stride)) != 0) return rc;
    if (stride != SPHSM_PARTICLE_STRIDE || slab)  // keep the caller's other bytes: round-trip through the device image
        CU(cudaMemcpyAsync(h->d_aos, particles, (size_t)n * stride, cudaMemcpyHostToDevice, h->stream));
    const int nown = h->dp.own_end - h->dp.own_begin;
    if (nown > 0) LAUNCH(k_soa_to_aos, cdiv(nown, 256), 256, h->dp.own_begin, nown, h->cur, h->d_aos, stride);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(particles, h->d_aos, (size_t)n * stride, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return SPHSM_OK;
}

extern "C" int sphsm_download_positions(sphsm_handle *h, float *xyz, int n) {
    if (!h || !xyz || n < 0) return SPHSM_ERR_INVALID;
    const bool slab = h->dp.slab_on != 0;  // slab mode: n is the GLOBAL count, only owned particles are written
    if (n > (slab ? h->prm.capacity : h->n)) return SPHSM_ERR_INVALID;
    CU(cudaSetDevice(h->prm.device));
    int rc;
    if ((rc = ensure_tmp(h, (size_t)std::max(std::max(h->n, n), 1) * 3)) != 0) return rc;
    if (slab) CU(cudaMemcpyAsync(h->d_tmp, xyz, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    const int nown = h->dp.own_end - h->dp.own_begin;
    if (nown > 0) LAUNCH(k_positions_out, cdiv(nown, 256), 256, h->dp.own_begin, nown, h->cur, h->d_tmp);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(xyz, h->d_tmp, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return SPHSM_OK;
}

static int stim_list(sphsm_handle *h, const float *centres, int m, float radius, float strength) {
    if (m <= 0 || h->n <= 0) return SPHSM_OK;
    int rc;
    if ((rc = ensure_tmp(h, (size_t)m * 3)) != 0) return rc;
    CU(cudaMemcpyAsync(h->d_tmp, centres, (size_t)m * 3 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    LAUNCH(k_set_stim_list, cdiv(h->n, 256), 256, h->n, h->cur, h->d_tmp, m, radius, strength);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->stream));
    return SPHSM_OK;
}

extern "C" int sphsm_set_stim(sphsm_handle *h, float cx, float cy, float cz, float radius, float strength) {
    if (!h) return SPHSM_ERR_INVALID;
    CU(cudaSetDevice(h->prm.device));
    const float c[3] = {cx, cy, cz};
    return stim_list(h, c, 1, radius, strength);
}

extern "C" int sphsm_stim_mesh(sphsm_handle *h, const float *xyz, int n) {
    if (!h || (!xyz && n > 0)) return SPHSM_ERR_INVALID;
    CU(cudaSetDevice(h->prm.device));
    int rc = stim_list(h, xyz, n, 0.01f, h->prm.stim_strength);  // cpp:753
    if (rc) return rc;
    if (h->n > 0) LAUNCH(k_fix_rule, cdiv(h->n, 256), 256, h->n, h->cur, 0, h->prm.diagnostics);
    CU(cudaGetLastError());
    h->rest_dirty = true;
    return SPHSM_OK;
}

extern "C" int sphsm_stim_cube(sphsm_handle *h, const float *xyz, int n) {
    if (!h || (!xyz && n > 0)) return SPHSM_ERR_INVALID;
    CU(cudaSetDevice(h->prm.device));
    std::vector<float> sel;
    for (int i = 0; i < n; i++) {  // the position filter of cpp:727 (double-literal comparisons)
        const float px = xyz[3 * i], pz = xyz[3 * i + 2];
        if (((double)px >= 0.45 && (double)px <= 0.48) || ((double)px > 1.0 && pz <= 1.05f)) {
            sel.push_back(px); sel.push_back(xyz[3 * i + 1]); sel.push_back(pz);
        }
    }
    int rc = stim_list(h, sel.data(), (int)(sel.size() / 3), 0.001f, h->prm.stim_strength);  // cpp:728
    if (rc) return rc;
    if (h->n > 0) LAUNCH(k_fix_rule, cdiv(h->n, 256), 256, h->n, h->cur, 1, h->prm.diagnostics);
    CU(cudaGetLastError());
    h->rest_dirty = true;
    return SPHSM_OK;
}

extern "C" int sphsm_stim_off(sphsm_handle *h) {
    if (!h) return SPHSM_ERR_INVALID;
    CU(cudaSetDevice(h->prm.device));
    if (h->n > 0) LAUNCH(k_stim_off, cdiv(h->n, 256), 256, h->n, h->cur);
    CU(cudaGetLastError());
    return SPHSM_OK;
}

extern "C" int sphsm_set_masks(sphsm_handle *h, const uint8_t *fixed, const float *stim, int n) {
    // slab mode: the arrays are indexed by ORIGINAL (global) id, so they hold the global particle count
    if (!h || n != (h->dp.slab_on ? h->n_global : h->n)) return fail(h, SPHSM_ERR_INVALID, "set_masks needs exactly num_particles entries");
    if (n == 0 || (!fixed && !stim)) return SPHSM_OK;
    CU(cudaSetDevice(h->prm.device));
    int rc;
    if ((rc = ensure_tmp(h, (size_t)n)) != 0) return rc;
    if ((rc = ensure_itmp(h, (size_t)(n + 3) / 4)) != 0) return rc;
    if (stim) CU(cudaMemcpyAsync(h->d_tmp, stim, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    if (fixed) CU(cudaMemcpyAsync(h->d_itmp, fixed, (size_t)n, cudaMemcpyHostToDevice, h->stream));
    if (h->n > 0)
        LAUNCH(k_set_masks, cdiv(h->n, 256), 256, h->n, h->cur, fixed ? (const uint8_t *)h->d_itmp : nullptr, stim ? h->d_tmp : nullptr,
               h->prm.diagnostics);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->stream));
    if (fixed) h->rest_dirty = true;
    return SPHSM_OK;
}

// ---------------------------------------------------------------------------------------------------
// State snapshot / restart (the reference has none).  File = header + sphsm_params + the particles in
// the reference's own Particle layout (132 B, all 33 fields, original order), i.e. exactly what Get_Paticles() shows.
struct SnapshotHeader {
    char magic[8];  // "SPHSMB2\0"
    uint32_t version, header_bytes, params_bytes, stride;
    int32_t n, total_steps;
};
extern "C" int sphsm_save_state(sphsm_handle *h, const char *path) {
    if (!h || !path) return SPHSM_ERR_INVALID;
    if (h->dp.slab_on) return fail(h, SPHSM_ERR_INVALID, "snapshots are written from a single-GPU handle");
    std::vector<uint8_t> buf((size_t)std::max(h->n, 1) * SPHSM_PARTICLE_STRIDE);
    int rc = h->n > 0 ? sphsm_download_aos(h, buf.data(), h->n, SPHSM_PARTICLE_STRIDE) : SPHSM_OK;
    if (rc) return rc;
    SnapshotHeader hd;
    memset(&hd, 0, sizeof(hd));
    memcpy(hd.magic, "SPHSMB2", 8);
    hd.version = 1; hd.header_bytes = sizeof(hd); hd.params_bytes = sizeof(sphsm_params); hd.stride = SPHSM_PARTICLE_STRIDE;
    hd.n = h->n; hd.total_steps = h->total_steps;
    FILE *f = fopen(path, "wb");
    if (!f) return fail(h, SPHSM_ERR_INVALID, "cannot open the snapshot file for writing");
    bool ok = fwrite(&hd, sizeof(hd), 1, f) == 1 && fwrite(&h->prm, sizeof(sphsm_params), 1, f) == 1 &&
              (h->n == 0 || fwrite(buf.data(), (size_t)h->n * SPHSM_PARTICLE_STRIDE, 1, f) == 1);
    ok = (fclose(f) == 0) && ok;
    return ok ? SPHSM_OK : fail(h, SPHSM_ERR_INVALID, "short write on the snapshot file");
}
// Restores particles, tunable parameters and the step counter into an existing handle (its capacity, device and world
// stay its own: the snapshot must fit, and its world must match).
extern "C" int sphsm_load_state(sphsm_handle *h, const char *path) {
    if (!h || !path) return SPHSM_ERR_INVALID;
    if (h->dp.slab_on) return fail(h, SPHSM_ERR_INVALID, "load the snapshot before sphsm_comm_set_slab");
    FILE *f = fopen(path, "rb");
    if (!f) return fail(h, SPHSM_ERR_INVALID, "cannot open the snapshot file");
    SnapshotHeader hd;
    sphsm_params sp;
    int rc = SPHSM_OK;
    std::vector<uint8_t> buf;
    if (fread(&hd, sizeof(hd), 1, f) != 1 || memcmp(hd.magic, "SPHSMB2", 8) != 0 || hd.version != 1 || hd.header_bytes != sizeof(hd) ||
        hd.params_bytes != sizeof(sphsm_params) || hd.stride != SPHSM_PARTICLE_STRIDE || hd.n < 0)
        rc = fail(h, SPHSM_ERR_INVALID, "not a snapshot of this library version");
    else if (fread(&sp, sizeof(sp), 1, f) != 1)
        rc = fail(h, SPHSM_ERR_INVALID, "truncated snapshot");
    else if (hd.n > h->prm.capacity)
        rc = fail(h, SPHSM_ERR_CAPACITY, "snapshot holds more particles than this handle's capacity");
    else if (sp.world[0] != h->prm.world[0] || sp.world[1] != h->prm.world[1] || sp.world[2] != h->prm.world[2] || sp.kernel_h != h->prm.kernel_h)
        rc = fail(h, SPHSM_ERR_INVALID, "snapshot was taken in a different world / kernel size");
    else {
        buf.resize((size_t)std::max(hd.n, 1) * SPHSM_PARTICLE_STRIDE);
        if (hd.n > 0 && fread(buf.data(), (size_t)hd.n * SPHSM_PARTICLE_STRIDE, 1, f) != 1) rc = fail(h, SPHSM_ERR_INVALID, "truncated snapshot");
    }
    fclose(f);
    if (rc) return rc;
    sp.device = h->prm.device; sp.capacity = h->prm.capacity; sp.slab_axis = h->prm.slab_axis; sp.strict = h->prm.strict;
    sp.diagnostics = h->prm.diagnostics;
    if ((rc = sphsm_set_params(h, &sp)) != 0) return rc;
    if ((rc = sphsm_upload_aos(h, buf.data(), hd.n, SPHSM_PARTICLE_STRIDE)) != 0) return rc;
    h->total_steps = hd.total_steps;
    return SPHSM_OK;
}

// ---------------------------------------------------------------------------------------------------
// Asynchronous I/O.  The host arrays must be page-locked for the copies to overlap and must stay untouched until
// sphsm_io_wait (or sphsm_sync) returns.  Input copies run on their own stream into their own staging and the kernel that
// applies them waits for the copy; output is gathered on the compute stream and copied out on a second copy stream, so a
// caller that loops { set_masks_async; step; download_*_async } has step k+1 computing while the results of step k cross
// PCIe one way and the inputs of step k+2 cross it the other way.
static int ensure_io_in(sphsm_handle *h, size_t n) {
    if (h->io_in_cap >= n) return SPHSM_OK;
    CU(cudaStreamSynchronize(h->h2d_stream));
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(h->io_in_f); cudaFree(h->io_in_b);
    h->io_in_f = nullptr; h->io_in_b = nullptr; h->io_in_cap = 0;
    CU(cudaMalloc(&h->io_in_f, n * sizeof(float)));
    CU(cudaMalloc(&h->io_in_b, n));
    h->io_in_cap = n;
    return SPHSM_OK;
}
static int ensure_io_out(sphsm_handle *h, size_t n) {
    if (h->io_out_cap >= n) return SPHSM_OK;
    CU(cudaStreamSynchronize(h->d2h_stream));
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(h->io_out_f); cudaFree(h->io_out_i);
    h->io_out_f = nullptr; h->io_out_i = nullptr; h->io_out_cap = 0;
    CU(cudaMalloc(&h->io_out_f, n * 3 * sizeof(float)));
    CU(cudaMalloc(&h->io_out_i, n * sizeof(int)));
    h->io_out_cap = n;
    return SPHSM_OK;
}

extern "C" int sphsm_set_masks_async(sphsm_handle *h, const uint8_t *fixed, const float *stim, int n) {
    if (!h || n != (h->dp.slab_on ? h->n_global : h->n)) return fail(h, SPHSM_ERR_INVALID, "set_masks needs exactly num_particles entries");
    if (n == 0 || (!fixed && !stim)) return SPHSM_OK;
    CU(cudaSetDevice(h->prm.device));
    int rc;
    if ((rc = ensure_io_in(h, (size_t)n)) != 0) return rc;
    CU(cudaStreamWaitEvent(h->h2d_stream, h->ev_in_free, 0));  // the previous call's kernel has consumed the staging
    if (stim) CU(cudaMemcpyAsync(h->io_in_f, stim, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, h->h2d_stream));
    if (fixed) CU(cudaMemcpyAsync(h->io_in_b, fixed, (size_t)n, cudaMemcpyHostToDevice, h->h2d_stream));
    CU(cudaEventRecord(h->ev_in_ready, h->h2d_stream));
    CU(cudaStreamWaitEvent(h->stream, h->ev_in_ready, 0));
    if (h->n > 0)
        LAUNCH(k_set_masks, cdiv(h->n, 256), 256, h->n, h->cur, fixed ? (const uint8_t *)h->io_in_b : nullptr, stim ? h->io_in_f : nullptr,
               h->prm.diagnostics);
    CU(cudaGetLastError());
    CU(cudaEventRecord(h->ev_in_free, h->stream));
    if (fixed) h->rest_dirty = true;
    return SPHSM_OK;
}

static int io_copy_out(sphsm_handle *h, int *ids, float *xyz, size_t count) {
    CU(cudaEventRecord(h->ev_out_ready, h->stream));
    CU(cudaStreamWaitEvent(h->d2h_stream, h->ev_out_ready, 0));
    if (ids) CU(cudaMemcpyAsync(ids, h->io_out_i, count * sizeof(int), cudaMemcpyDeviceToHost, h->d2h_stream));
    CU(cudaMemcpyAsync(xyz, h->io_out_f, count * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->d2h_stream));
    CU(cudaEventRecord(h->ev_out_done, h->d2h_stream));
    return SPHSM_OK;
}

extern "C" int sphsm_download_positions_async(sphsm_handle *h, float *xyz, int n) {
    if (!h || !xyz || n < 0) return SPHSM_ERR_INVALID;
    if (h->dp.slab_on) return fail(h, SPHSM_ERR_INVALID, "slab mode: use sphsm_download_owned_async");
    if (n > h->n) return SPHSM_ERR_INVALID;
    if (n == 0) return SPHSM_OK;
    CU(cudaSetDevice(h->prm.device));
    int rc;
    if ((rc = ensure_io_out(h, (size_t)h->n)) != 0) return rc;
    CU(cudaStreamWaitEvent(h->stream, h->ev_out_done, 0));  // the previous copy has left the staging
    LAUNCH(k_positions_out, cdiv(h->n, 256), 256, 0, h->n, h->cur, h->io_out_f);
    CU(cudaGetLastError());
    return io_copy_out(h, nullptr, xyz, (size_t)n);
}

extern "C" int sphsm_download_owned_async(sphsm_handle *h, int *ids, float *xyz, int cap, int *count) {
    if (!h || !ids || !xyz || !count || cap < 0) return SPHSM_ERR_INVALID;
    CU(cudaSetDevice(h->prm.device));
    const int first = h->dp.own_begin, nown = h->dp.own_end - h->dp.own_begin;
    *count = nown;
    if (nown > cap) return fail(h, SPHSM_ERR_CAPACITY, "output arrays smaller than the number of owned particles");
    if (nown == 0) return SPHSM_OK;
    int rc;
    if ((rc = ensure_io_out(h, (size_t)std::max(nown, h->prm.capacity / std::max(h->nranks, 1) + 65536))) != 0) return rc;
    if ((size_t)nown > h->io_out_cap && (rc = ensure_io_out(h, (size_t)nown)) != 0) return rc;
    CU(cudaStreamWaitEvent(h->stream, h->ev_out_done, 0));
    LAUNCH(k_mg_owned_out, cdiv(nown, 256), 256, first, nown, h->cur, h->io_out_i, h->io_out_f);
    CU(cudaGetLastError());
    return io_copy_out(h, ids, xyz, (size_t)nown);
}

extern "C" int sphsm_io_wait(sphsm_handle *h) {
    if (!h) return SPHSM_ERR_INVALID;
    CU(cudaSetDevice(h->prm.device));
    CU(cudaStreamSynchronize(h->h2d_stream));
    CU(cudaStreamSynchronize(h->d2h_stream));
    return SPHSM_OK;
}

// ---------------------------------------------------------------------------------------------------
// the step
struct GroupTimer {  // records an event at each kernel-group boundary while profiling
    sphsm_handle *h;
    int idx = 0;
    int groups[SPHSM_NUM_KERNEL_GROUPS + 2];
    long long l0;
    explicit GroupTimer(sphsm_handle *hh) : h(hh) {
        l0 = h->launches;
        if (h->profiling) cudaEventRecord(h->ev[0], h->stream);
    }
    void end_group(int g) {
        if (!h->profiling) return;
        groups[idx] = g;
        h->group_launches[g] += (int)(h->launches - l0);
        l0 = h->launches;
        idx++;
        cudaEventRecord(h->ev[idx], h->stream);
    }
    void finish() {
        if (!h->profiling) return;
        cudaEventSynchronize(h->ev[idx]);
        for (int k = 0; k < idx; k++) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, h->ev[k], h->ev[k + 1]);
            h->group_ms[groups[k]] += ms;
        }
    }
};

static void swap_sets(sphsm_handle *h, bool all) {
    std::swap(h->cur.P, h->alt.P); std::swap(h->cur.VEL, h->alt.VEL); std::swap(h->cur.O, h->alt.O);
    std::swap(h->cur.E, h->alt.E); std::swap(h->cur.ID, h->alt.ID); std::swap(h->cur.PB, h->alt.PB);
    if (all) {
        std::swap(h->cur.C, h->alt.C); std::swap(h->cur.V, h->alt.V); std::swap(h->cur.S, h->alt.S);
        std::swap(h->cur.ACC, h->alt.ACC); std::swap(h->cur.GOAL, h->alt.GOAL); std::swap(h->cur.PV, h->alt.PV);
    }
}

// Find_neighbors: hash -> radix sort -> cell table -> reorder.  grid_sort() is the first half (keys + sorted
// permutation), grid_finish() the second (cell table + gather into the new slot order).  The fast path runs the
// shape-matching sums and solve BETWEEN the two halves (they do not depend on slot order) so that the gather can apply
// stage 2's per-particle map while the values are in registers (k_reorder_goal).
static bool use_counting_sort(const sphsm_handle *h) {
    const int mode = h->prm.reserved[2];  // 0 auto, 1 LSD radix sort, 2 counting sort
    if (mode == 1) return false;
    if (mode == 2) return true;
    return (long long)h->dp.num_cells <= 8ll * std::max(h->n, 1) + (1ll << 20);
}

// one pass over the full key: count per cell -> scan (= the cell table) -> scatter -> canonical in-cell order
static int grid_sort_counting(sphsm_handle *h, GroupTimer *gt) {
    const int n = h->n, m = h->dp.num_cells + 1;  // cells + the limbo bucket
    const int tiles = cdiv(m + 1, SCAN_TILE);
    if (h->counts_ready) h->counts_ready = false;  // pass B filed keys, ranks and counts of these positions while it held them
    else LAUNCH(k_cell_count, cdiv(n, 256), 256, h->dp, h->cur.P, h->keys[0], h->keys[1], h->cell_count);
    if (gt) gt->end_group(KG_HASH);
    LAUNCH(k_scan_tile_sums, tiles, SCAN_THREADS, h->cell_count, m, h->tile_sums);
    LAUNCH(k_scan_tile_offsets, 1, 1024, h->tile_sums, tiles, h->big_count);
    LAUNCH(k_scan_apply, tiles, SCAN_THREADS, h->cell_count, m, h->tile_sums, h->cell_start);
    LAUNCH(k_cell_scatter, cdiv(n, 256), 256, n, h->keys[0], h->keys[1], h->cell_start, h->vals[0]);
    LAUNCH(k_cell_sort_ids, cdiv(h->dp.num_cells, 256), 256, h->cell_start, h->vals[0], h->cur.ID, h->dp.num_cells, h->big_cells, h->big_count);
    LAUNCH(k_cell_sort_big, 64, 256, h->cell_start, h->vals[0], h->vals[1], h->cur.ID, h->big_cells, h->big_count);
    h->sorted_buf = 0;
    h->bounds_ready = true;
    if (gt) gt->end_group(KG_SORT);
    return SPHSM_OK;
}

static int grid_sort(sphsm_handle *h, GroupTimer *gt) {
    const int n = h->n;
    if (use_counting_sort(h)) return grid_sort_counting(h, gt);
    drop_counts(h);
    h->bounds_ready = false;
    const int passes = h->sort_passes;
    const int tiles = cdiv(n, SORT_TILE);
    if (!h->dry_run) {  // (a replayed graph carries its own copies of these nodes)
        CU(cudaMemsetAsync(h->ghist, 0, MAX_SORT_PASSES * RADIX * sizeof(uint32_t), h->stream));
        CU(cudaMemsetAsync(h->tile_state, 0, (size_t)passes * tiles * RADIX * sizeof(uint32_t), h->stream));
        CU(cudaMemsetAsync(h->tile_counter, 0, MAX_SORT_PASSES * sizeof(uint32_t), h->stream));
    }
    LAUNCH(k_hash, std::min(cdiv(n, 256), 8 * 148), 256, h->dp, h->cur.P, h->keys[0], h->ghist, passes);
    if (gt) gt->end_group(KG_HASH);
    int src = 0;
    for (int k = 0; k < passes; k++) {
        LAUNCH(k_radix_pass, tiles, SORT_THREADS, h->keys[src], k == 0 ? nullptr : h->vals[src], h->keys[src ^ 1], h->vals[src ^ 1], n,
               k * RADIX_BITS, h->ghist + k * RADIX, h->tile_state + (size_t)k * tiles * RADIX, h->tile_counter + k);
        src ^= 1;
    }
    h->sorted_buf = src;
    // in-cell order = ascending original index: the reference's bucket order (strict mode), and the canonical order that
    // makes both sides of a slab face hold the shared plane identically (slab mode; reserved[1] forces it on one GPU)
    // (the slab step orders only the planes on either side of its faces, once the plane boundaries are known)
    if (h->prm.strict || h->prm.reserved[1])
        LAUNCH(k_cell_order_fix, cdiv(n, 128), 128, h->keys[src], h->vals[src], h->cur.ID, n, (uint32_t)h->dp.num_cells, 0, n);
    if (gt) gt->end_group(KG_SORT);
    return SPHSM_OK;
}

// fuse_goal: 0 = plain gather; 1 / 2 = gather + goal / predicted / corrected velocity (2 also stores GOAL and PV)
// n_dev != nullptr: the live count is read from device memory (grid sized for h->n, an upper bound)
static int grid_finish(sphsm_handle *h, GroupTimer *gt, int fuse_goal, bool bounds_done = false, const int *n_dev = nullptr) {
    const int n = h->n, src = h->sorted_buf;
    if (!bounds_done && !h->bounds_ready) LAUNCH(k_cell_bounds, cdiv(n + 1, 256), 256, h->keys[src], h->cell_start, n, h->dp.num_cells);
    if (fuse_goal) {
        if (fuse_goal == 2) LAUNCH(k_reorder_goal<true>, cdiv(n, 256), 256, h->dp, h->vals[src], h->cur, h->alt, h->sm, n_dev);
        else LAUNCH(k_reorder_goal<false>, cdiv(n, 256), 256, h->dp, h->vals[src], h->cur, h->alt, h->sm, n_dev);
        swap_sets(h, false);
        std::swap(h->cur.C, h->alt.C);
        std::swap(h->cur.GOAL, h->alt.GOAL);
        std::swap(h->cur.PV, h->alt.PV);
    } else {
        const bool all = h->prm.diagnostics || h->inter_live;
        LAUNCH(k_reorder, cdiv(n, 256), 256, n, h->vals[src], h->cur, h->alt, all ? 1 : 0);
        swap_sets(h, all);
    }
    if (gt) gt->end_group(KG_GRID);
    CU(cudaGetLastError());
    h->grid_valid = true;
    h->slot_of_valid = false;
    return SPHSM_OK;
}

static int build_grid(sphsm_handle *h, GroupTimer *gt) {
    if (h->n == 0) { h->grid_valid = true; return SPHSM_OK; }
    int rc;
    if ((rc = grid_sort(h, gt)) != 0) return rc;
    return grid_finish(h, gt, 0);
}

static int ensure_slot_of(sphsm_handle *h) {
    if (h->slot_of_valid || h->n == 0) return SPHSM_OK;
    LAUNCH(k_slot_of, cdiv(h->n, 256), 256, h->n, h->cur.ID, h->slot_of);
    CU(cudaGetLastError());
    h->slot_of_valid = true;
    return SPHSM_OK;
}

// sums over particles end in h->totals; in slab mode the host combines them across ranks (ncclAllReduce) between parts
static int comm_allreduce(sphsm_handle *h, int count);

// the sums run BEFORE the gather, over the unsorted arrays: in slab mode that extent includes the message regions
static DevParams moment_params(sphsm_handle *h) {
    DevParams d = h->dp;
    if (h->mom_n) d.n = h->mom_n;
    return d;
}
static int rest_part1(sphsm_handle *h) {
    const int B = h->red_blocks;
    LAUNCH(k_rest_pass1, B, 256, moment_params(h), h->cur.P, h->cur.O, h->partial);
    LAUNCH(k_sum_partials_par, 5, 256, h->partial, B, 5, h->totals);
    return SPHSM_OK;
}
static int rest_part2(sphsm_handle *h) {
    const int B = h->red_blocks;
    LAUNCH(k_rest_finalize1, 1, 1, h->totals, h->sm);
    LAUNCH(k_rest_pass2, dim3(B, 10), 256, moment_params(h), h->cur.P, h->cur.O, h->sm, h->partial);
    for (int r = 0; r < 10; r++) LAUNCH(k_sum_partials_par, 9, 256, h->partial + (size_t)r * B * 9, B, 9, h->totals + r * 9);
    return SPHSM_OK;
}
static int rest_part3(sphsm_handle *h) {
    LAUNCH(k_rest_finalize2, 1, 1, h->totals, h->sm, h->scratch);
    CU(cudaGetLastError());
    h->rest_dirty = false;
    return SPHSM_OK;
}
static int moments_part(sphsm_handle *h) {
    // one partial per block and a 33-double block reduction each: keep >= 2048 particles per block (at a slab's 1M
    // particles the full 8 x SMs grid spent most of its 32 us in the reductions)
    // Slab mode: every particle is summed by the rank that integrated it last step, i.e. over that rank's owned slot range
    // as it stood BEFORE this step's exchange (migrants on their way out included, arrivals not): each particle exactly
    // once across ranks, and the sums need neither the exchange nor the sort, so they start with the step.
    DevParams d = h->dp;
    int off = 0;
    if (d.slab_on) {
        off = h->mom_begin;
        d.n = h->mom_end - h->mom_begin;
        d.slab_on = 0;
    }
    const int B = std::max(1, std::min(h->red_blocks, cdiv(std::max(d.n, 1), 2048)));
    if (h->dp.quadratic) LAUNCH(k_moments<9>, B, 256, d, h->cur.P + off, h->cur.O + off, h->sm, h->partial);
    else LAUNCH(k_moments<3>, B, 256, d, h->cur.P + off, h->cur.O + off, h->sm, h->partial);
    const int nacc = h->dp.quadratic ? 33 : 15;
    LAUNCH(k_sum_partials_par, nacc, 256, h->partial, B, nacc, h->totals);
    if (h->comm_mode == 1) LAUNCH(k_store_double, 1, 1, h->totals + nacc, h->local_error ? 1.0 : 0.0);  // see local_error
    return SPHSM_OK;
}
// the per-step moment allreduce (NCCL mode: + the error flag, copied back to the host for the next read-back to look at)
static int moment_allreduce(sphsm_handle *h) {
    const int nacc = h->dp.quadratic ? 33 : 15;
    if (h->comm_mode != 1 || h->nranks == 1) return comm_allreduce(h, nacc);
    int rc = comm_allreduce(h, nacc + 1);
    if (rc) return rc;
    CU(cudaMemcpyAsync(h->h_flag, h->totals + nacc, sizeof(double), cudaMemcpyDeviceToHost, h->launch_stream));
    CU(cudaEventRecord(h->ev_flag, h->launch_stream));
    h->flag_pending = true;
    return SPHSM_OK;
}

static int rest_moments(sphsm_handle *h) {
    int rc;
    if ((rc = rest_part1(h)) != 0 || (rc = comm_allreduce(h, 5)) != 0) return rc;
    if ((rc = rest_part2(h)) != 0 || (rc = comm_allreduce(h, 90)) != 0) return rc;
    return rest_part3(h);
}

// calculate_corrected_velocity
// the fast path's shape-matching transform of this step: moment sums (any slot order) + the single-thread solve
static int sm_transform_fast(sphsm_handle *h) {
    int rc;
    if (h->rest_dirty && (rc = rest_moments(h)) != 0) return rc;
    if ((rc = moments_part(h)) != 0 || (rc = moment_allreduce(h)) != 0) return rc;
    LAUNCH(k_sm_solve, 1, 1, h->dp, h->totals, h->sm);
    return SPHSM_OK;
}

template <bool STRICT>
static int corrected_velocity(sphsm_handle *h, bool diag, GroupTimer *gt, int store = 7) {
    const int n = h->n;
    if (n == 0) return SPHSM_OK;
    int rc;
    if (n > 1 && (store & 3)) {  // projectPositions returns early for <= 1 particle, cpp:236
        if (STRICT) {
            if ((rc = ensure_slot_of(h)) != 0) return rc;
            LAUNCH(k_sm_strict, 1, 1, h->dp, h->cur.P, h->cur.O, h->slot_of, h->sm, h->scratch);
        } else if ((rc = sm_transform_fast(h)) != 0) return rc;
    }
    if (gt) gt->end_group(KG_MOMENTS);
    const int keep_goal = n <= 1;  // projectPositions returned early: mGoalPos keeps its previous value
    if (diag || keep_goal) LAUNCH((k_goal_cvel<STRICT, true>), cdiv(n, 256), 256, h->dp, h->cur, h->sm, keep_goal, store);
    else LAUNCH((k_goal_cvel<STRICT, false>), cdiv(n, 256), 256, h->dp, h->cur, h->sm, 0, store);
    if (gt) gt->end_group(KG_GOAL);
    CU(cudaGetLastError());
    return SPHSM_OK;
}

template <bool STRICT>
static int run_stage(sphsm_handle *h, int stage) {
    const int n = h->n;
    int rc;
    if (stage < SPHSM_STAGE_FIND_NEIGHBORS || stage > SPHSM_STAGE_PROJECT_POSITIONS) return fail(h, SPHSM_ERR_INVALID, "unknown stage id");
    if (n == 0) return SPHSM_OK;
    if ((stage == 3 || stage == 4 || stage == 6) && !h->grid_valid && (rc = build_grid(h, nullptr)) != 0) return rc;
    switch (stage) {
        case SPHSM_STAGE_FIND_NEIGHBORS:
            return build_grid(h, nullptr);
        case SPHSM_STAGE_CORRECTED_VELOCITY:
            return corrected_velocity<STRICT>(h, true, nullptr);
        case SPHSM_STAGE_EXTERNAL_FORCES:  // predicted_vel only
            return corrected_velocity<STRICT>(h, true, nullptr, 4);
        case SPHSM_STAGE_PROJECT_POSITIONS:  // mGoalPos only
            return corrected_velocity<STRICT>(h, true, nullptr, 2);
        case SPHSM_STAGE_INTERMEDIATE_VELOCITY:
            LAUNCH(k_refresh_derived, cdiv(n, 256), 256, n, h->cur, 1, 0);
            LAUNCH((k_pass_a<STRICT, false, true>), cdiv(n, 128), 128, h->dp, h->cur, h->cell_start);
            break;
        case SPHSM_STAGE_DENSITY_PRESSURE:
            LAUNCH((k_pass_a<STRICT, true, false>), cdiv(n, 128), 128, h->dp, h->cur, h->cell_start);
            break;
        case SPHSM_STAGE_CELL_MODEL:
            LAUNCH(k_cell_model<STRICT>, cdiv(n, 256), 256, h->dp, h->cur);
            break;
        case SPHSM_STAGE_FORCE:
            LAUNCH(k_refresh_derived, cdiv(n, 256), 256, n, h->cur, 0, 1);
            LAUNCH((k_pass_b<STRICT, PB_FORCE_ONLY>), cdiv(n, 128), 128, h->dp, h->cur, (float4 *)nullptr, h->cell_start);
            break;
        case SPHSM_STAGE_UPDATE:
            drop_counts(h);
            LAUNCH(k_update<STRICT>, cdiv(n, 256), 256, h->dp, h->cur);
            h->grid_valid = false;
            break;
        default:
            return fail(h, SPHSM_ERR_INVALID, "unknown stage id");
    }
    CU(cudaGetLastError());
    return SPHSM_OK;
}

// the fast-path neighbour passes over the owned slot range (count = own_end - own_begin)
// Small particle sets (the reference's own ~5k-particle inputs) take one warp per particle (sphsm_pass4w.cuh).  The choice
// follows the GLOBAL particle count, so that a slab rank and the single-GPU run of the same set use the same kernels (the
// bit-level slab parity depends on identical summation order).  SPHSM_WARP_PATH=0 disables it, SPHSM_WARP_PATH_MAX moves the limit.
// The limit is a particle count because that is all the host knows; what actually decides is candidates per stencil row:
// measured with the limit lifted, a 64k LATTICE (3-6 candidates per row, most lanes idle) runs pass A / B in 56 / 108 us on
// this path against 16 / 21 us on the thread path, while the reference's meshes (45 per row) gain 10x.  Hence the second
// condition: at least 3 particles per occupied cell, estimated on the host from the positions as they were handed in
// (note_host_positions; the reference's sets have 4.9-5.1, lattices of spacing 0.9 h have 1.4).
static bool warp_path(const sphsm_handle *h) {
    static const bool off = getenv("SPHSM_WARP_PATH") && atoi(getenv("SPHSM_WARP_PATH")) == 0;
    if (off || g_pass_gen < 4) return false;
    static const int limit = getenv("SPHSM_WARP_PATH_MAX") ? atoi(getenv("SPHSM_WARP_PATH_MAX")) : WARP_PATH_MAX;
    const int n = h->dp.slab_on ? h->n_global : h->n;
    if (n > limit || h->host_cells.empty()) return false;
    return (double)n >= 3.0 * (double)h->host_cells.size();  // >= 3 particles per occupied cell: rows long enough for a warp
}
// slots [begin, end) minus the hole [hole_b, hole_e) (generation-4 kernels only)
static int launch_pass_a(sphsm_handle *h, int begin, int end, int hole_b = 0, int hole_e = 0) {
    const int count = end - begin - (hole_e - hole_b);
    if (count <= 0) return SPHSM_OK;
    DevParams d = h->dp;
    d.own_begin = begin; d.own_end = end; d.hole_begin = hole_b; d.hole_len = hole_e - hole_b;
    if (warp_path(h)) LAUNCH(k_pass_a4w, cdiv((long long)count * 32, PTW), PTW, d, h->cur, h->cell_start, count);
    else if (g_pass_gen == 5) LAUNCH(k_pass_a5, cdiv(cdiv(count, 2), PT5), PT5, d, h->d_dp, h->cur, h->cell_start, count);
    else if (g_pass_gen == 4) LAUNCH(k_pass_a4, cdiv(count, PT4), PT4, d, h->d_dp, h->cur, h->cell_start);
    else if (g_pass_gen == 2) LAUNCH(k_pass_a2, cdiv(count, PT), PT, d, h->d_dp, h->cur, h->cell_start);
    else LAUNCH(k_pass_a3, cdiv(count, PT), PT, d, h->d_dp, h->cur, h->cell_start);
    return SPHSM_OK;
}
static int launch_pass_b(sphsm_handle *h, int begin, int end, bool diag, int hole_b = 0, int hole_e = 0, bool file_counts = false) {
    uint32_t *nk = file_counts ? h->keys[0] : nullptr, *nr = file_counts ? h->keys[1] : nullptr, *ncnt = file_counts ? h->cell_count : nullptr;
    const int count = end - begin - (hole_e - hole_b);
    if (count <= 0) return SPHSM_OK;
    DevParams d = h->dp;
    d.own_begin = begin; d.own_end = end; d.hole_begin = hole_b; d.hole_len = hole_e - hole_b;
    if (warp_path(h)) {
        if (diag) LAUNCH(k_pass_b4w<true>, cdiv((long long)count * 32, PTW), PTW, d, h->cur, h->alt.P, h->cell_start, nk, nr, ncnt, count);
        else LAUNCH(k_pass_b4w<false>, cdiv((long long)count * 32, PTW), PTW, d, h->cur, h->alt.P, h->cell_start, nk, nr, ncnt, count);
    } else if (g_pass_gen == 4 || g_pass_gen == 5) {
        if (diag) LAUNCH(k_pass_b4<true>, cdiv(count, PT4), PT4, d, h->d_dp, h->cur, h->alt.P, h->cell_start, nk, nr, ncnt);
        else LAUNCH(k_pass_b4<false>, cdiv(count, PT4), PT4, d, h->d_dp, h->cur, h->alt.P, h->cell_start, nk, nr, ncnt);
    } else if (g_pass_gen == 2) {
        if (diag) LAUNCH(k_pass_b2<true>, cdiv(count, PT), PT, d, h->d_dp, h->cur, h->alt.P, h->cell_start);
        else LAUNCH(k_pass_b2<false>, cdiv(count, PT), PT, d, h->d_dp, h->cur, h->alt.P, h->cell_start);
    } else {
        if (diag) LAUNCH(k_pass_b3<true>, cdiv(count, PT), PT, d, h->d_dp, h->cur, h->alt.P, h->cell_start);
        else LAUNCH(k_pass_b3<false>, cdiv(count, PT), PT, d, h->d_dp, h->cur, h->alt.P, h->cell_start);
    }
    return SPHSM_OK;
}

// one fused step: grid, shape matching, pass A, pass B
template <bool STRICT>
static int fused_step(sphsm_handle *h) {
    const int n = h->n;
    int rc;
    if (n == 0) return SPHSM_OK;
    const bool diag = h->prm.diagnostics != 0;
    if (memcmp(&h->dp, &h->dp_uploaded, sizeof(DevParams)) != 0) {  // n or a tunable changed since the last upload
        h->dp_uploaded = h->dp;
        CU(cudaMemcpyAsync(h->d_dp, &h->dp_uploaded, sizeof(DevParams), cudaMemcpyHostToDevice, h->stream));
    }
    GroupTimer gt(h);
    if (!STRICT && n > 1) {
        // sort -> shape-matching transform (slot-order independent) -> cell table + gather fused with stage 2's map
        // the moment sums and the solve only read the not-yet-sorted arrays: they run on the side stream beside the sort
        // and rejoin before the gather applies the transform (kept in line while the per-group timers are on)
        const bool fork = !h->rest_dirty && !h->profiling;
        if (fork) {
            if (!h->dry_run) {
                CU(cudaEventRecord(h->ev_fork, h->stream));
                CU(cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
            }
            h->launch_stream = h->side_stream;
            rc = sm_transform_fast(h);
            h->launch_stream = h->stream;
            if (rc) return rc;
            if (!h->dry_run) CU(cudaEventRecord(h->ev_join, h->side_stream));
        }
        if ((rc = grid_sort(h, &gt)) != 0) return rc;
        if (fork) {
            if (!h->dry_run) CU(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
        } else if ((rc = sm_transform_fast(h)) != 0) return rc;
        gt.end_group(KG_MOMENTS);
        if ((rc = grid_finish(h, &gt, diag ? 2 : 1)) != 0) return rc;
    } else {
        if ((rc = build_grid(h, &gt)) != 0) return rc;
        if ((rc = corrected_velocity<STRICT>(h, diag, &gt)) != 0) return rc;
    }
    if (STRICT) {
        LAUNCH((k_pass_a<STRICT, true, true>), cdiv(n, 128), 128, h->dp, h->cur, h->cell_start);
        gt.end_group(KG_PASS_A);
        if (diag) LAUNCH((k_pass_b<STRICT, PB_FUSED_DIAG>), cdiv(n, 128), 128, h->dp, h->cur, h->alt.P, h->cell_start);
        else LAUNCH((k_pass_b<STRICT, PB_FUSED>), cdiv(n, 128), 128, h->dp, h->cur, h->alt.P, h->cell_start);
    } else {
        if ((rc = launch_pass_a(h, 0, n)) != 0) return rc;  // single GPU: own range = [0, n)
        gt.end_group(KG_PASS_A);
        // the counting sort of the NEXT step starts inside pass B: each thread files the key / rank / count of the position it
        // has just integrated (valid until anything else moves particles: drop_counts)
        const bool file_counts = g_pass_gen >= 4 && h->comm_mode == 0 && use_counting_sort(h);
        if ((rc = launch_pass_b(h, 0, n, diag, 0, 0, file_counts)) != 0) return rc;
        h->counts_ready = file_counts;
    }
    std::swap(h->cur.P, h->alt.P);
    gt.end_group(KG_PASS_B);
    CU(cudaGetLastError());
    gt.finish();
    h->grid_valid = false;
    h->inter_live = false;
    return SPHSM_OK;
}

// the staged step with an event pair around every stage (what the class's d_* timers report)
template <bool STRICT>
static int timed_staged_step(sphsm_handle *h) {
    for (int st = 1; st <= 7; st++) {
        CU(cudaEventRecord(h->ev[0], h->stream));
        int rc = run_stage<STRICT>(h, st);
        if (rc) return rc;
        CU(cudaEventRecord(h->ev[1], h->stream));
        CU(cudaEventSynchronize(h->ev[1]));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
        h->stage_time[st - 1] += ms * 1e-3;
    }
    h->inter_live = false;
    return SPHSM_OK;
}

static int mg_step_nccl(sphsm_handle *h);  // the slab step (below)

// Small single-GPU steps are launch-latency bound (13 dependent launches of 3-5 us for a few microseconds of work each at the
// reference's own ~5k particles; still 2-4 % of the step at 1-2M), so the fast step is captured into a CUDA graph and replayed.  A step's launch sequence and
// arguments are a function of the handle's state only (no data-dependent host decisions on one GPU): that state — buffer
// pointers of both ping-pong sets, the device parameter block, the sort / counting flags — is the graph's signature.  A
// signature seen for the second time is captured (the ping-pong gives two signatures in steady state); on a hit the host
// runs the step's bookkeeping with launches suppressed (dry_run) and launches the graph.  Any mutator that changes what a
// step would launch changes the signature, so a stale graph cannot be picked.  params.reserved[4] = 1 turns graphs off.
static const int GRAPH_MAX_N = getenv("SPHSM_GRAPH_MAX_N") ? atoi(getenv("SPHSM_GRAPH_MAX_N")) : (1 << 22);  // measured: -22 % at 5k, -3.6 % at 1M, -2.2 % at 2M particles
static bool graph_eligible(const sphsm_handle *h) {
    static const bool env_off = getenv("SPHSM_NO_GRAPH") != nullptr;
    return !env_off && h->comm_mode == 0 && !h->prm.strict && !h->profiling && !h->stage_timing && !g_sync_debug && h->prm.reserved[4] != 1 && h->n > 1 &&
           h->n <= GRAPH_MAX_N && !h->rest_dirty && memcmp(&h->dp, &h->dp_uploaded, sizeof(DevParams)) == 0;
}
static std::string step_signature(const sphsm_handle *h) {
    std::string sig;
    auto put = [&](const void *ptr, size_t bytes) { sig.append(reinterpret_cast<const char *>(ptr), bytes); };
    put(&h->cur, sizeof(Arrays));
    put(&h->alt, sizeof(Arrays));
    put(&h->dp, sizeof(DevParams));
    put(&h->prm, sizeof(sphsm_params));
    const void *ptrs[] = {h->cell_start, h->cell_count, h->tile_sums, h->keys[0], h->keys[1], h->vals[0], h->vals[1], h->big_cells, h->big_count,
                          h->sm, h->partial, h->totals, h->d_dp, h->ghist, h->tile_state, h->tile_counter, h->scratch};
    put(ptrs, sizeof(ptrs));
    const int flags[] = {h->counts_ready, h->bounds_ready, h->sorted_buf, g_pass_gen, h->red_blocks, h->sort_passes, (int)h->grid_valid};
    put(flags, sizeof(flags));
    return sig;
}
static int graph_step(sphsm_handle *h) {
    if (!graph_eligible(h)) return fused_step<false>(h);
    const std::string sig = step_signature(h);
    for (auto &gx : h->graphs) {
        if (gx.sig == sig) {
            h->dry_run = true;
            const int rc = fused_step<false>(h);
            h->dry_run = false;
            if (rc) return rc;
            CU(cudaGraphLaunch(gx.exec, h->stream));
            return SPHSM_OK;
        }
    }
    bool seen = false;
    for (auto &x : h->seen_sigs) seen = seen || x == sig;
    if (!seen) {
        if (h->seen_sigs.size() >= 16) h->seen_sigs.clear();
        h->seen_sigs.push_back(sig);
        return fused_step<false>(h);
    }
    // second sighting: capture this step (it executes when the graph is launched below)
    cudaGraph_t graph = nullptr;
    CU(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed));
    const int rc = fused_step<false>(h);
    cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
    if (rc || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        return rc ? rc : fail(h, SPHSM_ERR_CUDA, "CUDA graph capture of the step failed");
    }
    cudaGraphExec_t exec = nullptr;
    ce = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) return fail(h, SPHSM_ERR_CUDA, "cudaGraphInstantiate failed");
    if (h->graphs.size() >= 8) {
        for (auto &gx : h->graphs) cudaGraphExecDestroy(gx.exec);
        h->graphs.clear();
    }
    h->graphs.push_back({sig, exec});
    CU(cudaGraphLaunch(exec, h->stream));
    return SPHSM_OK;
}

extern "C" int sphsm_step(sphsm_handle *h, int nsteps) {
    if (!h || nsteps < 0) return SPHSM_ERR_INVALID;
    CU(cudaSetDevice(h->prm.device));
    if (h->comm_mode == 2) return fail(h, SPHSM_ERR_COMM, "a local group steps through sphsm_step_group");
    if (h->failed) return fail(h, SPHSM_ERR_COMM, "the slab group stopped after a step error; re-upload the particle set and apply the slab again");
    CU(cudaEventRecord(h->ev_step0, h->stream));
    if (h->comm_mode == 1) {
        for (int s = 0; s < nsteps; s++) {
            int rc = mg_step_nccl(h);
            if (rc) return rc;
        }
        CU(cudaEventRecord(h->ev_step1, h->stream));
        return SPHSM_OK;
    }
    for (int s = 0; s < nsteps; s++) {
        int rc;
        if (h->stage_timing) rc = h->prm.strict ? timed_staged_step<true>(h) : timed_staged_step<false>(h);
        else rc = h->prm.strict ? fused_step<true>(h) : graph_step(h);
        if (rc) return rc;
        h->total_steps++;
    }
    CU(cudaEventRecord(h->ev_step1, h->stream));
    return SPHSM_OK;
}

extern "C" int sphsm_stage(sphsm_handle *h, int stage) {
    if (!h) return SPHSM_ERR_INVALID;
    CU(cudaSetDevice(h->prm.device));
    if (stage == SPHSM_STAGE_STEP) return sphsm_step(h, 1);
    h->inter_live = true;
    return h->prm.strict ? run_stage<true>(h, stage) : run_stage<false>(h, stage);
}

extern "C" int sphsm_sync(sphsm_handle *h) {
    if (!h) return SPHSM_ERR_INVALID;
    CU(cudaSetDevice(h->prm.device));
    CU(cudaStreamSynchronize(h->h2d_stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaStreamSynchronize(h->d2h_stream));
    CU(cudaGetLastError());
    return SPHSM_OK;
}

extern "C" int sphsm_last_step_ms(sphsm_handle *h, float *ms) {
    if (!h || !ms) return SPHSM_ERR_INVALID;
    CU(cudaSetDevice(h->prm.device));
    CU(cudaEventSynchronize(h->ev_step1));
    CU(cudaEventElapsedTime(ms, h->ev_step0, h->ev_step1));
    h->last_step_ms = *ms;
    return SPHSM_OK;
}

extern "C" int sphsm_profile_step(sphsm_handle *h, int nsteps, float out_ms[SPHSM_NUM_KERNEL_GROUPS]) {
    if (!h || !out_ms || nsteps <= 0) return SPHSM_ERR_INVALID;
    CU(cudaSetDevice(h->prm.device));
    for (int g = 0; g < SPHSM_NUM_KERNEL_GROUPS; g++) { h->group_ms[g] = 0.f; h->group_launches[g] = 0; }
    h->profiling = true;
    int rc = SPHSM_OK;
    for (int s = 0; s < nsteps && rc == SPHSM_OK; s++) {
        if (h->comm_mode == 1) { rc = mg_step_nccl(h); continue; }
        rc = h->prm.strict ? fused_step<true>(h) : fused_step<false>(h);
        if (rc == SPHSM_OK) h->total_steps++;
    }
    h->profiling = false;
    for (int g = 0; g < SPHSM_NUM_KERNEL_GROUPS; g++) out_ms[g] = h->group_ms[g] / (float)nsteps;
    return rc;
}

extern "C" const char *sphsm_kernel_group_name(int g) { return (g >= 0 && g < SPHSM_NUM_KERNEL_GROUPS) ? kGroupNames[g] : ""; }

extern "C" int sphsm_num_particles(sphsm_handle *h) { return h ? h->n : SPHSM_ERR_INVALID; }
extern "C" int sphsm_num_cells(sphsm_handle *h) { return h ? h->dp.g[0] * h->dp.g[1] * h->dp.g[2] : SPHSM_ERR_INVALID; }
extern "C" int sphsm_grid_size(sphsm_handle *h, int out3[3]) {
    if (!h || !out3) return SPHSM_ERR_INVALID;
    for (int a = 0; a < 3; a++) out3[a] = h->dp.g[a];
    return SPHSM_OK;
}
extern "C" int sphsm_total_time_steps(sphsm_handle *h) { return h ? h->total_steps : SPHSM_ERR_INVALID; }

extern "C" int sphsm_enable_stage_timing(sphsm_handle *h, int on) {
    if (!h) return SPHSM_ERR_INVALID;
    h->stage_timing = on != 0;
    return SPHSM_OK;
}
extern "C" int sphsm_get_stage_times(sphsm_handle *h, double out7[7]) {
    if (!h || !out7) return SPHSM_ERR_INVALID;
    for (int k = 0; k < 7; k++) out7[k] = h->stage_time[k];
    return SPHSM_OK;
}
extern "C" int sphsm_get_launch_count(sphsm_handle *h, long long *launches) {
    if (!h || !launches) return SPHSM_ERR_INVALID;
    *launches = h->launches;
    return SPHSM_OK;
}
extern "C" int sphsm_reset_launch_count(sphsm_handle *h) {
    if (!h) return SPHSM_ERR_INVALID;
    h->launches = 0;
    return SPHSM_OK;
}

// ---------------------------------------------------------------------------------------------------
// inspection
extern "C" int sphsm_get_cells_csr(sphsm_handle *h, int *cell_start, int *indices) {
    if (!h || !cell_start || !indices) return SPHSM_ERR_INVALID;
    CU(cudaSetDevice(h->prm.device));
    int rc;
    if (!h->grid_valid && (rc = build_grid(h, nullptr)) != 0) return rc;
    const DevParams &d = h->dp;
    const int n = h->n, ncell = d.g[0] * d.g[1] * d.g[2];
    std::vector<int> cs((size_t)d.num_cells + 2);
    std::vector<int> ids(std::max(n, 1));
    CU(cudaMemcpyAsync(cs.data(), h->cell_start, cs.size() * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(ids.data(), h->cur.ID, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    // internal key -> the reference's hash x + Gx*(y + Gy*z) (cpp:142); stable counting sort keeps slot order
    std::vector<int> ref_hash(n, -1);  // slots of the limbo bucket (outside the grid) stay -1
    std::fill(cell_start, cell_start + ncell + 1, 0);
    for (int k = 0; k < d.num_cells; k++) {
        if (cs[k + 1] == cs[k]) continue;
        int c[3];
        c[d.perm[0]] = k % d.ga - 1;  // (the table has a border cell on either side of these two axes)
        c[d.perm[1]] = (k / d.ga) % d.gb - 1;
        c[d.perm[2]] = k / (d.ga * d.gb) + d.c_off;
        const int rh = c[0] + d.g[0] * (c[1] + d.g[1] * c[2]);
        for (int s = cs[k]; s < cs[k + 1]; s++) ref_hash[s] = rh;
        cell_start[rh + 1] += cs[k + 1] - cs[k];
    }
    for (int c = 0; c < ncell; c++) cell_start[c + 1] += cell_start[c];
    std::vector<int> cursor(cell_start, cell_start + ncell);
    for (int s = 0; s < n; s++)
        if (ref_hash[s] >= 0) indices[cursor[ref_hash[s]]++] = ids[s];
    // the reference's bucket order is ascending particle index
    for (int c = 0; c < ncell; c++) std::sort(indices + cell_start[c], indices + cell_start[c + 1]);
    return SPHSM_OK;
}

extern "C" int sphsm_get_neighbor_sets(sphsm_handle *h, int kind, const int *query, int n_query, int cap, int *counts, int *indices) {
    if (!h || !query || !counts || !indices || n_query < 0 || cap <= 0 || kind < 0 || kind > 3) return SPHSM_ERR_INVALID;
    CU(cudaSetDevice(h->prm.device));
    for (int i = 0; i < n_query; i++)
        if (query[i] < 0 || query[i] >= h->n) return fail(h, SPHSM_ERR_INVALID, "query index out of range");
    if (n_query == 0) return SPHSM_OK;
    int rc;
    if (!h->grid_valid && (rc = build_grid(h, nullptr)) != 0) return rc;
    if ((rc = ensure_slot_of(h)) != 0) return rc;
    if ((rc = ensure_itmp(h, (size_t)n_query * (2 + (size_t)cap))) != 0) return rc;
    int *d_query = h->d_itmp, *d_counts = h->d_itmp + n_query, *d_idx = h->d_itmp + 2 * (size_t)n_query;
    CU(cudaMemcpyAsync(d_query, query, (size_t)n_query * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    LAUNCH(k_neighbor_sets, cdiv(n_query, 64), 64, h->dp, h->cur, h->cell_start, h->slot_of, d_query, n_query, kind, cap, d_counts, d_idx);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(counts, d_counts, (size_t)n_query * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(indices, d_idx, (size_t)n_query * cap * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < n_query; i++) std::sort(indices + (size_t)i * cap, indices + (size_t)i * cap + std::min(counts[i], cap));
    return SPHSM_OK;
}

extern "C" int sphsm_get_sm_transform(sphsm_handle *h, float cm[3], float ocm[3], float xform[27]) {
    if (!h || !cm || !ocm || !xform) return SPHSM_ERR_INVALID;
    CU(cudaSetDevice(h->prm.device));
    SmState s;
    CU(cudaMemcpyAsync(&s, h->sm, sizeof s, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    memcpy(cm, s.cm, sizeof s.cm);
    memcpy(ocm, s.ocm, sizeof s.ocm);
    memcpy(xform, s.xform, sizeof s.xform);
    return SPHSM_OK;
}

// ---------------------------------------------------------------------------------------------------
// multi-GPU slab layer, host side.  NCCL is resolved at run time (dlopen), so single-GPU hosts need no NCCL.
typedef struct { char internal[128]; } nccl_unique_id;
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(nccl_unique_id *) = nullptr;
    int (*CommInitRank)(void **, int, nccl_unique_id, int) = nullptr;
    int (*CommSplit)(void *, int, int, void **, void *) = nullptr;  // optional (NCCL >= 2.18)
    int (*CommDestroy)(void *) = nullptr;
    int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
enum { NCCL_CHAR = 0, NCCL_FLOAT = 7, NCCL_DOUBLE = 8, NCCL_SUM = 0 };  // ncclDataType_t / ncclRedOp_t values (nccl.h)

static int load_nccl(sphsm_handle *h) {
    if (g_nccl.lib) return SPHSM_OK;
    const char *names[] = {getenv("SPHSM_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *lib = nullptr;
    for (const char *nm : names)
        if (nm && (lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL)) != nullptr) break;
    if (!lib) return fail(h, SPHSM_ERR_COMM, "libnccl.so.2 not found (set SPHSM_NCCL_LIB)");
    bool ok = true;
    auto sym = [&](const char *nm) { void *f = dlsym(lib, nm); if (!f) ok = false; return f; };
    g_nccl.GetUniqueId = (int (*)(nccl_unique_id *))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void **, int, nccl_unique_id, int))sym("ncclCommInitRank");
    g_nccl.CommSplit = (int (*)(void *, int, int, void **, void *))dlsym(lib, "ncclCommSplit");
    g_nccl.CommDestroy = (int (*)(void *))sym("ncclCommDestroy");
    g_nccl.Send = (int (*)(const void *, size_t, int, int, void *, cudaStream_t))sym("ncclSend");
    g_nccl.Recv = (int (*)(void *, size_t, int, int, void *, cudaStream_t))sym("ncclRecv");
    g_nccl.AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))sym("ncclAllReduce");
    g_nccl.GroupStart = (int (*)())sym("ncclGroupStart");
    g_nccl.GroupEnd = (int (*)())sym("ncclGroupEnd");
    g_nccl.GetErrorString = (const char *(*)(int))sym("ncclGetErrorString");
    if (!ok) return fail(h, SPHSM_ERR_COMM, "libnccl is missing a required entry point");
    g_nccl.lib = lib;
    g_nccl_destroy = g_nccl.CommDestroy;
    return SPHSM_OK;
}
#define NC(call)                                                                                        \
    do {                                                                                                \
        int r_ = (call);                                                                                \
        if (r_ != 0) {                                                                                  \
            std::string m_ = std::string(#call) + " failed: " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?"); \
            if (h) h->err = m_; else g_create_error = m_;                                               \
            return SPHSM_ERR_COMM;                                                                      \
        }                                                                                               \
    } while (0)

static int comm_alloc(sphsm_handle *h) {
    if (h->msg_send[0]) return SPHSM_OK;
    const int cap = h->send_cap;  // fixed at create (array room was allocated for it)
    for (int k = 0; k < 2; k++) {
        CU(cudaMalloc(&h->msg_send[k], msg_bytes(cap)));
        CU(cudaMalloc(&h->msg_recv[k], msg_bytes(cap)));
        CU(cudaMemset(h->msg_send[k], 0, 16));
        CU(cudaMemset(h->msg_recv[k], 0, 16));
    }
    CU(cudaMalloc(&h->d_err, 4 * sizeof(int)));
    CU(cudaMemset(h->d_err, 0, 4 * sizeof(int)));
    CU(cudaMalloc(&h->d_meta, 8 * sizeof(int)));
    CU(cudaMallocHost(&h->h_meta, 8 * sizeof(int)));
    return SPHSM_OK;
}

extern "C" int sphsm_comm_unique_id(void *id128) {
    sphsm_handle *h = nullptr;
    if (!id128) return SPHSM_ERR_INVALID;
    int rc = load_nccl(nullptr);
    if (rc) return rc;
    nccl_unique_id id;
    NC(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return SPHSM_OK;
}

extern "C" int sphsm_comm_init(sphsm_handle *h, int nranks, int rank, const void *id128) {
    if (!h || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return SPHSM_ERR_INVALID;
    if (h->prm.slab_axis < 0) return fail(h, SPHSM_ERR_INVALID, "create the handle with params.slab_axis = 0, 1 or 2 for multi-GPU");
    if (h->prm.strict) return fail(h, SPHSM_ERR_INVALID, "strict mode is single-GPU only");
    CU(cudaSetDevice(h->prm.device));
    int rc = load_nccl(h);
    if (rc) return rc;
    nccl_unique_id id;
    memcpy(&id, id128, sizeof id);
    NC(g_nccl.CommInitRank(&h->nccl_comm, nranks, id, rank));
    h->nccl_comm_red = h->nccl_comm;
    // SPHSM_SPLIT_COMM=1: allreduce on a second communicator.  Off by default: measured at 8 GPUs / 8M it gained nothing (the
    // allreduce queued behind exchange 1 still ends before the sort does) and the two NCCL kernels then share the SMs.
    if (g_nccl.CommSplit && getenv("SPHSM_SPLIT_COMM")) NC(g_nccl.CommSplit(h->nccl_comm, 0, rank, &h->nccl_comm_red, nullptr));
    h->comm_mode = 1; h->nranks = nranks; h->rank = rank;
    return comm_alloc(h);
}

extern "C" int sphsm_comm_init_local(sphsm_handle **hs, int nranks) {
    if (!hs || nranks < 1) return SPHSM_ERR_INVALID;
    for (int r = 0; r < nranks; r++) {
        sphsm_handle *h = hs[r];
        if (!h) return SPHSM_ERR_INVALID;
        if (h->prm.slab_axis < 0) return fail(h, SPHSM_ERR_INVALID, "create the handle with params.slab_axis = 0, 1 or 2 for multi-GPU");
        if (h->prm.strict) return fail(h, SPHSM_ERR_INVALID, "strict mode is single-GPU only");
        if (h->prm.device != hs[0]->prm.device || h->prm.capacity != hs[0]->prm.capacity)
            return fail(h, SPHSM_ERR_INVALID, "a local group shares one device and one capacity");
        CU(cudaSetDevice(h->prm.device));
        h->comm_mode = 2; h->nranks = nranks; h->rank = r;
        int rc = comm_alloc(h);
        if (rc) return rc;
    }
    return SPHSM_OK;
}

// read the plane boundaries back (one 32-byte copy + stream sync) and set n / owned range from them
static bool g_host_prof_early() { static const bool v = getenv("SPHSM_HOST_PROF") != nullptr; return v; }
static double now_us_early() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}
static int slab_meta_launch(sphsm_handle *h) {
    const DevParams &d = h->dp;
    LAUNCH(k_mg_meta, 1, 32, h->cell_start, d.num_cells, d.ga * d.gb, d.gcl, h->d_err, h->d_meta);
    CU(cudaMemcpyAsync(h->h_meta, h->d_meta, 8 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaEventRecord(h->ev_meta, h->stream));
    return SPHSM_OK;
}
static int slab_meta_read(sphsm_handle *h) {
    const double tw = g_host_prof_early() ? now_us_early() : 0.0;
    CU(cudaEventSynchronize(h->ev_meta));  // (not the stream: work queued behind the read-back keeps running)
    if (tw != 0.0) h->meta_wait_us += now_us_early() - tw;
    const int *m = h->h_meta;
    const char *what = m[5] ? "a particle crossed more than one cell plane in one step (or left the slab window)"
                       : m[6] ? "halo message overflow: raise params.reserved[0] (halo capacity)" : nullptr;
    const bool defer = h->comm_mode == 1 && h->nranks > 1 && h->slab_applied;  // inside an NCCL step: see local_error
    if (what && !defer) return fail(h, SPHSM_ERR_COMM, what);
    if (what && !h->local_error) {
        h->local_error = 1;
        h->local_error_msg = what;
    }
    if (defer && h->flag_pending) {
        CU(cudaEventSynchronize(h->ev_flag));
        h->flag_pending = false;
        if (*h->h_flag != 0.0) h->peer_error = true;
    }
    h->n = m[0];
    h->dp.n = m[0];
    h->dp.own_begin = m[1];
    h->b2 = m[2];
    h->b3 = m[3];
    h->dp.own_end = m[4];
    return SPHSM_OK;
}
static int slab_meta(sphsm_handle *h) {
    int rc = slab_meta_launch(h);
    return rc ? rc : slab_meta_read(h);
}

extern "C" int sphsm_comm_set_slab(sphsm_handle *h, int cell_lo, int cell_hi) {
    if (!h) return SPHSM_ERR_INVALID;
    if (!h->comm_mode) return fail(h, SPHSM_ERR_COMM, "sphsm_comm_init first");
    if (cell_lo < 0 || cell_hi > h->dp.gc || cell_hi - cell_lo < 1) return fail(h, SPHSM_ERR_INVALID, "slab must hold at least one cell plane of the grid");
    CU(cudaSetDevice(h->prm.device));
    DevParams &d = h->dp;
    if (!d.slab_on) h->n_global = h->n;
    d.slab_lo = cell_lo; d.slab_hi = cell_hi;
    d.c_off = cell_lo - 1; d.gcl = cell_hi - cell_lo + 2;
    d.num_cells = d.ga * d.gb * d.gcl;
    d.slab_on = 1;
    int rc;
    if ((rc = setup_grid_buffers(h)) != 0) return rc;
    // keep only the owned planes: dead entries sort into the limbo bucket and fall off the end
    if (h->n > 0) {
        d.own_begin = 0; d.own_end = h->n;
        LAUNCH(k_mg_filter, cdiv(h->n, 256), 256, h->dp, h->cur);
        h->inter_live = true;
        if ((rc = build_grid(h, nullptr)) != 0) return rc;
        if ((rc = slab_meta(h)) != 0) return rc;
    }
    h->grid_valid = false;
    h->slab_applied = true;
    h->local_error = 0; h->peer_error = false; h->failed = false; h->flag_pending = false;
    if (h->d_err) CU(cudaMemsetAsync(h->d_err, 0, 4 * sizeof(int), h->stream));
    return SPHSM_OK;
}

extern "C" int sphsm_comm_info(sphsm_handle *h, int out[8]) {
    if (!h || !out) return SPHSM_ERR_INVALID;
    out[0] = h->comm_mode; out[1] = h->nranks; out[2] = h->rank; out[3] = h->n;
    out[4] = h->dp.own_begin; out[5] = h->dp.own_end; out[6] = h->send_cap; out[7] = h->dp.slab_on;
    return SPHSM_OK;
}

extern "C" int sphsm_download_owned(sphsm_handle *h, int *ids, float *xyz, int cap, int *count) {
    if (!h || !ids || !xyz || !count || cap < 0) return SPHSM_ERR_INVALID;
    CU(cudaSetDevice(h->prm.device));
    const int first = h->dp.own_begin, nown = h->dp.own_end - h->dp.own_begin;
    *count = nown;
    if (nown > cap) return fail(h, SPHSM_ERR_CAPACITY, "output arrays smaller than the number of owned particles");
    if (nown == 0) return SPHSM_OK;
    int rc;
    if ((rc = ensure_tmp(h, (size_t)nown * 3)) != 0 || (rc = ensure_itmp(h, (size_t)nown)) != 0) return rc;
    LAUNCH(k_mg_owned_out, cdiv(nown, 256), 256, first, nown, h->cur, h->d_itmp, h->d_tmp);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(ids, h->d_itmp, (size_t)nown * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(xyz, h->d_tmp, (size_t)nown * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return SPHSM_OK;
}

// ---- collectives: NCCL (one process per GPU) --------------------------------------------------------------------
static int comm_allreduce(sphsm_handle *h, int count) {
    if (h->comm_mode != 1 || h->nranks == 1) return SPHSM_OK;  // single GPU; the local group sums between phases
    NC(g_nccl.AllReduce(h->totals, h->totals, (size_t)count, NCCL_DOUBLE, NCCL_SUM, h->nccl_comm_red, h->launch_stream));
    return SPHSM_OK;
}
static int nccl_exchange1(sphsm_handle *h) {
    const size_t bytes = msg_bytes(h->send_cap);
    NC(g_nccl.GroupStart());
    if (h->rank > 0) {
        NC(g_nccl.Send(h->msg_send[0], bytes, NCCL_CHAR, h->rank - 1, h->nccl_comm, h->stream));
        NC(g_nccl.Recv(h->msg_recv[0], bytes, NCCL_CHAR, h->rank - 1, h->nccl_comm, h->stream));
    }
    if (h->rank < h->nranks - 1) {
        NC(g_nccl.Send(h->msg_send[1], bytes, NCCL_CHAR, h->rank + 1, h->nccl_comm, h->stream));
        NC(g_nccl.Recv(h->msg_recv[1], bytes, NCCL_CHAR, h->rank + 1, h->nccl_comm, h->stream));
    }
    NC(g_nccl.GroupEnd());
    return SPHSM_OK;
}
// boundary planes' pass-A results: V = (inter_vel, m/dens) and S = (pres, Vm), contiguous slot ranges on both sides
static int nccl_exchange2(sphsm_handle *h, cudaStream_t st) {
    const int ob = h->dp.own_begin, oe = h->dp.own_end, n = h->n;
    NC(g_nccl.GroupStart());
    if (h->rank > 0) {
        NC(g_nccl.Send(h->cur.V + ob, (size_t)(h->b2 - ob) * 4, NCCL_FLOAT, h->rank - 1, h->nccl_comm, st));
        NC(g_nccl.Send(h->cur.S + ob, (size_t)(h->b2 - ob) * 2, NCCL_FLOAT, h->rank - 1, h->nccl_comm, st));
        NC(g_nccl.Recv(h->cur.V, (size_t)ob * 4, NCCL_FLOAT, h->rank - 1, h->nccl_comm, st));
        NC(g_nccl.Recv(h->cur.S, (size_t)ob * 2, NCCL_FLOAT, h->rank - 1, h->nccl_comm, st));
    }
    if (h->rank < h->nranks - 1) {
        NC(g_nccl.Send(h->cur.V + h->b3, (size_t)(oe - h->b3) * 4, NCCL_FLOAT, h->rank + 1, h->nccl_comm, st));
        NC(g_nccl.Send(h->cur.S + h->b3, (size_t)(oe - h->b3) * 2, NCCL_FLOAT, h->rank + 1, h->nccl_comm, st));
        NC(g_nccl.Recv(h->cur.V + oe, (size_t)(n - oe) * 4, NCCL_FLOAT, h->rank + 1, h->nccl_comm, st));
        NC(g_nccl.Recv(h->cur.S + oe, (size_t)(n - oe) * 2, NCCL_FLOAT, h->rank + 1, h->nccl_comm, st));
    }
    NC(g_nccl.GroupEnd());
    const int halo = ob + (n - oe);  // VN (the dense copy of V.w) of the halo slots is rebuilt locally
    if (halo > 0) {
        cudaStream_t keep = h->launch_stream;
        h->launch_stream = st;
        int rc = [&]() -> int { LAUNCH(k_mg_halo_vn, cdiv(halo, 256), 256, ob, oe, n - oe, h->cur.V, h->cur.VN); return SPHSM_OK; }();
        h->launch_stream = keep;
        if (rc) return rc;
    }
    return SPHSM_OK;
}

// ---- the slab step as phases; every phase ends in the collective named by *coll ----------------------------------------
enum { COLL_NONE = 0, COLL_EXCH1, COLL_ALLREDUCE, COLL_EXCH2, COLL_DONE, COLL_ALLREDUCE_MOMENTS };
static const int MG_PHASES = 6;

static int mg_forked_allreduce(sphsm_handle *h);
static int mg_phase(sphsm_handle *h, int phase, int *coll, int *count) {
    const bool diag = h->prm.diagnostics != 0;
    const bool has_left = h->rank > 0, has_right = h->rank < h->nranks - 1;
    int rc;
    *coll = COLL_NONE; *count = 0;
    switch (phase) {
        case 0: {  // classify + pack
            if (h->profiling) h->gt = new GroupTimer(h);
            h->mom_begin = h->dp.own_begin;
            h->mom_end = h->dp.own_end;
            h->moments_forked = false;
            if (h->comm_mode == 1 && !h->rest_dirty && !h->profiling) {
                // the moment sums, their allreduce and the solve only need last step's owned slots: they run on the side
                // stream beside the exchange, the hash and the sort, and rejoin before the gather applies the transform
                CU(cudaEventRecord(h->ev_fork, h->stream));
                CU(cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
                h->launch_stream = h->side_stream;
                // (NCCL runs one communicator's operations in issue order whatever their streams: the allreduce is issued
                // after exchange 1, in mg_forked_allreduce, so that the exchange does not queue behind the sums)
                rc = moments_part(h);
                h->launch_stream = h->stream;
                if (rc) return rc;
                h->moments_forked = true;
                h->allreduce_pending = true;
                if (h->nccl_comm_red != h->nccl_comm) {  // own communicator: nothing to queue behind
                    if ((rc = mg_forked_allreduce(h)) != 0) return rc;
                    h->allreduce_pending = false;
                }
            }
            CU(cudaMemsetAsync(h->msg_send[0], 0, 16, h->stream));
            CU(cudaMemsetAsync(h->msg_send[1], 0, 16, h->stream));
            if (h->n > 0)
                LAUNCH(k_mg_classify, cdiv(h->n, 256), 256, h->dp, h->cur, has_left ? 1 : 0, has_right ? 1 : 0, msg_view(h->msg_send[0], h->send_cap),
                       msg_view(h->msg_send[1], h->send_cap), h->send_cap, h->d_err);
            if (h->gt) h->gt->end_group(KG_OTHER);
            *coll = COLL_EXCH1;
            return SPHSM_OK;
        }
        case 1: {  // unpack arrivals, hash + sort everything, cell table, plane boundaries
            const int n0 = h->n, cap = h->send_cap;
            if (n0 + 2 * cap > h->alloc_n) return fail(h, SPHSM_ERR_CAPACITY, "capacity too small for the halo arrivals");
            LAUNCH(k_mg_unpack, cdiv(2 * cap, 256), 256, n0, h->cur, has_left ? 1 : 0, has_right ? 1 : 0, msg_view(h->msg_recv[0], cap),
                   msg_view(h->msg_recv[1], cap), cap);
            h->n = n0 + 2 * cap;
            h->dp.n = h->n;
            h->mom_n = h->n;
            if (h->gt) h->gt->end_group(KG_OTHER);
            if ((rc = grid_sort(h, h->gt)) != 0) return rc;
            if (!h->bounds_ready) LAUNCH(k_cell_bounds, cdiv(h->n + 1, 256), 256, h->keys[h->sorted_buf], h->cell_start, h->n, h->dp.num_cells);
            h->reordered = false;
            if (h->moments_forked && h->bounds_ready) {
                // the gather needs the live count only as a bound: it is queued behind the read-back with the count taken
                // from device memory, so the GPU is busy while the host waits for the plane boundaries
                if ((rc = slab_meta_launch(h)) != 0) return rc;
                CU(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
                h->moments_forked = false;
                if ((rc = grid_finish(h, h->gt, diag ? 2 : 1, true, h->d_meta)) != 0) return rc;
                h->reordered = true;
                if ((rc = slab_meta_read(h)) != 0) return rc;
            } else if ((rc = slab_meta(h)) != 0) return rc;  // n = live slots from here on
            if (!h->bounds_ready && !h->prm.reserved[1] && h->n > 0) {  // (the counting sort leaves every cell in canonical order)
                // canonical in-cell order (ascending original id) where two ranks must agree slot by slot: the halo
                // plane and the owned plane on either side of each face (arrivals were appended in atomic order)
                const int src = h->sorted_buf, ob = h->dp.own_begin, oe = h->dp.own_end;
                const bool one = h->b3 <= h->b2;  // slab of one or two planes: the ranges meet
                const int r0 = 0, c0 = one ? h->n : (has_left ? h->b2 : 0);
                const int r1 = h->b3, c1 = one ? 0 : (has_right ? h->n - h->b3 : 0);
                (void)ob; (void)oe;
                if (c0 > 0) LAUNCH(k_cell_order_fix, cdiv(c0, 128), 128, h->keys[src], h->vals[src], h->cur.ID, h->n, (uint32_t)h->dp.num_cells, r0, c0);
                if (c1 > 0) LAUNCH(k_cell_order_fix, cdiv(c1, 128), 128, h->keys[src], h->vals[src], h->cur.ID, h->n, (uint32_t)h->dp.num_cells, r1, c1);
            }
            if (h->gt) h->gt->end_group(KG_GRID);
            if (h->rest_dirty) {
                if ((rc = rest_part1(h)) != 0) return rc;
                *coll = COLL_ALLREDUCE; *count = 5;
            }
            return SPHSM_OK;
        }
        case 2:
            if (h->rest_dirty) {
                if ((rc = rest_part2(h)) != 0) return rc;
                *coll = COLL_ALLREDUCE; *count = 90;
            }
            return SPHSM_OK;
        case 3:
            if (h->moments_forked || h->reordered) return SPHSM_OK;
            if (h->rest_dirty && (rc = rest_part3(h)) != 0) return rc;
            if ((rc = moments_part(h)) != 0) return rc;
            *coll = COLL_ALLREDUCE_MOMENTS; *count = h->dp.quadratic ? 33 : 15;
            return SPHSM_OK;
        case 4: {  // solve, gather + stage 2, pass A
            if (memcmp(&h->dp, &h->dp_uploaded, sizeof(DevParams)) != 0) {
                h->dp_uploaded = h->dp;
                CU(cudaMemcpyAsync(h->d_dp, &h->dp_uploaded, sizeof(DevParams), cudaMemcpyHostToDevice, h->stream));
            }
            if (h->reordered) {
            } else if (h->moments_forked) CU(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
            else LAUNCH(k_sm_solve, 1, 1, h->dp, h->totals, h->sm);
            h->moments_forked = false;
            h->mom_n = 0;
            if (h->gt) h->gt->end_group(KG_MOMENTS);
            if (h->n > 0 && !h->reordered && (rc = grid_finish(h, h->gt, diag ? 2 : 1, true)) != 0) return rc;
            const int ob = h->dp.own_begin, oe = h->dp.own_end;
            // NCCL mode with at least three owned planes: pass A on the two boundary planes first, their V / S records travel
            // on the side stream while the interior planes are computed here (and pass B's interior after them)
            h->split = h->comm_mode == 1 && h->nranks > 1 && !h->profiling && h->b2 < h->b3 && g_pass_gen >= 4;
            if (h->split) {
                // side stream (high priority): pass A on the two boundary planes -> exchange 2 -> pass B on them;
                // main stream: pass A, then pass B on the interior planes.  Cross dependencies: pass B's interior reads the
                // boundary planes' pass-A records (ev_bnd), pass B's boundary reads the interior's (ev_int).
                CU(cudaEventRecord(h->ev_fork, h->stream));
                CU(cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
                h->launch_stream = h->side_stream;
                rc = launch_pass_a(h, ob, oe, h->b2, h->b3);
                h->launch_stream = h->stream;
                if (rc) return rc;
                CU(cudaEventRecord(h->ev_bnd, h->side_stream));
                if ((rc = launch_pass_a(h, h->b2, h->b3)) != 0) return rc;  // (queued before the NCCL calls: they take host time)
                CU(cudaEventRecord(h->ev_int, h->stream));
                CU(cudaStreamWaitEvent(h->stream, h->ev_bnd, 0));
                if ((rc = launch_pass_b(h, h->b2, h->b3, diag)) != 0) return rc;
                if (g_host_prof_early() && h->pev[4]) CU(cudaEventRecord(h->pev[4], h->side_stream));
                rc = nccl_exchange2(h, h->side_stream);
                if (g_host_prof_early() && h->pev[5]) CU(cudaEventRecord(h->pev[5], h->side_stream));
                return rc;
            }
            if ((rc = launch_pass_a(h, ob, oe)) != 0) return rc;
            if (h->gt) h->gt->end_group(KG_PASS_A);
            *coll = COLL_EXCH2;
            return SPHSM_OK;
        }
        case 5: {  // pass B on the owned slots
            if (h->gt) h->gt->end_group(KG_OTHER);  // exchange 2
            const int ob = h->dp.own_begin, oe = h->dp.own_end;
            if (h->split) {  // interior planes need no halo record; the boundary planes wait for exchange 2 (stream order)
                CU(cudaStreamWaitEvent(h->side_stream, h->ev_int, 0));  // (pass B's interior was queued in phase 4)
                h->launch_stream = h->side_stream;
                rc = launch_pass_b(h, ob, oe, diag, h->b2, h->b3);
                h->launch_stream = h->stream;
                if (rc) return rc;
                CU(cudaEventRecord(h->ev_join, h->side_stream));
                CU(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
            } else if ((rc = launch_pass_b(h, ob, oe, diag)) != 0) return rc;
            std::swap(h->cur.P, h->alt.P);
            if (h->gt) {
                h->gt->end_group(KG_PASS_B);
                h->gt->finish();
                delete h->gt;
                h->gt = nullptr;
            }
            CU(cudaGetLastError());
            h->grid_valid = false;
            h->inter_live = false;
            h->total_steps++;
            *coll = COLL_DONE;
            return SPHSM_OK;
        }
    }
    return fail(h, SPHSM_ERR_INVALID, "bad phase");
}

static int mg_check(sphsm_handle *h) {
    if (!h->slab_applied) return fail(h, SPHSM_ERR_COMM, "sphsm_comm_set_slab must be applied after the particle set is uploaded");
    if (h->stage_timing) return fail(h, SPHSM_ERR_INVALID, "stage timing is single-GPU only");
    return SPHSM_OK;
}

// second half of the forked moment chain: allreduce + solve on the side stream, then the join event
static int mg_forked_allreduce(sphsm_handle *h) {
    h->launch_stream = h->side_stream;
    int rc = moment_allreduce(h);
    if (!rc) rc = [&]() -> int { LAUNCH(k_sm_solve, 1, 1, h->dp, h->totals, h->sm); return SPHSM_OK; }();
    h->launch_stream = h->stream;
    if (rc) return rc;
    CU(cudaEventRecord(h->ev_join, h->side_stream));
    return SPHSM_OK;
}

// SPHSM_HOST_PROF=1: host-side time of the slab step per phase (kernel launches / NCCL calls / the read-back wait), printed
// by rank 0 every 64 steps — tells a launch-bound step from a device-bound one
static const bool g_host_prof = getenv("SPHSM_HOST_PROF") != nullptr;
static double now_us() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}
static int mg_step_nccl(sphsm_handle *h) {
    int rc, coll, count;
    if ((rc = mg_check(h)) != 0) return rc;
    static double acc[MG_PHASES + 1][2];
    static int steps_seen = 0;
    for (int ph = 0; ph < MG_PHASES; ph++) {
        const double t0 = g_host_prof ? now_us() : 0.0;
        if ((rc = mg_phase(h, ph, &coll, &count)) != 0) return rc;
        const double t1 = g_host_prof ? now_us() : 0.0;
        if (coll == COLL_EXCH1) {
            if (g_host_prof) {
                for (int k = 0; k < 8; k++)
                    if (!h->pev[k]) CU(cudaEventCreate(&h->pev[k]));
                CU(cudaEventRecord(h->pev[0], h->stream));
            }
            rc = nccl_exchange1(h);
            if (g_host_prof) CU(cudaEventRecord(h->pev[1], h->stream));
            if (g_host_prof && h->moments_forked) CU(cudaEventRecord(h->pev[2], h->side_stream));
            if (!rc && h->moments_forked && h->allreduce_pending) rc = mg_forked_allreduce(h);
            h->allreduce_pending = false;
            if (g_host_prof && h->moments_forked) CU(cudaEventRecord(h->pev[3], h->side_stream));
        }
        else if (coll == COLL_ALLREDUCE) rc = comm_allreduce(h, count);
        else if (coll == COLL_ALLREDUCE_MOMENTS) rc = moment_allreduce(h);
        else if (coll == COLL_EXCH2) rc = nccl_exchange2(h, h->stream);
        if (rc) return rc;
        if (g_host_prof) {
            acc[ph][0] += t1 - t0;
            acc[ph][1] += now_us() - t1;
        }
    }
    if (h->peer_error) {  // some rank (maybe this one) failed in the previous step: every rank stops here
        h->failed = true;
        return fail(h, SPHSM_ERR_COMM, h->local_error ? h->local_error_msg.c_str()
                                                      : "another rank of the slab group reported a step error (its sphsm_last_error has the cause)");
    }
    if (g_host_prof && h->pev[5] && h->split) {  // device-side durations of the three collectives (this serialises the steps)
        CU(cudaStreamSynchronize(h->stream));
        CU(cudaStreamSynchronize(h->side_stream));
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->pev[0], h->pev[1]) == cudaSuccess) h->pacc[0] += ms * 1e3;
        if (cudaEventElapsedTime(&ms, h->pev[2], h->pev[3]) == cudaSuccess) h->pacc[1] += ms * 1e3;
        if (cudaEventElapsedTime(&ms, h->pev[4], h->pev[5]) == cudaSuccess) h->pacc[2] += ms * 1e3;
        if (cudaEventElapsedTime(&ms, h->pev[0], h->pev[5]) == cudaSuccess) h->pacc[3] += ms * 1e3;
        cudaGetLastError();
    }
    if (g_host_prof && ++steps_seen % 64 == 0) {
        fprintf(stderr, "[sphsm dev prof rank %d, cumulative us over %d steps] exch1 %.0f allreduce+solve %.0f exch2+vn %.0f exch1-start..exch2-end %.0f\n",
                h->rank, steps_seen, h->pacc[0], h->pacc[1], h->pacc[2], h->pacc[3]);
    }
    if (g_host_prof && steps_seen % 64 == 0 && h->rank == 0) {
        fprintf(stderr, "[sphsm host prof, us/step over %d steps] ", steps_seen);
        double tot = 0;
        for (int ph = 0; ph < MG_PHASES; ph++) {
            fprintf(stderr, "ph%d %.1f+%.1f  ", ph, acc[ph][0] / steps_seen, acc[ph][1] / steps_seen);
            tot += acc[ph][0] + acc[ph][1];
        }
        fprintf(stderr, "total %.1f (read-back wait %.1f)\n", tot / steps_seen, h->meta_wait_us / steps_seen);
    }
    return SPHSM_OK;
}

// virtual ranks: the same phases in lockstep over handles that share one device; collectives are device copies
extern "C" int sphsm_step_group(sphsm_handle **hs, int nranks, int nsteps) {
    if (!hs || nranks < 1 || nsteps < 0) return SPHSM_ERR_INVALID;
    for (int r = 0; r < nranks; r++) {
        sphsm_handle *h = hs[r];
        if (!h || h->comm_mode != 2 || h->nranks != nranks || h->rank != r) return fail(h, SPHSM_ERR_COMM, "not the local group made by sphsm_comm_init_local");
        int rc = mg_check(h);
        if (rc) return rc;
    }
    sphsm_handle *h = hs[0];
    CU(cudaSetDevice(h->prm.device));
    std::vector<int> coll(nranks), count(nranks);
    std::vector<double> sum(128), part(128);
    for (int s = 0; s < nsteps; s++) {
        for (int ph = 0; ph < MG_PHASES; ph++) {
            for (int r = 0; r < nranks; r++) {
                int rc = mg_phase(hs[r], ph, &coll[r], &count[r]);
                if (rc) return rc;
                if (coll[r] != coll[0] || count[r] != count[0]) return fail(hs[r], SPHSM_ERR_COMM, "ranks disagree on the phase program");
            }
            for (int r = 0; r < nranks; r++) CU(cudaStreamSynchronize(hs[r]->stream));
            if (coll[0] == COLL_EXCH1) {
                const size_t bytes = msg_bytes(h->send_cap);
                for (int r = 0; r < nranks; r++) {
                    if (r > 0) CU(cudaMemcpy(hs[r]->msg_recv[0], hs[r - 1]->msg_send[1], bytes, cudaMemcpyDeviceToDevice));
                    if (r < nranks - 1) CU(cudaMemcpy(hs[r]->msg_recv[1], hs[r + 1]->msg_send[0], bytes, cudaMemcpyDeviceToDevice));
                }
            } else if (coll[0] == COLL_ALLREDUCE || coll[0] == COLL_ALLREDUCE_MOMENTS) {
                const int c = count[0];
                std::fill(sum.begin(), sum.end(), 0.0);
                for (int r = 0; r < nranks; r++) {
                    CU(cudaMemcpy(part.data(), hs[r]->totals, c * sizeof(double), cudaMemcpyDeviceToHost));
                    for (int k = 0; k < c; k++) sum[k] += part[k];
                }
                for (int r = 0; r < nranks; r++) CU(cudaMemcpy(hs[r]->totals, sum.data(), c * sizeof(double), cudaMemcpyHostToDevice));
            } else if (coll[0] == COLL_EXCH2) {
                for (int r = 0; r + 1 < nranks; r++) {  // face between rank r (left) and rank r + 1 (right)
                    sphsm_handle *a = hs[r], *b = hs[r + 1];
                    const int na = a->dp.own_end - a->b3, nb_halo = b->dp.own_begin;       // a's last owned plane -> b's left halo
                    const int nb = b->b2 - b->dp.own_begin, na_halo = a->n - a->dp.own_end;  // b's first owned plane -> a's right halo
                    if (na != nb_halo || nb != na_halo) return fail(a, SPHSM_ERR_COMM, "boundary plane populations differ across a slab face");
                    CU(cudaMemcpy(b->cur.V, a->cur.V + a->b3, (size_t)na * sizeof(float4), cudaMemcpyDeviceToDevice));
                    CU(cudaMemcpy(b->cur.S, a->cur.S + a->b3, (size_t)na * sizeof(float2), cudaMemcpyDeviceToDevice));
                    CU(cudaMemcpy(b->cur.VN, a->cur.VN + a->b3, (size_t)na * sizeof(float), cudaMemcpyDeviceToDevice));
                    CU(cudaMemcpy(a->cur.V + a->dp.own_end, b->cur.V + b->dp.own_begin, (size_t)nb * sizeof(float4), cudaMemcpyDeviceToDevice));
                    CU(cudaMemcpy(a->cur.S + a->dp.own_end, b->cur.S + b->dp.own_begin, (size_t)nb * sizeof(float2), cudaMemcpyDeviceToDevice));
                    CU(cudaMemcpy(a->cur.VN + a->dp.own_end, b->cur.VN + b->dp.own_begin, (size_t)nb * sizeof(float), cudaMemcpyDeviceToDevice));
                }
            }
        }
    }
    return SPHSM_OK;
}

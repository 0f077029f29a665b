// api.cu -- libcalipso_b200.so: kernels (one CTA per problem instance) + the C ABI of include/calipso_b200.h.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo (see calipso_b200/build.py).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/calipso_b200.h"
#include "device_newton.h"
#include "host_setup.h"

using namespace cb200;

// phase counters: shared-memory accumulators -> the instance's slots (only when profiling is on)
#define PROF_FLUSH(ptr)                                                                  \
    do {                                                                                 \
        __syncthreads();                                                                 \
        if ((ptr) && threadIdx.x < PROF_COUNT) (ptr)[threadIdx.x] += cb_prof[threadIdx.x]; \
    } while (0)

static_assert(sizeof(cb200_options) == sizeof(Options), "cb200_options must mirror cb200::Options");
static_assert(sizeof(Phase) == 6 * sizeof(int) && sizeof(ChainDesc) == 32, "shared-memory caches assume these layouts");
static_assert((int)CB200_S_COUNT == (int)S_COUNT && (int)CB200_I_COUNT == (int)I_COUNT, "slot enums out of sync");

// ---------------------------------------------------------------------------------------------------- batch view
struct Batch {
    int count, F;
    long long ksize;
    double *w, *cand, *step, *res, *err, *corr, *tmp;
    double *grad, *gyx, *hzx, *g, *h, *Wv, *Gv, *Cv, *Wf, *Gr, *prod, *bgrad, *lambda;
    double *panels, *D, *Dinv, *kx, *Lcsr, *xs, *rs, *xp, *mgrad, *q, *g0, *h0, *filter, *krylov, *scal;
    double *Aval, *rhs;
    int *istat;
    long long *prof;
    int scratch_doubles;
    const int *order;      // k_lq_step: CTA -> instance (null: identity); a scheduling hint, see cb200_lq_set_order
    __device__ __forceinline__ Inst inst(const DevProblem &P, int b) const
    {
        Inst I;
        const long long t = P.total, n = P.n, m = P.m, p = P.p, N = P.N;
        I.w = w + b * t; I.cand = cand + b * t; I.step = step + b * t; I.res = res + b * t; I.err = err + b * t;
        I.corr = corr + b * t; I.tmp = tmp + b * t;
        I.grad = grad + b * n; I.gyx = gyx + b * n; I.hzx = hzx + b * n;
        I.g = g + b * m; I.h = h + b * p;
        I.Wv = Wv + b * (long long)P.nnzW; I.Gv = Gv + b * (long long)P.nnzG; I.Cv = Cv + b * (long long)P.nnzC;
        I.Wf = Wf + b * (long long)P.nnzWf; I.Gr = Gr + b * (long long)P.nnzG;
        I.prod = prod + b * p; I.bgrad = bgrad + b * p; I.lambda = lambda + b * m;
        I.panels = panels + b * P.panel_total; I.D = D + b * N; I.Dinv = Dinv + b * N;
        I.kx = kx + b * P.kx_total;
        I.Lcsr = Lcsr + b * P.lcsr_total;
        I.prof = prof ? prof + b * (long long)PROF_COUNT : nullptr;
        I.xs = xs + b * N; I.rs = rs + b * N; I.xp = xp + b * N; I.mgrad = mgrad + b * N;
        I.q = q + b * n; I.g0 = g0 + b * m; I.h0 = h0 + b * p;
        I.filter = filter + b * 4LL * F;
        I.krylov = krylov ? krylov + b * ksize : nullptr;
        I.scal = scal + b * (long long)S_COUNT;
        I.istat = istat + b * (long long)I_COUNT;
        return I;
    }
};

#ifndef CB_THREADS
#define CB_THREADS 256
#endif
#define CB_THREADS_WIDE 512
#ifndef CB_MIN_CTAS
#define CB_MIN_CTAS 3   // registers capped at 85 per thread: three CTAs per SM
#endif
#ifndef CB_THREADS_NARROW
#define CB_THREADS_NARROW 192   // the four-CTA plan (symbolic.cpp, analyze_auto): 4 x 192 threads, the same 85-register cap
#endif
#ifndef CB_NARROW_CTAS
#define CB_NARROW_CTAS 4
#endif
// resident CTAs the heavy kernels are compiled for, by threads per CTA
#define CB_CTAS_FOR(T) ((T) == CB_THREADS ? CB_MIN_CTAS : ((T) == CB_THREADS_NARROW ? CB_NARROW_CTAS : 1))
#define CTX_SETUP                                                                                          \
    if (threadIdx.x < 32) cb_prof[threadIdx.x] = 0;                                                        \
    {                                                                                                      \
        const int nch = P.nbig <= CB_MAX_CHAIN ? P.nbig : 0;                                               \
        if (threadIdx.x == 0) cb_chain_n = nch;                                                            \
        for (int e = threadIdx.x; e < 2 * nch; e += blockDim.x)                                            \
            cb_chain[e] = reinterpret_cast<const int4 *>(P.bdesc)[e];                                      \
        const int nph = P.nphases <= CB_MAX_PHASES ? P.nphases : 0;                                        \
        if (threadIdx.x == 0) cb_phase_n = nph;                                                            \
        for (int e = threadIdx.x; e < 6 * nph; e += blockDim.x)                                            \
            cb_phase[e] = reinterpret_cast<const int *>(P.phases)[e];                                      \
        for (int e = threadIdx.x; e < nph + 1 && nph > 0; e += blockDim.x) cb_pphase[e] = P.pphase_ptr[e]; \
    }                                                                                                      \
    if (threadIdx.x == 0) {                                                                                \
        mbar_init(&cb_bars[0], 1);                                                                         \
        mbar_init(&cb_bars[1], 1);                                                                         \
        cb_bar_uses = 0;                                                                                   \
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");                               \
    }                                                                                                      \
    __syncthreads();                                                                                       \
    Ctx ctx{(int)threadIdx.x, (int)blockDim.x, 0, cb_red, B.scratch_doubles > 0 ? cb_dyn_smem : nullptr, cb_bars, &cb_bar_uses};
#define KERNEL_PROLOGUE                                   \
    CTX_SETUP                                             \
    const int b = blockIdx.x;                             \
    if (b >= B.count) return;                             \
    Inst I = B.inst(P, b);

__global__ void __launch_bounds__(CB_THREADS, CB_MIN_CTAS) k_expand(const __grid_constant__ DevProblem P, const __grid_constant__ Batch B)
{
    KERNEL_PROLOGUE
    expand_values(ctx, P, I);
}

__global__ void __launch_bounds__(CB_THREADS, CB_MIN_CTAS) k_cone(const __grid_constant__ DevProblem P, const __grid_constant__ Batch B, int flags, int at_candidate)
{
    KERNEL_PROLOGUE
    cone_eval(ctx, P, I, at_candidate ? I.cand : I.w, flags & 1, (flags >> 1) & 1, (flags >> 2) & 1);
}

__global__ void __launch_bounds__(CB_THREADS, CB_MIN_CTAS) k_residual(const __grid_constant__ DevProblem P, const __grid_constant__ Batch B)
{
    KERNEL_PROLOGUE
    residual_eval(ctx, P, I);
    double th = constraint_violation(ctx, P, I, I.w);
    if (ctx.tid == 0) I.scal[S_THETA] = th;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, CB_CTAS_FOR(THREADS)) k_search_direction(const __grid_constant__ DevProblem P, const __grid_constant__ Batch B, Options o)
{
    KERNEL_PROLOGUE
    int st = search_direction(ctx, P, I, o);
    if (ctx.tid == 0) I.istat[I_STATUS] = st;
    PROF_FLUSH(I.prof);
}

__global__ void __launch_bounds__(CB_THREADS, CB_MIN_CTAS) k_cone_search(const __grid_constant__ DevProblem P, const __grid_constant__ Batch B, Options o)
{
    KERNEL_PROLOGUE
    int st = cone_search(ctx, P, I, o);
    // candidate x, r with the cone step size (solve.jl:224-229)
    double a = I.scal[S_STEP_SIZE];
    PAR_FOR(i, P.n + P.m) I.cand[i] = I.w[i] - a * I.step[i];
    if (ctx.tid == 0 && st != ST_OK) I.istat[I_STATUS] = st;
}

__global__ void __launch_bounds__(CB_THREADS, CB_MIN_CTAS) k_apply_step(const __grid_constant__ DevProblem P, const __grid_constant__ Batch B)
{
    KERNEL_PROLOGUE
    const int n = P.n, m = P.m, p = P.p, N = P.N;
    const double a = I.scal[S_STEP_SIZE];
    double *w = I.w, *c = I.cand;
    // candidate primals for the (possibly reduced) step size, then the update of solve.jl:309-326
    PAR_FOR(i, N) w[i] = w[i] - a * I.step[i];
    PAR_FOR(i, m + p) w[N + i] = w[N + i] - a * I.step[N + i];
    PAR_FOR(i, p) w[N + m + p + i] = c[N + m + p + i];
    ctx.sync();
    cone_eval(ctx, P, I, w, 0, 0, 1);
    double ev = scope_max(ctx, m, [&](int i) { return fabs(I.g[i]); });
    double cv = scope_max(ctx, p, [&](int i) { return fabs(I.prod[i]); });
    if (ctx.tid == 0) { I.scal[S_EQUALITY_VIOLATION] = ev; I.scal[S_CONE_PRODUCT_VIOLATION] = cv; }
    (void)n;
}

// filter line search over uploaded candidates (solve.jl:224-306), see cb200_filter_search
__global__ void __launch_bounds__(CB_THREADS, CB_MIN_CTAS) k_filter_search(const __grid_constant__ DevProblem P, const __grid_constant__ Batch B, Options o,
                                                                            int first, int count, const double *fc, const double *gc,
                                                                            const double *hc, int *accepted)
{
    KERNEL_PROLOGUE
    const int a = filter_search(ctx, P, I, o, first, count, fc + (long long)b * count, gc + (long long)b * count * P.m,
                                hc + (long long)b * count * P.p);
    if (ctx.tid == 0) accepted[b] = a;
}

__global__ void __launch_bounds__(CB_THREADS, CB_MIN_CTAS) k_filter_reset(const __grid_constant__ DevProblem P, const __grid_constant__ Batch B, Options o)
{
    KERNEL_PROLOGUE
    filter_reset(ctx, I, o);
}

__global__ void __launch_bounds__(CB_THREADS, CB_MIN_CTAS) k_jtimes(const __grid_constant__ DevProblem P, const __grid_constant__ Batch B)
{   // tmp <- J * err  (err used as the input vector)
    KERNEL_PROLOGUE
    jacobian_times(ctx, P, I, I.err, I.tmp);
}

__global__ void __launch_bounds__(CB_THREADS, CB_MIN_CTAS) k_lq_evaluate(const __grid_constant__ DevProblem P, const __grid_constant__ Batch B, int flags, int at_candidate)
{
    KERNEL_PROLOGUE
    lq_evaluate(ctx, P, I, at_candidate ? I.cand : I.w, flags);
}

__global__ void __launch_bounds__(CB_THREADS, CB_MIN_CTAS) k_lq_begin(const __grid_constant__ DevProblem P, const __grid_constant__ Batch B, Options o, int warmstart)
{
    KERNEL_PROLOGUE
    solve_begin_lq(ctx, P, I, o, warmstart);
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, CB_CTAS_FOR(THREADS)) k_lq_step(const __grid_constant__ DevProblem P, const __grid_constant__ Batch B, Options o, int iterations)
{
    CTX_SETUP
    if ((int)blockIdx.x >= B.count) return;
    const int b = B.order ? B.order[blockIdx.x] : (int)blockIdx.x;      // (CTAs are dispatched in blockIdx order)
    Inst I = B.inst(P, b);
    // `iterations` Newton iterations of this instance inside one launch: instances are independent, so nothing forces
    // them into lock-step -- a CTA whose instance converges (or fails) leaves at once and the next instance of the batch
    // takes its place on the SM instead of waiting for the slowest instance of every iteration
    for (int k = 0; k < iterations; k++) {
        solve_step_lq(ctx, P, I, o);             // (ends with a CTA barrier: the state below is uniform)
        if (I.istat[I_CONVERGED] != 0) break;
    }
    PROF_FLUSH(I.prof);
}

// LinearSolver seam: factor the generic matrix / solve in place.  These two are the "KKT LDL^T solve" whose HBM
// roofline bench.py reports (B_unit).
template <int THREADS>
__global__ void __launch_bounds__(THREADS, CB_CTAS_FOR(THREADS)) k_ldl_factor(const __grid_constant__ DevProblem P, const __grid_constant__ Batch B, int assemble_generic)
{
    CTX_SETUP
    const int b = blockIdx.x;
    if (b >= B.count) return;
    double *pan = B.panels + b * P.panel_total;
    double *D = B.D + b * (long long)P.N, *Dinv = B.Dinv + b * (long long)P.N;
    int *istat = B.istat + b * (long long)I_COUNT;
    (void)assemble_generic;
    const double *Ax = B.Aval + b * (long long)P.nnzA;
    ldl_factor(ctx, P, pan, D, Dinv, KSrc{Ax, Ax, Ax, Ax}, B.Lcsr + b * P.lcsr_total, istat, B.prof ? B.prof + b * (long long)PROF_COUNT : nullptr);
    PROF_FLUSH(B.prof ? B.prof + b * (long long)PROF_COUNT : nullptr);
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, CB_CTAS_FOR(THREADS)) k_ldl_solve(const __grid_constant__ DevProblem P, const __grid_constant__ Batch B)
{
    CTX_SETUP
    const int b = blockIdx.x;
    if (b >= B.count) return;
    double *rhs = B.rhs + b * (long long)P.N;
    ldl_solve(ctx, P, B.panels + b * P.panel_total, B.D + b * (long long)P.N, B.Dinv + b * (long long)P.N,
              B.kx + b * P.kx_total, B.Lcsr + b * P.lcsr_total, rhs, rhs, B.xp + b * (long long)P.N, B.istat + b * (long long)I_COUNT,
              B.prof ? B.prof + b * (long long)PROF_COUNT : nullptr);
    PROF_FLUSH(B.prof ? B.prof + b * (long long)PROF_COUNT : nullptr);
}

// KKT path: assemble + factor with the current regularisation (no inertia loop) -- used by the roofline bench
template <int THREADS>
__global__ void __launch_bounds__(THREADS, CB_CTAS_FOR(THREADS)) k_kkt_factor_solve(const __grid_constant__ DevProblem P, const __grid_constant__ Batch B, int nsolves)
{
    KERNEL_PROLOGUE
    ProfTimer pt{I.prof, 0};
    pt.start();
    kkt_entries(ctx, P, I);
    pt.stop(PROF_ASSEMBLE);
    ldl_factor(ctx, P, I.panels, I.D, I.Dinv, KSrc{I.Wv, I.Gr, I.Cv, I.kx}, I.Lcsr, I.istat, I.prof);
    for (int k = 0; k < nsolves; k++) direction_symmetric(ctx, P, I, I.res, I.step);
    PROF_FLUSH(I.prof);
}

// differentiate!: one factorisation, num_parameters reduced solves with recovery, sign flip (differentiate.jl:13-57).  The
// factorisation is k_kkt_factor_solve with nsolves = 0; the right-hand sides are independent, so every (instance,
// parameter) pair gets its own CTA -- a single problem with many parameters fills the GPU instead of running its columns
// one after the other on one SM -- with private reduced-rhs / solution / scratch vectors (work: [pairs][3][N]).
template <int THREADS>
__global__ void __launch_bounds__(THREADS, CB_CTAS_FOR(THREADS)) k_differentiate(const __grid_constant__ DevProblem P, const __grid_constant__ Batch B, int nparam,
                                                                         const double *H, double *S, double *work)
{
    CTX_SETUP
    const int pair = blockIdx.x;
    if (pair >= B.count * nparam) return;
    const int b = pair / nparam;
    Inst I = B.inst(P, b);
    I.rs = work + (long long)pair * 3 * P.N;
    I.xs = I.rs + P.N;
    I.xp = I.xs + P.N;
    I.istat = nullptr;                       // (the per-instance counters are not shared between the CTAs of an instance)
    I.prof = nullptr;
    const double *rhs = H + (long long)pair * P.total;
    double *out = S + (long long)pair * P.total;
    reduced_rhs(ctx, P, I, rhs, I.rs);
    ldl_solve(ctx, P, I.panels, I.D, I.Dinv, I.kx, I.Lcsr, I.rs, I.xs, I.xp, nullptr, nullptr);
    recover_step(ctx, P, I, rhs, I.xs, out);
    PAR_FOR(k, P.total) out[k] = -1.0 * out[k];
}

// evaluate!'s cache scatter: grid = (chunks of the pattern, instances); coalesced index reads and value
// writes, gathered cache reads
__global__ void __launch_bounds__(256) k_scatter(int nnz, int ncaches, const int *__restrict__ idx,
                                                 const double *__restrict__ caches, long long cache_total,
                                                 double *__restrict__ out, long long out_len, int first_instance)
{
    Ctx ctx{(int)threadIdx.x, (int)blockDim.x, 0, nullptr, nullptr, nullptr, nullptr};
    const long long b = first_instance + blockIdx.y;
    scatter_caches(ctx, nnz, ncaches, idx, caches + b * cache_total, out + b * out_len, blockIdx.x * blockDim.x,
                   gridDim.x * blockDim.x);
}

// stage-level scatter of the trajectory-optimisation front end: grid = (chunks of the output, instances);
// coalesced writes, gathered cache reads in program order
__global__ void __launch_bounds__(256) k_stage_gather(int nout, int accumulate, const int *__restrict__ ptr, const int *__restrict__ src,
                                                      const double *__restrict__ caches, long long cache_total,
                                                      double *__restrict__ out, long long out_len, int first_instance)
{
    Ctx ctx{(int)threadIdx.x, (int)blockDim.x, 0, nullptr, nullptr, nullptr, nullptr};
    const long long b = first_instance + blockIdx.y;
    stage_gather(ctx, nout, accumulate, ptr, src, caches + b * cache_total, out + b * out_len, blockIdx.x * blockDim.x,
                 gridDim.x * blockDim.x);
}

__global__ void k_count_states(Batch B, long long *counts)
{
    __shared__ int c[4];
    if (threadIdx.x < 4) c[threadIdx.x] = 0;
    __syncthreads();
    for (int b = threadIdx.x; b < B.count; b += blockDim.x) atomicAdd(&c[B.istat[b * (long long)I_COUNT + I_CONVERGED] & 3], 1);
    __syncthreads();
    if (threadIdx.x < 4) counts[threadIdx.x] = c[threadIdx.x];
}

// ---------------------------------------------------------------------------------------------------- host side
static thread_local std::string g_err;
static int fail(const std::string &m) { g_err = m; return -1; }
#define CUDA_OK(x)                                                                                      \
    do {                                                                                                \
        cudaError_t e_ = (x);                                                                           \
        if (e_ != cudaSuccess) return fail(std::string(#x) + ": " + cudaGetErrorString(e_));            \
    } while (0)

struct ArrayDesc { double *ptr; long long len; };

struct NcclId { char internal[128]; };
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
};
static NcclApi g_nccl;

struct cb200_handle {
    int device = 0, batch = 0;
    bool generic = false;
    cudaStream_t stream = nullptr;
    HostProblem hp;        // KKT handles
    Symbolic gsym;         // LinearSolver-seam handles
    const Symbolic &sym() const { return generic ? gsym : hp.sym; }
    DevProblem P{};
    Batch B{};
    Options opt{};
    std::vector<void *> allocs;
    ArrayDesc arr[CB200_NUM_ARRAYS]{};
    long long *d_counts = nullptr, *h_counts = nullptr;
    size_t smem_bytes = 0;
    long long *prof_store = nullptr;
    void *comm = nullptr;
    int nranks = 1;
    int nnzW = 0, nnzG = 0, nnzC = 0;
    bool wide = false;          // heavy kernels with CB_THREADS_WIDE threads per instance (small batches)
    int resident_ctas = 0;      // CTAs of the heavy kernels that actually fit an SM (occupancy query at creation)
    struct Scatter { ScatterPlan plan; const int *d_idx = nullptr; double *d_caches = nullptr; } scatter[3];   // W, G, C
    struct Stage { StagePlan plan; const int *d_ptr = nullptr, *d_src = nullptr; double *d_caches = nullptr; } stage[5];   // grad f, (g'y)_x, (h'z)_x, g, h
    bool values_dirty = true;   // W or G values changed since the row-ordered copies were refreshed
    int *d_order = nullptr;     // cb200_lq_set_order
    double *d_cand = nullptr;   // cb200_filter_search: candidates' callback outputs
    double *d_diff = nullptr;   // cb200_differentiate: right-hand sides, results, solve scratch
    size_t diff_bytes = 0;
    cudaStream_t copy_stream = nullptr;   // cb200_differentiate: the upload of the right-hand sides runs beside the factorisation
    cudaEvent_t copy_done = nullptr;
    size_t cand_bytes = 0;
};

extern "C" const char *cb200_last_error(void) { return g_err.c_str(); }

extern "C" int cb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" void cb200_options_default(cb200_options *o)
{   // options.jl:6-59
    memset(o, 0, sizeof(*o));
    o->max_outer_iterations = 10; o->max_residual_iterations = 100; o->max_residual_line_search = 25;
    o->max_cone_line_search = 25; o->iterative_refinement = 1; o->max_iterative_refinement = 10;
    o->min_iterative_refinement = 1; o->scaling_line_search = 0.5; o->iterative_refinement_tolerance = 1.0e-10;
    o->central_path_initial = 1.0; o->central_path_update_tolerance = 10.0; o->central_path_scaling = 0.2;
    o->central_path_exponent = 1.5; o->penalty_initial = 1.0; o->penalty_scaling = 10.0; o->dual_initial = 0.0;
    o->residual_tolerance = 1.0e-4; o->optimality_tolerance = 1.0e-4; o->slack_tolerance = 1.0e-4;
    o->equality_tolerance = 1.0e-4; o->complementarity_tolerance = 1.0e-4; o->min_regularization = 1.0e-20;
    o->primal_regularization_initial = 1.0e-7; o->dual_regularization_initial = 1.0e-7;
    o->max_regularization = 1.0e40; o->dual_regularization = 1.0e-8; o->dual_regularization_exponent = 0.25;
    o->scaling_regularization_initial = 100.0; o->scaling_regularization = 8.0;
    o->scaling_regularization_last = 1.0 / 3.0; o->max_penalty = 1.0e8; o->violation_tolerance = 1.0e-5;
    o->violation_exponent = 1.1; o->merit_tolerance = 1.0e-5; o->merit_exponent = 2.3; o->armijo_tolerance = 1.0e-4;
    o->machine_tolerance = 1.0e-16; o->max_filter = 1000; o->gmres_restart = 30; o->gmres_max_cycles = 10;
}

template <class T> static const T *upload(cb200_handle *h, const std::vector<T> &v, bool &ok)
{
    void *d = nullptr;
    size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    if (cudaMalloc(&d, bytes) != cudaSuccess) { ok = false; return nullptr; }
    h->allocs.push_back(d);
    if (!v.empty() && cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) ok = false;
    return (const T *)d;
}

static double *dalloc(cb200_handle *h, long long per_instance, bool &ok, bool zero = true)
{
    void *d = nullptr;
    size_t bytes = (size_t)std::max<long long>(per_instance * h->batch, 1) * sizeof(double);
    if (cudaMalloc(&d, bytes) != cudaSuccess) { ok = false; return nullptr; }
    h->allocs.push_back(d);
    if (zero && cudaMemset(d, 0, bytes) != cudaSuccess) ok = false;
    return (double *)d;
}

static int common_init(cb200_handle *h, int device)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail("no CUDA device: libcalipso_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail("invalid device index");
    h->device = device;
    CUDA_OK(cudaSetDevice(device));
    CUDA_OK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CUDA_OK(cudaMalloc(&h->d_counts, 4 * sizeof(long long)));
    CUDA_OK(cudaMallocHost(&h->h_counts, 4 * sizeof(long long)));
    return 0;
}

static const int BIG_TASK_THRESHOLD = 3000;

static bool finish_batch(cb200_handle *h)
{   // profile counters + dynamic shared memory of the factor/solve kernels
    void *d = nullptr;
    size_t bytes = (size_t)h->batch * PROF_COUNT * sizeof(long long);
    if (cudaMalloc(&d, bytes) != cudaSuccess || cudaMemset(d, 0, bytes) != cudaSuccess) return false;
    h->allocs.push_back(d);
    h->prof_store = (long long *)d;   // counters stay off (B.prof == nullptr) until cb200_get_profile is first called
    h->B.prof = nullptr;
    {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
        h->wide = h->batch <= sms || h->sym().ctas_per_sm == 1;      // one CTA per SM anyway: give it 512 threads
        if (const char *e = getenv("CB200_THREADS")) h->wide = atoi(e) > CB_THREADS;   // testing / tuning override
    }
    h->B.scratch_doubles = h->sym().scratch_doubles;
    h->smem_bytes = (size_t)h->B.scratch_doubles * sizeof(double);
    if (!h->wide) {   // what the hardware grants at this shared-memory size (a plan that overshoots its budget would silently lose a CTA per SM)
        int n = 0;
        const int T = h->sym().threads;
        auto query = [&](auto kernel) {
            cudaFuncAttributes fa;
            if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess) return 0;
            if ((size_t)fa.maxDynamicSharedSizeBytes < h->smem_bytes &&
                cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes) != cudaSuccess) return 0;
            int k = 0;
            return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k, kernel, T, h->smem_bytes) == cudaSuccess ? k : 0;
        };
        n = T == CB_THREADS_NARROW ? query(k_kkt_factor_solve<CB_THREADS_NARROW>) : query(k_kkt_factor_solve<CB_THREADS>);
        if (n == 0) cudaGetLastError();
        h->resident_ctas = n;
    }
    return true;
}   // multiply-adds above which a supernode gets the whole CTA

extern "C" cb200_handle *cb200_create(int batch, int n, int m, int p, int q_nn, int nsoc, const int *soc_dims,
                                      const int *Wp, const int *Wi, const int *Gp, const int *Gi, const int *Cp,
                                      const int *Ci, const int *perm, const cb200_options *options, int device)
{
    if (batch <= 0 || n <= 0 || m < 0 || p < 0) { fail("invalid dimensions"); return nullptr; }
    cb200_handle *h = new cb200_handle();
    h->batch = batch;
    if (common_init(h, device)) { delete h; return nullptr; }
    cb200_options dflt;
    cb200_options_default(&dflt);
    memcpy(&h->opt, options ? options : &dflt, sizeof(Options));
    DevProblem &P = h->P;
    {
        std::string msg = h->hp.build(n, m, p, q_nn, nsoc, soc_dims, Wp, Wi, Gp, Gi, Cp, Ci, perm, BIG_TASK_THRESHOLD);
        if (!msg.empty()) { fail(msg); cb200_destroy(h); return nullptr; }
    }
    const int nnzW = h->hp.nnzW, nnzG = h->hp.nnzG, nnzC = h->hp.nnzC, N = h->hp.N;
    h->nnzW = nnzW; h->nnzG = nnzG; h->nnzC = nnzC;
    bool ok = true;
    fill_problem(P, h->hp, [&](const auto &v) { return upload(h, v, ok); });
    // per-instance storage
    Batch &B = h->B;
    const long long T = P.total;
    B.count = batch; B.F = h->opt.max_filter;
    const int mr = h->opt.gmres_restart;
    B.ksize = mr > 0 ? (long long)(mr + 1) * T + (long long)(mr + 1) * mr + 4LL * mr + 8 : 0;
    B.w = dalloc(h, T, ok); B.cand = dalloc(h, T, ok); B.step = dalloc(h, T, ok); B.res = dalloc(h, T, ok);
    B.err = dalloc(h, T, ok); B.corr = dalloc(h, T, ok); B.tmp = dalloc(h, T, ok);
    B.grad = dalloc(h, n, ok); B.gyx = dalloc(h, n, ok); B.hzx = dalloc(h, n, ok);
    B.g = dalloc(h, m, ok); B.h = dalloc(h, p, ok);
    B.Wv = dalloc(h, nnzW, ok); B.Gv = dalloc(h, nnzG, ok); B.Cv = dalloc(h, nnzC, ok);
    B.Wf = dalloc(h, P.nnzWf, ok); B.Gr = dalloc(h, nnzG, ok);
    B.prod = dalloc(h, p, ok); B.bgrad = dalloc(h, p, ok); B.lambda = dalloc(h, m, ok);
    B.panels = dalloc(h, P.panel_total, ok); B.D = dalloc(h, N, ok); B.Dinv = dalloc(h, N, ok);
    B.kx = dalloc(h, P.kx_total, ok); B.Lcsr = dalloc(h, P.lcsr_total, ok);
    B.xs = dalloc(h, N, ok); B.rs = dalloc(h, N, ok); B.xp = dalloc(h, N, ok); B.mgrad = dalloc(h, N, ok);
    B.q = dalloc(h, n, ok); B.g0 = dalloc(h, m, ok); B.h0 = dalloc(h, p, ok);
    B.filter = dalloc(h, 4LL * B.F, ok);
    B.krylov = B.ksize ? dalloc(h, B.ksize, ok, false) : nullptr;
    B.scal = dalloc(h, S_COUNT, ok);
    B.Aval = nullptr; B.rhs = nullptr;
    {
        void *d = nullptr;
        size_t bytes = (size_t)batch * I_COUNT * sizeof(int);
        if (cudaMalloc(&d, bytes) != cudaSuccess || cudaMemset(d, 0, bytes) != cudaSuccess) ok = false;
        h->allocs.push_back(d);
        B.istat = (int *)d;
    }
    ok = ok && finish_batch(h);
    if (!ok) { fail(std::string("device allocation/upload failed: ") + cudaGetErrorString(cudaGetLastError())); cb200_destroy(h); return nullptr; }
    ArrayDesc *a = h->arr;
    a[CB200_POINT] = {B.w, T}; a[CB200_CANDIDATE] = {B.cand, T}; a[CB200_STEP] = {B.step, T};
    a[CB200_RESIDUAL] = {B.res, T}; a[CB200_GRADIENT] = {B.grad, n}; a[CB200_EQ_DUAL_GRAD] = {B.gyx, n};
    a[CB200_CONE_DUAL_GRAD] = {B.hzx, n}; a[CB200_EQUALITY] = {B.g, m}; a[CB200_CONE] = {B.h, p};
    a[CB200_W_VALUES] = {B.Wv, nnzW}; a[CB200_G_VALUES] = {B.Gv, nnzG}; a[CB200_C_VALUES] = {B.Cv, nnzC};
    a[CB200_CONE_PRODUCT] = {B.prod, p}; a[CB200_BARRIER_GRADIENT] = {B.bgrad, p}; a[CB200_DUAL] = {B.lambda, m};
    a[CB200_LQ_Q] = {B.q, n}; a[CB200_LQ_G0] = {B.g0, m}; a[CB200_LQ_H0] = {B.h0, p};
    a[CB200_SCALARS] = {B.scal, S_COUNT}; a[CB200_MERIT_GRADIENT] = {B.mgrad, N};
    a[CB200_RESIDUAL_SYMMETRIC] = {B.rs, N}; a[CB200_STEP_SYMMETRIC] = {B.xs, N}; a[CB200_PIVOTS] = {B.D, N};
    a[CB200_MATRIX_VALUES] = {nullptr, 0}; a[CB200_RHS] = {nullptr, 0}; a[CB200_PANELS] = {B.panels, P.panel_total};
    // solver.jl:81-86,125-127 initial scalars
    std::vector<double> sc((size_t)batch * S_COUNT, 0.0);
    for (int b = 0; b < batch; b++) { sc[(size_t)b * S_COUNT + S_KAPPA] = 0.1; sc[(size_t)b * S_COUNT + S_TAU] = 0.99; sc[(size_t)b * S_COUNT + S_RHO] = 10.0; }
    cudaMemcpy(B.scal, sc.data(), sc.size() * sizeof(double), cudaMemcpyHostToDevice);
    return h;
}

extern "C" cb200_handle *cb200_ldl_create(int batch, int N, const int *Ap, const int *Ai, const int *perm, int device)
{
    if (batch <= 0 || N <= 0) { fail("invalid dimensions"); return nullptr; }
    cb200_handle *h = new cb200_handle();
    h->batch = batch;
    h->generic = true;
    if (common_init(h, device)) { delete h; return nullptr; }
    cb200_options dflt;
    cb200_options_default(&dflt);
    memcpy(&h->opt, &dflt, sizeof(Options));
    const char *msg = h->gsym.analyze_auto(N, Ap, Ai, perm, BIG_TASK_THRESHOLD);
    if (msg[0]) { fail(msg); cb200_destroy(h); return nullptr; }
    DevProblem &P = h->P;
    bool ok = true;
    fill_symbolic(P, h->gsym, [&](const auto &v) { return upload(h, v, ok); });
    P.nnzA = Ap[N];
    Batch &B = h->B;
    B.count = batch;
    B.panels = dalloc(h, P.panel_total, ok); B.D = dalloc(h, N, ok); B.Dinv = dalloc(h, N, ok);
    B.kx = dalloc(h, P.kx_total, ok); B.Lcsr = dalloc(h, P.lcsr_total, ok);
    B.xp = dalloc(h, N, ok); B.Aval = dalloc(h, P.nnzA, ok); B.rhs = dalloc(h, N, ok);
    {
        void *d = nullptr;
        size_t bytes = (size_t)batch * I_COUNT * sizeof(int);
        if (cudaMalloc(&d, bytes) != cudaSuccess || cudaMemset(d, 0, bytes) != cudaSuccess) ok = false;
        h->allocs.push_back(d);
        B.istat = (int *)d;
    }
    ok = ok && finish_batch(h);
    if (!ok) { fail("device allocation/upload failed"); cb200_destroy(h); return nullptr; }
    h->arr[CB200_PIVOTS] = {B.D, N};
    h->arr[CB200_MATRIX_VALUES] = {B.Aval, P.nnzA};
    h->arr[CB200_RHS] = {B.rhs, N};
    h->arr[CB200_PANELS] = {B.panels, P.panel_total};
    return h;
}

extern "C" void cb200_destroy(cb200_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    for (void *p : h->allocs) cudaFree(p);
    if (h->d_counts) cudaFree(h->d_counts);
    if (h->h_counts) cudaFreeHost(h->h_counts);
    if (h->copy_done) cudaEventDestroy(h->copy_done);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" int cb200_info(const cb200_handle *h, long long *out)
{
    const Symbolic &S = h->sym();
    out[0] = S.N; out[1] = h->P.total; out[2] = S.nnzA; out[3] = S.nnzL; out[4] = S.ns; out[5] = S.nlevels;
    out[6] = (long long)S.phases.size(); out[7] = S.max_w; out[8] = S.max_nrow; out[9] = S.panel_total;
    out[10] = S.flops; out[11] = h->batch; out[12] = h->P.n; out[13] = h->P.m; out[14] = h->P.p;
    out[15] = (long long)h->nnzW + h->nnzG + h->nnzC;
    return 0;
}

extern "C" int cb200_path_info(const cb200_handle *h, long long *out)
{
    const Symbolic &S = h->sym();
    out[0] = S.solve_smem; out[1] = (h->resident_ctas > 0 && h->resident_ctas < S.ctas_per_sm) ? h->resident_ctas : S.ctas_per_sm; out[2] = (long long)S.scratch_doubles * 8; out[3] = S.n_cta_tasks;
    out[4] = S.n_generic_cta_tasks; out[5] = (long long)S.big.size() <= CB_MAX_CHAIN; out[6] = (long long)S.phases.size() <= CB_MAX_PHASES;
    out[7] = h->wide ? CB_THREADS_WIDE : h->sym().threads;
    return 0;
}

extern "C" int cb200_amd_order(int N, const int *Ap, const int *Ai, int *perm)
{
    if (N < 0 || !Ap || !Ai || !perm) return fail("cb200_amd_order: invalid arguments");
    for (int j = 0; j < N; j++)
        for (int q = Ap[j]; q < Ap[j + 1]; q++)
            if (Ai[q] < 0 || Ai[q] >= N || (q > Ap[j] && Ai[q] <= Ai[q - 1])) return fail("cb200_amd_order: rows must be sorted, unique and in range");
    std::vector<int> p;
    amd_order(N, Ap, Ai, p);
    std::copy(p.begin(), p.end(), perm);
    return 0;
}

extern "C" int cb200_get_symbolic(const cb200_handle *h, int *perm, int *etree, int *Lnz)
{
    const Symbolic &S = h->sym();
    if (perm) memcpy(perm, S.perm.data(), sizeof(int) * S.N);
    if (etree) memcpy(etree, S.etree.data(), sizeof(int) * S.N);
    if (Lnz) memcpy(Lnz, S.Lnz.data(), sizeof(int) * S.N);
    return 0;
}

extern "C" int cb200_get_factor(cb200_handle *h, int instance, int *Lp, int *Li, double *Lx, double *D)
{
    if (instance < 0 || instance >= h->batch) return fail("instance out of range");
    const Symbolic &S = h->sym();
    CUDA_OK(cudaSetDevice(h->device));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    std::vector<double> pan((size_t)S.panel_total);
    CUDA_OK(cudaMemcpy(pan.data(), h->B.panels + (long long)instance * S.panel_total, sizeof(double) * pan.size(), cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(D, h->B.D + (long long)instance * S.N, sizeof(double) * S.N, cudaMemcpyDeviceToHost));
    extract_factor(S, pan.data(), Lp, Li, Lx);
    return 0;
}

static int check_array(cb200_handle *h, int which, int first, int count)
{
    if (which < 0 || which >= CB200_NUM_ARRAYS || !h->arr[which].ptr) return fail("array not available on this handle");
    if (first < 0 || count < 0 || first + count > h->batch) return fail("instance range out of bounds");
    return 0;
}

extern "C" int cb200_set_array(cb200_handle *h, int which, const double *host, int first, int count)
{
    if (check_array(h, which, first, count)) return -1;
    CUDA_OK(cudaSetDevice(h->device));
    const ArrayDesc &a = h->arr[which];
    CUDA_OK(cudaMemcpyAsync(a.ptr + (long long)first * a.len, host, sizeof(double) * a.len * count, cudaMemcpyHostToDevice, h->stream));
    if (which == CB200_W_VALUES || which == CB200_G_VALUES) h->values_dirty = true;
    return 0;
}

extern "C" int cb200_initialize(cb200_handle *h, const double *guess_host, int first, int count)
{
    if (h->generic) return fail("not available on a LinearSolver-seam handle");
    if (check_array(h, CB200_POINT, first, count)) return -1;
    if (count == 0) return 0;
    CUDA_OK(cudaSetDevice(h->device));
    const ArrayDesc &a = h->arr[CB200_POINT];
    const size_t row = sizeof(double) * (size_t)h->hp.n;
    CUDA_OK(cudaMemcpy2DAsync(a.ptr + (long long)first * a.len, sizeof(double) * a.len, guess_host, row, row, count,
                              cudaMemcpyHostToDevice, h->stream));
    return 0;
}

extern "C" int cb200_get_array(cb200_handle *h, int which, double *host, int first, int count)
{
    if (check_array(h, which, first, count)) return -1;
    CUDA_OK(cudaSetDevice(h->device));
    const ArrayDesc &a = h->arr[which];
    CUDA_OK(cudaMemcpyAsync(host, a.ptr + (long long)first * a.len, sizeof(double) * a.len * count, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int cb200_get_array_async(cb200_handle *h, int which, double *host, int first, int count)
{
    if (check_array(h, which, first, count)) return -1;
    CUDA_OK(cudaSetDevice(h->device));
    const ArrayDesc &a = h->arr[which];
    CUDA_OK(cudaMemcpyAsync(host, a.ptr + (long long)first * a.len, sizeof(double) * a.len * count, cudaMemcpyDeviceToHost, h->stream));
    return 0;
}

extern "C" int cb200_get_stats_async(cb200_handle *h, int *host, int first, int count)
{
    if (first < 0 || count < 0 || first + count > h->batch) return fail("instance range out of bounds");
    CUDA_OK(cudaSetDevice(h->device));
    CUDA_OK(cudaMemcpyAsync(host, h->B.istat + (long long)first * I_COUNT, sizeof(int) * I_COUNT * count, cudaMemcpyDeviceToHost, h->stream));
    return 0;
}

extern "C" int cb200_get_stats(cb200_handle *h, int *host, int first, int count)
{
    if (first < 0 || count < 0 || first + count > h->batch) return fail("instance range out of bounds");
    CUDA_OK(cudaSetDevice(h->device));
    CUDA_OK(cudaMemcpyAsync(host, h->B.istat + (long long)first * I_COUNT, sizeof(int) * I_COUNT * count, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int cb200_get_profile(cb200_handle *h, long long *host, int reset)
{   // [batch][16] cycle counters (see PROF_* in device_core.h)
    CUDA_OK(cudaSetDevice(h->device));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    size_t bytes = (size_t)h->batch * PROF_COUNT * sizeof(long long);
    if (!h->B.prof) {      // first call switches the counters on (they cost a global read-modify-write per phase)
        h->B.prof = h->prof_store;
        CUDA_OK(cudaMemset(h->B.prof, 0, bytes));
    }
    if (host) CUDA_OK(cudaMemcpy(host, h->B.prof, bytes, cudaMemcpyDeviceToHost));
    if (reset) CUDA_OK(cudaMemset(h->B.prof, 0, bytes));
    return 0;
}

extern "C" int cb200_array_length(const cb200_handle *h, int which)
{
    if (which < 0 || which >= CB200_NUM_ARRAYS || !h->arr[which].ptr) return -1;
    return (int)h->arr[which].len;
}
extern "C" void *cb200_device_ptr(cb200_handle *h, int which)
{
    if (which < 0 || which >= CB200_NUM_ARRAYS) return nullptr;
    return h->arr[which].ptr;
}
extern "C" int cb200_values_changed(cb200_handle *h) { h->values_dirty = true; return 0; }
extern "C" void *cb200_stream(cb200_handle *h) { return (void *)h->stream; }
extern "C" int cb200_synchronize(cb200_handle *h)
{
    CUDA_OK(cudaSetDevice(h->device));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
}
extern "C" int cb200_set_options(cb200_handle *h, const cb200_options *o)
{
    if (o->max_filter != h->opt.max_filter || o->gmres_restart != h->opt.gmres_restart)
        return fail("max_filter / gmres_restart are fixed at creation");
    memcpy(&h->opt, o, sizeof(Options));
    return 0;
}

#define LAUNCH(kernel, ...)                                                   \
    do {                                                                      \
        CUDA_OK(cudaSetDevice(h->device));                                    \
        kernel<<<h->batch, CB_THREADS, 0, h->stream>>>(__VA_ARGS__);          \
        CUDA_OK(cudaGetLastError());                                          \
    } while (0)
// kernels that factor or solve get the CTA work area (panel + Y staging) as dynamic shared memory
// (the opt-in is a per-device function attribute: remembered per device, atomically -- handles on several GPUs and
// several host threads may share one process)
#define CB_MAX_DEVICES 64
#define LAUNCH_SMEM_T(kernel, T, ...)                                                                     \
    do {                                                                                                  \
        static std::atomic<size_t> configured[CB_MAX_DEVICES];                                            \
        const int dv_ = h->device < CB_MAX_DEVICES ? h->device : CB_MAX_DEVICES - 1;                      \
        if (h->device >= CB_MAX_DEVICES - 1 || h->smem_bytes > configured[dv_].load()) {                  \
            CUDA_OK(cudaFuncSetAttribute(kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes)); \
            size_t cur_ = configured[dv_].load();                                                         \
            while (cur_ < h->smem_bytes && !configured[dv_].compare_exchange_weak(cur_, h->smem_bytes)) {} \
        }                                                                                                 \
        kernel<T><<<h->batch, T, h->smem_bytes, h->stream>>>(__VA_ARGS__);                                \
    } while (0)
// Small batches (at most one CTA per SM anyway) run the heavy kernels with CB_THREADS_WIDE threads per instance: the
// parallel phases (gathers, tensor-core tiles, bulk solve passes, J v) get twice the warps, the register cap doubles.
#define LAUNCH_SMEM(kernel, ...)                                                                          \
    do {                                                                                                  \
        CUDA_OK(cudaSetDevice(h->device));                                                                \
        if (h->wide) LAUNCH_SMEM_T(kernel, CB_THREADS_WIDE, __VA_ARGS__);                                 \
        else if (h->sym().threads == CB_THREADS_NARROW) LAUNCH_SMEM_T(kernel, CB_THREADS_NARROW, __VA_ARGS__); \
        else LAUNCH_SMEM_T(kernel, CB_THREADS, __VA_ARGS__);                                              \
        CUDA_OK(cudaGetLastError());                                                                      \
    } while (0)
#define NEED_KKT() if (h->generic) return fail("not available on a LinearSolver-seam handle")
// kernels that read the row-ordered copies of W / G refresh them first when cb200_set_array changed the values
#define FRESH_VALUES()                                          \
    do {                                                        \
        if (h->values_dirty) {                                  \
            LAUNCH(k_expand, h->P, h->B);                       \
            h->values_dirty = false;                            \
        }                                                       \
    } while (0)

extern "C" int cb200_cone(cb200_handle *h, int flags, int at_candidate) { NEED_KKT(); LAUNCH(k_cone, h->P, h->B, flags, at_candidate); return 0; }
extern "C" int cb200_residual(cb200_handle *h) { NEED_KKT(); LAUNCH(k_residual, h->P, h->B); return 0; }
extern "C" int cb200_search_direction(cb200_handle *h) { NEED_KKT(); FRESH_VALUES(); LAUNCH_SMEM(k_search_direction, h->P, h->B, h->opt); return 0; }
extern "C" int cb200_cone_search(cb200_handle *h) { NEED_KKT(); LAUNCH(k_cone_search, h->P, h->B, h->opt); return 0; }
extern "C" int cb200_apply_step(cb200_handle *h) { NEED_KKT(); LAUNCH(k_apply_step, h->P, h->B); return 0; }
extern "C" int cb200_lq_evaluate(cb200_handle *h, int flags, int at_candidate) { NEED_KKT(); FRESH_VALUES(); LAUNCH(k_lq_evaluate, h->P, h->B, flags, at_candidate); return 0; }
extern "C" int cb200_lq_begin(cb200_handle *h, int warmstart) { NEED_KKT(); FRESH_VALUES(); LAUNCH(k_lq_begin, h->P, h->B, h->opt, warmstart); return 0; }
extern "C" int cb200_lq_step(cb200_handle *h, int iterations)
{
    NEED_KKT();
    FRESH_VALUES();
    // one launch carries all `iterations` passes of an instance (CB200_LQ_LOCKSTEP=1: one launch per pass, every instance
    // in lock-step -- the same arithmetic; kept for A/B timing)
    const char *ls = getenv("CB200_LQ_LOCKSTEP");
    const bool lockstep = ls && atoi(ls) != 0;
    if (lockstep) {
        for (int k = 0; k < iterations; k++) LAUNCH_SMEM(k_lq_step, h->P, h->B, h->opt, 1);
    } else if (iterations > 0) {
        LAUNCH_SMEM(k_lq_step, h->P, h->B, h->opt, iterations);
    }
    return 0;
}
extern "C" int cb200_filter_reset(cb200_handle *h) { NEED_KKT(); LAUNCH(k_filter_reset, h->P, h->B, h->opt); return 0; }

extern "C" int cb200_filter_search(cb200_handle *h, int first, int count, const double *f_host, const double *g_host,
                                   const double *h_host, int *accepted_host)
{
    NEED_KKT();
    if (first < 0 || count <= 0 || !f_host || !accepted_host) return fail("cb200_filter_search: invalid arguments");
    CUDA_OK(cudaSetDevice(h->device));
    const size_t B = (size_t)h->batch, m = (size_t)h->P.m, p = (size_t)h->P.p, c = (size_t)count;
    const size_t nf = B * c, ng = B * c * m, nh = B * c * p;
    const size_t need = (nf + ng + nh) * sizeof(double) + B * sizeof(int);
    if (need > h->cand_bytes) {      // staging buffer of the candidates' callback outputs, grown on demand
        CUDA_OK(cudaStreamSynchronize(h->stream));
        if (h->d_cand) {
            h->allocs.erase(std::remove(h->allocs.begin(), h->allocs.end(), (void *)h->d_cand), h->allocs.end());
            cudaFree(h->d_cand);
            h->d_cand = nullptr;
            h->cand_bytes = 0;
        }
        void *d = nullptr;
        CUDA_OK(cudaMalloc(&d, need));
        h->allocs.push_back(d);
        h->d_cand = (double *)d;
        h->cand_bytes = need;
    }
    double *df = h->d_cand, *dg = df + nf, *dh = dg + ng;
    int *dacc = (int *)(dh + nh);
    CUDA_OK(cudaMemcpyAsync(df, f_host, nf * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (ng) CUDA_OK(cudaMemcpyAsync(dg, g_host, ng * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (nh) CUDA_OK(cudaMemcpyAsync(dh, h_host, nh * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    LAUNCH(k_filter_search, h->P, h->B, h->opt, first, count, df, dg, dh, dacc);
    CUDA_OK(cudaMemcpyAsync(accepted_host, dacc, B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int cb200_lq_set_order(cb200_handle *h, const int *order)
{
    NEED_KKT();
    CUDA_OK(cudaSetDevice(h->device));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    if (!order) { h->B.order = nullptr; return 0; }
    std::vector<char> seen((size_t)h->batch, 0);
    for (int i = 0; i < h->batch; i++) {
        if (order[i] < 0 || order[i] >= h->batch || seen[(size_t)order[i]]) return fail("cb200_lq_set_order: not a permutation of the instances");
        seen[(size_t)order[i]] = 1;
    }
    if (!h->d_order) {
        void *d = nullptr;
        CUDA_OK(cudaMalloc(&d, sizeof(int) * (size_t)h->batch));
        h->allocs.push_back(d);
        h->d_order = (int *)d;
    }
    CUDA_OK(cudaMemcpy(h->d_order, order, sizeof(int) * (size_t)h->batch, cudaMemcpyHostToDevice));
    h->B.order = h->d_order;
    return 0;
}

extern "C" int cb200_kkt_factor_solve(cb200_handle *h, int nsolves) { NEED_KKT(); FRESH_VALUES(); LAUNCH_SMEM(k_kkt_factor_solve, h->P, h->B, nsolves); return 0; }

extern "C" int cb200_differentiate(cb200_handle *h, int nparam, const double *H_host, double *S_host)
{
    NEED_KKT();
    if (nparam <= 0) return 0;
    FRESH_VALUES();
    CUDA_OK(cudaSetDevice(h->device));
    const size_t pairs = (size_t)h->batch * (size_t)nparam;
    const size_t bytes = sizeof(double) * pairs * (size_t)h->P.total, wbytes = sizeof(double) * pairs * 3 * (size_t)h->P.N;
    // work area (right-hand sides, results, per-column solve scratch): kept with the handle and grown on demand -- a
    // cudaMalloc / cudaFree pair per call costs more than the solves (5-400 ms measured, tools/r2_diff_probe.py)
    const size_t need = 2 * bytes + wbytes;
    if (need > h->diff_bytes) {
        CUDA_OK(cudaStreamSynchronize(h->stream));
        if (h->d_diff) {
            h->allocs.erase(std::remove(h->allocs.begin(), h->allocs.end(), (void *)h->d_diff), h->allocs.end());
            cudaFree(h->d_diff);
            h->d_diff = nullptr;
            h->diff_bytes = 0;
        }
        void *d = nullptr;
        if (cudaMalloc(&d, need) != cudaSuccess) return fail("cb200_differentiate: device allocation failed");
        h->allocs.push_back(d);
        h->d_diff = (double *)d;
        h->diff_bytes = need;
    }
    double *dH = h->d_diff, *dS = dH + pairs * (size_t)h->P.total, *dW = dS + pairs * (size_t)h->P.total;
    int rc = 0;
    do {
        // the right-hand sides are uploaded on a second stream while the factorisation (which does not read them) runs
        if (!h->copy_stream && (cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
                                cudaEventCreateWithFlags(&h->copy_done, cudaEventDisableTiming) != cudaSuccess)) { rc = fail("stream creation failed"); break; }
        if (cudaEventRecord(h->copy_done, h->stream) != cudaSuccess ||                      // (the work area may still be in use)
            cudaStreamWaitEvent(h->copy_stream, h->copy_done, 0) != cudaSuccess ||
            cudaMemcpyAsync(dH, H_host, bytes, cudaMemcpyHostToDevice, h->copy_stream) != cudaSuccess ||
            cudaEventRecord(h->copy_done, h->copy_stream) != cudaSuccess) { rc = fail("H2D failed"); break; }
        if (cb200_kkt_factor_solve(h, 0)) { rc = -1; break; }          // assemble + factor at the current point and regularisation
        if (cudaStreamWaitEvent(h->stream, h->copy_done, 0) != cudaSuccess) { rc = fail("cudaStreamWaitEvent failed"); break; }
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
        const bool wide = pairs <= (size_t)sms || h->sym().ctas_per_sm == 1;
        if (wide) {
            if (cudaFuncSetAttribute(k_differentiate<CB_THREADS_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes) != cudaSuccess) { rc = fail("cudaFuncSetAttribute failed"); break; }
            k_differentiate<CB_THREADS_WIDE><<<(unsigned)pairs, CB_THREADS_WIDE, h->smem_bytes, h->stream>>>(h->P, h->B, nparam, dH, dS, dW);
        } else if (h->sym().threads == CB_THREADS_NARROW) {
            if (cudaFuncSetAttribute(k_differentiate<CB_THREADS_NARROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes) != cudaSuccess) { rc = fail("cudaFuncSetAttribute failed"); break; }
            k_differentiate<CB_THREADS_NARROW><<<(unsigned)pairs, CB_THREADS_NARROW, h->smem_bytes, h->stream>>>(h->P, h->B, nparam, dH, dS, dW);
        } else {
            if (cudaFuncSetAttribute(k_differentiate<CB_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes) != cudaSuccess) { rc = fail("cudaFuncSetAttribute failed"); break; }
            k_differentiate<CB_THREADS><<<(unsigned)pairs, CB_THREADS, h->smem_bytes, h->stream>>>(h->P, h->B, nparam, dH, dS, dW);
        }
        if (cudaGetLastError() != cudaSuccess) { rc = fail("k_differentiate launch failed"); break; }
        if (cudaMemcpyAsync(S_host, dS, bytes, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) { rc = fail("D2H failed"); break; }
        cudaError_t e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) { rc = fail(std::string("cb200_differentiate: ") + cudaGetErrorString(e)); break; }
    } while (0);
    return rc;
}

static int scatter_slot(int which)
{
    return which == CB200_W_VALUES ? 0 : which == CB200_G_VALUES ? 1 : which == CB200_C_VALUES ? 2 : -1;
}

extern "C" int cb200_scatter_plan(cb200_handle *h, int which, int ncaches, const int *cache_len, const int *rows, const int *cols)
{
    NEED_KKT();
    const int slot = scatter_slot(which);
    if (slot < 0) return fail("cb200_scatter_plan: which must be CB200_W_VALUES, CB200_G_VALUES or CB200_C_VALUES");
    if (slot > 0 && ncaches != 1) return fail("cb200_scatter_plan: the G and C values have one cache each");
    const HostProblem &H = h->hp;
    cb200_handle::Scatter &sc = h->scatter[slot];
    std::string err = slot == 0 ? sc.plan.build(H.n, H.n, H.Wp.data(), H.Wi.data(), true, ncaches, cache_len, rows, cols)
                    : slot == 1 ? sc.plan.build(H.m, H.n, H.Gp.data(), H.Gi.data(), false, ncaches, cache_len, rows, cols)
                                : sc.plan.build(H.p, H.n, H.Cp.data(), H.Ci.data(), false, ncaches, cache_len, rows, cols);
    CUDA_OK(cudaSetDevice(h->device));
    CUDA_OK(cudaStreamSynchronize(h->stream));      // a previous plan may still be in use
    for (void *old : {(void *)sc.d_idx, (void *)sc.d_caches})      // a new plan replaces the previous one
        if (old) {
            h->allocs.erase(std::remove(h->allocs.begin(), h->allocs.end(), old), h->allocs.end());
            cudaFree(old);
        }
    sc.d_idx = nullptr;
    sc.d_caches = nullptr;
    if (!err.empty()) { sc.plan = ScatterPlan(); return fail(err); }
    bool ok = true;
    sc.d_idx = upload(h, sc.plan.idx, ok);
    sc.d_caches = dalloc(h, sc.plan.cache_total, ok, false);
    if (!ok) return fail("cb200_scatter_plan: device allocation failed");
    return 0;
}

extern "C" void *cb200_scatter_buffer(cb200_handle *h, int which)
{
    const int slot = scatter_slot(which);
    return slot < 0 ? nullptr : (void *)h->scatter[slot].d_caches;
}

extern "C" int cb200_scatter(cb200_handle *h, int which, const double *caches_host, int first, int count)
{
    NEED_KKT();
    const int slot = scatter_slot(which);
    if (slot < 0 || !h->scatter[slot].d_idx) return fail("cb200_scatter: no scatter plan for this array");
    if (check_array(h, which, first, count)) return -1;
    if (count == 0) return 0;
    const cb200_handle::Scatter &sc = h->scatter[slot];
    const ArrayDesc &a = h->arr[which];
    CUDA_OK(cudaSetDevice(h->device));
    if (caches_host)
        CUDA_OK(cudaMemcpyAsync(sc.d_caches + (long long)first * sc.plan.cache_total, caches_host,
                                sizeof(double) * sc.plan.cache_total * count, cudaMemcpyHostToDevice, h->stream));
    if (sc.plan.nnz > 0) {
        const int chunks = std::max(1, std::min((sc.plan.nnz + 1023) / 1024, 64));     // four entries per thread or more
        for (int c0 = 0; c0 < count; c0 += 65535) {       // gridDim.y limit
            k_scatter<<<dim3(chunks, std::min(count - c0, 65535)), 256, 0, h->stream>>>(
                sc.plan.nnz, sc.plan.ncaches, sc.d_idx, sc.d_caches, sc.plan.cache_total, a.ptr, a.len, first + c0);
            CUDA_OK(cudaGetLastError());
        }
    }
    if (which == CB200_W_VALUES || which == CB200_G_VALUES) h->values_dirty = true;
    return 0;
}

static int stage_slot(int which)
{
    return which == CB200_GRADIENT ? 0 : which == CB200_EQ_DUAL_GRAD ? 1 : which == CB200_CONE_DUAL_GRAD ? 2
         : which == CB200_EQUALITY ? 3 : which == CB200_CONE ? 4 : -1;
}

extern "C" int cb200_stage_plan(cb200_handle *h, int which, int accumulate, int count, const int *dst)
{
    NEED_KKT();
    const int slot = stage_slot(which);
    if (slot < 0) return fail("cb200_stage_plan: which must be CB200_GRADIENT, CB200_EQ_DUAL_GRAD, CB200_CONE_DUAL_GRAD, CB200_EQUALITY or CB200_CONE");
    if (count > 0 && !dst) return fail("cb200_stage_plan: no index list");
    cb200_handle::Stage &sg = h->stage[slot];
    std::string err = sg.plan.build((int)h->arr[which].len, accumulate != 0, count, dst);
    CUDA_OK(cudaSetDevice(h->device));
    CUDA_OK(cudaStreamSynchronize(h->stream));      // a previous plan may still be in use
    for (void *old : {(void *)sg.d_ptr, (void *)sg.d_src, (void *)sg.d_caches})      // a new plan replaces the previous one
        if (old) {
            h->allocs.erase(std::remove(h->allocs.begin(), h->allocs.end(), old), h->allocs.end());
            cudaFree(old);
        }
    sg.d_ptr = sg.d_src = nullptr;
    sg.d_caches = nullptr;
    if (!err.empty()) { sg.plan = StagePlan(); return fail(err); }
    bool ok = true;
    sg.d_ptr = upload(h, sg.plan.ptr, ok);
    sg.d_src = upload(h, sg.plan.src, ok);
    sg.d_caches = dalloc(h, sg.plan.cache_total, ok, false);
    if (!ok) return fail("cb200_stage_plan: device allocation failed");
    return 0;
}

extern "C" void *cb200_stage_buffer(cb200_handle *h, int which)
{
    const int slot = stage_slot(which);
    return slot < 0 ? nullptr : (void *)h->stage[slot].d_caches;
}

extern "C" int cb200_stage_scatter(cb200_handle *h, int which, const double *caches_host, int first, int count)
{
    NEED_KKT();
    const int slot = stage_slot(which);
    if (slot < 0 || !h->stage[slot].d_ptr) return fail("cb200_stage_scatter: no stage plan for this array");
    if (check_array(h, which, first, count)) return -1;
    if (count == 0) return 0;
    const cb200_handle::Stage &sg = h->stage[slot];
    const ArrayDesc &a = h->arr[which];
    CUDA_OK(cudaSetDevice(h->device));
    if (caches_host && sg.plan.cache_total > 0)
        CUDA_OK(cudaMemcpyAsync(sg.d_caches + (long long)first * sg.plan.cache_total, caches_host,
                                sizeof(double) * sg.plan.cache_total * count, cudaMemcpyHostToDevice, h->stream));
    if (sg.plan.nout > 0) {
        const int chunks = std::max(1, std::min((sg.plan.nout + 1023) / 1024, 64));
        for (int c0 = 0; c0 < count; c0 += 65535) {       // gridDim.y limit
            k_stage_gather<<<dim3(chunks, std::min(count - c0, 65535)), 256, 0, h->stream>>>(
                sg.plan.nout, sg.plan.accumulate, sg.d_ptr, sg.d_src, sg.d_caches, sg.plan.cache_total, a.ptr, a.len, first + c0);
            CUDA_OK(cudaGetLastError());
        }
    }
    return 0;
}

extern "C" int cb200_jacobian_times(cb200_handle *h, const double *v_host, double *out_host)
{
    NEED_KKT();
    FRESH_VALUES();
    CUDA_OK(cudaSetDevice(h->device));
    size_t bytes = sizeof(double) * h->P.total * (size_t)h->batch;
    CUDA_OK(cudaMemcpyAsync(h->B.err, v_host, bytes, cudaMemcpyHostToDevice, h->stream));
    LAUNCH(k_jtimes, h->P, h->B);
    CUDA_OK(cudaMemcpyAsync(out_host, h->B.tmp, bytes, cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int cb200_allreduce_counts(cb200_handle *h, long long *counts)
{
    CUDA_OK(cudaSetDevice(h->device));
    k_count_states<<<1, 256, 0, h->stream>>>(h->B, h->d_counts);
    CUDA_OK(cudaGetLastError());
    if (h->comm) {
        int rc = g_nccl.AllReduce(h->d_counts, h->d_counts, 4, /*ncclInt64*/ 4, /*ncclSum*/ 0, h->comm, h->stream);
        if (rc != 0) return fail("ncclAllReduce failed with code " + std::to_string(rc));
    }
    CUDA_OK(cudaMemcpyAsync(h->h_counts, h->d_counts, 4 * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    for (int k = 0; k < 4; k++) counts[k] = h->h_counts[k];
    return 0;
}

extern "C" int cb200_lq_solve(cb200_handle *h, int max_steps, int check_every, long long *counts, int *steps_done)
{
    NEED_KKT();
    if (check_every <= 0) check_every = 1;
    int done = 0;
    long long c[4] = {h->batch, 0, 0, 0};
    const bool wide0 = h->wide;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    while (done < max_steps) {
        int chunk = std::min(check_every, max_steps - done);
        if (cb200_lq_step(h, chunk)) { h->wide = wide0; return -1; }
        done += chunk;
        if (cb200_allreduce_counts(h, c)) { h->wide = wide0; return -1; }
        if (c[0] == 0) break;
        // The tail of a batched solve: once the instances still running (converged ones exit at once) fit one CTA per SM, the
        // 512-thread instantiations can take over.  Opt-in (CB200_TAIL_WIDE=1): the reductions of a 512-thread CTA associate
        // differently, so an instance's iterates would depend on when the switch happens (batch, check interval, ranks), and
        // measured on B200 it buys little -- the late iterations of the hard instances that form the tail are dominated by
        // refinement solves, 1.35x faster at 512 threads (profiles/r2_tail_experiments.txt).
        static const bool tail_wide = getenv("CB200_TAIL_WIDE") && atoi(getenv("CB200_TAIL_WIDE")) != 0;
        if (tail_wide && !getenv("CB200_THREADS") && c[0] <= (long long)sms * h->nranks) h->wide = true;
    }
    h->wide = wide0;
    if (counts) for (int k = 0; k < 4; k++) counts[k] = c[k];
    if (steps_done) *steps_done = done;
    return 0;
}

// ---------------------------------------------------------------------------------------------------- LinearSolver seam
extern "C" int cb200_ldl_factorize(cb200_handle *h)
{
    if (!h->generic) return fail("cb200_ldl_factorize needs a handle from cb200_ldl_create");
    LAUNCH_SMEM(k_ldl_factor, h->P, h->B, 1);
    return 0;
}
extern "C" int cb200_ldl_solve(cb200_handle *h)
{
    if (!h->generic) return fail("cb200_ldl_solve needs a handle from cb200_ldl_create");
    LAUNCH_SMEM(k_ldl_solve, h->P, h->B);
    return 0;
}
extern "C" int cb200_ldl_inertia(cb200_handle *h, int *out)
{
    std::vector<int> st((size_t)h->batch * I_COUNT);
    if (cb200_get_stats(h, st.data(), 0, h->batch)) return -1;
    for (int b = 0; b < h->batch; b++)
        for (int k = 0; k < 3; k++) out[3 * b + k] = st[(size_t)b * I_COUNT + k];
    return 0;
}
extern "C" int cb200_ldl_linear_solve(cb200_handle *h, const double *Ax, const double *bvec, double *x, int factorize)
{
    if (!h->generic) return fail("cb200_ldl_linear_solve needs a handle from cb200_ldl_create");
    if (factorize) {
        if (cb200_set_array(h, CB200_MATRIX_VALUES, Ax, 0, h->batch)) return -1;
        if (cb200_ldl_factorize(h)) return -1;
    }
    if (cb200_set_array(h, CB200_RHS, bvec, 0, h->batch)) return -1;
    if (cb200_ldl_solve(h)) return -1;
    return cb200_get_array(h, CB200_RHS, x, 0, h->batch);
}

// ---------------------------------------------------------------------------------------------------- NCCL (dlopen'd)
static int load_nccl()
{
    if (g_nccl.lib) return 0;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) return fail(std::string("cannot load NCCL: ") + dlerror());
    g_nccl.GetUniqueId = (int (*)(void *))dlsym(g_nccl.lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void **, int, NcclId, int))dlsym(g_nccl.lib, "ncclCommInitRank");
    g_nccl.AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))dlsym(g_nccl.lib, "ncclAllReduce");
    g_nccl.CommDestroy = (int (*)(void *))dlsym(g_nccl.lib, "ncclCommDestroy");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) return fail("NCCL symbols missing");
    return 0;
}

extern "C" int cb200_nccl_unique_id(char *out128)
{
    if (load_nccl()) return -1;
    NcclId id;
    int rc = g_nccl.GetUniqueId(&id);
    if (rc != 0) return fail("ncclGetUniqueId failed with code " + std::to_string(rc));
    memcpy(out128, id.internal, 128);
    return 0;
}

extern "C" int cb200_comm_init(cb200_handle *h, int rank, int nranks, const char *unique_id128)
{
    if (load_nccl()) return -1;
    CUDA_OK(cudaSetDevice(h->device));
    NcclId id;
    memcpy(id.internal, unique_id128, 128);
    int rc = g_nccl.CommInitRank(&h->comm, nranks, id, rank);
    if (rc != 0) { h->comm = nullptr; return fail("ncclCommInitRank failed with code " + std::to_string(rc)); }
    h->nranks = nranks;
    return 0;
}

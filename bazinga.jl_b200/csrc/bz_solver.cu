// Host side of libbazinga_b200.so: contexts, problems, the AL-gradient provider dispatch, the
// PANOC+ / ALPS drivers and the C-ABI entry points declared in include/bazinga_b200.h.
//
// Host responsibilities are deliberately thin: enqueue the fixed kernel cycle, poll one status
// word, and run the handful of per-OUTER-iteration decisions that need a host round trip
// (alps.jl:59-102 -- one sync per outer iteration).
#include "bz_internal.cuh"
#include "bz_dense.cuh"
#include "bz_vec.cuh"
#include "bz_comm.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <cstdlib>

static thread_local std::string g_last_error;   // bz_last_error(NULL): per host thread

cudaError_t bz_pool_alloc(bz_ctx* ctx, void** ptr, size_t bytes) {
    auto it = ctx->pool.find(bytes);
    if (it != ctx->pool.end()) {
        *ptr = it->second;
        ctx->pool.erase(it);
        ctx->pool_bytes -= bytes;
    } else {
        cudaError_t e = cudaMalloc(ptr, bytes);
        if (e != cudaSuccess && !ctx->pool.empty()) {      // out of memory with blocks parked: give them back and retry
            cudaGetLastError();
            bz_pool_release(ctx);
            e = cudaMalloc(ptr, bytes);
        }
        if (e != cudaSuccess) return e;
    }
    ctx->pool_live[*ptr] = bytes;
    return cudaSuccess;
}
void bz_pool_free(bz_ctx* ctx, void* ptr) {
    if (!ptr) return;
    auto it = ctx->pool_live.find(ptr);
    if (it == ctx->pool_live.end()) { cudaFree(ptr); return; }
    const size_t bytes = it->second;
    ctx->pool_live.erase(it);
    if (bytes > BZ_POOL_MAX_BLOCK || ctx->pool_bytes + bytes > BZ_POOL_MAX_BYTES) { cudaFree(ptr); return; }
    ctx->pool.emplace(bytes, ptr);
    ctx->pool_bytes += bytes;
}
void bz_pool_release(bz_ctx* ctx) {
    for (auto& kv : ctx->pool) cudaFree(kv.second);
    ctx->pool.clear();
    ctx->pool_bytes = 0;
}

int bz_fail(bz_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    g_last_error = msg;
    return code;
}

struct bz_problem {
    bz_ctx* ctx = nullptr;
    bz_problem_desc desc{};
    long long n = 0, m = 0, ml = 0;   // nx, global ny, local rows
    long long nv = 0;                 // length of the PANOC+ variable: nx, or nx + ny for als (desc.slack_formulation)
    bool slack = false, slack_active = false;
    // oracle data
    double* cA = nullptr; bool own_cA = false; double* cb = nullptr;
    double* fA = nullptr; double* fb = nullptr; long long fm = 0;
    double* f_ones = nullptr; double* f_zeros = nullptr; double* tm1_f = nullptr;
    double* g_u = nullptr;
    // AL m-vectors (local rows)
    double *mu = nullptr, *y = nullptr, *muy = nullptr, *cx = nullptr, *s = nullptr, *yupd = nullptr;
    double* rmu = nullptr;   // 1 ./ mu (fused obstacle kernel), refreshed by every k_al_update
    double *tm1 = nullptr, *tm2 = nullptr, *tm3 = nullptr;
    // n-vectors
    BzVecs v{};
    double *Ec = nullptr, *Ef = nullptr;
    // provider partial buffers
    double *Pc = nullptr, *Pf = nullptr, *lxpart = nullptr, *fxpart = nullptr;
    int c_clusters = 0, c_rpc = 0, f_clusters = 0, f_rpc = 0, nlx = 0, nfx = 0, nfx2 = 0;
    bool c_wide = false, f_wide = false;   // wide path (nx > 16384)
    bool c_grid = false, f_grid = false;   // ... served by the grid-wide single-pass kernel (bz_grid.cu)
    unsigned long long* gridD = nullptr;   // its exchange ring and tag sequence base (constraint pass)
    unsigned long long* gseq = nullptr;
    unsigned long long* gridD_f = nullptr; // ... of the LeastSquares pass
    unsigned long long* gseq_f = nullptr;
    int c_nlx = 0, f_nlx = 0;              // lx partial counts of the dense passes
    double* part = nullptr;
    unsigned int* counters = nullptr;   // [8]
    BzTeam team{};
    BzState* st = nullptr;
    BzState* st_host = nullptr;         // pinned
    int* status_host = nullptr;         // pinned, [2]
    cudaEvent_t poll_ev[2]{};
    BzGParams g{};
    BzDParams D{};
    BzObstacle ob{};
    bool sharded = false;
    BzShard shard{};
    // slab-decomposed obstacle grid (cross-rank exchange over peer memory)
    bool lspec = false;   // dense problems: the L step of the next iteration runs speculatively beside G(z) (bz_lspec.cuh)
    int c_clusters_z = 0, c_rpc_z = 0;      // ... on this many row clusters (the rest of the SMs carry the L step)
    int lspec_cs = 0;                       // CTAs of the speculative L step
    unsigned long long* lspec_buf = nullptr;
    bool rotate = false;  // fused problems large enough for the cooperative L kernel: x/x_prev, res/res_prev, z/z_curr swap roles
    bool fused = false;   // obstacle problems on one GPU: the cycle is [L, X] with the fused kernel of bz_obfused.cu
    bool slab = false;
    unsigned long long* slab_base = nullptr;   // [mailbox | seq | x | z | xeps | yupd]
    size_t slab_bytes = 0;
    BzXr xr_host{};
    void* peer_ptr[XR_MAXR] = {};
    long long n_global = 0;
    // dominant-kernel timing
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> dom_events;
    long long dom_counter = 0;
};

static const double* const* src_ptrs(bz_problem* p, const double* (&buf)[3]) {
    buf[0] = p->v.x; buf[1] = p->v.z; buf[2] = p->v.xeps;
    return buf;
}

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
extern "C" int bz_abi_version(void) { return BZ_ABI_VERSION; }

extern "C" const char* bz_last_error(bz_ctx* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

extern "C" int bz_device_count(int* n) {
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) { *n = 0; return bz_fail(nullptr, BZ_ERR_CUDA, cudaGetErrorString(e)); }
    *n = c;
    return BZ_OK;
}

extern "C" int bz_ctx_create(int device, bz_ctx** out) {
    if (!out) return bz_fail(nullptr, BZ_ERR_ARG, "bz_ctx_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return bz_fail(nullptr, BZ_ERR_CUDA,
                       std::string("no CUDA device (this library has no CPU path): ") +
                           cudaGetErrorString(e));
    if (device < 0 || device >= count) return bz_fail(nullptr, BZ_ERR_ARG, "bad device index");
    bz_ctx* ctx = new bz_ctx();
    ctx->device = device;
    BZ_CUDA(ctx, cudaSetDevice(device));
    cudaDeviceProp prop;
    BZ_CUDA(ctx, cudaGetDeviceProperties(&prop, device));
    ctx->num_sms = prop.multiProcessorCount;
    if (prop.major < 10) {
        std::string msg = "device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                          "; this library is built for sm_100a only";
        delete ctx;
        return bz_fail(nullptr, BZ_ERR_CUDA, msg);
    }
    BZ_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    if (const char* e = getenv("BZ_L2_PREFETCH_ROWS")) ctx->l2_prefetch_rows = atoi(e);
    if (const char* e = getenv("BZ_EVICT_FIRST")) ctx->evict_first = atoi(e);
    if (const char* e = getenv("BZ_PDL")) ctx->pdl = atoi(e);
    if (const char* e = getenv("BZ_FORCE_WIDE")) ctx->force_wide = atoi(e);
    if (const char* e = getenv("BZ_WIDE_TWO_PASS")) ctx->wide_two_pass = atoi(e);
    if (const char* e = getenv("BZ_GRID_VARIANT")) ctx->grid_variant = atoi(e);
    if (const char* e = getenv("BZ_GRID_TEAM")) ctx->grid_team = atoi(e);
    if (const char* e = getenv("BZ_BATCH_BUDGET")) ctx->batch_budget = atoi(e);
    if (const char* e = getenv("BZ_BATCH_CONCURRENT")) ctx->batch_concurrent = atoi(e);
    if (const char* e = getenv("BZ_BATCH_P2CTAS")) ctx->batch_p2ctas = atoi(e);
    if (const char* e = getenv("BZ_BATCH_DRAIN")) ctx->batch_drain = atoi(e);
    if (const char* e = getenv("BZ_LSPEC")) ctx->lspec = atoi(e);
    if (const char* e = getenv("BZ_OBSTACLE_ROTATE")) ctx->obstacle_rotate = atoi(e);
    if (const char* e = getenv("BZ_SLAB_FUSED")) ctx->slab_fused = atoi(e);
    if (const char* e = getenv("BZ_OBSTACLE_FUSED")) ctx->obstacle_fused = atoi(e);
    if (const char* e = getenv("BZ_PROF_STRIDE")) ctx->prof_stride = atoi(e) > 0 ? atoi(e) : 1;
    *out = ctx;
    return BZ_OK;
}

extern "C" int bz_ctx_destroy(bz_ctx* ctx) {
    if (!ctx) return BZ_OK;
    cudaSetDevice(ctx->device);
    bz_comm_destroy(ctx);
    for (auto ev : ctx->ev_pool) cudaEventDestroy(ev);
    if (ctx->batch_arena) cudaFree(ctx->batch_arena);
    if (ctx->xpose_buf) cudaFree(ctx->xpose_buf);
    bz_pool_release(ctx);
    for (void* q : ctx->pinned_pool) cudaFreeHost(q);
    if (ctx->timer0) { cudaEventDestroy(ctx->timer0); cudaEventDestroy(ctx->timer1); }
    if (ctx->solve_ev0) { cudaEventDestroy(ctx->solve_ev0); cudaEventDestroy(ctx->solve_ev1); }
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return BZ_OK;
}

extern "C" int bz_ctx_set_profiling(bz_ctx* ctx, int on) {
    if (!ctx) return BZ_ERR_ARG;
    ctx->profiling = on != 0;
    return BZ_OK;
}

extern "C" int bz_ctx_sync(bz_ctx* ctx) {
    if (!ctx) return BZ_ERR_ARG;
    BZ_CUDA(ctx, cudaSetDevice(ctx->device));
    BZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BZ_OK;
}

extern "C" int bz_ctx_timer_start(bz_ctx* ctx) {
    if (!ctx) return BZ_ERR_ARG;
    BZ_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->timer0) { BZ_CUDA(ctx, cudaEventCreate(&ctx->timer0)); BZ_CUDA(ctx, cudaEventCreate(&ctx->timer1)); }
    BZ_CUDA(ctx, cudaEventRecord(ctx->timer0, ctx->stream));
    return BZ_OK;
}
extern "C" int bz_ctx_timer_stop(bz_ctx* ctx, double* ms) {
    if (!ctx || !ms || !ctx->timer0) return BZ_ERR_ARG;
    BZ_CUDA(ctx, cudaSetDevice(ctx->device));
    BZ_CUDA(ctx, cudaEventRecord(ctx->timer1, ctx->stream));
    BZ_CUDA(ctx, cudaEventSynchronize(ctx->timer1));
    float t = 0.f;
    BZ_CUDA(ctx, cudaEventElapsedTime(&t, ctx->timer0, ctx->timer1));
    *ms = t;
    return BZ_OK;
}

extern "C" void bz_default_options(bz_options* o) {
    // alps.jl:14-25 and SURVEY Appendix A.1
    o->tol = 1e-6; o->tol_prim = -1.0; o->tol_dual = -1.0; o->inner_tol = -1.0;
    o->maxit = 100;
    o->theta_penalty = 0.8; o->kappa_penalty = 0.5; o->kappa_tol = 0.1;
    o->subsolver_maxit = 1000000000LL;
    o->alpha = 0.95; o->beta = 0.5; o->minimum_gamma = 1e-7;
    o->max_backtracks = 20; o->lbfgs_memory = 5;
    o->sub_maxit = 1000;
    o->max_inner_total = 0;
    o->verbose = 0;
    o->dual_safeguard = nullptr;        // default_dual_safeguard! on the device
    o->dual_safeguard_user = nullptr;
}

static cudaEvent_t ctx_event(bz_ctx* ctx) {
    if (ctx->ev_used == ctx->ev_pool.size()) {
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        ctx->ev_pool.push_back(ev);
    }
    return ctx->ev_pool[ctx->ev_used++];
}

// ---------------------------------------------------------------------------------------------
// problems
// ---------------------------------------------------------------------------------------------
template <typename T>
static int dev_alloc(bz_ctx* ctx, T** ptr, size_t count) {
    *ptr = nullptr;
    if (count == 0) count = 1;
    BZ_CUDA(ctx, bz_pool_alloc(ctx, (void**)ptr, count * sizeof(T)));
    BZ_CUDA(ctx, cudaMemsetAsync(*ptr, 0, count * sizeof(T), ctx->stream));
    return BZ_OK;
}
#define BZ_TRY(x) do { int rc__ = (x); if (rc__ != BZ_OK) return rc__; } while (0)

static int upload_matrix(bz_ctx* ctx, const double* host, int layout, long long rows, long long cols,
                         double* dst_rowmajor) {
    const size_t bytes = (size_t)rows * cols * sizeof(double);
    if (layout == BZ_ROW_MAJOR) {
        BZ_CUDA(ctx, cudaMemcpyAsync(dst_rowmajor, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    } else {
        // column-major (Julia) input: staged in a grow-only buffer owned by the context and transposed on the device
        if (ctx->xpose_bytes < bytes) {
            BZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (ctx->xpose_buf) cudaFree(ctx->xpose_buf);
            ctx->xpose_buf = nullptr; ctx->xpose_bytes = 0;
            BZ_CUDA(ctx, cudaMalloc(&ctx->xpose_buf, bytes));
            ctx->xpose_bytes = bytes;
        }
        double* tmp = static_cast<double*>(ctx->xpose_buf);
        BZ_CUDA(ctx, cudaMemcpyAsync(tmp, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
        BZ_TRY(bz_launch_transpose(ctx, tmp, dst_rowmajor, rows, cols));
    }
    return BZ_OK;
}

static void dense_partition(bz_ctx* ctx, long long rows, int* clusters, int* rpc) {
    int maxc = ctx->num_sms / 2;
    if (maxc < 1) maxc = 1;
    long long r = (rows + maxc - 1) / maxc;
    if (r < 1) r = 1;
    *rpc = (int)r;
    *clusters = (int)((rows + r - 1) / r);
    if (*clusters < 1) *clusters = 1;
}

extern "C" int bz_problem_destroy(bz_problem* p) {
    if (!p) return BZ_OK;
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    bz_ctx* ctx = p->ctx;
    if (p->own_cA) bz_pool_free(ctx, p->cA);
    if (p->slab_base) {   // x, z, xeps, yupd, mu, 1 ./ mu, mu .* y live inside the exported slab
        p->v.x = p->v.z = p->v.xeps = p->v.x_prev = p->v.z_curr = nullptr;
        p->yupd = p->mu = p->rmu = p->muy = nullptr;
        for (int r = 0; r < XR_MAXR; ++r)
            if (p->peer_ptr[r]) cudaIpcCloseMemHandle(p->peer_ptr[r]);
        ctx->pool_live.erase(p->slab_base);
        cudaFree(p->slab_base);      // (exported through cudaIpc: never recycled)
    }
    double* frees[] = {p->rmu, p->tm1_f, p->cb, p->fA, p->fb, p->f_ones, p->f_zeros, p->g_u, p->mu, p->y, p->muy, p->cx,
                       p->s, p->yupd, p->tm1, p->tm2, p->tm3, p->v.x, p->v.z, p->v.xeps, p->v.x_prev,
                       p->v.x_d, p->v.z_curr, p->v.res, p->v.res_prev, p->v.grad_x, p->v.grad_z,
                       p->v.d, p->v.sM, p->v.yM, p->Ec, p->Ef, p->Pc, p->Pf, p->lxpart, p->fxpart,
                       p->part, p->team.slots};
    for (double* f : frees) bz_pool_free(ctx, f);
    bz_pool_free(ctx, p->counters);
    bz_pool_free(ctx, p->gridD);
    bz_pool_free(ctx, p->gseq);
    bz_pool_free(ctx, p->lspec_buf);
    bz_pool_free(ctx, p->v.xs);
    bz_pool_free(ctx, p->gridD_f);
    bz_pool_free(ctx, p->gseq_f);
    bz_pool_free(ctx, p->team.bars);
    bz_pool_free(ctx, p->st);
    if (p->st_host) ctx->pinned_pool.push_back(p->st_host);
    if (p->status_host) ctx->pinned_pool.push_back(p->status_host);
    for (auto& e : p->poll_ev) if (e) cudaEventDestroy(e);
    bz_shard_free(p->ctx, &p->shard);
    delete p;
    return BZ_OK;
}

extern "C" int bz_problem_create(bz_ctx* ctx, const bz_problem_desc* d, bz_problem** out) {
    if (!ctx || !d || !out) return bz_fail(ctx, BZ_ERR_ARG, "bz_problem_create: NULL argument");
    *out = nullptr;
    BZ_CUDA(ctx, cudaSetDevice(ctx->device));
    if (d->nx <= 0 || d->ny <= 0) return bz_fail(ctx, BZ_ERR_ARG, "nx and ny must be positive");
    bz_problem* p = new bz_problem();
    p->ctx = ctx;
    p->desc = *d;
    p->n = d->nx;
    p->m = d->ny;
    p->ml = d->ny_local > 0 ? d->ny_local : d->ny;
    // row-sharded: a communicator exists (bz_comm_init; one rank is allowed and exercises the same code path)
    p->sharded = d->ny_local > 0 && ctx->nccl_comm != nullptr;
    auto fail = [&](int code, const std::string& msg) {
        bz_problem_destroy(p);
        return bz_fail(ctx, code, msg);
    };
    const long long n = p->n, ml = p->ml;

    // ---- catalogue check (H2: closed catalogue, unknown combinations raise; no CPU fallback)
    const bool d_pair = bz_d_is_pair(d->d_kind);
    if (d->d_kind == BZ_D_PORTFOLIO && (d->ny != 2 || d->c_kind != BZ_C_DENSE_AFFINE || d->ny_local != 0))
        return fail(BZ_ERR_ARG, "SetP (portfolio) is [rho, inf) x {1}: it needs ny == 2 and an unsharded dense affine c");
    switch (d->c_kind) {
    case BZ_C_DENSE_AFFINE:
        if (d_pair) return fail(BZ_ERR_UNSUPPORTED, "pair sets with a dense affine c are not in the fused path yet");
        if (!(d->f_kind == BZ_F_ZERO || d->f_kind == BZ_F_LEAST_SQUARES || d->f_kind == BZ_F_QUADRATIC))
            return fail(BZ_ERR_UNSUPPORTED, "f kind not available with a dense affine c");
        break;
    case BZ_C_IDENTITY:
        if (d->ny != d->nx) return fail(BZ_ERR_ARG, "identity c needs ny == nx");
        if (!(d->f_kind == BZ_F_ZERO || d->f_kind == BZ_F_LEAST_SQUARES || d->f_kind == BZ_F_QUADRATIC))
            return fail(BZ_ERR_UNSUPPORTED, "f kind not available with the identity c");
        if (d_pair && d->d_pair_stride != 1 && d->d_pair_stride * 2 != d->ny)
            return fail(BZ_ERR_ARG, "pair stride must be 1 or ny/2");
        break;
    case BZ_C_OBSTACLE:
        if (!((d->f_kind == BZ_F_OBSTACLE_L1 && d->g_kind == BZ_G_OBSTACLE_L1) ||
              (d->f_kind == BZ_F_OBSTACLE_L2 && d->g_kind == BZ_G_OBSTACLE_L2)) || d->d_kind != BZ_D_ZERO)
            return fail(BZ_ERR_UNSUPPORTED, "obstacle c is built for the obstacle_l1 and obstacle_l2 variants (D = {0})");
        if (d->nx != 3 * d->ny) return fail(BZ_ERR_ARG, "obstacle: nx must be 3*ny");
        if (d->grid_y * d->grid_x != 0 && d->grid_y * d->grid_x != d->ny)
            return fail(BZ_ERR_ARG, "obstacle: grid_y*grid_x must equal ny");
        break;
    case BZ_C_OBSTACLE_RED:
        if (!((d->f_kind == BZ_F_OBSTACLE_L1 && d->g_kind == BZ_G_OBSTACLE_RED_L1) ||
              (d->f_kind == BZ_F_OBSTACLE_L2 && d->g_kind == BZ_G_NONNEG)) || d->d_kind != BZ_D_CC)
            return fail(BZ_ERR_UNSUPPORTED, "reduced obstacle c is built for the obstacle_l1red and obstacle_l2red variants (D = CC pairs)");
        if (d->nx != d->ny || d->nx % 2) return fail(BZ_ERR_ARG, "obstacle_l1red: nx = ny = 2N");
        if (d->d_pair_stride * 2 != d->ny) return fail(BZ_ERR_ARG, "obstacle_l1red: the CC pairs are (i, i+N)");
        if (d->grid_y * d->grid_x != 0 && d->grid_y * d->grid_x * 2 != d->ny)
            return fail(BZ_ERR_ARG, "obstacle_l1red: grid_y*grid_x must equal N");
        break;
    default:
        return fail(BZ_ERR_UNSUPPORTED,
                    "c kind is only available in the batched tiny-instance path (bz_batch_solve)");
    }
    p->nv = p->n;
    if (d->slack_formulation) {   // als: PANOC+ on [x; s]
        if (!(d->c_kind == BZ_C_DENSE_AFFINE || d->c_kind == BZ_C_IDENTITY) || d_pair || d->ny_local > 0 || d->slab_nranks > 1)
            return fail(BZ_ERR_UNSUPPORTED, "als (slack formulation) is built for dense-affine / identity c with an elementwise D on one GPU");
        p->slack = true;
        p->nv = p->n + p->m;
    }
    if (d->slab_nranks > 1) {
        if (d->c_kind != BZ_C_OBSTACLE || d->grid_x <= 0 || d->grid_y <= 0)
            return fail(BZ_ERR_UNSUPPORTED, "slab decomposition is built for the 2-D obstacle grid only");
        if (d->slab_nranks > XR_MAXR || d->slab_rank < 0 || d->slab_rank >= d->slab_nranks)
            return fail(BZ_ERR_ARG, "slab: bad rank / more than 8 ranks");
        if (d->nx_global != 3 * d->ny * (long long)d->slab_nranks)
            return fail(BZ_ERR_ARG, "slab: nx_global must be 3 * ny_local * slab_nranks (equal slabs)");
        p->slab = true;
        p->n_global = d->nx_global;
    }
    if ((d->g_kind == BZ_G_OBSTACLE_L1 || d->g_kind == BZ_G_OBSTACLE_L2) && d->c_kind != BZ_C_OBSTACLE)
        return fail(BZ_ERR_UNSUPPORTED, "obstacle g needs the obstacle c");
    if (d->g_kind == BZ_G_OBSTACLE_RED_L1 && d->c_kind != BZ_C_OBSTACLE_RED)
        return fail(BZ_ERR_UNSUPPORTED, "reduced obstacle g needs the reduced obstacle c");
    if ((d->g_kind == BZ_G_NORML1_BOX || d->g_kind == BZ_G_NORML0_BOX)) {
        if (!d->g_u) return fail(BZ_ERR_ARG, "g_u is required for the *_BOX prox kinds");
        if (d->g_lambda < 0) return fail(BZ_ERR_ARG, "parameter λ must be nonnegative");
        for (long long j = 0; j < n; ++j)
            if (d->g_u[j] < 0) return fail(BZ_ERR_ARG, "vector u must have nonnegative entries");
    }
    if (d->g_kind == BZ_G_NORML1_NONNEG && d->g_lambda < 0)
        return fail(BZ_ERR_ARG, "λ must be nonnegative");
    if (d->g_kind == BZ_G_NORMLP_BOX || d->g_kind == BZ_G_NORMLP_NONNEG) {   // normLpBox.jl:16-29, normLpNonneg.jl:15-24
        if (d->g_lo <= 0) return fail(BZ_ERR_ARG, "p must be positive");
        if (d->g_lo >= 1) return fail(BZ_ERR_ARG, "p must be smaller than one");
        if (d->g_lambda < 0) return fail(BZ_ERR_ARG, "alpha must be nonnegative");
        if (d->g_kind == BZ_G_NORMLP_BOX) {
            if (!d->g_u) return fail(BZ_ERR_ARG, "g_u is required for the *_BOX prox kinds");
            for (long long j = 0; j < n; ++j)
                if (d->g_u[j] < 0) return fail(BZ_ERR_ARG, "vector u must have nonnegative entries");
        }
    }

    int rc = BZ_OK;
#define PTRY(x) do { rc = (x); if (rc != BZ_OK) { bz_problem_destroy(p); return rc; } } while (0)
    // ---- parameters
    p->g.kind = d->g_kind; p->g.lambda = d->g_lambda; p->g.lo = d->g_lo; p->g.hi = d->g_hi;
    p->g.u = nullptr; p->g.N = d->ny;
    p->D.kind = d->d_kind; p->D.lo = d->d_lo; p->D.hi = d->d_hi;
    p->D.stride = d->d_pair_stride > 0 ? d->d_pair_stride : 1;
    p->ob.N = d->c_kind == BZ_C_OBSTACLE_RED ? d->ny / 2 : d->ny; p->ob.gy = d->grid_y; p->ob.gx = d->grid_x;
    p->ob.l2 = d->f_kind == BZ_F_OBSTACLE_L2 ? 1 : 0;
    if (d->c_kind == BZ_C_OBSTACLE_RED) p->g.N = d->ny / 2;
    if (d->g_u) {
        PTRY(dev_alloc(ctx, &p->g_u, n));
        if (cudaMemcpyAsync(p->g_u, d->g_u, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
            return fail(BZ_ERR_CUDA, "copy of g_u to the device failed");
        p->g.u = p->g_u;
    }
    // ---- vectors
    if (p->slab) {
        // one exportable allocation: mailbox ring, sequence word, and the vectors whose boundary rows the
        // neighbouring ranks read through the peer mapping (same layout on every rank)
        const size_t head = (size_t)XR_MAIL_WORDS + 8;
        // (mu, 1 ./ mu, mu .* y too: the fused forward-backward kernel recomputes a radius-4 neighbourhood from raw rows)
        // (x_prev and z_curr too: with buffer rotation they take the roles of x and z on every other iteration)
        p->slab_bytes = (head + 5 * (size_t)n + 4 * (size_t)ml) * sizeof(double);
        PTRY(dev_alloc(ctx, &p->slab_base, p->slab_bytes / sizeof(double)));
        double* vecs = reinterpret_cast<double*>(p->slab_base + head);
        p->v.x = vecs; p->v.z = vecs + n; p->v.xeps = vecs + 2 * n; p->yupd = vecs + 3 * n;
        p->mu = vecs + 3 * n + ml; p->rmu = vecs + 3 * n + 2 * ml; p->muy = vecs + 3 * n + 3 * ml;
        p->v.x_prev = vecs + 3 * n + 4 * ml; p->v.z_curr = vecs + 4 * n + 4 * ml;
    }
    double** nvecs[] = {&p->v.x, &p->v.z, &p->v.xeps, &p->v.x_prev, &p->v.x_d, &p->v.z_curr, &p->v.res,
                        &p->v.res_prev, &p->v.grad_x, &p->v.grad_z, &p->v.d, &p->Ec, &p->Ef};
    for (double** q : nvecs) if (!*q) PTRY(dev_alloc(ctx, q, p->nv));
    // L-BFGS history: BZ_MAXMEM ring slots + one spare (the cooperative L kernel writes the candidate pair there)
    PTRY(dev_alloc(ctx, &p->v.sM, (size_t)(BZ_MAXMEM + 1) * p->nv));
    PTRY(dev_alloc(ctx, &p->v.yM, (size_t)(BZ_MAXMEM + 1) * p->nv));
    p->v.n = p->nv;
    double** mvecs[] = {&p->mu, &p->y, &p->muy, &p->cx, &p->s, &p->yupd, &p->tm1, &p->tm2, &p->tm3};
    for (double** q : mvecs) if (!*q) PTRY(dev_alloc(ctx, q, ml));
    const int maxgrid = ctx->num_sms * 8;
    size_t npart = (size_t)4 * maxgrid;
    if (d->c_kind == BZ_C_OBSTACLE && (!p->slab || ctx->slab_fused) && !p->slack && ctx->obstacle_fused && p->nv < (1LL << 31)) {   // (32-bit element indices)
        p->fused = true;
        // (n <= 32768 runs the cluster L kernel, whose vectors live in registers.  In slab mode the ranks rotate in lock-step
        // and both buffers of a pair are inside the exported slab, so a neighbour finds the current x at the same offset.)
        p->rotate = p->nv > 32768 && ctx->obstacle_rotate;
        npart = std::max(npart, (size_t)bz_obfused_partials(ctx, p->ob));
        if (!p->rmu) PTRY(dev_alloc(ctx, &p->rmu, ml));
    }
    PTRY(dev_alloc(ctx, &p->part, npart));
    PTRY(dev_alloc(ctx, &p->lxpart, (size_t)maxgrid));
    PTRY(dev_alloc(ctx, &p->fxpart, (size_t)2 * maxgrid));
    PTRY(dev_alloc(ctx, &p->counters, 8));
    PTRY(dev_alloc(ctx, &p->team.bars, 2 * BZ_MAXBAR));
    p->team.max_ctas = ctx->num_sms * BZ_L_CTAS_PER_SM;
    PTRY(dev_alloc(ctx, &p->team.slots, (size_t)2 * BZ_MAXBAR * p->team.max_ctas * 4));
    p->team.parity = 0;
    PTRY(dev_alloc(ctx, &p->st, 1));
#define PCUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return fail(BZ_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); } while (0)
    static_assert(sizeof(BzState) <= BZ_PINNED_BLOCK, "pinned blocks hold a BzState");
    for (void** q : {(void**)&p->st_host, (void**)&p->status_host}) {
        if (!ctx->pinned_pool.empty()) { *q = ctx->pinned_pool.back(); ctx->pinned_pool.pop_back(); }
        else PCUDA(cudaMallocHost(q, BZ_PINNED_BLOCK));
    }
    for (auto& e : p->poll_ev) PCUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));

    // ---- f data
    if (d->f_kind == BZ_F_LEAST_SQUARES) {
        if (!d->f_A || !d->f_b || d->f_m <= 0) return fail(BZ_ERR_ARG, "LeastSquares needs f_A, f_b, f_m");
        p->fm = d->f_m;
        PTRY(dev_alloc(ctx, &p->fA, (size_t)p->fm * n));
        PTRY(upload_matrix(ctx, d->f_A, d->f_layout, p->fm, n, p->fA));
        PTRY(dev_alloc(ctx, &p->fb, p->fm));
        PCUDA(cudaMemcpyAsync(p->fb, d->f_b, p->fm * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        PTRY(dev_alloc(ctx, &p->f_ones, p->fm));
        PTRY(dev_alloc(ctx, &p->f_zeros, p->fm));
        PTRY(dev_alloc(ctx, &p->tm1_f, p->fm));
        PTRY(bz_launch_fill(ctx, p->f_ones, p->fm, 1.0));
        p->f_wide = n > DA_MAX_N || ctx->force_wide;
        BzGridGeom gg;
        if (bz_grid_geometry(ctx, p->fA, p->fm, n, n, &gg)) { p->f_grid = true; p->f_clusters = p->f_nlx = gg.T; }
        else if (p->f_wide) bz_wide_geometry(ctx, p->fm, n, &p->f_nlx, &p->f_clusters, &p->f_rpc);
        else { dense_partition(ctx, p->fm, &p->f_clusters, &p->f_rpc); p->f_nlx = p->f_clusters; }
        PTRY(dev_alloc(ctx, &p->Pf, (size_t)p->f_clusters * n));
    } else if (d->f_kind == BZ_F_QUADRATIC) {
        if (!d->f_A || !d->f_b) return fail(BZ_ERR_ARG, "Quadratic needs f_A (Q) and f_b (q)");
        PTRY(dev_alloc(ctx, &p->fA, (size_t)n * n));
        PTRY(upload_matrix(ctx, d->f_A, d->f_layout, n, n, p->fA));
        PTRY(dev_alloc(ctx, &p->fb, n));
        PCUDA(cudaMemcpyAsync(p->fb, d->f_b, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    // ---- c data
    if (d->c_kind == BZ_C_DENSE_AFFINE) {
        if (!d->c_b) return fail(BZ_ERR_ARG, "dense affine c needs c_b");
        PTRY(dev_alloc(ctx, &p->cb, ml));
        PCUDA(cudaMemcpyAsync(p->cb, d->c_b, ml * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        if (d->c_A_is_device) {
            if (!d->c_A) return fail(BZ_ERR_ARG, "c_A_is_device set but c_A is NULL");
            p->cA = const_cast<double*>(d->c_A);
            p->own_cA = false;
        } else {
            PTRY(dev_alloc(ctx, &p->cA, (size_t)ml * n));
            p->own_cA = true;
            if (d->c_gen) {
                PTRY(bz_launch_randn_fill(ctx, d->c_gen_seed, 0,
                                          (uint64_t)d->row_begin * (uint64_t)n, ml * n, p->cA));
            } else {
                if (!d->c_A) return fail(BZ_ERR_ARG, "dense affine c needs c_A");
                PTRY(upload_matrix(ctx, d->c_A, d->c_layout, ml, n, p->cA));
            }
        }
        p->c_wide = n > DA_MAX_N || ctx->force_wide;
        BzGridGeom gg;
        if (bz_grid_geometry(ctx, p->cA, ml, n, n, &gg)) { p->c_grid = true; p->c_clusters = p->c_nlx = gg.T; }
        else if (p->c_wide) bz_wide_geometry(ctx, ml, n, &p->c_nlx, &p->c_clusters, &p->c_rpc);
        else { dense_partition(ctx, ml, &p->c_clusters, &p->c_rpc); p->c_nlx = p->c_clusters; }
        PTRY(dev_alloc(ctx, &p->Pc, (size_t)p->c_clusters * n));
        if (p->sharded) PTRY(bz_shard_alloc(ctx, &p->shard, n));
        // speculative L step beside G(z): plain alps on one GPU with the cluster kernels (n <= 16384) and f = Zero
        if (ctx->lspec && !p->c_wide && !p->c_grid && !p->sharded && !p->slack && d->f_kind == BZ_F_ZERO &&
            d->g_kind != BZ_G_OBSTACLE_L1 && d->g_kind != BZ_G_OBSTACLE_L2 &&     // (their F step is not the elementwise one that commits)
            p->nv <= (long long)LS_MAXCS * LS_THREADS * 16) {
            int cs = 1;
            while (cs < LS_MAXCS && (long long)cs * LS_THREADS * 8 < p->nv) cs *= 2;     // the cluster size of k_lstep_cluster
            const int maxc = ctx->num_sms / 2, extra = (cs + 1) / 2;
            if (maxc > extra) {
                p->lspec = true;
                p->lspec_cs = cs;
                p->c_clusters_z = p->c_clusters;
                p->c_rpc_z = p->c_rpc;
                if (p->c_clusters + extra > maxc) {          // give up a few row clusters so that the whole grid stays co-resident
                    const int want = maxc - extra;
                    p->c_rpc_z = (int)((ml + want - 1) / want);
                    p->c_clusters_z = (int)((ml + p->c_rpc_z - 1) / p->c_rpc_z);
                }
                PTRY(dev_alloc(ctx, &p->v.xs, p->nv));
                PTRY(dev_alloc(ctx, &p->lspec_buf, (size_t)32 * LS_MAXCS * 2 * 2));
            }
        }
    }
    {   // a usable L-BFGS state before the first solve (the parity entry points may be called right away)
        memset(p->st_host, 0, sizeof(BzState));
        p->st_host->next_op = OP_DONE;
        p->st_host->H = 1.0;
        for (int i = 0; i < BZ_MAXMEM; ++i) p->st_host->slot[i] = i;
        p->st_host->spare = BZ_MAXMEM;
        PCUDA(cudaMemcpyAsync(p->st, p->st_host, sizeof(BzState), cudaMemcpyHostToDevice, ctx->stream));
    }
    // the LeastSquares pass and the constraint pass may pick different team geometries: each has its own exchange
    // ring, arrival counters and row-sequence bases
    if (p->c_grid) {
        PTRY(dev_alloc(ctx, &p->gridD, (size_t)BZ_GRID_EXCHANGE_WORDS));
        PTRY(dev_alloc(ctx, &p->gseq, 160));   // one row-sequence base per team
    }
    if (p->f_grid) {
        PTRY(dev_alloc(ctx, &p->gridD_f, (size_t)BZ_GRID_EXCHANGE_WORDS));
        PTRY(dev_alloc(ctx, &p->gseq_f, 160));
    }
#undef PTRY
    PCUDA(cudaStreamSynchronize(ctx->stream));
#undef PCUDA
    *out = p;
    return BZ_OK;
}

extern "C" int bz_problem_upload_dense(bz_problem* p, const double* c_A, const double* c_b) {
    if (!p || !c_b) return BZ_ERR_ARG;
    bz_ctx* ctx = p->ctx;
    BZ_CUDA(ctx, cudaSetDevice(ctx->device));
    if (p->desc.c_kind != BZ_C_DENSE_AFFINE)
        return bz_fail(ctx, BZ_ERR_ARG, "bz_problem_upload_dense: problem has no dense affine c");
    if (c_A) {   // c_A == NULL: keep the matrix (adopted or generated on the device), replace b only
        if (!p->own_cA) return bz_fail(ctx, BZ_ERR_ARG, "bz_problem_upload_dense: problem does not own its dense A");
        BZ_TRY(upload_matrix(ctx, c_A, p->desc.c_layout, p->ml, p->n, p->cA));
    }
    BZ_CUDA(ctx, cudaMemcpyAsync(p->cb, c_b, p->ml * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    BZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // the caller owns c_A / c_b again when this returns
    return BZ_OK;
}

static int fetch_state(bz_problem* p);
static int push_state(bz_problem* p);

extern "C" int bz_problem_xr_handle(bz_problem* p, void* handle64) {
    if (!p || !handle64) return BZ_ERR_ARG;
    bz_ctx* ctx = p->ctx;
    BZ_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!p->slab) return bz_fail(ctx, BZ_ERR_ARG, "bz_problem_xr_handle: not a slab-decomposed problem");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t hnd;
    BZ_CUDA(ctx, cudaIpcGetMemHandle(&hnd, p->slab_base));
    memcpy(handle64, &hnd, 64);
    return BZ_OK;
}

extern "C" int bz_problem_xr_attach(bz_problem* p, const void* handles64) {
    if (!p || !handles64) return BZ_ERR_ARG;
    bz_ctx* ctx = p->ctx;
    BZ_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!p->slab) return bz_fail(ctx, BZ_ERR_ARG, "bz_problem_xr_attach: not a slab-decomposed problem");
    const int R = p->desc.slab_nranks, me = p->desc.slab_rank;
    const size_t head = (size_t)XR_MAIL_WORDS + 8;
    BzXr xr{};
    xr.nranks = R; xr.rank = me;
    for (int r = 0; r < R; ++r) {
        unsigned long long* base = p->slab_base;
        if (r != me) {
            cudaIpcMemHandle_t hnd;
            memcpy(&hnd, (const char*)handles64 + 64 * (size_t)r, 64);
            void* ptr = nullptr;
            BZ_CUDA(ctx, cudaIpcOpenMemHandle(&ptr, hnd, cudaIpcMemLazyEnablePeerAccess));
            p->peer_ptr[r] = ptr;
            base = reinterpret_cast<unsigned long long*>(ptr);
        }
        xr.mail[r] = base;
        xr.slab[r] = reinterpret_cast<double*>(base);
    }
    xr.seq = p->slab_base + XR_MAIL_WORDS;
    (void)head;
    p->xr_host = xr;
    BZ_TRY(fetch_state(p));
    p->st_host->xr = xr;
    p->st_host->n_global = p->n_global;
    BZ_TRY(push_state(p));
    BZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BZ_OK;
}

// ---------------------------------------------------------------------------------------------
// AL-gradient provider: enqueue the kernels that evaluate gradient!(dlx, alFun, .) at the target
// vector (auglagfun.jl:73-86).  `force_target` < 0: predicated on the state machine.
// ---------------------------------------------------------------------------------------------
static BzGradSrc grad_src(bz_problem* p) {
    BzGradSrc gs{};
    const bz_problem_desc& d = p->desc;
    gs.fx_scale = 1.0; gs.fx_scale2 = 0.0;
    if (d.f_kind == BZ_F_LEAST_SQUARES) {
        gs.Pf = p->Pf; gs.npf = p->f_clusters;
        gs.fxpart = p->fxpart; gs.nfx = p->f_nlx; gs.fx_scale = 0.5;
    } else if (d.f_kind == BZ_F_QUADRATIC) {
        gs.Ef = p->Ef;
        gs.fxpart = p->fxpart; gs.nfx = p->nfx; gs.fx_scale = 1.0;
    } else if (d.f_kind == BZ_F_OBSTACLE_L1 || d.f_kind == BZ_F_OBSTACLE_L2) {
        gs.fxpart = p->fxpart; gs.nfx = p->nfx; gs.fx_scale = 0.5;
        gs.fxpart2 = p->fxpart + p->nfx; gs.nfx2 = p->nfx; gs.fx_scale2 = -1.0;
    }
    if (d.c_kind == BZ_C_DENSE_AFFINE) {
        if (p->sharded) { gs.Ec = p->shard.reduced; gs.lxpart = p->shard.reduced + p->n; gs.nlx = 1; }
        else { gs.Pc = p->Pc; gs.npc = p->c_clusters; gs.lxpart = p->lxpart; gs.nlx = p->c_nlx; }
    } else if (d.c_kind == BZ_C_IDENTITY) {
        gs.Ec = p->yupd; gs.lxpart = p->lxpart; gs.nlx = p->nlx;
    } else {
        gs.Ec = p->Ec; gs.lxpart = p->lxpart; gs.nlx = p->nlx;
    }
    if (p->slack_active) { gs.nx = p->n; gs.tail = p->yupd; }   // als: gradient = [dfx + jtv; -yupd]
    return gs;
}

static int launch_provider(bz_problem* p, int force_target, bool gz_slot = false) {
    bz_ctx* ctx = p->ctx;
    const bz_problem_desc& d = p->desc;
    const double* srcbuf[3];
    const double* const* src = src_ptrs(p, srcbuf);
    // ---- smooth cost f
    if (d.f_kind == BZ_F_LEAST_SQUARES) {
        BzDenseArgs a{};
        a.A = p->fA; a.lda = p->n; a.b = p->fb; a.mu = p->f_ones; a.muy = p->f_zeros;
        a.src[0] = src[0]; a.src[1] = src[1]; a.src[2] = src[2];
        a.st = p->st; a.force_target = force_target; a.m = p->fm; a.n = p->n;
        a.rows_per_cluster = p->f_rpc; a.dkind = BZ_D_ZERO; a.dlo = 0; a.dhi = 0;
        a.cx = nullptr; a.s = nullptr; a.yupd = nullptr;
        a.P = p->Pf; a.lxpart = p->fxpart; a.counter = p->counters + 0; a.set_after = 0;
        a.l2_prefetch_rows = ctx->l2_prefetch_rows;
        a.evict_first = ctx->evict_first;
        a.D = p->gridD_f; a.gseq = p->gseq_f;
        if (p->f_grid) BZ_TRY(bz_launch_grid_algrad(ctx, a));
        else if (p->f_wide) { a.yupd = p->tm1_f; BZ_TRY(bz_launch_dense_wide(ctx, a)); }
        else BZ_TRY(bz_launch_dense_algrad(ctx, a, p->f_clusters));
    } else if (d.f_kind == BZ_F_QUADRATIC) {
        p->nfx = std::min<long long>(ctx->num_sms, (p->n + 7) / 8);
        if (p->nfx < 1) p->nfx = 1;
        BZ_TRY(bz_launch_quad_grad(ctx, p->fA, p->fb, p->st, force_target, src, p->n, p->Ef,
                                   p->fxpart, p->nfx));
    }
    // ---- constraint part
    if (d.c_kind == BZ_C_DENSE_AFFINE) {
        BzDenseArgs a{};
        a.A = p->cA; a.lda = p->n; a.b = p->cb; a.mu = p->mu; a.muy = p->muy;
        a.src[0] = src[0]; a.src[1] = src[1]; a.src[2] = src[2];
        a.st = p->st; a.force_target = force_target; a.m = p->ml; a.n = p->n;
        a.rows_per_cluster = p->c_rpc; a.dkind = p->D.kind; a.dlo = p->D.lo; a.dhi = p->D.hi;
        a.cx = p->cx; a.s = p->s; a.yupd = p->yupd;
        a.P = p->Pc; a.lxpart = p->lxpart; a.counter = p->counters + 1; a.set_after = 1;
        a.l2_prefetch_rows = ctx->l2_prefetch_rows;
        a.evict_first = ctx->evict_first;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        const bool timed = ctx->profiling && (p->dom_counter++ % ctx->prof_stride == 0);
        if (timed) { e0 = ctx_event(ctx); e1 = ctx_event(ctx); cudaEventRecord(e0, ctx->stream); }
        a.D = p->gridD; a.gseq = p->gseq;
        a.slack = p->slack_active ? 1 : 0; a.y = p->y;
        a.is_cpass = 1;
        a.x_shadow = p->lspec ? p->v.xs : nullptr;
        int nclus = p->c_clusters;
        if (gz_slot && p->lspec && force_target < 0) {       // the second provider launch of a cycle: G(z) + speculative L
            a.ls.enabled = 1; a.ls.dense_clusters = p->c_clusters_z; a.ls.cs = p->lspec_cs; a.ls.v = p->v; a.ls.xs = p->v.xs; a.ls.xbuf = p->lspec_buf;
            a.rows_per_cluster = p->c_rpc_z;
            nclus = p->c_clusters_z;
        }
        if (p->c_grid) BZ_TRY(bz_launch_grid_algrad(ctx, a));
        else if (p->c_wide) BZ_TRY(bz_launch_dense_wide(ctx, a));
        else BZ_TRY(bz_launch_dense_algrad(ctx, a, nclus));
        if (timed) { cudaEventRecord(e1, ctx->stream); p->dom_events.push_back({e0, e1}); }
        if (p->sharded)
            BZ_TRY(bz_shard_allreduce(ctx, &p->shard, p->Pc, p->c_clusters, p->lxpart, p->c_nlx, p->n));
    } else if (d.c_kind == BZ_C_IDENTITY) {
        p->nlx = bz_vec_grid(ctx, bz_d_is_pair(p->D.kind) ? p->n / 2 : p->n);
        BZ_TRY(bz_launch_identity_c(ctx, p->st, force_target, src, p->n, p->mu, p->muy, p->D, p->cx,
                                    p->s, p->yupd, p->lxpart, p->counters + 1, 1, p->slack_active ? 1 : 0, p->y));
    } else if (d.c_kind == BZ_C_OBSTACLE) {
        p->nlx = p->nfx = bz_obstacle_partials(ctx, p->ob);
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        const bool timed = ctx->profiling && (p->dom_counter++ % ctx->prof_stride == 0);
        if (timed) { e0 = ctx_event(ctx); e1 = ctx_event(ctx); cudaEventRecord(e0, ctx->stream); }
        BZ_TRY(bz_launch_obstacle(ctx, p->st, force_target, src, p->ob, p->mu, p->muy, p->cx, p->s,
                                  p->yupd, p->Ec, p->lxpart, p->fxpart, p->counters + 1, 1));
        if (timed) { cudaEventRecord(e1, ctx->stream); p->dom_events.push_back({e0, e1}); }
    } else if (d.c_kind == BZ_C_OBSTACLE_RED) {
        p->nlx = p->nfx = bz_obstacle_partials(ctx, p->ob);
        BZ_TRY(bz_launch_obstacle_red(ctx, p->st, force_target, src, p->ob, p->mu, p->muy, p->cx, p->s,
                                      p->yupd, p->Ec, p->lxpart, p->fxpart, p->counters + 1, 1));
    }
    return BZ_OK;
}

// g parameters as the F step sees them: in als mode the entries >= nx of the stacked variable are projected onto D
static BzGParams g_params(const bz_problem* p) {
    BzGParams g = p->g;
    if (p->slack_active) { g.snx = p->n; g.sdk = p->D.kind; g.slo = p->D.lo; g.shi = p->D.hi; }
    return g;
}

// one pass of the fixed cycle [L, G, F, G, D]; fused problems: [L, X]
static int launch_cycle(bz_problem* p) {
    bz_ctx* ctx = p->ctx;
    BzGradSrc gs = grad_src(p);
    BzTeam tm = p->team;
    if (p->fused) {
        BZ_TRY(bz_launch_lstep(ctx, p->st, p->v, tm, p->counters + 2, -1, p->slab, p->rotate));
        p->team.parity ^= 1;
        return bz_launch_obstacle_fused(ctx, p->st, p->v, p->ob, g_params(p), p->mu, p->rmu, p->muy, p->part, p->counters + 3, 0,
                                        nullptr, nullptr);
    }
    BZ_TRY(bz_launch_lstep(ctx, p->st, p->v, tm, p->counters + 2, -1, p->slab));
    p->team.parity ^= 1;
    BZ_TRY(launch_provider(p, -1));
    gs = grad_src(p);
    BZ_TRY(bz_launch_fstep(ctx, p->st, p->v, gs, g_params(p), p->part, p->counters + 3, -1));
    BZ_TRY(launch_provider(p, -1, true));
    BZ_TRY(bz_launch_dstep(ctx, p->st, p->v, gs, p->part, p->counters + 4));
    return BZ_OK;
}

// PANOCplus on the current (mu, y): x must be in p->v.x.  Returns with the solution in p->v.z and
// the AugLagFun side-effect state at z (Appendix A.7).
static int panoc_run(bz_problem* p, const bz_options* o, double tol, long long maxit) {
    bz_ctx* ctx = p->ctx;
    if (o->lbfgs_memory < 0 || o->lbfgs_memory > BZ_MAXMEM)
        return bz_fail(ctx, BZ_ERR_ARG, "lbfgs_memory must be in [0, 16]");
    if (maxit < 1) maxit = 1;
    BZ_TRY(bz_launch_panoc_begin(ctx, p->st, tol, maxit, o->alpha, o->beta, o->minimum_gamma,
                                 o->max_backtracks, o->lbfgs_memory, p->fused ? 1 : 0));
    if (p->fused) {
        // Appendix A.4 with the unfused kernels, once per inner solve: gradient at x, at x .+ 1, the Lipschitz estimate;
        // from MODE_INITFB on everything is the X kernel
        for (int k = 0; k < 2; ++k) {
            BZ_TRY(launch_provider(p, -1));
            BzGradSrc gs0 = grad_src(p);
            BZ_TRY(bz_launch_fstep(ctx, p->st, p->v, gs0, g_params(p), p->part, p->counters + 3, -1));
        }
    }
    const int cycles_per_batch = 4;
    int slot = 0, inflight = 0;
    p->status_host[0] = p->status_host[1] = IST_RUNNING;
    for (;;) {
        for (int c = 0; c < cycles_per_batch; ++c) BZ_TRY(launch_cycle(p));
        BZ_CUDA(ctx, cudaMemcpyAsync(&p->status_host[slot], &p->st->inner_status, sizeof(int),
                                     cudaMemcpyDeviceToHost, ctx->stream));
        BZ_CUDA(ctx, cudaEventRecord(p->poll_ev[slot], ctx->stream));
        inflight++;
        slot ^= 1;
        if (inflight == 2) {
            BZ_CUDA(ctx, cudaEventSynchronize(p->poll_ev[slot]));   // the older batch
            inflight--;
            if (p->status_host[slot] != IST_RUNNING) break;
        }
    }
    if (p->rotate) {   // buffer rotation: make the host's pointers agree with the roles the device ended with
        BZ_TRY(fetch_state(p));
        const int rot = p->st_host->rot;
        if (rot & 1) std::swap(p->v.x, p->v.x_prev);
        if (rot & 2) std::swap(p->v.res, p->v.res_prev);
        if (rot & 4) std::swap(p->v.z, p->v.z_curr);
        if (rot) BZ_CUDA(ctx, cudaMemsetAsync(&p->st->rot, 0, sizeof(int), ctx->stream));
    }
    // fused problems never materialise c(z), s, yupd inside the loop: one unfused AL-gradient at the returned point
    // leaves the AugLagFun side-effect state the outer step reads (alps.jl:68,74)
    if (p->fused) BZ_TRY(launch_provider(p, TGT_Z));
    BZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BZ_OK;
}

static int fetch_state(bz_problem* p) {
    bz_ctx* ctx = p->ctx;
    BZ_CUDA(ctx, cudaMemcpyAsync(p->st_host, p->st, sizeof(BzState), cudaMemcpyDeviceToHost, ctx->stream));
    BZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BZ_OK;
}
static int push_state(bz_problem* p) {
    bz_ctx* ctx = p->ctx;
    BZ_CUDA(ctx, cudaMemcpyAsync(p->st, p->st_host, sizeof(BzState), cudaMemcpyHostToDevice, ctx->stream));
    return BZ_OK;
}
static int h2d(bz_problem* p, double* dst, const double* src, long long count) {
    BZ_CUDA(p->ctx, cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyHostToDevice, p->ctx->stream));
    return BZ_OK;
}
static int d2h(bz_problem* p, double* dst, const double* src, long long count) {
    if (!dst) return BZ_OK;
    BZ_CUDA(p->ctx, cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyDeviceToHost, p->ctx->stream));
    return BZ_OK;
}
static int d2d(bz_problem* p, double* dst, const double* src, long long count) {
    BZ_CUDA(p->ctx, cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyDeviceToDevice, p->ctx->stream));
    return BZ_OK;
}

static void collect_dominant(bz_problem* p, bz_result* res) {
    // Only every prof_stride-th launch of the dominant kernel is bracketed by CUDA events (an event
    // record between two kernels disables their programmatic overlap, so bracketing all of them
    // would perturb the step).  Launches that returned at the state-machine check take a few
    // microseconds; the ones that really ran are those above 30 % of the longest.
    std::vector<float> ms;
    ms.reserve(p->dom_events.size());
    for (auto& e : p->dom_events) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, e.first, e.second) == cudaSuccess) ms.push_back(t);
    }
    std::sort(ms.begin(), ms.end(), [](float a, float b) { return a > b; });
    double sum = 0.0;
    long long k = 0;
    for (float t : ms)
        if (t > 0.3f * ms[0]) { sum += t; ++k; }
    res->dominant_kernel_ms = sum;
    res->dominant_kernel_launches = k;
    p->dom_events.clear();
    p->ctx->ev_used = 0;
}

// ---------------------------------------------------------------------------------------------
// alps (src/algorithms/alps.jl:7-117) and als (src/algorithms/als.jl:7-120): one outer-loop driver.
// als runs the same loop on the stacked variable [x; s] with AugLagFunSlack / NonsmoothCostFunSlack; the two
// differ in three places only (marked `slack`).
// ---------------------------------------------------------------------------------------------
static int outer_solve(bz_problem* p, bool slack, const double* x0, const double* y0, const bz_options* o,
                       double* x, double* y, double* s, double* mu, bz_result* res) {
    bz_ctx* ctx = p->ctx;
    p->slack_active = false;   // als.jl:38-44: the initial prox / c(x) / proj_D / mu use the plain oracles
    const auto t0 = std::chrono::steady_clock::now();
    const long long launches0 = ctx->n_launches;
    p->dom_events.clear();
    ctx->ev_used = 0;
    const double tol = o->tol;
    const double tol_prim = o->tol_prim >= 0 ? o->tol_prim : tol;
    const double tol_dual = o->tol_dual >= 0 ? o->tol_dual : tol;
    double inner_tol = o->inner_tol >= 0 ? o->inner_tol : std::cbrt(tol_dual);
    const long long n = p->n, ml = p->ml;
    typedef void (*safeguard_fn)(double*, const double*, int64_t, void*);
    const safeguard_fn user_safeguard = reinterpret_cast<safeguard_fn>(o->dual_safeguard);
    std::vector<double> hy, hcx;
    if (user_safeguard) { hy.resize(ml); hcx.resize(ml); }

    // ---- initialise (alps.jl:31-46)
    BZ_TRY(h2d(p, p->v.z_curr, x0, n));
    BZ_TRY(h2d(p, p->y, y0 + (p->sharded ? p->desc.row_begin : 0), ml));
    if (!ctx->solve_ev0) { BZ_CUDA(ctx, cudaEventCreate(&ctx->solve_ev0)); BZ_CUDA(ctx, cudaEventCreate(&ctx->solve_ev1)); }
    BZ_CUDA(ctx, cudaEventRecord(ctx->solve_ev0, ctx->stream));   // device_ms: the solve between the H2D of x0, y0 and the D2H of the results
    BzState* h = p->st_host;
    memset(h, 0, sizeof(BzState));
    h->xr = p->xr_host;            // slab mode: peer mapping (zero otherwise)
    h->n_global = p->n_global;
    h->next_op = OP_DONE;
    h->H = 1.0;
    for (int i = 0; i < BZ_MAXMEM; ++i) h->slot[i] = i;
    h->spare = BZ_MAXMEM;
    h->inner_tol = inner_tol; h->tol_prim = tol_prim; h->tol_dual = tol_dual;
    h->theta_penalty = o->theta_penalty; h->kappa_penalty = o->kappa_penalty; h->kappa_tol = o->kappa_tol;
    h->outer_maxit = o->maxit; h->subsolver_maxit = o->subsolver_maxit;
    h->n_prox = 1;   // the prox!(x, gFun, x0, eps(T)) of alps.jl:38
    BZ_TRY(push_state(p));
    BZ_TRY(bz_launch_prox(ctx, p->st, p->v.z_curr, BZ_EPS, p->v.x, n, p->g, p->part, p->counters + 5));
    BZ_TRY(bz_launch_fill(ctx, p->mu, ml, 1.0));
    BZ_TRY(bz_launch_fill(ctx, p->muy, ml, 0.0));
    BZ_TRY(launch_provider(p, TGT_X));       // cx = c(x), f(x)
    {
        BzGradSrc gs = grad_src(p);
        BZ_TRY(bz_launch_assemble(ctx, p->st, gs, n, p->v.grad_x, p->part, p->counters + 6));
    }
    BZ_TRY(fetch_state(p));
    double objx = h->al_fx + h->g_z;         // objx = f(x) + gFun.gz
    h->objx = objx;
    BZ_TRY(push_state(p));
    BZ_TRY(bz_launch_init_mu(ctx, p->st, p->cx, p->s, p->mu, ml, p->D, 1, nullptr));   // also s = proj_D(cx)
    res->mu0_min = res->mu0_max = NAN;
    res->inner_tol0 = inner_tol;
    if (o->verbose) {   // alps.jl:52-55: the two @info lines report min/max of the initial mu and the inner tolerance
        std::vector<double> hm(ml);
        BZ_TRY(d2h(p, hm.data(), p->mu, ml));
        BZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        // Julia's minimum / maximum propagate NaN
        double lo = hm[0], hi = hm[0];
        for (long long j = 1; j < ml; ++j) {
            lo = (lo != lo || hm[j] != hm[j]) ? NAN : std::min(lo, hm[j]);
            hi = (hi != hi || hm[j] != hm[j]) ? NAN : std::max(hi, hm[j]);
        }
        res->mu0_min = lo; res->mu0_max = hi;
    }
    if (slack) p->slack_active = true;    // fSlack / gSlack from here on (als.jl:46-47)

    int tot_it = 0;
    long long tot_inner = 0;
    bool solved = false, tired = tot_it >= o->maxit, broken = std::isnan(objx);
    double norm_res_prim = NAN;
    long long n_grad = 0, n_prox = 0;
    bool can_stop = solved || tired || broken;
    while (!can_stop) {
        tot_it++;
        if (user_safeguard) {   // dual_safeguard(y, cx), alps.jl:62 -- a host callback, one round trip per OUTER iteration
            BZ_TRY(d2h(p, hy.data(), p->y, ml));
            BZ_TRY(d2h(p, hcx.data(), p->cx, ml));
            BZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            user_safeguard(hy.data(), hcx.data(), (int64_t)ml, o->dual_safeguard_user);
            BZ_TRY(h2d(p, p->y, hy.data(), ml));
        }
        BZ_TRY(bz_launch_al_update(ctx, p->st, p->mu, p->y, p->muy, ml, user_safeguard ? 0 : 1, 1, p->part, p->counters + 5, p->rmu));
        if (p->sharded) BZ_TRY(bz_comm_allreduce(ctx, &p->st->musqy, 1, BZ_COMM_SUM));
        long long maxit_inner = o->sub_maxit;
        if (o->max_inner_total > 0)
            maxit_inner = std::max<long long>(1, std::min<long long>(maxit_inner, o->max_inner_total - tot_inner));
        if (slack) BZ_TRY(d2d(p, p->v.x + n, p->s, ml));         // xSlack = [x; s], als.jl:69-70
        BZ_TRY(panoc_run(p, o, inner_tol, maxit_inner));
        if (slack) {
            BZ_TRY(d2d(p, p->v.x, p->v.z, n + ml));             // xSlack .= sub_sol; x, s = its two parts
            BZ_TRY(d2d(p, p->s, p->v.z + n, ml));
            // y .+= (cx .- s) ./ mu with cx = c(x) as left by the last gradient at the returned point (als.jl:81-87)
            BZ_TRY(bz_launch_multiplier_als(ctx, p->st, p->cx, p->mu, p->y, p->s, ml, p->part, p->counters + 5));
        } else {
            BZ_TRY(d2d(p, p->v.x, p->v.z, n));                  // x .= sub_sol
            BZ_TRY(bz_launch_multiplier(ctx, p->st, p->cx, p->muy, p->mu, p->y, p->s, ml, p->D,
                                        p->sharded ? 0 : 1, p->part, p->counters + 5));
            if (p->sharded) {   // ||cx - s||_inf over all row blocks, then the same scalar logic on every rank
                BZ_TRY(bz_comm_allreduce(ctx, &p->st->out[0], 1, BZ_COMM_MAX));
                BZ_TRY(bz_launch_outer_logic(ctx, p->st));
            }
        }
        BZ_TRY(fetch_state(p));                               // the one sync per outer iteration
        if (h->err_flag == 3) return bz_fail(ctx, BZ_ERR_COMM, "slab mode: a peer rank did not answer a cross-rank exchange");
        if (h->err_flag == 4) return bz_fail(ctx, BZ_ERR_CUDA, "grid-wide dense pass: a CTA of the team did not arrive (the grid is not co-resident: is the device shared?)");
        if (h->err_flag) return bz_fail(ctx, BZ_ERR_NONPOSITIVE_MU, "parameters `mu` must be positive");
        objx = h->objx;
        tot_inner = h->tot_inner_it;
        norm_res_prim = h->norm_res_prim;
        solved = h->solved; tired = h->tired; broken = h->broken;
        n_grad = h->n_grad; n_prox = h->n_prox;
        can_stop = solved || tired || broken;
        if (o->max_inner_total > 0 && tot_inner >= o->max_inner_total) can_stop = true;
        if (!can_stop && h->scale_mu) BZ_TRY(bz_launch_scale(ctx, p->mu, ml, o->kappa_penalty));
        inner_tol = h->inner_tol;
    }
    BZ_CUDA(ctx, cudaEventRecord(ctx->solve_ev1, ctx->stream));
    BZ_TRY(d2h(p, x, p->v.x, n));
    BZ_TRY(d2h(p, y, p->y, ml));
    BZ_TRY(d2h(p, s, p->s, ml));
    BZ_TRY(d2h(p, mu, p->mu, ml));
    BZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    {
        float dms = 0.f;
        cudaEventElapsedTime(&dms, ctx->solve_ev0, ctx->solve_ev1);
        res->device_ms = dms;
    }
    res->tot_it = tot_it;
    res->tot_inner_it = tot_inner;
    res->status = solved ? 0 : tired ? 1 : broken ? 2 : 3;
    res->inner_tol = inner_tol;
    res->norm_res_prim = norm_res_prim;
    res->objective = objx;
    res->n_grad = n_grad;
    res->n_prox = n_prox;
    res->n_launches = ctx->n_launches - launches0;
    res->dominant_kernel_ms = 0.0;
    res->dominant_kernel_launches = 0;
    if (ctx->profiling) collect_dominant(p, res);
    for (int k = 0; k < 4; ++k) {
        res->kind_ms[k] = 1e-6 * (double)h->tsum[k];
        res->kind_launches[k] = (int64_t)h->tcnt[k];
    }
    res->elapsed_time =
        std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return BZ_OK;
}

extern "C" int bz_alps_solve(bz_problem* p, const double* x0, const double* y0, const bz_options* o,
                             double* x, double* y, double* s, double* mu, bz_result* res) {
    if (!p || !x0 || !y0 || !o || !res) return bz_fail(p ? p->ctx : nullptr, BZ_ERR_ARG, "bz_alps_solve: NULL argument");
    BZ_CUDA(p->ctx, cudaSetDevice(p->ctx->device));
    if (p->slack) return bz_fail(p->ctx, BZ_ERR_ARG, "bz_alps_solve: the problem was created with slack_formulation (use bz_als_solve)");
    return outer_solve(p, false, x0, y0, o, x, y, s, mu, res);
}

extern "C" int bz_als_solve(bz_problem* p, const double* x0, const double* y0, const bz_options* o,
                            double* x, double* y, double* s, double* mu, bz_result* res) {
    if (!p || !x0 || !y0 || !o || !res) return bz_fail(p ? p->ctx : nullptr, BZ_ERR_ARG, "bz_als_solve: NULL argument");
    BZ_CUDA(p->ctx, cudaSetDevice(p->ctx->device));
    if (!p->slack) return bz_fail(p->ctx, BZ_ERR_ARG, "bz_als_solve: create the problem with slack_formulation = 1");
    return outer_solve(p, true, x0, y0, o, x, y, s, mu, res);
}

// ---------------------------------------------------------------------------------------------
// subsolver slot and parity entry points
// ---------------------------------------------------------------------------------------------
static int set_mu_y(bz_problem* p, const double* mu, const double* y) {
    bz_ctx* ctx = p->ctx;
    BZ_TRY(h2d(p, p->mu, mu, p->ml));
    BZ_TRY(h2d(p, p->y, y, p->ml));
    BZ_TRY(fetch_state(p));
    p->st_host->err_flag = 0;
    BZ_TRY(push_state(p));
    BZ_TRY(bz_launch_al_update(ctx, p->st, p->mu, p->y, p->muy, p->ml, 0, 0, p->part, p->counters + 5, p->rmu));
    if (p->sharded) BZ_TRY(bz_comm_allreduce(ctx, &p->st->musqy, 1, BZ_COMM_SUM));   // 0.5*sum(mu.*y.^2) over all row blocks
    BZ_TRY(fetch_state(p));
    if (p->st_host->err_flag) return bz_fail(ctx, BZ_ERR_NONPOSITIVE_MU, "parameters `mu` must be positive");
    return BZ_OK;
}

extern "C" int bz_auglag_update(bz_problem* p, const double* mu, const double* y) {
    if (!p || !mu || !y) return BZ_ERR_ARG;
    BZ_CUDA(p->ctx, cudaSetDevice(p->ctx->device));
    return set_mu_y(p, mu, y);
}

extern "C" int bz_auglag_grad(bz_problem* p, const double* x, double* dlx, double* lx, double* fx,
                              double* cx, double* s, double* yupd) {
    if (!p || !x) return BZ_ERR_ARG;
    bz_ctx* ctx = p->ctx;
    BZ_CUDA(ctx, cudaSetDevice(ctx->device));
    if (p->slack) p->slack_active = true;       // gradient!(dFxs, F::AugLagFunSlack, xs): nx + ny entries in and out
    BZ_TRY(h2d(p, p->v.xeps, x, p->nv));
    BZ_TRY(launch_provider(p, TGT_XEPS));
    BzGradSrc gs = grad_src(p);
    BZ_TRY(bz_launch_assemble(ctx, p->st, gs, p->nv, p->v.d, p->part, p->counters + 6));
    BZ_TRY(fetch_state(p));
    if (lx) *lx = p->st_host->out[0];
    if (fx) *fx = p->st_host->out[1];
    BZ_TRY(d2h(p, dlx, p->v.d, p->nv));
    BZ_TRY(d2h(p, cx, p->cx, p->ml));
    BZ_TRY(d2h(p, s, p->s, p->ml));
    BZ_TRY(d2h(p, yupd, p->yupd, p->ml));
    BZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BZ_OK;
}

extern "C" int bz_auglag_state(bz_problem* p, double* cx, double* s, double* yupd, double* muy,
                               double* musqy) {
    if (!p) return BZ_ERR_ARG;
    bz_ctx* ctx = p->ctx;
    BZ_CUDA(ctx, cudaSetDevice(ctx->device));
    BZ_TRY(fetch_state(p));
    if (musqy) *musqy = p->st_host->musqy;
    BZ_TRY(d2h(p, cx, p->cx, p->ml));
    BZ_TRY(d2h(p, s, p->s, p->ml));
    BZ_TRY(d2h(p, yupd, p->yupd, p->ml));
    BZ_TRY(d2h(p, muy, p->muy, p->ml));
    BZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BZ_OK;
}

extern "C" int bz_panocplus_solve(bz_problem* p, const double* mu, const double* y, const double* x0,
                                  const bz_options* o, double tol, double* x_out, int64_t* sub_it,
                                  double* fx, double* gz) {
    if (!p || !mu || !y || !x0 || !o) return BZ_ERR_ARG;
    bz_ctx* ctx = p->ctx;
    BZ_CUDA(ctx, cudaSetDevice(ctx->device));
    BZ_TRY(set_mu_y(p, mu, y));
    BZ_TRY(h2d(p, p->v.x, x0, p->n));
    BZ_TRY(panoc_run(p, o, tol, o->sub_maxit));
    BZ_TRY(fetch_state(p));
    if (sub_it) *sub_it = p->st_host->iter;
    if (fx) *fx = p->st_host->al_fx;
    if (gz) *gz = p->st_host->g_z;
    BZ_TRY(d2h(p, x_out, p->v.z, p->n));
    BZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BZ_OK;
}

extern "C" int bz_panoc_state(bz_problem* p, double* out8) {
    if (!p || !out8) return BZ_ERR_ARG;
    BZ_CUDA(p->ctx, cudaSetDevice(p->ctx->device));
    BZ_TRY(fetch_state(p));
    const BzState* h = p->st_host;
    out8[0] = h->gamma; out8[1] = h->tau; out8[2] = h->f_x; out8[3] = h->f_z; out8[4] = h->g_z;
    out8[5] = h->stopnorm; out8[6] = (double)h->iter; out8[7] = h->H;
    return BZ_OK;
}

extern "C" int bz_prox(bz_problem* p, const double* x, double gamma, double* z, double* gz) {
    if (!p || !x) return BZ_ERR_ARG;
    bz_ctx* ctx = p->ctx;
    BZ_CUDA(ctx, cudaSetDevice(ctx->device));
    BZ_TRY(h2d(p, p->v.xeps, x, p->n));
    BZ_TRY(bz_launch_prox(ctx, p->st, p->v.xeps, gamma, p->v.d, p->n, p->g, p->part, p->counters + 5));
    BZ_TRY(fetch_state(p));
    if (gz) *gz = p->st_host->out[0];
    BZ_TRY(d2h(p, z, p->v.d, p->n));
    BZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BZ_OK;
}

extern "C" int bz_proj(bz_problem* p, const double* v, double* z) {
    if (!p || !v) return BZ_ERR_ARG;
    bz_ctx* ctx = p->ctx;
    BZ_CUDA(ctx, cudaSetDevice(ctx->device));
    BZ_TRY(h2d(p, p->tm1, v, p->ml));
    BZ_TRY(bz_launch_proj(ctx, p->tm1, p->tm2, p->ml, p->D));
    BZ_TRY(d2h(p, z, p->tm2, p->ml));
    BZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BZ_OK;
}

extern "C" int bz_fb_step(bz_problem* p, const double* x, const double* grad, double gamma, double* z,
                          double* res, double* gz, double* dot_grad_res, double* res_sq) {
    if (!p || !x || !grad) return BZ_ERR_ARG;
    bz_ctx* ctx = p->ctx;
    BZ_CUDA(ctx, cudaSetDevice(ctx->device));
    BZ_TRY(h2d(p, p->v.x, x, p->n));
    BZ_TRY(h2d(p, p->v.grad_x, grad, p->n));
    BZ_TRY(fetch_state(p));
    p->st_host->gamma = gamma;
    p->st_host->grad_x_fresh = 0;
    p->st_host->alpha = 0.95;
    BZ_TRY(push_state(p));
    BzGradSrc gs = grad_src(p);
    BZ_TRY(bz_launch_fstep(ctx, p->st, p->v, gs, p->g, p->part, p->counters + 3, MODE_ITER));
    BZ_TRY(fetch_state(p));
    if (gz) *gz = p->st_host->out[0];
    if (dot_grad_res) *dot_grad_res = p->st_host->out[1];
    if (res_sq) *res_sq = p->st_host->out[2];
    BZ_TRY(d2h(p, z, p->v.z, p->n));
    BZ_TRY(d2h(p, res, p->v.res, p->n));
    BZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BZ_OK;
}

extern "C" int bz_obstacle_fused_step(bz_problem* p, const double* x, double gamma, double* z, double* res,
                                      double* grad_x, double* grad_z, double* scalars9) {
    if (!p || !x) return BZ_ERR_ARG;
    bz_ctx* ctx = p->ctx;
    BZ_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!p->fused) return bz_fail(ctx, BZ_ERR_UNSUPPORTED, "bz_obstacle_fused_step: the problem does not run the fused obstacle kernel");
    BZ_TRY(h2d(p, p->v.x, x, p->n));
    BZ_TRY(fetch_state(p));
    p->st_host->gamma = gamma;
    BZ_TRY(push_state(p));
    // grad_x / grad_z are debug outputs of the kernel: staged in x_d / d (free outside a solve)
    BZ_TRY(bz_launch_obstacle_fused(ctx, p->st, p->v, p->ob, g_params(p), p->mu, p->rmu, p->muy, p->part, p->counters + 3, 1,
                                    p->v.x_d, p->v.d));
    BZ_TRY(fetch_state(p));
    if (scalars9) for (int k = 0; k < 9; ++k) scalars9[k] = p->st_host->out[k];
    BZ_TRY(d2h(p, z, p->v.z, p->n));
    BZ_TRY(d2h(p, res, p->v.res, p->n));
    BZ_TRY(d2h(p, grad_x, p->v.x_d, p->n));
    BZ_TRY(d2h(p, grad_z, p->v.d, p->n));
    BZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BZ_OK;
}

extern "C" int bz_lbfgs_reset(bz_problem* p, int memory) {
    if (!p || memory < 0 || memory > BZ_MAXMEM) return BZ_ERR_ARG;
    bz_ctx* ctx = p->ctx;
    BZ_CUDA(ctx, cudaSetDevice(ctx->device));
    BZ_TRY(fetch_state(p));
    BzState* h = p->st_host;
    h->mem = memory; h->currmem = 0; h->curridx = 0; h->H = 1.0;
    for (int i = 0; i < BZ_MAXMEM; ++i) h->slot[i] = i;
    h->spare = BZ_MAXMEM;
    h->gamma = 1.0; h->alpha = 0.95; h->beta = 0.5;
    BZ_TRY(push_state(p));
    BZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BZ_OK;
}

static int lbfgs_call(bz_problem* p, int skip_update) {
    bz_ctx* ctx = p->ctx;
    BZ_TRY(fetch_state(p));
    p->st_host->skip_update = skip_update;
    BZ_TRY(push_state(p));
    BzTeam tm = p->team;
    BZ_TRY(bz_launch_lstep(ctx, p->st, p->v, tm, p->counters + 2, LROLE_NEWITER, p->slab));
    p->team.parity ^= 1;
    return BZ_OK;
}

extern "C" int bz_lbfgs_update(bz_problem* p, const double* s, const double* y) {
    if (!p || !s || !y) return BZ_ERR_ARG;
    bz_ctx* ctx = p->ctx;
    BZ_CUDA(ctx, cudaSetDevice(ctx->device));
    // update!(H, s, y) with s = x - x_prev, y = res - res_prev: feed x_prev = res_prev = 0
    BZ_TRY(h2d(p, p->v.x, s, p->n));
    BZ_TRY(h2d(p, p->v.res, y, p->n));
    BZ_CUDA(ctx, cudaMemsetAsync(p->v.x_prev, 0, p->n * sizeof(double), ctx->stream));
    BZ_CUDA(ctx, cudaMemsetAsync(p->v.res_prev, 0, p->n * sizeof(double), ctx->stream));
    BZ_TRY(lbfgs_call(p, 0));
    BZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BZ_OK;
}

extern "C" int bz_lbfgs_apply(bz_problem* p, const double* v, double* d) {
    if (!p || !v) return BZ_ERR_ARG;
    bz_ctx* ctx = p->ctx;
    BZ_CUDA(ctx, cudaSetDevice(ctx->device));
    // d = H * v: the kernel computes H * (-res), so feed res = -v
    std::vector<double> neg(p->n);
    for (long long j = 0; j < p->n; ++j) neg[j] = -v[j];
    BZ_TRY(h2d(p, p->v.res, neg.data(), p->n));
    BZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    BZ_TRY(lbfgs_call(p, 1));
    BZ_TRY(d2h(p, d, p->v.d, p->n));
    BZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BZ_OK;
}

extern "C" int bz_multiplier_update(bz_problem* p, const double* cx, const double* mu,
                                    const double* y_in, double* y_out, double* s_out,
                                    double* norm_res_prim) {
    if (!p || !cx || !mu || !y_in) return BZ_ERR_ARG;
    bz_ctx* ctx = p->ctx;
    BZ_CUDA(ctx, cudaSetDevice(ctx->device));
    BZ_TRY(set_mu_y(p, mu, y_in));
    BZ_TRY(h2d(p, p->cx, cx, p->ml));
    BZ_TRY(bz_launch_multiplier(ctx, p->st, p->cx, p->muy, p->mu, p->y, p->s, p->ml, p->D, 0, p->part,
                                p->counters + 5));
    BZ_TRY(fetch_state(p));
    if (norm_res_prim) *norm_res_prim = p->st_host->out[0];
    BZ_TRY(d2h(p, y_out, p->y, p->ml));
    BZ_TRY(d2h(p, s_out, p->s, p->ml));
    BZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BZ_OK;
}

extern "C" int bz_default_penalty(bz_problem* p, const double* cx, const double* proj_cx, double objx,
                                  double* mu) {
    if (!p || !cx || !proj_cx || !mu) return BZ_ERR_ARG;
    bz_ctx* ctx = p->ctx;
    BZ_CUDA(ctx, cudaSetDevice(ctx->device));
    BZ_TRY(h2d(p, p->tm1, cx, p->ml));
    BZ_TRY(h2d(p, p->tm2, proj_cx, p->ml));
    BZ_TRY(fetch_state(p));
    p->st_host->out[7] = objx;
    BZ_TRY(push_state(p));
    BZ_TRY(bz_launch_init_mu(ctx, p->st, p->tm1, p->tm2, p->tm3, p->ml, p->D, 0, &p->st->out[7]));
    BZ_TRY(d2h(p, mu, p->tm3, p->ml));
    BZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BZ_OK;
}

extern "C" int bz_randn_fill(bz_ctx* ctx, uint64_t seed, uint64_t stream, uint64_t start, int64_t count,
                             double* out) {
    if (!ctx || !out || count < 0) return BZ_ERR_ARG;
    BZ_CUDA(ctx, cudaSetDevice(ctx->device));
    double* d = nullptr;
    BZ_CUDA(ctx, cudaMalloc((void**)&d, (size_t)std::max<int64_t>(count, 1) * sizeof(double)));
    int rc = bz_launch_randn_fill(ctx, seed, stream, start, count, d);
    if (rc == BZ_OK) {
        cudaError_t e = cudaMemcpyAsync(out, d, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = bz_fail(ctx, BZ_ERR_CUDA, cudaGetErrorString(e));
    }
    cudaFree(d);
    return rc;
}

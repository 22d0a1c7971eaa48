/*
 * bazinga_b200.h -- C ABI of libbazinga_b200.so, the B200 (sm_100a) implementation of the
 * PANOC+/ALPS hot path of aldma/Bazinga.jl.
 *
 * This is the drop-in boundary: every entry point is what a Julia `ccall` (or the ctypes harness
 * in bazinga.jl_b200/) binds in place of the reference function cited next to it.  Plain pointers
 * and sizes only; all arrays are caller-owned HOST buffers of IEEE fp64 unless a parameter says
 * "device".  Every function returns 0 on success and a negative bz_status on failure; the message
 * is available from bz_last_error().  Nothing unwinds across this boundary.  There is no CPU
 * fallback: a machine without a CUDA device gets BZ_ERR_CUDA from bz_ctx_create.
 *
 * Threading: a bz_ctx is not thread-safe (the reference is single-threaded and mutates its oracle
 * structs, src/utilities/auglagfun.jl:11-26).  Use one ctx per host thread / per GPU.
 *
 * Device: the library expects to have the GPU to itself (one process per GPU).  Some kernels make CTAs of one launch --
 * or two kernels of one call -- wait for each other and are sized to be co-resident on an idle device: the grid-wide
 * dense pass for nx > 16384, the speculative L step that rides on the dense pass, the cooperative L kernel, the slab
 * mode's cross-rank exchanges (in the stencil passes, the fused obstacle kernel and the L kernel), and the persistent
 * warp-per-instance kernel of bz_batch_solve, which runs beside the thread-per-solve kernel and is fed by it.  Every
 * such wait is bounded (4 to 20 s): on a shared device a call fails with BZ_ERR_CUDA / BZ_ERR_COMM instead of hanging,
 * and the problem object must then be destroyed.  Tuning knobs read once per context from the environment (defaults are
 * the measured best): BZ_LSPEC, BZ_OBSTACLE_FUSED, BZ_OBSTACLE_ROTATE, BZ_SLAB_FUSED, BZ_BATCH_BUDGET,
 * BZ_BATCH_CONCURRENT, BZ_BATCH_P2CTAS, BZ_BATCH_DRAIN, BZ_PROF_STRIDE.
 */
#ifndef BAZINGA_B200_H
#define BAZINGA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BZ_ABI_VERSION 5
#define BZ_LBFGS_MAX_MEMORY 16

typedef struct bz_ctx bz_ctx;
typedef struct bz_problem bz_problem;
typedef struct bz_batch bz_batch;
typedef struct bz_bdense bz_bdense;

typedef enum {
    BZ_OK = 0,
    BZ_ERR_ARG = -1,         /* bad argument (the reference's `error(...)` guards) */
    BZ_ERR_CUDA = -2,        /* CUDA runtime failure, or no device */
    BZ_ERR_UNSUPPORTED = -3, /* oracle combination not in the catalogue: no CPU fallback exists */
    BZ_ERR_COMM = -4,        /* NCCL failure */
    BZ_ERR_NONPOSITIVE_MU = -5 /* "parameters `mu` must be positive", auglagfun.jl:33-34,92-93 */
} bz_status;

/* ---- oracle catalogue: the dispatch keys of the drop-in ------------------- */
typedef enum { /* smooth cost f: f(x), gradient!(dfx,f,x)->fx */
    BZ_F_ZERO = 0,          /* ProximalOperators.Zero / src/proxoperators/zero.jl:12-20 */
    BZ_F_LEAST_SQUARES = 1, /* ProximalOperators.LeastSquares(A,b), test/problems/test_verbose.jl:22 */
    BZ_F_QUADRATIC = 2,     /* ProximalOperators.Quadratic(Q,q), test/problems/test_nonconvex_qp.jl:14 */
    BZ_F_ROSENBROCK = 3,    /* demo/rosenbrock.jl:39-50 (f_p[0] = w) */
    BZ_F_SQDIST2 = 4,       /* demo/eitheror.jl:58-68 SmoothCostOR (f_p[0..1] = c) */
    BZ_F_LINEAR2 = 5,       /* demo/mpvca.jl:77-87 SmoothCostMPVCA (f_p[0..1] = c) */
    BZ_F_OBSTACLE_L1 = 6,   /* demo/obstacle.jl:37-47 SmoothCostObstacleL1 */
    BZ_F_OBSTACLE_L2 = 7    /* demo/obstacle.jl:21-35 SmoothCostObstacleL2: 0.5||x[1:2N]||^2 - sum(x2), dfx = [x1; x2 - 1; 0] */
} bz_f_kind;

typedef enum { /* nonsmooth cost g: prox!(z,g,x,gamma)->g(z) */
    BZ_G_ZERO = 0,          /* zero.jl:22-25; ProximalOperators.IndFree/Zero; demo/eitheror.jl:71-74 */
    BZ_G_NORM_L1 = 1,       /* ProximalOperators.NormL1(lambda), demo/basispursuit.jl:63 */
    BZ_G_NORM_L0 = 2,       /* ProximalOperators.NormL0(lambda), demo/basispursuit.jl:64 */
    BZ_G_IND_BOX = 3,       /* ProximalOperators.IndBox(lo,hi), test_nonconvex_qp.jl:15 */
    BZ_G_NORML1_BOX = 4,    /* src/proxoperators/normL1Box.jl:30-39 (u vector) */
    BZ_G_NORML1_NONNEG = 5, /* src/proxoperators/normL1Nonneg.jl:29-42 */
    BZ_G_NORML0_BOX = 6,    /* src/proxoperators/normL0Box.jl:33-58 (u vector) */
    BZ_G_ROSENBROCK = 7,    /* demo/rosenbrock.jl:52-64 */
    BZ_G_NONNEG = 8,        /* demo/mpvca.jl:89-93 */
    BZ_G_OBSTACLE_L1 = 9,   /* demo/obstacle.jl:64-81 */
    BZ_G_NORMLP_BOX = 10,   /* src/proxoperators/normLpBox.jl:37-97: alpha*sum x^p on [0,u], g_lambda = alpha, g_lo = p */
    BZ_G_NORMLP_NONNEG = 11, /* src/proxoperators/normLpNonneg.jl:36-90: alpha*sum x^p on x >= 0, g_lambda = alpha, g_lo = p */
    BZ_G_OBSTACLE_RED_L1 = 12, /* demo/obstacle.jl:83-91 NonsmoothCostObstacleRedL1 (nx = 2N) */
    BZ_G_OBSTACLE_L2 = 13   /* demo/obstacle.jl:49-62 NonsmoothCostObstacleL2: max(0, x), then the complementarity zeroing; value 0.
                             * (obstacle_l2red's NonsmoothCostObstacleL2Red, :121-126, is max(0, x): BZ_G_NONNEG) */
} bz_g_kind;

typedef enum { /* constraint function c: eval!(cx,c,x), jtprod!(jtv,c,x,v) */
    BZ_C_IDENTITY = 0,      /* test/definitions/identityFunction.jl:3-13 */
    BZ_C_DENSE_AFFINE = 1,  /* demo/basispursuit.jl:38-49: cx = A*x - b, jtv = A'*v */
    BZ_C_ROSENBROCK = 2,    /* demo/rosenbrock.jl:66-74 */
    BZ_C_EITHEROR = 3,      /* demo/eitheror.jl:92-104 ConstraintOR */
    BZ_C_MPVCA = 4,         /* demo/mpvca.jl:110-118 ConstraintMPVCA */
    BZ_C_OBSTACLE = 5,      /* demo/obstacle.jl:92-113 ConstraintObstacle (1-D, or 2-D grid) */
    BZ_C_OBSTACLE_RED = 6   /* demo/obstacle.jl:128-146 ConstraintObstacleRed: nx = ny = 2N, D = CC pairs (i, i+N) */
} bz_c_kind;

typedef enum { /* closed set D: proj!(z,D,v) */
    BZ_D_ZERO = 0,     /* src/projections/zeroSet.jl:17-20 */
    BZ_D_FREE = 1,     /* src/projections/freeSet.jl:17-20 */
    BZ_D_BOX = 2,      /* ClosedSet(IndBox) = IndicatorSet, src/projections/indicatorSet.jl:8-11 */
    BZ_D_EITHEROR = 3, /* pairs, src/projections/orConstraints.jl:7-17 */
    BZ_D_XOR = 4,      /* pairs, src/projections/orConstraints.jl:24-36 */
    BZ_D_VC = 5,       /* pairs, src/projections/vanishingConstraints.jl:27-46 */
    BZ_D_CC = 6,       /* pairs, src/projections/complementarityConstraints.jl:8-20 */
    BZ_D_PORTFOLIO = 7 /* demo/portfolio.jl:52-63 SetP = [rho, inf) x {1}: z[1] = max(rho, v[1]), z[2] = 1; d_lo = rho, ny = 2 */
} bz_d_kind;

typedef enum { BZ_ROW_MAJOR = 0, BZ_COL_MAJOR = 1 /* Julia's native layout */ } bz_layout;

/* Problem description: the four oracles of `alps(f, g, c, D, x0, y0; ...)`, alps.jl:7-13. */
typedef struct {
    int32_t f_kind, g_kind, c_kind, d_kind;
    int64_t nx, ny;
    /* f parameters.  LEAST_SQUARES: f_A is f_m x nx, f_b has f_m entries.  QUADRATIC: f_A = Q
     * (nx x nx), f_b = q.  Closed forms use f_p. */
    const double* f_A;
    const double* f_b;
    int64_t f_m;
    int32_t f_layout;
    double f_p[4];
    /* g parameters */
    double g_lambda;
    double g_lo, g_hi;
    const double* g_u; /* nx entries for the *_BOX kinds */
    /* c parameters.  DENSE_AFFINE: c_A is ny_local x nx (see sharding), c_b has ny_local entries.
     * If c_gen != 0 the library generates A = bz_randn(c_gen_seed, global element index) on the
     * device instead of reading c_A (the 68.7 GB instance cannot live on the host);
     * c_b must still be supplied. */
    const double* c_A;
    const double* c_b;
    int32_t c_layout;
    int32_t c_A_is_device; /* adopt a device pointer (row-major only), caller keeps it alive */
    int32_t c_gen;
    uint64_t c_gen_seed;
    int64_t grid_y, grid_x; /* OBSTACLE: 0,0 = 1-D SymTridiagonal(2,-1); else 2-D 5-point grid */
    /* D parameters */
    double d_lo, d_hi;
    int64_t d_pair_stride; /* pair sets: 1 = adjacent (2k,2k+1); N = (i, i+N) */
    /* Row sharding of a DENSE_AFFINE instance over ranks.  ny is the GLOBAL row
     * count; this rank owns rows [row_begin, row_begin + ny_local).  ny_local = 0 means unsharded. */
    int64_t row_begin, ny_local;
    /* Slab decomposition of the 2-D OBSTACLE grid over ranks: slab_nranks > 1 means
     * nx, ny, grid_y describe THIS rank's slab (grid_y local grid rows of grid_x columns, equal on all
     * ranks; rank r owns global rows [r*grid_y, (r+1)*grid_y)), nx_global = 3 * grid_x * grid_y *
     * slab_nranks.  After bz_problem_create every rank exports its slab (bz_problem_xr_handle), the
     * handles are all-gathered by the host and attached (bz_problem_xr_attach); from then on the
     * kernels exchange boundary rows and reduction scalars over peer memory (NVLink), no host sync. */
    int32_t slab_nranks, slab_rank;
    int64_t nx_global;
    /* != 0: the problem will be solved by bz_als_solve (src/algorithms/als.jl): PANOCplus runs on the stacked
     * variable [x; s] (nx + ny entries) with AugLagFunSlack / NonsmoothCostFunSlack
     * (src/utilities/auglagfunslack.jl).  Dense-affine and identity c, elementwise D (ZeroSet, FreeSet, box).
     * bz_auglag_grad then takes / returns nx + ny entries (gradient! of AugLagFunSlack). */
    int32_t slack_formulation;
} bz_problem_desc;

/* Options of `alps` (alps.jl:14-25) plus the PANOCplus closure options the demos bake in
 * (demo/basispursuit.jl:69-78). Use bz_default_options() and override. */
typedef struct {
    double tol, tol_prim, tol_dual, inner_tol; /* inner_tol < 0 means cbrt(tol_dual) */
    int32_t maxit;
    double theta_penalty, kappa_penalty, kappa_tol;
    int64_t subsolver_maxit; /* alps.jl:25,70 */
    /* PANOCplus */
    double alpha, beta, minimum_gamma;
    int32_t max_backtracks, lbfgs_memory;
    int64_t sub_maxit;       /* PANOCplus(maxit = ...) */
    int64_t max_inner_total; /* >0: stop after this many inner iterations in total (bench sampling) */
    int32_t verbose;         /* alps.jl:21,52-55: fills bz_result.mu0_min / mu0_max / inner_tol0 (the two @info lines) */
    /* alps.jl:23,62 `dual_safeguard(y, cx)`: NULL = default_dual_safeguard! (clamp to +-1e20, on the device).
     * Otherwise a host callback void f(double* y, const double* cx, int64_t ny_local, void* user) -- a Julia
     * @cfunction or a ctypes CFUNCTYPE -- called once per OUTER iteration on host copies of y and cx; it edits y
     * in place.  (One host round trip per outer iteration; the inner loop is untouched.) */
    void* dual_safeguard;
    void* dual_safeguard_user;
} bz_options;

typedef struct {
    int32_t tot_it;
    int64_t tot_inner_it;
    double elapsed_time;
    int32_t status; /* 0 :first_order, 1 :max_iter, 2 :exception, 3 :unknown (alps.jl:105-113) */
    double inner_tol;
    double norm_res_prim;
    double objective;
    int64_t n_grad, n_prox;       /* AL-gradient evaluations and prox calls */
    int64_t n_launches;           /* kernels launched by the library for this call */
    double dominant_kernel_ms;    /* CUDA-event time summed over the dominant kernel's launches */
    int64_t dominant_kernel_launches;
    /* device-stopwatch breakdown per kernel kind: 0 AL-gradient provider, 1 L-BFGS step, 2 forward-
     * backward step, 3 decision step -- summed over the launches that really ran */
    double kind_ms[4];
    int64_t kind_launches[4];
    /* verbose (alps.jl:52-55): extrema of the initial penalty vector and the initial inner tolerance */
    double mu0_min, mu0_max, inner_tol0;
    /* device time of the solve proper (CUDA events on the library's stream): after the H2D of x0, y0, before the D2H
     * of x, y, s, mu -- elapsed_time minus the host transfers */
    double device_ms;
} bz_result;

/* ---- context ------------------------------------------------------------------------------------ */
int bz_abi_version(void);
int bz_ctx_create(int device, bz_ctx** out);
int bz_ctx_destroy(bz_ctx* ctx);
const char* bz_last_error(bz_ctx* ctx); /* ctx may be NULL: last error of a failed bz_ctx_create */
int bz_ctx_set_profiling(bz_ctx* ctx, int on); /* CUDA events around the dominant kernel */
int bz_ctx_sync(bz_ctx* ctx);
/* device-side stopwatch on the library's stream (CUDA events): start, then stop returns ms */
int bz_ctx_timer_start(bz_ctx* ctx);
int bz_ctx_timer_stop(bz_ctx* ctx, double* ms);
int bz_device_count(int* n);
void bz_default_options(bz_options* opt);

/* ---- problems -------------------------------------------------------------------------------------
 * bz_problem_create copies the oracle data to the device (or adopts / generates it) and allocates
 * all solver state.  Replaces constructing the Julia oracle structs + AugLagFun (auglagfun.jl:32-54). */
int bz_problem_create(bz_ctx* ctx, const bz_problem_desc* desc, bz_problem** out);
int bz_problem_destroy(bz_problem* p);
/* Re-upload the dense constraint data of an existing problem (host -> device), for end-to-end timing.
 * c_A == NULL keeps the matrix (e.g. one generated on the device from a seed) and replaces b only. */
int bz_problem_upload_dense(bz_problem* p, const double* c_A, const double* c_b);
/* Slab mode: export this rank's slab (64 bytes, a cudaIpcMemHandle_t) / attach the slabs of all ranks
 * (slab_nranks x 64 bytes in rank order; this rank's own entry is ignored). */
int bz_problem_xr_handle(bz_problem* p, void* handle64);
int bz_problem_xr_attach(bz_problem* p, const void* handles64);

/* alps(f,g,c,D,x0,y0; ...) -> (x, y, tot_it, tot_inner_it, elapsed_time, status, inner_tol,
 * norm_res_prim, s, mu), src/algorithms/alps.jl:7-117.  x0,y0 are read-only. */
int bz_alps_solve(bz_problem* p, const double* x0, const double* y0, const bz_options* opt,
                  double* x, double* y, double* s, double* mu, bz_result* res);

/* als(f,g,c,D,x0,y0; ...) -> the same 10-tuple, src/algorithms/als.jl:7-120 (desc.slack_formulation must be set). */
int bz_als_solve(bz_problem* p, const double* x0, const double* y0, const bz_options* opt,
                 double* x, double* y, double* s, double* mu, bz_result* res);

/* The `subsolver` plugin slot, alps.jl:64,66: PANOCplus on (alFun, gFun) from x0 with fixed mu, y.
 * Leaves the AugLagFun / NonsmoothCostFun side-effect state (fx, gz, cx, s, yupd) at the returned
 * point, readable with bz_auglag_state. */
int bz_panocplus_solve(bz_problem* p, const double* mu, const double* y, const double* x0,
                       const bz_options* opt, double tol, double* x_out, int64_t* sub_it,
                       double* fx, double* gz);
int bz_auglag_state(bz_problem* p, double* cx, double* s, double* yupd, double* muy, double* musqy);
/* The PANOCplus iterate's scalars after the last bz_panocplus_solve / solve (what upstream's verbose table and
 * `state` hold): out8 = {gamma, tau, f_x, f_z, g_z, ||res/gamma - grad_x + grad_z||_inf, k, H}.
 * With PANOCplus(maxit = k) this is the k-th entry of the oracle's per-iteration trace (tests/golden/panoc_traces.json). */
int bz_panoc_state(bz_problem* p, double* out8);

/* ---- per-kernel parity entry points ---------------------------------------------------------------- */
/* AugLagUpdate!(al, mu, y), auglagfun.jl:91-101 (returns BZ_ERR_NONPOSITIVE_MU like the reference) */
int bz_auglag_update(bz_problem* p, const double* mu, const double* y);
/* lx = gradient!(dlx, al, x), auglagfun.jl:73-86; also returns the cached cx, s, yupd, fx
 * (any output pointer may be NULL) */
int bz_auglag_grad(bz_problem* p, const double* x, double* dlx, double* lx, double* fx,
                   double* cx, double* s, double* yupd);
/* gz = prox!(z, g, x, gamma), nonsmoothcostfun.jl:17-22 + the g kinds above */
int bz_prox(bz_problem* p, const double* x, double gamma, double* z, double* gz);
/* proj!(z, D, v) */
int bz_proj(bz_problem* p, const double* v, double* z);
/* forward-backward step: y = x - gamma*grad; z = prox(y); res = x - z; returns g(z), <grad,res>,
 * ||res||^2 */
int bz_fb_step(bz_problem* p, const double* x, const double* grad, double gamma, double* z,
               double* res, double* gz, double* dot_grad_res, double* res_sq);
/* One fused forward-backward evaluation of an obstacle problem (c kind OBSTACLE on one GPU; csrc/bz_obfused.cu) at x with
 * step gamma, against the current (mu, y): what PANOCplus computes between two direction steps --
 * grad_x = gradient!(., alFun, x) (auglagfun.jl:73-86), z = prox!(., g, x - gamma*grad_x, gamma), res = x - z,
 * grad_z = gradient!(., alFun, z).  scalars9 = {alFun(x), f(x), <grad_x,res>, ||res||^2, g(z), alFun(z), f(z),
 * ||res/gamma - grad_x + grad_z||_inf, the same norm at gamma/2}.  Output pointers may be NULL. */
int bz_obstacle_fused_step(bz_problem* p, const double* x, double gamma, double* z, double* res,
                           double* grad_x, double* grad_z, double* scalars9);
/* L-BFGS operator: reset / update!(s, y) / d = H*v */
int bz_lbfgs_reset(bz_problem* p, int memory);
int bz_lbfgs_update(bz_problem* p, const double* s, const double* y);
int bz_lbfgs_apply(bz_problem* p, const double* v, double* d);
/* multiplier update y = (c(x)+mu*y - proj_D(c(x)+mu*y))/mu and ||c(x)-s||_inf, alps.jl:72-84 */
int bz_multiplier_update(bz_problem* p, const double* cx, const double* mu, const double* y_in,
                         double* y_out, double* s_out, double* norm_res_prim);
/* default_penalty_parameter!, safeguards.jl:13-18 */
int bz_default_penalty(bz_problem* p, const double* cx, const double* proj_cx, double objx,
                       double* mu);
/* synthetic data generator shared bit-for-bit with the oracle (bz_randn_like): fills out[0..count)
 * with entries for element indices [start, start+count) */
int bz_randn_fill(bz_ctx* ctx, uint64_t seed, uint64_t stream, uint64_t start, int64_t count,
                  double* out);

/* ---- batched tiny instances (the demos' `for i = 1:ntests` loop) --------------------
 * kind: 0 rosenbrock (demo/rosenbrock.jl; nx 2, ny 2), 1 either-or, 2 xor (demo/eitheror.jl; 2, 4), 3 mpvca
 * (demo/mpvca.jl; 2, 4), and the demos' slack variants 4 eitheror_fullslack, 5 xor_fullslack (6, 4),
 * 6 mpvca_slack (4, 2), 7 mpvca_fullslack (6, 4).  x0 is batch x nx (instance-major), y0 batch x ny.  One full alps solve per
 * instance, on the device, no collectives.  Two launches: every instance on its own thread up to a budget of
 * AL-gradient evaluations, then the few instances over the budget once more, one warp each, with the line search
 * evaluated across the lanes (same results bit for bit; csrc/bz_batched.cu). */
int bz_batch_solve(bz_ctx* ctx, int kind, int64_t batch, const double* x0, const double* y0,
                   const bz_options* opt, double* x, double* y, int32_t* tot_it,
                   int64_t* tot_inner_it, int32_t* status, double* norm_res_prim, double* kernel_ms);
/* The same batch through `als` (src/algorithms/als.jl:7-120): PANOCplus on the stacked variable [x; s] with
 * AugLagFunSlack / NonsmoothCostFunSlack (src/utilities/auglagfunslack.jl); same arguments and outputs. */
int bz_batch_solve_als(bz_ctx* ctx, int kind, int64_t batch, const double* x0, const double* y0,
                   const bz_options* opt, double* x, double* y, int32_t* tot_it,
                   int64_t* tot_inner_it, int32_t* status, double* norm_res_prim, double* kernel_ms);

/* ---- batched shared-A mode of the dense AL-gradient (csrc/bz_bdense.cu) ---------------------------------
 * Up to 8 instances of the dense-affine problem that share the matrix A (demo/basispursuit.jl:38-49 over a batch of
 * x, b, mu, y): R = A X and G = A' W become skinny GEMMs on the fp64 tensor cores (DMMA m8n8k4), A is read once per
 * pass for the whole batch.  bz_bdense_auglag_grad replaces `batch` calls of gradient!(dlx_k, alFun_k, x_k)
 * (auglagfun.jl:73-86, f = Zero): X is batch x n, b / mu / y / cx / yupd are batch x m, dlx is batch x n, all
 * instance-major; lx has batch entries.  reps > 1 repeats the two kernels and returns their average device time. */
int bz_bdense_create(bz_ctx* ctx, int64_t m, int64_t n, const double* A, int32_t layout, int32_t gen, uint64_t gen_seed,
                     bz_bdense** out);
int bz_bdense_destroy(bz_bdense* h);
int bz_bdense_auglag_grad(bz_bdense* h, int32_t batch, const double* X, const double* b, const double* mu, const double* y,
                          int32_t d_kind, double d_lo, double d_hi, double* dlx, double* lx, double* cx, double* yupd,
                          int32_t reps, double* us_per_call);

/* ---- matrix prox operators: src/proxoperators/rank.jl:28-33 `prox!(Y, f::Rank, X, gamma)`,
 * nuclearNorm.jl:29-34 `prox!(Y, f::NuclearNorm, X, gamma)`, schattenNormLp.jl:36-42 `prox!(Y, f::SchattenNormLpPower, X,
 * gamma)` (with solve_lp_quasi_norm_subproblem, :60-97).  The reference factorises with LAPACK (`svd(X)`); here a one-sided
 * Jacobi iteration, one CTA per matrix of the batch (csrc/bz_matprox.cu).
 *   X, Y    batch matrices of rows x cols doubles each, columns contiguous (Julia's layout, and what
 *           check_and_reshape_as_matrix makes of a vector, rank.jl:36-44); rows, cols in 1..64
 *   lambda  Rank / NuclearNorm: lambda >= 0;  Schatten: alpha >= 0 and 0 < p < 1 (the reference's constructor errors)
 *   gz      batch values: what the reference's prox! returns (lambda*count, lambda*sum, alpha*sum sigma^p of the result)
 *   sigma   (may be NULL) batch x cols: singular values of every X, descending, zeros beyond min(rows, cols)
 *   sweeps  (may be NULL) batch: Jacobi sweeps used */
typedef enum { BZ_MATPROX_RANK = 0, BZ_MATPROX_NUCLEAR = 1, BZ_MATPROX_SCHATTEN_LP = 2 } bz_matprox_kind;
int bz_prox_matrix(bz_ctx* ctx, int kind, double lambda, double p, double gamma, int rows, int cols, int64_t batch,
                   const double* X, double* Y, double* gz, double* sigma, int32_t* sweeps);

/* Measurement aid for the batched path (not HBM-bound): the device's fp64 FMA issue peak in TFLOP/s, from a
 * microbenchmark of independent dependent-FMA chains. */
int bz_fp64_fma_peak(bz_ctx* ctx, double* tflops);

/* ---- row-sharded multi-GPU: one process per GPU --------------------------------------- */
int bz_comm_unique_id(void* id128); /* 128-byte NCCL unique id, generate on rank 0 and broadcast */
int bz_comm_init(bz_ctx* ctx, const void* id128, int nranks, int rank);
int bz_comm_destroy(bz_ctx* ctx);
/* Measurement aid: average device time (us) of one allreduce(sum) of `count` doubles on the library's communicator
 * and stream -- the collective of the row-sharded AL-gradient (count = nx + 8).  Collective: call on every rank. */
int bz_comm_bench_allreduce(bz_ctx* ctx, int64_t count, int reps, double* us_per_call);

#ifdef __cplusplus
}
#endif
#endif /* BAZINGA_B200_H */

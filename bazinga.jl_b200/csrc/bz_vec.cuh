// Launch interfaces of bz_vec.cu (n-space / m-space vector kernels of the PANOC+ state machine)
#pragma once
#include "bz_internal.cuh"

#define BZ_MAXBAR 48          // team barriers per L launch (1 + 2*memory + slack)
#define BZ_VEC_THREADS 256
#define BZ_L_THREADS 512
#define BZ_L_CTAS_PER_SM 2     // co-resident CTAs per SM of the cooperative L kernel (64 regs x 512 thr: the double2 passes need them)

struct BzVecs {
    double *x, *z, *xeps, *x_prev, *x_d, *z_curr, *res, *res_prev, *grad_x, *grad_z, *d;
    double *sM, *yM;   // mem x n ring buffers (+ one spare slot)
    double* xs;        // shadow x + d of the speculative L step (bz_lspec.cuh), or nullptr
    long long n;
};

// software grid barrier + all-CTA reduction used by the cooperative L kernel
struct BzTeam {
    unsigned int* bars;   // [2][BZ_MAXBAR]
    double* slots;        // [2][BZ_MAXBAR][max_ctas][4]
    int max_ctas;
    int parity;
};

struct BzObstacle {      // demo/obstacle.jl ConstraintObstacle (+ 2-D extension)
    long long N, gy, gx; // gy = gx = 0: 1-D SymTridiagonal(2,-1)
    int l2;              // f = SmoothCostObstacleL2 (adds 0.5||x1||^2: dfx block 1 = x1), else SmoothCostObstacleL1
};

int bz_vec_grid(const bz_ctx* ctx, long long n);
int bz_l_grid(const bz_ctx* ctx, long long n);

int bz_launch_fstep(bz_ctx* ctx, BzState* st, const BzVecs& v, const BzGradSrc& gs,
                    const BzGParams& g, double* part, unsigned int* counter, int force_mode);
int bz_launch_dstep(bz_ctx* ctx, BzState* st, const BzVecs& v, const BzGradSrc& gs, double* part,
                    unsigned int* counter);
int bz_launch_lstep(bz_ctx* ctx, BzState* st, const BzVecs& v, const BzTeam& tm,
                    unsigned int* counter, int force_role, bool slab, bool rotate = false);
int bz_launch_assemble(bz_ctx* ctx, BzState* st, const BzGradSrc& gs, long long n, double* out,
                       double* part, unsigned int* counter);

int bz_launch_identity_c(bz_ctx* ctx, BzState* st, int force_target, const double* const* src,
                         long long n, const double* mu, const double* muy, const BzDParams& D,
                         double* cx, double* s, double* yupd, double* lxpart, unsigned int* counter,
                         int set_after, int slack = 0, const double* y = nullptr);
int bz_launch_obstacle(bz_ctx* ctx, BzState* st, int force_target, const double* const* src,
                       const BzObstacle& ob, const double* mu, const double* muy, double* cx,
                       double* s, double* yupd, double* Ec, double* lxpart, double* fxpart,
                       unsigned int* counter, int set_after);

int bz_launch_obstacle_red(bz_ctx* ctx, BzState* st, int force_target, const double* const* src,
                           const BzObstacle& ob, const double* mu, const double* muy, double* cx,
                           double* s, double* yupd, double* Ec, double* lxpart, double* fxpart,
                           unsigned int* counter, int set_after);
int bz_obstacle_partials(const bz_ctx* ctx, const BzObstacle& ob);
int bz_launch_prox(bz_ctx* ctx, BzState* st, const double* in, double gamma, double* out, long long n,
                   const BzGParams& g, double* part, unsigned int* counter);
int bz_launch_proj(bz_ctx* ctx, const double* in, double* out, long long m, const BzDParams& D);
int bz_launch_al_update(bz_ctx* ctx, BzState* st, double* mu, double* y, double* muy, long long m,
                        int safeguard, int count_outer, double* part, unsigned int* counter, double* rmu = nullptr);
int bz_launch_multiplier(bz_ctx* ctx, BzState* st, const double* cx, const double* muy,
                         const double* mu, double* y, double* s, long long m, const BzDParams& D,
                         int outer_logic, double* part, unsigned int* counter);
int bz_launch_multiplier_als(bz_ctx* ctx, BzState* st, const double* cx, const double* mu, double* y, const double* s,
                             long long m, double* part, unsigned int* counter);
int bz_launch_init_mu(bz_ctx* ctx, BzState* st, const double* cx, double* s, double* mu, long long m,
                      const BzDParams& D, int compute_s, const double* objx_ptr);
int bz_launch_outer_logic(bz_ctx* ctx, BzState* st);
int bz_launch_scale(bz_ctx* ctx, double* v, long long m, double factor);
int bz_launch_fill(bz_ctx* ctx, double* v, long long m, double value);
int bz_launch_panoc_begin(bz_ctx* ctx, BzState* st, double tol, long long maxit, double alpha,
                          double beta, double gamma_min, int max_backtracks, int mem, int fused);
// fused forward-backward evaluation of the obstacle problems (bz_obfused.cu): [G, F, G, D] in one pass
int bz_obfused_partials(bz_ctx* ctx, const BzObstacle& ob);
int bz_launch_obstacle_fused(bz_ctx* ctx, BzState* st, const BzVecs& v, const BzObstacle& ob, const BzGParams& g,
                             const double* mu, const double* rmu, const double* muy, double* part, unsigned int* counter,
                             int force, double* dbg_gx, double* dbg_gz);

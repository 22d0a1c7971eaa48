// Vector kernels of the PANOC+ state machine and of the ALPS outer step.
//
// The inner iteration (ProximalAlgorithms.PANOCplus driven from
// src/algorithms/alps.jl:64-66) is executed as a fixed kernel cycle  [L, G, F, G, D]  whose
// members look at BzState::next_op and return immediately when it is not their turn:
//   L  k_lstep   L-BFGS update + two-loop direction + x_d = x + d   (or the tau-interpolation)
//   G            the AL-gradient provider (bz_dense.cu, k_identity_c, k_obstacle_*)
//   F  k_fstep   assemble grad_x, forward-backward step  y = x - gamma*grad; z = prox(y); res
//   D  k_dstep   assemble grad_z, f(z); gamma-halving / sufficient-decrease / stop decisions
// so the host never needs the outcome of a branch; it only polls BzState::inner_status.
//
// This file is compiled with -fmad=false: the reference's Julia broadcasts are not contracted
// into FMAs, and the discontinuous operators (H4) make that matter.
#include "bz_vecdev.cuh"
#include "bz_fdiv.cuh"
#include <cooperative_groups.h>


int bz_vec_grid(const bz_ctx* ctx, long long n) {
    long long nb = (n + BZ_VEC_THREADS - 1) / BZ_VEC_THREADS;
    long long cap = (long long)ctx->num_sms * 8;
    if (nb < 1) nb = 1;
    if (nb > cap) nb = cap;
    return (int)nb;
}
int bz_l_grid(const bz_ctx* ctx, long long n) {
    // cooperative launch: every CTA must be co-resident; BZ_L_CTAS_PER_SM CTAs of 512 threads fit on an SM
    long long nb = (n + 4 * BZ_L_THREADS - 1) / (4 * BZ_L_THREADS);
    if (nb < 1) nb = 1;
    if (nb > (long long)ctx->num_sms * BZ_L_CTAS_PER_SM) nb = (long long)ctx->num_sms * BZ_L_CTAS_PER_SM;
    return (int)nb;
}

// ---------------------------------------------------------------------------------------------
// AL-gradient providers for c = Identity and c = Obstacle
// ---------------------------------------------------------------------------------------------
// cx = x; t = cx + mu*y; s = proj_D(t); yupd = (t - s)/mu; jtv = yupd     (identityFunction.jl)
__global__ void __launch_bounds__(BZ_VEC_THREADS)
    k_identity_c(BzState* st, int force_target, const double* x0, const double* x1, const double* x2,
                 long long n, const double* __restrict__ mu, const double* __restrict__ muy,
                 BzDParams D, double* cx, double* s, double* yupd, double* lxpart,
                 unsigned int* counter, int set_after, int slack, const double* __restrict__ yv) {
    __shared__ double scratch[32];
    int target = force_target;
    if (target < 0) {
        if (__ldcg(&st->next_op) != OP_G) return;
        target = __ldcg(&st->grad_target);
    }
    const double* x = target == 0 ? x0 : (target == 1 ? x1 : x2);
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long j0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double a[1] = {0.0};
    if (bz_d_is_pair(D.kind)) {
        const long long np = n / 2;
        for (long long p = j0; p < np; p += stride) {
            const long long i1 = D.stride == 1 ? 2 * p : p, i2 = D.stride == 1 ? 2 * p + 1 : p + D.stride;
            const double c1 = x[i1], c2 = x[i2];
            const double t1 = c1 + muy[i1], t2 = c2 + muy[i2];
            double s1, s2;
            bz_proj_pair(D.kind, t1, t2, s1, s2);
            const double u1 = t1 - s1, u2 = t2 - s2;
            a[0] += (u1 * u1) / mu[i1];
            a[0] += (u2 * u2) / mu[i2];
            cx[i1] = c1; cx[i2] = c2; s[i1] = s1; s[i2] = s2;
            yupd[i1] = u1 / mu[i1]; yupd[i2] = u2 / mu[i2];
        }
    } else {
        for (long long j = j0; j < n; j += stride) {
            const double c = x[j];
            if (slack) {                                 // als, auglagfunslack.jl:88-92: input [x; s]
                const double sv = x[n + j];
                const double tt = (c + muy[j]) - sv;
                a[0] += (tt * tt) / mu[j];
                cx[j] = c; s[j] = sv; yupd[j] = yv[j] + (c - sv) / mu[j];
                continue;
            }
            const double t = c + muy[j];
            const double sv = bz_proj_elem(D.kind, D.lo, D.hi, t, j);
            const double u = t - sv;
            a[0] += (u * u) / mu[j];
            cx[j] = c; s[j] = sv; yupd[j] = u / mu[j];
        }
    }
    bz_block_sum<1>(a, scratch);
    if (threadIdx.x == 0) lxpart[blockIdx.x] = a[0];
    if (bz_last_cta(counter, gridDim.x) && threadIdx.x == 0 && set_after) {
        st->next_op = st->after_g;
        __threadfence();
    }
}

int bz_launch_identity_c(bz_ctx* ctx, BzState* st, int force_target, const double* const* src,
                         long long n, const double* mu, const double* muy, const BzDParams& D,
                         double* cx, double* s, double* yupd, double* lxpart, unsigned int* counter,
                         int set_after, int slack, const double* y) {
    k_identity_c<<<bz_vec_grid(ctx, n), BZ_VEC_THREADS, 0, ctx->stream>>>(
        st, force_target, src[0], src[1], src[2], n, mu, muy, D, cx, s, yupd, lxpart, counter,
        set_after, slack, y);
    ctx->n_launches++;
    BZ_CUDA(ctx, cudaGetLastError());
    return BZ_OK;
}

// Obstacle stencil provider (demo/obstacle.jl:92-113; 1-D SymTridiagonal(2,-1), or the 2-D 5-point
// extension of config C4 with 4 on the diagonal and Dirichlet boundaries).  The grid is walked as
// (row, column) -- blockIdx.y strides over rows, blockIdx.x * 256 + threadIdx.x is the column -- so
// no integer division is needed and every access is unit-stride; the vertical neighbours come from
// rows that the same and neighbouring CTAs touch (L2).  The 1-D operator is the grid (1, N).
struct StencilGeom {
    long long GY, GX;     // rows, columns
    double diag;          // 2 (1-D) or 4 (2-D)
};
__device__ __forceinline__ StencilGeom bz_geom(const BzObstacle& ob) {
    StencilGeom g;
    if (ob.gx == 0) { g.GY = 1; g.GX = ob.N; g.diag = 2.0; }
    else { g.GY = ob.gy; g.GX = ob.gx; g.diag = 4.0; }
    return g;
}
// (A v)(r, c): same operation order as the oracle's _stencil_apply (diag*v, then -left, -right, -up, -down).
// `up` / `dn`: the neighbouring ranks' boundary rows in slab mode (null: Dirichlet zero beyond the grid)
__device__ __forceinline__ double bz_stencil_rc(const double* __restrict__ v, long long r, long long c,
                                                const StencilGeom& g, const double* up = nullptr,
                                                const double* dn = nullptr) {
    const long long i = r * g.GX + c;
    double out = g.diag * v[i];
    if (c > 0) out -= v[i - 1];
    if (c + 1 < g.GX) out -= v[i + 1];
    if (r > 0) out -= v[i - g.GX];
    else if (up) out -= up[c];
    if (r + 1 < g.GY) out -= v[i + g.GX];
    else if (dn) out -= dn[c];
    return out;
}

// boundary rows of the neighbours' copy of local vector `v` (a vector inside the exported slab), block offset
// `off` (0 or N), grid GY x GX: up = last row of rank-1, dn = first row of rank+1
__device__ __forceinline__ void bz_xr_halo(const BzXr& xr, const double* v, long long off, const StencilGeom& g,
                                           const double*& up, const double*& dn) {
    up = dn = nullptr;
    if (xr.nranks <= 1) return;
    const long long rel = v - xr.slab[xr.rank];
    if (xr.rank > 0) up = xr.slab[xr.rank - 1] + rel + off + (g.GY - 1) * g.GX;
    if (xr.rank + 1 < xr.nranks) dn = xr.slab[xr.rank + 1] + rel + off;
}

// pass 1: cx = x1 + A x2 - x3; yupd = (cx + mu*y - 0)/mu; partials of lx and of f(x)
__global__ void __launch_bounds__(BZ_VEC_THREADS, 4)
    k_obstacle_c(BzState* st, int force_target, const double* x0, const double* x1p,
                 const double* x2p, BzObstacle ob, const double* __restrict__ mu,
                 const double* __restrict__ muy, double* __restrict__ cx, double* __restrict__ s,
                 double* __restrict__ yupd, double* lxpart, double* fxpart) {
    __shared__ double scratch[3 * 32];
    bz_pdl_wait();
    bz_pdl_launch_dependents();
    int target = force_target;
    if (target < 0) {
        if (__ldcg(&st->next_op) != OP_G) return;
        target = __ldcg(&st->grad_target);
        if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) bz_tic(st, 0);
    }
    const double* __restrict__ x = target == 0 ? x0 : (target == 1 ? x1p : x2p);
    const long long N = ob.N;
    const StencilGeom g = bz_geom(ob);
    const double *up, *dn;
    if (st->xr.nranks > 1) bz_xr_barrier_all(st, *st->xr.seq);
    bz_xr_halo(st->xr, x, N, g, up, dn);
    double a[3] = {0.0, 0.0, 0.0};
    // two grid rows per trip, every load of both rows before the first store (the outputs are __restrict__):
    // twice the bytes in flight per thread -- these passes are DRAM-latency-bound, not issue-bound
    for (long long r = blockIdx.y; r < g.GY; r += 2LL * gridDim.y) {
        const long long r2 = r + gridDim.y;
        const bool two = r2 < g.GY;
        for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < g.GX;
             c += (long long)gridDim.x * blockDim.x) {
            const long long i = r * g.GX + c, i2 = r2 * g.GX + c;
            const double x1a = x[i], x2a = x[N + i];
            const double cva = (x1a + bz_stencil_rc(x + N, r, c, g, up, dn)) - x[2 * N + i];
            const double mua = mu[i], muya = muy[i];
            double x1b = 0.0, x2b = 0.0, cvb = 0.0, mub = 1.0, muyb = 0.0;
            if (two) {
                x1b = x[i2]; x2b = x[N + i2];
                cvb = (x1b + bz_stencil_rc(x + N, r2, c, g, up, dn)) - x[2 * N + i2];
                mub = mu[i2]; muyb = muy[i2];
            }
            {
                const double t = cva + muya;
                const double u = t - 0.0;            // SetObstacle: z .= 0 (demo/obstacle.jl:116-119)
                a[0] += (u * u) / mua;
                cx[i] = cva; s[i] = 0.0; yupd[i] = u / mua;
                if (ob.l2) a[1] += x1a * x1a;        // SmoothCostObstacleL2: 0.5*norm(x[1:2N])^2 - sum(x2)  (:24-26)
                a[1] += x2a * x2a;                   // f = 0.5*sum(x2.^2) - sum(x2)  (:40-42)
                a[2] += x2a;
            }
            if (two) {
                const double t = cvb + muyb;
                const double u = t - 0.0;
                a[0] += (u * u) / mub;
                cx[i2] = cvb; s[i2] = 0.0; yupd[i2] = u / mub;
                if (ob.l2) a[1] += x1b * x1b;
                a[1] += x2b * x2b;
                a[2] += x2b;
            }
        }
    }
    bz_block_sum<3>(a, scratch);
    const int nb = gridDim.x * gridDim.y, b = blockIdx.y * gridDim.x + blockIdx.x;
    if (threadIdx.x == 0) {
        lxpart[b] = a[0];
        fxpart[b] = a[1];
        fxpart[nb + b] = a[2];
    }
}
// pass 2: grad = dfx + jtv with dfx = [0; x2-1; 0], jtv = [w; A w; -w]
__global__ void __launch_bounds__(BZ_VEC_THREADS, 4)
    k_obstacle_jt(BzState* st, int force_target, const double* x0, const double* x1p,
                  const double* x2p, BzObstacle ob, const double* __restrict__ yupd, double* __restrict__ Ec,
                  unsigned int* counter, int set_after) {
    bz_pdl_wait();
    bz_pdl_launch_dependents();
    int target = force_target;
    if (target < 0) {
        if (__ldcg(&st->next_op) != OP_G) return;
        target = __ldcg(&st->grad_target);
    }
    const double* __restrict__ x = target == 0 ? x0 : (target == 1 ? x1p : x2p);
    const long long N = ob.N;
    const StencilGeom g = bz_geom(ob);
    const double *up, *dn;
    const unsigned long long xseq = st->xr.nranks > 1 ? *st->xr.seq : 0ULL;
    if (st->xr.nranks > 1) bz_xr_barrier_all(st, xseq + 1ULL);   // every rank's yupd (pass 1) is complete
    bz_xr_halo(st->xr, yupd, 0, g, up, dn);
    for (long long r = blockIdx.y; r < g.GY; r += 2LL * gridDim.y) {   // two rows per trip, loads before stores
        const long long r2 = r + gridDim.y;
        const bool two = r2 < g.GY;
        for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < g.GX;
             c += (long long)gridDim.x * blockDim.x) {
            const long long i = r * g.GX + c, i2 = r2 * g.GX + c;
            const double wa = yupd[i];
            const double ea = (x[N + i] - 1.0) + bz_stencil_rc(yupd, r, c, g, up, dn);
            double wb = 0.0, eb = 0.0;
            if (two) {
                wb = yupd[i2];
                eb = (x[N + i2] - 1.0) + bz_stencil_rc(yupd, r2, c, g, up, dn);
            }
            const double f1a = ob.l2 ? x[i] : 0.0, f1b = (ob.l2 && two) ? x[i2] : 0.0;   // dfx block 1: x1 (L2) or 0 (L1)
            Ec[i] = f1a + wa;
            Ec[N + i] = ea;
            Ec[2 * N + i] = 0.0 + (-wa);
            if (two) {
                Ec[i2] = f1b + wb;
                Ec[N + i2] = eb;
                Ec[2 * N + i2] = 0.0 + (-wb);
            }
        }
    }
    if (bz_last_cta(counter, gridDim.x * gridDim.y) && threadIdx.x == 0) {
        if (st->xr.nranks > 1) *st->xr.seq = xseq + 2ULL;   // the two barriers of this gradient
        if (set_after) {
            st->next_op = st->after_g;
            if (force_target < 0) bz_toc(st, 0);
        }
        __threadfence();
    }
}

// ---------------------------------------------------------------------------------------------
// Reduced obstacle variant (demo/obstacle.jl:128-168, obstacle_l1red): x = [x1; x2] (2N),
//   c(x) = [x1 + A x2; x2],  D = complementarity pairs (i, i+N) (project_onto_CC_set!),
//   jtv = [v1; A v1 + v2],  dfx = [0; x2 - 1]
// pass 1: cx, t = cx + mu*y, s = proj_CC pairwise, yupd = (t - s)/mu, partials of lx and f
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BZ_VEC_THREADS, 4)
    k_obstacle_red_c(BzState* st, int force_target, const double* x0, const double* x1p,
                     const double* x2p, BzObstacle ob, const double* __restrict__ mu,
                     const double* __restrict__ muy, double* __restrict__ cx, double* __restrict__ s,
                     double* __restrict__ yupd, double* lxpart, double* fxpart) {
    __shared__ double scratch[3 * 32];
    bz_pdl_wait();
    bz_pdl_launch_dependents();
    int target = force_target;
    if (target < 0) {
        if (__ldcg(&st->next_op) != OP_G) return;
        target = __ldcg(&st->grad_target);
        if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) bz_tic(st, 0);
    }
    const double* __restrict__ x = target == 0 ? x0 : (target == 1 ? x1p : x2p);
    const long long N = ob.N;
    const StencilGeom g = bz_geom(ob);
    double a[3] = {0.0, 0.0, 0.0};
    for (long long r = blockIdx.y; r < g.GY; r += gridDim.y) {
        for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < g.GX;
             c += (long long)gridDim.x * blockDim.x) {
            const long long i = r * g.GX + c;
            const double x1 = x[i], x2 = x[N + i];
            const double c1 = x1 + bz_stencil_rc(x + N, r, c, g);     // cx[1:N] = x1 + A x2
            const double c2 = x2;                                      // cx[N+1:2N] = x2
            const double m1 = mu[i], m2 = mu[N + i];
            const double t1 = c1 + muy[i], t2 = c2 + muy[N + i];
            double s1, s2;
            bz_proj_pair(BZ_D_CC, t1, t2, s1, s2);                     // SetObstacleRed, obstacle.jl:150-156
            const double u1 = t1 - s1, u2 = t2 - s2;
            a[0] += (u1 * u1) / m1;
            a[0] += (u2 * u2) / m2;
            cx[i] = c1; cx[N + i] = c2;
            s[i] = s1; s[N + i] = s2;
            yupd[i] = u1 / m1; yupd[N + i] = u2 / m2;
            if (ob.l2) a[1] += x1 * x1;          // SmoothCostObstacleL2 (:24-26)
            a[1] += x2 * x2;                     // f = 0.5*sum(x2.^2) - sum(x2)  (:40-42)
            a[2] += x2;
        }
    }
    bz_block_sum<3>(a, scratch);
    const int nb = gridDim.x * gridDim.y, b = blockIdx.y * gridDim.x + blockIdx.x;
    if (threadIdx.x == 0) {
        lxpart[b] = a[0];
        fxpart[b] = a[1];
        fxpart[nb + b] = a[2];
    }
}
// pass 2: grad = dfx + jtv = [w1; (x2 - 1) + (A w1 + w2)]
__global__ void __launch_bounds__(BZ_VEC_THREADS, 4)
    k_obstacle_red_jt(BzState* st, int force_target, const double* x0, const double* x1p,
                      const double* x2p, BzObstacle ob, const double* __restrict__ yupd, double* __restrict__ Ec,
                      unsigned int* counter, int set_after) {
    bz_pdl_wait();
    bz_pdl_launch_dependents();
    int target = force_target;
    if (target < 0) {
        if (__ldcg(&st->next_op) != OP_G) return;
        target = __ldcg(&st->grad_target);
    }
    const double* __restrict__ x = target == 0 ? x0 : (target == 1 ? x1p : x2p);
    const long long N = ob.N;
    const StencilGeom g = bz_geom(ob);
    for (long long r = blockIdx.y; r < g.GY; r += gridDim.y) {
        for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < g.GX;
             c += (long long)gridDim.x * blockDim.x) {
            const long long i = r * g.GX + c;
            const double w1 = yupd[i];
            const double e2 = (x[N + i] - 1.0) + (bz_stencil_rc(yupd, r, c, g) + yupd[N + i]);
            Ec[i] = (ob.l2 ? x[i] : 0.0) + w1;
            Ec[N + i] = e2;
        }
    }
    if (bz_last_cta(counter, gridDim.x * gridDim.y) && threadIdx.x == 0 && set_after) {
        st->next_op = st->after_g;
        if (force_target < 0) bz_toc(st, 0);
        __threadfence();
    }
}

// launch geometry of the two passes (also the number of lx / fx partials)
static dim3 obstacle_grid(const bz_ctx* ctx, const BzObstacle& ob) {
    const long long GX = ob.gx == 0 ? ob.N : ob.gx, GY = ob.gx == 0 ? 1 : ob.gy;
    const long long cap = (long long)ctx->num_sms * 8;
    long long bx = (GX + BZ_VEC_THREADS - 1) / BZ_VEC_THREADS;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    long long by = cap / bx;
    if (by > GY) by = GY;
    if (by < 1) by = 1;
    return dim3((unsigned)bx, (unsigned)by);
}
int bz_obstacle_partials(const bz_ctx* ctx, const BzObstacle& ob) {
    const dim3 g = obstacle_grid(ctx, ob);
    return (int)(g.x * g.y);
}

int bz_launch_obstacle(bz_ctx* ctx, BzState* st, int force_target, const double* const* src,
                       const BzObstacle& ob, const double* mu, const double* muy, double* cx,
                       double* s, double* yupd, double* Ec, double* lxpart, double* fxpart,
                       unsigned int* counter, int set_after) {
    const dim3 grid = obstacle_grid(ctx, ob);
    BZ_TRY_RC(bz_launch_pdl(ctx, k_obstacle_c, grid, dim3(BZ_VEC_THREADS), 0, 0, st, force_target, src[0], src[1],
                            src[2], ob, mu, muy, cx, s, yupd, lxpart, fxpart));
    return bz_launch_pdl(ctx, k_obstacle_jt, grid, dim3(BZ_VEC_THREADS), 0, 0, st, force_target, src[0], src[1],
                         src[2], ob, yupd, Ec, counter, set_after);
}

int bz_launch_obstacle_red(bz_ctx* ctx, BzState* st, int force_target, const double* const* src,
                           const BzObstacle& ob, const double* mu, const double* muy, double* cx,
                           double* s, double* yupd, double* Ec, double* lxpart, double* fxpart,
                           unsigned int* counter, int set_after) {
    const dim3 grid = obstacle_grid(ctx, ob);
    BZ_TRY_RC(bz_launch_pdl(ctx, k_obstacle_red_c, grid, dim3(BZ_VEC_THREADS), 0, 0, st, force_target, src[0], src[1],
                            src[2], ob, mu, muy, cx, s, yupd, lxpart, fxpart));
    return bz_launch_pdl(ctx, k_obstacle_red_jt, grid, dim3(BZ_VEC_THREADS), 0, 0, st, force_target, src[0], src[1],
                         src[2], ob, yupd, Ec, counter, set_after);
}

// ---------------------------------------------------------------------------------------------
// standalone prox / proj, and the ALPS outer-step kernels
// ---------------------------------------------------------------------------------------------
template <bool HEAVY>
__global__ void __launch_bounds__(BZ_VEC_THREADS)
    k_prox(BzState* st, const double* in, double gamma, double* out, long long n, BzGParams g,
           double* part, unsigned int* counter) {
    __shared__ double scratch[32];
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long j0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double a[1] = {0.0};
    if (g.kind == BZ_G_OBSTACLE_L1 || g.kind == BZ_G_OBSTACLE_L2) {
        const long long N = g.N;
        const bool l1 = g.kind == BZ_G_OBSTACLE_L1;    // L2 (demo/obstacle.jl:52-62): no threshold on block 1, value 0
        for (long long i = j0; i < N; i += stride) {
            double z0 = l1 ? fmax(0.0, in[i] - gamma) : fmax(0.0, in[i]);
            double z1 = fmax(0.0, in[N + i]), z2 = fmax(0.0, in[2 * N + i]);
            if (z1 > z2) z2 = 0.0; else z1 = 0.0;
            out[i] = z0; out[N + i] = z1; out[2 * N + i] = z2;
            if (l1) a[0] += z0;
        }
    } else {
        for (long long j = j0; j < n; j += stride) out[j] = bz_prox_elem<HEAVY>(g, in[j], gamma, j, a[0]);
    }
    bz_block_sum<1>(a, scratch);
    if (threadIdx.x == 0) part[blockIdx.x] = a[0];
    if (!bz_last_cta(counter, gridDim.x)) return;
    const double s = bz_final_sum(part, gridDim.x, scratch);
    if (threadIdx.x == 0) {
        double v1[1] = {s};
        bz_xr_allreduce(st, v1, 1, 0u);       // slab mode: sum over ranks
        st->g_z = bz_g_value_scale(g) * v1[0];
        st->out[0] = st->g_z;
        __threadfence();
    }
}
int bz_launch_prox(bz_ctx* ctx, BzState* st, const double* in, double gamma, double* out, long long n,
                   const BzGParams& g, double* part, unsigned int* counter) {
    const long long items = (g.kind == BZ_G_OBSTACLE_L1 || g.kind == BZ_G_OBSTACLE_L2) ? g.N : n;
    if (bz_g_is_heavy(g.kind))
        k_prox<true><<<bz_vec_grid(ctx, items), BZ_VEC_THREADS, 0, ctx->stream>>>(st, in, gamma, out, n, g, part, counter);
    else
        k_prox<false><<<bz_vec_grid(ctx, items), BZ_VEC_THREADS, 0, ctx->stream>>>(st, in, gamma, out, n, g, part, counter);
    ctx->n_launches++;
    BZ_CUDA(ctx, cudaGetLastError());
    return BZ_OK;
}

__global__ void __launch_bounds__(BZ_VEC_THREADS)
    k_proj(const double* in, double* out, long long m, BzDParams D) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long j0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (bz_d_is_pair(D.kind)) {
        for (long long p = j0; p < m / 2; p += stride) {
            const long long i1 = D.stride == 1 ? 2 * p : p, i2 = D.stride == 1 ? 2 * p + 1 : p + D.stride;
            double s1, s2;
            bz_proj_pair(D.kind, in[i1], in[i2], s1, s2);
            out[i1] = s1; out[i2] = s2;
        }
    } else {
        for (long long j = j0; j < m; j += stride) out[j] = bz_proj_elem(D.kind, D.lo, D.hi, in[j], j);
    }
}
int bz_launch_proj(bz_ctx* ctx, const double* in, double* out, long long m, const BzDParams& D) {
    k_proj<<<bz_vec_grid(ctx, m), BZ_VEC_THREADS, 0, ctx->stream>>>(in, out, m, D);
    ctx->n_launches++;
    BZ_CUDA(ctx, cudaGetLastError());
    return BZ_OK;
}

// dual safeguard (safeguards.jl:2-10) + AugLagUpdate! (auglagfun.jl:91-101)
__global__ void __launch_bounds__(BZ_VEC_THREADS)
    k_al_update(BzState* st, double* mu, double* y, double* muy, long long m, int safeguard,
                int count_outer, double* part, unsigned int* counter, double* rmu) {
    __shared__ double scratch[2 * 32];
    const long long stride = (long long)gridDim.x * blockDim.x;
    double a[1] = {0.0}, bad[1] = {0.0};
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += stride) {
        double yj = y[j];
        if (safeguard) { yj = fmax(-1e20, fmin(yj, 1e20)); y[j] = yj; }
        const double mj = mu[j];
        if (!(mj > 0)) bad[0] = 1.0;      // any(mu .<= 0) (NaN counts as bad here)
        const double my = mj * yj;
        muy[j] = my;
        if (rmu) rmu[j] = frecip(mj);     // 1 ./ mu for the kernels that divide by mu with fdiv (bz_fdiv.cuh)
        a[0] += my * yj;
    }
    bz_block_sum<1>(a, scratch);
    bz_block_max<1>(bad, scratch);
    if (threadIdx.x == 0) { part[blockIdx.x] = a[0]; part[gridDim.x + blockIdx.x] = bad[0]; }
    if (!bz_last_cta(counter, gridDim.x)) return;
    const double s = bz_final_sum(part, gridDim.x, scratch);
    const double b = bz_final_sum(part + gridDim.x, gridDim.x, scratch);
    if (threadIdx.x == 0) {
        double v2[2] = {s, b};
        bz_xr_allreduce(st, v2, 2, 0u);       // slab mode: sums over ranks
        st->musqy = 0.5 * v2[0];
        if (v2[1] > 0.0) st->err_flag = 1;
        if (count_outer) st->tot_it += 1;
        __threadfence();
    }
}
int bz_launch_al_update(bz_ctx* ctx, BzState* st, double* mu, double* y, double* muy, long long m,
                        int safeguard, int count_outer, double* part, unsigned int* counter, double* rmu) {
    k_al_update<<<bz_vec_grid(ctx, m), BZ_VEC_THREADS, 0, ctx->stream>>>(st, mu, y, muy, m, safeguard,
                                                                         count_outer, part, counter, rmu);
    ctx->n_launches++;
    BZ_CUDA(ctx, cudaGetLastError());
    return BZ_OK;
}

// outer stopping / penalty / tolerance logic of alps.jl:67-100, run by one thread
__device__ inline void bz_outer_logic(BzState* st, double nrp) {
    const long long sub_it = st->iter;
    st->objx = st->al_fx + st->g_z;                           // alps.jl:68
    st->tot_inner_it += sub_it;
    const bool sub_solved = sub_it < st->subsolver_maxit;     // :70
    st->norm_res_prim_old = st->norm_res_prim;
    st->norm_res_prim = nrp;                                  // :84
    const bool solved = (st->inner_tol <= st->tol_dual && sub_solved) && (nrp <= st->tol_prim);
    const bool tired = st->tot_it >= st->outer_maxit;
    const bool broken = st->objx != st->objx;
    st->solved = solved; st->tired = tired; st->broken = broken;
    st->scale_mu = 0;
    if (!(solved || tired || broken)) {
        if (st->have_old && nrp > jl_max(st->theta_penalty * st->norm_res_prim_old, st->tol_prim))
            st->scale_mu = 1;                                 // mu .*= kappa_penalty (:96-97)
        st->inner_tol = jl_max(st->kappa_tol * st->inner_tol, st->tol_dual);   // :100
    }
    st->have_old = 1;
}
__global__ void k_outer_logic(BzState* st) { bz_outer_logic(st, st->out[0]); }
int bz_launch_outer_logic(bz_ctx* ctx, BzState* st) {
    k_outer_logic<<<1, 1, 0, ctx->stream>>>(st);
    ctx->n_launches++;
    BZ_CUDA(ctx, cudaGetLastError());
    return BZ_OK;
}

// multiplier update + primal residual + outer stopping / penalty / tolerance logic (alps.jl:67-100)
__global__ void __launch_bounds__(BZ_VEC_THREADS)
    k_multiplier(BzState* st, const double* __restrict__ cx, const double* __restrict__ muy,
                 const double* __restrict__ mu, double* y, double* s, long long m, BzDParams D,
                 int outer_logic, double* part, unsigned int* counter) {
    __shared__ double scratch[2 * 32];
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long j0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double mx = 0.0;
    if (bz_d_is_pair(D.kind)) {
        for (long long p = j0; p < m / 2; p += stride) {
            const long long i1 = D.stride == 1 ? 2 * p : p, i2 = D.stride == 1 ? 2 * p + 1 : p + D.stride;
            const double t1 = cx[i1] + muy[i1], t2 = cx[i2] + muy[i2];
            double s1, s2;
            bz_proj_pair(D.kind, t1, t2, s1, s2);
            s[i1] = s1; s[i2] = s2;
            y[i1] = (t1 - s1) / mu[i1];
            y[i2] = (t2 - s2) / mu[i2];
            mx = bz_nanmax(mx, cx[i1] - s1);
            mx = bz_nanmax(mx, cx[i2] - s2);
        }
    } else {
        for (long long j = j0; j < m; j += stride) {
            const double t = cx[j] + muy[j];                      // y .= cx .+ alFun.muy
            const double sv = bz_proj_elem(D.kind, D.lo, D.hi, t, j);  // proj!(s, D, y)
            s[j] = sv;
            y[j] = (t - sv) / mu[j];                              // y .-= s; y ./= mu
            mx = bz_nanmax(mx, cx[j] - sv);                       // norm(cx .- s, Inf)
        }
    }
    double a[1] = {mx != mx ? 0.0 : mx}, f[1] = {mx != mx ? 1.0 : 0.0};
    bz_block_max<1>(a, scratch);
    bz_block_max<1>(f, scratch);
    if (threadIdx.x == 0) part[blockIdx.x] = f[0] > 0.0 ? NANQ : a[0];
    if (!bz_last_cta(counter, gridDim.x)) return;
    double nrp = bz_final_nanmax(part, gridDim.x, scratch);
    if (threadIdx.x != 0) return;
    { double v1[1] = {nrp}; bz_xr_allreduce(st, v1, 1, 1u); nrp = v1[0]; }   // slab mode: max over ranks
    st->out[0] = nrp;
    if (outer_logic) bz_outer_logic(st, nrp);
    __threadfence();
}
// als.jl:83-87: y .= y .+ (cx .- s) ./ mu with the slack s of the sub-solution; norm(cx .- s, Inf); outer logic
__global__ void __launch_bounds__(BZ_VEC_THREADS)
    k_multiplier_als(BzState* st, const double* __restrict__ cx, const double* __restrict__ mu, double* y,
                     const double* __restrict__ s, long long m, double* part, unsigned int* counter) {
    __shared__ double scratch[2 * 32];
    const long long stride = (long long)gridDim.x * blockDim.x;
    double mx = 0.0;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += stride) {
        const double d = cx[j] - s[j];
        y[j] = y[j] + d / mu[j];
        mx = bz_nanmax(mx, d);
    }
    double a[1] = {mx != mx ? 0.0 : mx}, f[1] = {mx != mx ? 1.0 : 0.0};
    bz_block_max<1>(a, scratch);
    bz_block_max<1>(f, scratch);
    if (threadIdx.x == 0) part[blockIdx.x] = f[0] > 0.0 ? NANQ : a[0];
    if (!bz_last_cta(counter, gridDim.x)) return;
    const double nrp = bz_final_nanmax(part, gridDim.x, scratch);
    if (threadIdx.x != 0) return;
    st->out[0] = nrp;
    bz_outer_logic(st, nrp);
    __threadfence();
}
int bz_launch_multiplier_als(bz_ctx* ctx, BzState* st, const double* cx, const double* mu, double* y, const double* s,
                             long long m, double* part, unsigned int* counter) {
    k_multiplier_als<<<bz_vec_grid(ctx, m), BZ_VEC_THREADS, 0, ctx->stream>>>(st, cx, mu, y, s, m, part, counter);
    ctx->n_launches++;
    BZ_CUDA(ctx, cudaGetLastError());
    return BZ_OK;
}

int bz_launch_multiplier(bz_ctx* ctx, BzState* st, const double* cx, const double* muy,
                         const double* mu, double* y, double* s, long long m, const BzDParams& D,
                         int outer_logic, double* part, unsigned int* counter) {
    k_multiplier<<<bz_vec_grid(ctx, m), BZ_VEC_THREADS, 0, ctx->stream>>>(st, cx, muy, mu, y, s, m, D,
                                                                          outer_logic, part, counter);
    ctx->n_launches++;
    BZ_CUDA(ctx, cudaGetLastError());
    return BZ_OK;
}

// s = proj_D(cx) (compute_s) ; default_penalty_parameter!(mu, cx, s, objx)
// (alps.jl:41-42, safeguards.jl:13-18).  objx comes from *objx_ptr when given, else st->objx.
__global__ void __launch_bounds__(BZ_VEC_THREADS)
    k_init_mu(BzState* st, const double* __restrict__ cx, double* s, double* mu, long long m,
              BzDParams D, int compute_s, const double* objx_ptr) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long j0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const double objx = objx_ptr ? *objx_ptr : __ldcg(&st->objx);
    const double den = jl_max(1.0, objx);
    auto one = [&](long long j, double sv) {
        const double df = cx[j] - sv;
        double mj = jl_max(1.0, 0.5 * (df * df)) / den;   // max.(1.0, 0.5*(cx-proj).^2) ./ max(1.0, objx)
        mj *= 0.1;
        mj = jl_max(1e-8, (mj != mj) ? NANQ : fmin(mj, 1e8));
        mu[j] = mj;
    };
    if (compute_s && bz_d_is_pair(D.kind)) {
        for (long long p = j0; p < m / 2; p += stride) {
            const long long i1 = D.stride == 1 ? 2 * p : p, i2 = D.stride == 1 ? 2 * p + 1 : p + D.stride;
            double s1, s2;
            bz_proj_pair(D.kind, cx[i1], cx[i2], s1, s2);
            s[i1] = s1; s[i2] = s2;
            one(i1, s1); one(i2, s2);
        }
    } else {
        for (long long j = j0; j < m; j += stride) {
            double sv;
            if (compute_s) { sv = bz_proj_elem(D.kind, D.lo, D.hi, cx[j], j); s[j] = sv; }
            else sv = s[j];
            one(j, sv);
        }
    }
}
int bz_launch_init_mu(bz_ctx* ctx, BzState* st, const double* cx, double* s, double* mu, long long m,
                      const BzDParams& D, int compute_s, const double* objx_ptr) {
    k_init_mu<<<bz_vec_grid(ctx, m), BZ_VEC_THREADS, 0, ctx->stream>>>(st, cx, s, mu, m, D, compute_s,
                                                                       objx_ptr);
    ctx->n_launches++;
    BZ_CUDA(ctx, cudaGetLastError());
    return BZ_OK;
}

__global__ void k_scale(double* v, long long m, double f) {
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < m;
         j += (long long)gridDim.x * blockDim.x)
        v[j] *= f;
}
__global__ void k_fill(double* v, long long m, double f) {
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < m;
         j += (long long)gridDim.x * blockDim.x)
        v[j] = f;
}
int bz_launch_scale(bz_ctx* ctx, double* v, long long m, double factor) {
    k_scale<<<bz_vec_grid(ctx, m), BZ_VEC_THREADS, 0, ctx->stream>>>(v, m, factor);
    ctx->n_launches++;
    BZ_CUDA(ctx, cudaGetLastError());
    return BZ_OK;
}
int bz_launch_fill(bz_ctx* ctx, double* v, long long m, double value) {
    k_fill<<<bz_vec_grid(ctx, m), BZ_VEC_THREADS, 0, ctx->stream>>>(v, m, value);
    ctx->n_launches++;
    BZ_CUDA(ctx, cudaGetLastError());
    return BZ_OK;
}

__global__ void k_panoc_begin(BzState* st, double tol, long long maxit, double alpha, double beta,
                              double gamma_min, int max_backtracks, int mem, int fused) {
    st->fused = fused;
    st->tol = tol; st->maxit = maxit; st->alpha = alpha; st->beta = beta;
    st->gamma_min = gamma_min; st->max_backtracks = max_backtracks; st->mem = mem;
    st->iter = 0; st->inner_status = IST_RUNNING;
    st->currmem = 0; st->curridx = 0; st->H = 1.0;
    for (int i = 0; i < BZ_MAXMEM; ++i) st->slot[i] = i;
    st->spare = BZ_MAXMEM;
    st->rot = 0;
    st->tau = 1.0; st->k_bt = 1; st->skip_update = 1;
    st->mode = MODE_INIT0; st->grad_target = TGT_X; st->after_g = OP_F; st->grad_x_fresh = 1;
    st->l_role = LROLE_NEWITER;
    st->next_op = OP_G;
}
int bz_launch_panoc_begin(bz_ctx* ctx, BzState* st, double tol, long long maxit, double alpha,
                          double beta, double gamma_min, int max_backtracks, int mem, int fused) {
    k_panoc_begin<<<1, 1, 0, ctx->stream>>>(st, tol, maxit, alpha, beta, gamma_min, max_backtracks, mem, fused);
    ctx->n_launches++;
    BZ_CUDA(ctx, cudaGetLastError());
    return BZ_OK;
}

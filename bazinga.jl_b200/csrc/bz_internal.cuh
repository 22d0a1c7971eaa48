// Internal declarations shared by the translation units of libbazinga_b200.so.
// Device-side state machine, reduction helpers and the prox / projection functors.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <string>
#include <vector>
#include <map>
#include <unordered_map>

#include "../../include/bazinga_b200.h"

#define BZ_MAXMEM BZ_LBFGS_MAX_MEMORY
#define BZ_MAX_PARTS 1024          // max per-CTA scalar partial slots of any kernel
#define BZ_EPS 2.220446049250313e-16

// ---------------------------------------------------------------------------------------------
// Device-resident solver state.  All control flow of the PANOC+ inner loop lives here so that the
// host can enqueue the fixed kernel cycle [L, G, F, G, D] without knowing which branch is taken
// (<= 1 host sync per inner iteration; we need far fewer).
// ---------------------------------------------------------------------------------------------
enum BzOp : int { OP_L = 0, OP_G = 1, OP_F = 2, OP_D = 3, OP_DONE = 4, OP_X = 5 /* fused G+F+G+D (bz_obfused.cu) */ };
enum BzMode : int { MODE_INIT0 = 0, MODE_INIT1 = 1, MODE_INITFB = 2, MODE_ITER = 3 };
enum BzLRole : int { LROLE_NEWITER = 0, LROLE_INTERP = 1 };
enum BzTarget : int { TGT_X = 0, TGT_Z = 1, TGT_XEPS = 2 };
enum BzInnerStatus : int { IST_RUNNING = 0, IST_CONVERGED = 1, IST_MAXIT = 2 };

// ---------------------------------------------------------------------------------------------
// Cross-rank exchange over peer memory (slab-decomposed obstacle grid): one process per
// GPU, every rank maps every other rank's "slab" (cudaIpc) -- a mailbox ring plus the n-vectors whose
// boundary rows the neighbours read.  A scalar all-reduce is an all-gather of tagged words: rank r
// stores its K values into slot (seq % XR_RING, r) of EVERY rank's mailbox (NVLink P2P stores), each
// value as two 64-bit words {tag:32 | half of the double}; every aligned 8-byte word is single-copy
// atomic and validates itself, so relaxed system-scope accesses suffice.  All ranks then add the
// gathered values in rank order: bitwise identical results everywhere, which keeps the blind state
// machines of the ranks in lock-step.
// ---------------------------------------------------------------------------------------------
#define XR_MAXR 8
#define XR_K 12    /* values per exchange: the fused obstacle kernel all-gathers 9 sums + 2 stop norms at once */
#define XR_RING 64
#define XR_MAIL_WORDS (XR_RING * XR_MAXR * XR_K * 2)
struct BzXr {
    int nranks, rank;                    // nranks <= 1: disabled
    unsigned long long* mail[XR_MAXR];   // mailbox of every rank (peer-mapped; mail[rank] is local)
    double* slab[XR_MAXR];               // base of every rank's exported vectors (same layout on every rank)
    unsigned long long* seq;             // local: number of exchanges completed (identical on all ranks)
};

struct BzState {
    // ---- control words
    int next_op;
    int after_g;
    int mode;
    int l_role;
    int grad_target;
    int grad_x_fresh;   // F must assemble grad_x from the provider outputs (else reuse stored grad_x)
    int skip_update;    // first L of an inner solve: no L-BFGS update (A.4 -> A.5)
    int inner_status;
    int k_bt;
    int max_backtracks;
    int mem;            // L-BFGS memory M
    int currmem, curridx;  // curridx is 1-based like upstream, 0 = empty
    int slot[BZ_MAXMEM];   // cooperative L kernel: physical history slot of each ring position (a permutation)
    int spare;             // ... and the free physical slot the next candidate pair is written to
    int err_flag;       // nonpositive mu etc.
    // speculative L step (bz_lspec.cuh): shadow results of the L step that ran beside G(z), and the hand-shake with D / L
    int spec_ok, spec_commit, sp_currmem, sp_curridx, sp_pair;
    int npc_c;          // partial rows the last dense constraint pass wrote (0: the host-side count applies)
    unsigned long long lspec_seq;
    double sp_H, sp_ys[BZ_MAXMEM], sp_alphas[BZ_MAXMEM];
    int rot;            // fused problems: which physical buffers play x/x_prev (bit 0), res/res_prev (bit 1), z/z_curr (bit 2)
    int fused;          // the problem runs the fused forward-backward kernel X instead of [G, F, G, D] (bz_obfused.cu)
    long long n_global; // slab mode: global nx (0: the local n is the global one)
    long long iter;     // k of the IterativeAlgorithm driver (initial state counts as 1)
    long long maxit;
    long long n_grad, n_prox;
    // ---- PANOC+ scalars
    double gamma, tau, tol, alpha, beta, gamma_min;
    double f_x, f_z, g_z, f_up, dot_gr, res_nrm2sq, FBE_x, sigma, thresh, stopnorm, Lest;
    double H;
    double ys[BZ_MAXMEM];
    double alphas[BZ_MAXMEM];
    // ---- augmented-Lagrangian scalars (auglagfun.jl cached state)
    double musqy;       // 0.5*sum(mu.*y.^2)
    double al_fx;       // alFun.fx of the last gradient! call
    double al_lx;       // value of the last gradient! call
    // ---- outer loop scalars (alps.jl:59-102), updated by the multiplier kernel
    double objx, norm_res_prim, norm_res_prim_old, inner_tol, tol_prim, tol_dual;
    double theta_penalty, kappa_penalty, kappa_tol;
    int have_old, solved, tired, broken, scale_mu;
    int tot_it, outer_maxit;
    long long tot_inner_it, subsolver_maxit;
    // generic scalar outputs of the parity entry points
    double out[16];
    // device-side stopwatch per kernel kind (0 G dense, 1 L, 2 F, 3 D): %globaltimer at the start of a
    // launch that really runs and at its last-CTA epilogue (diagnostic breakdown; the roofline number
    // of the dominant kernel uses CUDA events, see bz_solver.cu)
    unsigned long long t0[4], tsum[4], tcnt[4];
    BzXr xr;
};

__device__ __forceinline__ unsigned long long bz_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void bz_tic(BzState* st, int kind) { st->t0[kind] = bz_globaltimer(); }
__device__ __forceinline__ void bz_toc(BzState* st, int kind) {
    st->tsum[kind] += bz_globaltimer() - st->t0[kind];
    st->tcnt[kind] += 1;
}

// Where the assembled gradient comes from:  grad = dfx_term + jtv_term  (auglagfun.jl:84), each
// term being 0, a vector, or a sum of per-CTA partial rows written by a dense pass.
struct BzGradSrc {
    const double* Pf; int npf;   // partial rows of dfx (LeastSquares pass), npf x n
    const double* Ef;            // dfx as a vector (Quadratic), or nullptr
    const double* Pc; int npc;   // partial rows of jtv (dense-affine constraint pass)
    const double* Ec;            // jtv as a vector (identity / stencil constraints)
    const double* lxpart; int nlx;    // sum -> sum(yupd.^2 ./ mu) (before the 0.5)
    const double* fxpart; int nfx;    // f(x) = fx_scale * sum(fxpart) + fx_scale2 * sum(fxpart2)
    double fx_scale;
    const double* fxpart2; int nfx2;
    double fx_scale2;
    // als (AugLagFunSlack, auglagfunslack.jl:93-95): the gradient has nx + ny entries, entries >= nx are -tail[j - nx]
    // (tail = yupd); nx is also the row stride of the partial arrays.  nx = 0: plain AugLagFun.
    long long nx; const double* tail;
};

// g / D parameter blocks passed by value to kernels
struct BzGParams {
    int kind;
    double lambda, lo, hi;
    const double* u;
    long long N;   // obstacle block size
    // als (NonsmoothCostFunSlack, auglagfunslack.jl:147-164): entries >= snx of the stacked variable are projected onto D
    long long snx; int sdk; double slo, shi;   // snx = 0: plain NonsmoothCostFun
};
struct BzDParams {
    int kind;
    double lo, hi;
    long long stride;  // pair stride (1 = adjacent)
};

// ---------------------------------------------------------------------------------------------
// Projections onto the 2-element sets, exact branch order of src/projections/*.jl
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline void bz_proj_pair(int kind, double x1, double x2, double& z1, double& z2) {
    switch (kind) {
    case BZ_D_EITHEROR:  // orConstraints.jl:7-17
        z1 = x1; z2 = x2;
        if (x1 < 0 && x2 < 0) { if (x1 > x2) z1 = 0.0; else z2 = 0.0; }
        break;
    case BZ_D_XOR:       // orConstraints.jl:24-36
        z1 = x1; z2 = x2;
        if (x1 * x2 > 0) {
            if (x1 > x2) { z1 = fmax(0.0, x1); z2 = fmin(0.0, x2); }
            else         { z1 = fmin(0.0, x1); z2 = fmax(0.0, x2); }
        }
        break;
    case BZ_D_VC:        // vanishingConstraints.jl:27-46
        z1 = 0.0; z2 = 0.0;
        if (x1 <= 0) { z2 = x2; }
        else if (x2 >= 0) { z1 = x1; z2 = x2; }
        else {
            double sm = x1 + x2;
            if (sm > 0) z1 = x1; else z2 = x2;   // sm < 0 and the set-valued tie both keep z2
        }
        break;
    case BZ_D_CC:        // complementarityConstraints.jl:8-20
        if (x1 > 0 && x2 > 0) { z1 = x1; z2 = x2; if (x2 > x1) z1 = 0.0; else z2 = 0.0; }
        else { z1 = fmax(0.0, x1); z2 = fmax(0.0, x2); }
        break;
    default: z1 = x1; z2 = x2;
    }
}

// elementwise sets (zeroSet.jl:17-20, freeSet.jl:17-20, indicatorSet.jl:8-11 -> IndBox clamp)
// `row` is the (global) row index: only SetP (demo/portfolio.jl:55-63: z .= cx; z[1] = max(rho, cx[1]); z[2] = 1) needs it
__host__ __device__ inline double bz_proj_elem(int kind, double lo, double hi, double v, long long row) {
    if (kind == BZ_D_ZERO) return 0.0;
    if (kind == BZ_D_BOX) return fmin(hi, fmax(lo, v));
    if (kind == BZ_D_PORTFOLIO) return row == 0 ? fmax(lo, v) : (row == 1 ? 1.0 : v);
    return v;
}
__host__ __device__ inline bool bz_d_is_pair(int kind) { return kind >= BZ_D_EITHEROR && kind <= BZ_D_CC; }

// Elementwise prox of the g kinds that do not couple entries.  Returns z and adds this entry's
// contribution to the value accumulator `acc` (sum |z|, sum z, or a nonzero count).
// HEAVY = false leaves out the Lp quasi-norm kinds (a Newton loop around pow(): ~110 registers, which would halve
// the occupancy of every kernel this is inlined into); the launchers pick the instantiation from g.kind.
__host__ __device__ inline bool bz_g_is_heavy(int kind) { return kind == BZ_G_NORMLP_BOX || kind == BZ_G_NORMLP_NONNEG; }
template <bool HEAVY = true>
__device__ inline double bz_prox_elem(const BzGParams& g, double x, double gamma, long long j,
                                      double& acc) {
    if constexpr (!HEAVY) {
        if (g.kind == BZ_G_NORMLP_BOX || g.kind == BZ_G_NORMLP_NONNEG) return x;   // never launched with these kinds
    }
    switch (g.kind) {
    case BZ_G_NORM_L1: {  // sign(x)*max(|x|-gamma*lambda,0)
        double gl = gamma * g.lambda;
        double a = fabs(x) - gl;
        double z = (a > 0.0 ? a : 0.0);
        z = (x > 0.0) ? z : ((x < 0.0) ? -z : 0.0 * z);
        acc += fabs(z);
        return z;
    }
    case BZ_G_NORM_L0: {  // keep iff |x| > sqrt(2*gamma*lambda)
        double thr = sqrt(2.0 * gamma * g.lambda);
        double z = (fabs(x) > thr) ? x : 0.0;
        acc += (z != 0.0) ? 1.0 : 0.0;
        return z;
    }
    case BZ_G_IND_BOX: return fmin(g.hi, fmax(g.lo, x));
    case BZ_G_NORML1_BOX: {  // normL1Box.jl:34-37
        double gl = gamma * g.lambda;
        double z = fmax(0.0, fmin(x - gl, g.u[j]));
        acc += z;
        return z;
    }
    case BZ_G_NORML1_NONNEG: {  // normL1Nonneg.jl:33-40
        double gl = gamma * g.lambda;
        if (x >= gl) { double z = x - gl; acc += z; return z; }
        return 0.0;
    }
    case BZ_G_NORML0_BOX: {  // normL0Box.jl:36-56 (threshold sqrt(gamma*lambda), returns x not u)
        double gl2 = gamma * g.lambda;
        double u = g.u[j];
        if (u == 0.0) return 0.0;
        if (x > sqrt(gl2)) {
            if (x > u) {
                double du = u - x;
                if (x * x > gl2 + du * du) { acc += 1.0; return x; }
                return 0.0;
            }
            acc += 1.0;
            return x;
        }
        return 0.0;
    }
    case BZ_G_ROSENBROCK: {  // demo/rosenbrock.jl:55-64: soft threshold on entry 1 only
        if (j != 0) return x;
        double gl = gamma * g.lambda;
        double z;
        if (fabs(x) <= gl) z = 0.0;
        else z = ((x > 0.0) ? 1.0 : -1.0) * (fabs(x) - gl);
        acc += fabs(z);
        return z;
    }
    case BZ_G_NONNEG: return fmax(0.0, x);   // demo/mpvca.jl:90-93
    case BZ_G_OBSTACLE_RED_L1: {             // demo/obstacle.jl:86-91: max(0, x1 - gamma) on block 1, identity on block 2
        if (j >= g.N) return x;
        const double z = fmax(0.0, x - gamma);
        acc += z;
        return z;
    }
    case BZ_G_NORMLP_BOX:       // normLpBox.jl:46-97: Newton on the scalar subproblem, then the global-minimum tests
    case BZ_G_NORMLP_NONNEG: {  // normLpNonneg.jl:43-85 (no upper bound, perturbation 1 instead of 0.1)
        if constexpr (!HEAVY) return x;
        const bool box = g.kind == BZ_G_NORMLP_BOX;
        const double p = g.lo, alpha = g.lambda * gamma;
        const double u = box ? g.u[j] : 0.0;
        if (x <= 0.0 || (box && u == 0.0)) return 0.0;
        const double zbar = pow(1 / (alpha * p * (1 - p)), 1 / (p - 2));
        const double psi_zbar = zbar + alpha * p * pow(zbar, p - 1);
        if (psi_zbar >= x) return 0.0;
        double z = zbar + (box ? 0.1 : 1.0);   // perturbation to the right
        for (int it = 0; it < 1000; ) {
            const double dphi = z - x + alpha * p * pow(z, p - 1);
            if (fabs(dphi) <= 1e-12) break;
            ++it;
            const double ddphi = 1 + alpha * p * (p - 1) * pow(z, p - 2);
            z -= dphi / ddphi;
        }
        const double phi_0 = 0.5 * (x * x);
        const double dz = z - x;
        const double phi_z = 0.5 * (dz * dz) + alpha * pow(z, p);
        if (phi_0 <= phi_z) return 0.0;
        if (box && z > u) {
            const double du = u - x;
            const double phi_u = 0.5 * (du * du) + alpha * pow(u, p);
            if (!(phi_u < phi_0)) return 0.0;
            z = u;
        }
        acc += pow(z, p);
        return z;
    }
    default: return x;  // BZ_G_ZERO
    }
}
// scale that turns the accumulated sum into g(z)
__host__ __device__ inline double bz_g_value_scale(const BzGParams& g) {
    switch (g.kind) {
    case BZ_G_NORM_L1: case BZ_G_NORM_L0: case BZ_G_NORML1_BOX: case BZ_G_NORML1_NONNEG:
    case BZ_G_NORML0_BOX: case BZ_G_ROSENBROCK: case BZ_G_NORMLP_BOX: case BZ_G_NORMLP_NONNEG: return g.lambda;
    case BZ_G_OBSTACLE_L1: case BZ_G_OBSTACLE_RED_L1: return 1.0;
    default: return 0.0;
    }
}

// ---------------------------------------------------------------------------------------------
// Reductions.  Everything is order-fixed (no float atomics): warp butterfly, then a fixed-order
// sum over warp partials, then a fixed-order sum over CTA partials by the last CTA to finish.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double bz_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double bz_warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum of K values per thread; result valid in every thread.  `scratch` needs
// K * 32 doubles of shared memory.  Contains two __syncthreads.
template <int K>
__device__ __forceinline__ void bz_block_sum(double (&v)[K], double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = bz_warp_sum(v[k]);
    __syncthreads();   // protect scratch reuse
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) scratch[k * 32 + warp] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double s = 0.0;
        for (int w = 0; w < nwarps; ++w) s += scratch[k * 32 + w];
        v[k] = s;
    }
}
template <int K>
__device__ __forceinline__ void bz_block_max(double (&v)[K], double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = bz_warp_max(v[k]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) scratch[k * 32 + warp] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double s = scratch[k * 32];
        for (int w = 1; w < nwarps; ++w) s = fmax(s, scratch[k * 32 + w]);
        v[k] = s;
    }
}

// "last CTA done" ticket: returns true in every thread of the CTA that finishes last.  The caller
// must have written its per-CTA partials to global memory before calling.  The counter is reset
// by the last CTA so the kernel can be launched again.
__device__ __forceinline__ bool bz_last_cta(unsigned int* counter, unsigned int nctas) {
    __shared__ int s_is_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        // acq_rel ticket: releases this thread's partial writes, acquires everybody else's for the last
        // CTA -- one instruction instead of membar.sc + atomic
        unsigned int t;
        asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(t) : "l"(counter) : "memory");
        s_is_last = (t == nctas - 1u);
        if (s_is_last) *counter = 0u;
    }
    __syncthreads();
    return s_is_last != 0;
}

__device__ __forceinline__ double bz_ldcg(const double* p) { return __ldcg(p); }

// fixed-order sum of a global partial array (read through L2, the writers were other CTAs)
__device__ inline double bz_sum_parts(const double* part, int n) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += __ldcg(part + i);
    return s;
}


// ---- cross-rank exchange (see BzXr).  err_flag = 3: a peer did not answer within ~4 s.
// fence = false: for a post that only carries scalars -- nothing this GPU's kernel wrote is read by a peer that sees it
__device__ __forceinline__ void bz_xr_post(const BzXr& xr, unsigned long long s, const double* v, int K, bool fence = true) {
    if (fence) __threadfence_system();   // what this GPU wrote before the post is visible to a peer that sees the post
    const unsigned long long tag = ((s + 1ULL) & 0xffffffffULL) << 32;
    const size_t off = (((size_t)(s % XR_RING) * XR_MAXR + xr.rank) * XR_K) * 2;
    for (int p = 0; p < xr.nranks; ++p) {
        unsigned long long* base = xr.mail[p] + off;
        for (int k = 0; k < K; ++k) {
            const unsigned long long bits = (unsigned long long)__double_as_longlong(v[k]);
            asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(base + 2 * k),
                         "l"(tag | (bits & 0xffffffffULL)), "l"(tag | (bits >> 32)) : "memory");
        }
    }
}
// gather the K values of every rank for exchange s from the local mailbox: out[r * XR_K + k]
__device__ inline void bz_xr_gather(BzState* st, unsigned long long s, int K, double* out) {
    const BzXr& xr = st->xr;
    const unsigned long long tag = ((s + 1ULL) & 0xffffffffULL) << 32;
    const unsigned long long* base = xr.mail[xr.rank] + ((size_t)(s % XR_RING) * XR_MAXR * XR_K) * 2;
    const long long t0 = clock64();
    for (int r = 0; r < xr.nranks; ++r)
        for (int k = 0; k < K; ++k) {
            const unsigned long long* q = base + ((size_t)r * XR_K + k) * 2;
            unsigned long long w0, w1;
            for (;;) {
                asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(q) : "memory");
                if ((((w0 ^ tag) | (w1 ^ tag)) >> 32) == 0) break;
                if (clock64() - t0 > 8000000000LL) { st->err_flag = 3; w0 = w1 = 0; break; }
            }
            out[r * XR_K + k] = __longlong_as_double((long long)((w0 & 0xffffffffULL) | (w1 << 32)));
        }
}
// All-reduce of K <= XR_K scalars by ONE thread per GPU (the thread that holds the local totals).
// ismax bit k set: NaN-propagating max (Julia's), else sum; both in rank order.
__device__ inline void bz_xr_allreduce(BzState* st, double* v, int K, unsigned int ismax, bool fence = true) {
    if (st->xr.nranks <= 1) return;
    const unsigned long long s = *st->xr.seq;
    bz_xr_post(st->xr, s, v, K, fence);
    double g[XR_MAXR * XR_K];
    bz_xr_gather(st, s, K, g);
    for (int k = 0; k < K; ++k) {
        double acc = g[k];
        for (int r = 1; r < st->xr.nranks; ++r) {
            const double w = g[r * XR_K + k];
            if ((ismax >> k) & 1u) acc = (acc != acc || w != w) ? __longlong_as_double(0x7ff8000000000000LL) : fmax(acc, w);
            else acc += w;
        }
        v[k] = acc;
    }
    *st->xr.seq = s + 1ULL;
}

// Slab mode (the grid rows are split over the ranks): every stencil pass starts with a
// cross-rank barrier -- after it the previous kernel of EVERY rank is complete, so the neighbours' boundary
// rows may be read through the peer mapping.  Exchange number s; CTA (0,0) posts, every CTA waits.
__device__ inline void bz_xr_barrier_all(BzState* st, unsigned long long s) {
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0 && blockIdx.y == 0) { double z[1] = {0.0}; bz_xr_post(st->xr, s, z, 1); }
        double g[XR_MAXR * XR_K];
        bz_xr_gather(st, s, 1, g);
        __threadfence_system();
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Host-side objects
// ---------------------------------------------------------------------------------------------
struct bz_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;     // bz_batch_solve: the warp-per-instance kernel runs beside phase 1
    std::string err;
    bool profiling = false;
    // NCCL (loaded lazily with dlopen; see bz_comm.cu)
    void* nccl_comm = nullptr;
    int nranks = 1, rank = 0;
    // event pool for the dominant-kernel timing
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    long long n_launches = 0;
    void* batch_arena = nullptr;          // device scratch of bz_batch_solve (grow-only)
    size_t batch_arena_bytes = 0;
    cudaEvent_t timer0 = nullptr, timer1 = nullptr;
    cudaEvent_t solve_ev0 = nullptr, solve_ev1 = nullptr;   // bz_result.device_ms
    int prof_stride = 31;       // bracket every n-th dominant-kernel launch with CUDA events (an event between two kernels costs their PDL
                                // overlap); odd, so that the two launches of an iteration -- G(x), G(z) beside the speculative L -- are sampled alike
    int force_wide = 0;         // testing: use the two-pass wide dense path for every shape (env BZ_FORCE_WIDE)
    // Device-memory cache of the context: bz_problem_destroy parks its blocks here and the next bz_problem_create of the
    // same shapes takes them back, so a create / solve / destroy cycle (what one `alps(f, g, c, D, x0, y0)` call is) pays
    // no cudaMalloc / cudaFree (both synchronise the device and cost milliseconds for the 0.5 GB matrix of config C2).
    std::multimap<size_t, void*> pool;             // free blocks by size
    std::unordered_map<void*, size_t> pool_live;   // blocks handed out (size)
    size_t pool_bytes = 0;                         // bytes parked in `pool`
    std::vector<void*> pinned_pool;                // parked pinned host blocks of BZ_PINNED_BLOCK bytes
    void* xpose_buf = nullptr;  // grow-only staging buffer for column-major (Julia) matrix uploads
    size_t xpose_bytes = 0;
    int batch_concurrent = 1;   // run the warp-per-instance kernel beside phase 1 on a second stream (env BZ_BATCH_CONCURRENT=0: one after the other)
    int batch_drain = -1;       // phase-1 threads left at which the remaining solves move to persistent warps (env BZ_BATCH_DRAIN; -1 = as many as there are such warps, 0 = never)
    unsigned batch_presized = 0; // bz_batch_solve: kernel pairs whose modules / local-memory reservation are in place
    int batch_p2ctas = 0;       // CTAs per SM of the warp-per-instance kernel (env BZ_BATCH_P2CTAS; 0 = 4)
    int batch_budget = 400;     // bz_batch_solve: AL-gradients an instance may use on its own thread before the warp-per-instance
                                // pass takes it over (env BZ_BATCH_BUDGET; 0 = single phase)
    int lspec = 1;              // dense problems (n <= 16384, f = Zero): the L step runs speculatively beside G(z) (env BZ_LSPEC=0: off)
    int slab_fused = 1;         // slab-decomposed obstacle grids run the fused kernel too, halo rows read from the neighbours (env BZ_SLAB_FUSED=0: the unfused cycle)
    int obstacle_rotate = 1;    // ... with buffer rotation in the cooperative L kernel (env BZ_OBSTACLE_ROTATE=0: copies)
    int obstacle_fused = 1;     // obstacle problems run the fused X kernel (env BZ_OBSTACLE_FUSED=0: the unfused cycle, for A/B runs)
    int pdl = 1;                // programmatic dependent launch inside the cycle (env BZ_PDL=0 disables)
    int evict_first = 1;        // stream A through L2 with evict-first (env BZ_EVICT_FIRST=0 disables)
    int l2_prefetch_rows = 0;   // tuning knob (env BZ_L2_PREFETCH_ROWS), see bz_dense.cu
    int wide_two_pass = 0;      // nx > 16384: use the two-pass kernels instead of the grid-wide single pass (env BZ_WIDE_TWO_PASS)
    int grid_team = 0;          // tuning knob of bz_grid.cu: CTAs per team (env BZ_GRID_TEAM)
    int grid_variant = 0;       // tuning knob of bz_grid.cu: warp-role split (env BZ_GRID_VARIANT)
};

int bz_fail(bz_ctx* ctx, int code, const std::string& msg);
#define BZ_POOL_MAX_BYTES (8ULL << 30)     // park at most this much; larger blocks (the 68.7 GB matrix of C5) go back to the driver
#define BZ_POOL_MAX_BLOCK (2ULL << 30)
#define BZ_PINNED_BLOCK 4096
cudaError_t bz_pool_alloc(bz_ctx* ctx, void** ptr, size_t bytes);
void bz_pool_free(bz_ctx* ctx, void* ptr);
void bz_pool_release(bz_ctx* ctx);

// Function attributes (dynamic shared memory limit, carve-out) are per DEVICE: a launcher configures its kernels the
// first time it runs on each device of the process, not once per process.
struct BzOncePerDevice {
    unsigned long long mask = 0;
    bool first(int dev) {
        const unsigned long long b = 1ULL << (dev & 63);
        if (mask & b) return false;
        mask |= b;
        return true;
    }
};

// Programmatic dependent launch (PDL): the kernels of the PANOC+ cycle are launched with
// programmatic stream serialization, so the next kernel's CTAs are scheduled while the previous
// kernel drains; each such kernel starts with bz_pdl_wait() (griddepcontrol.wait: everything the
// predecessor wrote is visible after it) and immediately allows its own dependents to pre-launch.
__device__ __forceinline__ void bz_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void bz_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }

template <typename... KArgs, typename... Args>
int bz_launch_pdl(bz_ctx* ctx, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, int cluster_x,
                  Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (ctx->pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (cluster_x > 0) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = cluster_x;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
    ctx->n_launches++;
    if (e != cudaSuccess) return bz_fail(ctx, BZ_ERR_CUDA, std::string("cudaLaunchKernelEx: ") + cudaGetErrorString(e));
    return BZ_OK;
}
#define BZ_TRY_RC(x) do { int rc__ = (x); if (rc__ != BZ_OK) return rc__; } while (0)
#define BZ_CUDA(ctx, call)                                                                   \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return bz_fail((ctx), BZ_ERR_CUDA,                                               \
                           std::string(#call) + ": " + cudaGetErrorString(e__));             \
    } while (0)

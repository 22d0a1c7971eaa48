// Batched tiny instances: the demos' sequential
// `for i = 1:ntests ... alps(...)` loop (demo/eitheror.jl:238-271, demo/mpvca.jl:241-281,
// demo/rosenbrock.jl:154-177) becomes one kernel launch in which every thread runs one complete
// `alps` solve -- outer augmented-Lagrangian loop (src/algorithms/alps.jl:59-102), PANOC+ inner loop
// with L-BFGS(<=8) and the closed-form oracles -- entirely in registers.
// No collectives, no host syncs inside a solve.  Not HBM-bound: I/O is 16 B in and ~100 B out per
// solve; the bound is fp64 dependent-issue latency and the iteration-count tail (H8).
//
// Compiled with -fmad=false: the arithmetic (order of operations, no contraction) is the same as
// oracle/tiny_oracle.c, so results and iteration counts are bit-identical to the oracle.
#include "bz_internal.cuh"
#include "bz_fdiv.cuh"

#define TB_MAXMEM 8
#define TB_THREADS 64
#define TB_QUEUE_TIMEOUT_NS 20000000000ull   // 20 s without a survivor or the end-of-phase-1 flag: the persistent warps give up
#define TB_CKPT 24           // doubles per hand-over record (x, s, y, mu, inner_tol, nrp, counters)
#ifndef TB_MINB
#define TB_MINB 5            // CTAs per SM the thread-per-solve kernel is compiled for: 168 registers instead of 224, a few spills,
#endif                       // 10 instead of 8 warps per SM beside the warp-per-instance kernel (either-or grid 21-24 -> 16-17 ms)
#define TB_HEAVY 10          // AL-gradients per inner iteration that make an instance a "heavy back-tracker"

namespace {

struct TbOpts {
    double tol_prim, tol_dual, inner_tol;
    int maxit;
    double theta_penalty, kappa_penalty, kappa_tol;
    long long subsolver_maxit;
    double alpha, beta, minimum_gamma;
    int max_backtracks, mem;
    long long sub_maxit;
};

__device__ __forceinline__ double jmax(double a, double b) {
    return (a != a || b != b) ? __longlong_as_double(0x7ff8000000000000LL) : fmax(a, b);
}

// ---- problem definitions: KIND 0 rosenbrock, 1 either-or, 2 xor, 3 mpvca, and the demos' slack variants
// 4 eitheror_fullslack, 5 xor_fullslack, 6 mpvca_slack, 7 mpvca_fullslack (x carries explicit slacks, D = {0})
template <int KIND> struct Prob;

template <> struct Prob<0> {   // demo/rosenbrock.jl:39-80 (w = 10, lambda = 1)
    static constexpr int NX = 2;
    static constexpr int NY = 2;
    __device__ static double f_grad(const double (&x)[2], double (&dfx)[2]) {
        const double w = 10.0;
        const double tmp = x[1] + 1 - (x[0] + 1) * (x[0] + 1);
        dfx[0] = -4 * w * tmp * (x[0] + 1);
        dfx[1] = 2 * w * tmp;
        return w * (tmp * tmp);
    }
    __device__ static void c_eval(const double (&x)[2], double (&cx)[NY]) {
        cx[0] = -x[0] - x[1];
        cx[1] = x[1] - x[0];
    }
    __device__ static void c_jt(const double (&x)[2], const double (&v)[NY], double (&jtv)[2]) {
        jtv[0] = -v[0] - v[1];
        jtv[1] = v[1] - v[0];
    }
    __device__ static void proj(const double (&v)[NY], double (&z)[NY]) {
        bz_proj_pair(BZ_D_EITHEROR, v[0], v[1], z[0], z[1]);
    }
    __device__ static double prox(const double (&x)[2], double gamma, double (&z)[2]) {
        const double gl = gamma * 1.0;
        if (fabs(x[0]) <= gl) z[0] = 0.0;
        else z[0] = (x[0] > 0 ? 1.0 : -1.0) * (fabs(x[0]) - gl);
        z[1] = x[1];
        return 1.0 * fabs(z[0]);
    }
};

template <int DK> struct ProbOr {   // demo/eitheror.jl:58-132 (c = (8,-3))
    static constexpr int NX = 2;
    static constexpr int NY = 4;
    __device__ static double f_grad(const double (&x)[2], double (&dfx)[2]) {
        const double d0 = x[0] - 8.0, d1 = x[1] - (-3.0);
        dfx[0] = 2 * d0;
        dfx[1] = 2 * d1;
        const double nr = sqrt(d0 * d0 + d1 * d1);
        return nr * nr;
    }
    __device__ static void c_eval(const double (&x)[2], double (&cx)[NY]) {
        cx[0] = 2 * x[1] - x[0] - 4;
        cx[1] = 2 - x[0];
        cx[2] = 4 * x[1] - x[0] * x[0];
        cx[3] = 10 - (x[0] - 3) * (x[0] - 3) - (x[1] - 1) * (x[1] - 1);
    }
    __device__ static void c_jt(const double (&x)[2], const double (&v)[NY], double (&jtv)[2]) {
        jtv[0] = -v[0] - v[1] - 2 * x[0] * v[2] + 2 * (3 - x[0]) * v[3];
        jtv[1] = 2 * v[0] + 4 * v[2] + 2 * (1 - x[1]) * v[3];
    }
    __device__ static void proj(const double (&v)[NY], double (&z)[NY]) {
        bz_proj_pair(DK, v[0], v[1], z[0], z[1]);
        bz_proj_pair(DK, v[2], v[3], z[2], z[3]);
    }
    __device__ static double prox(const double (&x)[2], double, double (&z)[2]) {
        z[0] = x[0];
        z[1] = x[1];
        return 0.0;
    }
};
template <> struct Prob<1> : ProbOr<BZ_D_EITHEROR> {};
template <> struct Prob<2> : ProbOr<BZ_D_XOR> {};

template <> struct Prob<3> {   // demo/mpvca.jl:77-150 (c = (4,2))
    static constexpr int NX = 2;
    static constexpr int NY = 4;
    __device__ static double f_grad(const double (&x)[2], double (&dfx)[2]) {
        dfx[0] = 4.0;
        dfx[1] = 2.0;
        return 4.0 * x[0] + 2.0 * x[1];
    }
    __device__ static void c_eval(const double (&x)[2], double (&cx)[NY]) {
        cx[0] = x[0];
        cx[1] = x[0] + x[1] - 5.0 * sqrt(2.0);
        cx[2] = x[1];
        cx[3] = x[0] + x[1] - 5.0;
    }
    __device__ static void c_jt(const double (&)[2], const double (&v)[NY], double (&jtv)[2]) {
        jtv[0] = v[0] + v[1] + v[3];
        jtv[1] = v[1] + v[2] + v[3];
    }
    __device__ static void proj(const double (&v)[NY], double (&z)[NY]) {
        bz_proj_pair(BZ_D_VC, v[0], v[1], z[0], z[1]);
        bz_proj_pair(BZ_D_VC, v[2], v[3], z[2], z[3]);
    }
    __device__ static double prox(const double (&x)[2], double, double (&z)[2]) {
        z[0] = fmax(0.0, x[0]);
        z[1] = fmax(0.0, x[1]);
        return 0.0;
    }
};

// demo/eitheror.jl:76-90,106-119: x = [x1, x2, s1..s4]; g projects (s1,s2), (s3,s4) onto EITHEROR / XOR; D = ZeroSet
template <int DK> struct ProbOrFullSlack {
    static constexpr int NX = 6;
    static constexpr int NY = 4;
    __device__ static double f_grad(const double (&x)[NX], double (&dfx)[NX]) {
        const double d0 = x[0] - 8.0, d1 = x[1] - (-3.0);
        dfx[0] = 2 * d0;
        dfx[1] = 2 * d1;
        dfx[2] = 0.0; dfx[3] = 0.0; dfx[4] = 0.0; dfx[5] = 0.0;
        const double nr = sqrt(d0 * d0 + d1 * d1);
        return nr * nr;
    }
    __device__ static void c_eval(const double (&x)[NX], double (&cx)[NY]) {
        cx[0] = 2 * x[1] - x[0] - 4 - x[2];
        cx[1] = 2 - x[0] - x[3];
        cx[2] = 4 * x[1] - x[0] * x[0] - x[4];
        cx[3] = 10 - (x[0] - 3) * (x[0] - 3) - (x[1] - 1) * (x[1] - 1) - x[5];
    }
    __device__ static void c_jt(const double (&x)[NX], const double (&v)[NY], double (&jtv)[NX]) {
        jtv[0] = -v[0] - v[1] - 2 * x[0] * v[2] + 2 * (3 - x[0]) * v[3];
        jtv[1] = 2 * v[0] + 4 * v[2] + 2 * (1 - x[1]) * v[3];
        jtv[2] = -v[0]; jtv[3] = -v[1]; jtv[4] = -v[2]; jtv[5] = -v[3];
    }
    __device__ static void proj(const double (&)[NY], double (&z)[NY]) { z[0] = 0.0; z[1] = 0.0; z[2] = 0.0; z[3] = 0.0; }
    __device__ static double prox(const double (&x)[NX], double, double (&z)[NX]) {
        z[0] = x[0];
        z[1] = x[1];
        bz_proj_pair(DK, x[2], x[3], z[2], z[3]);
        bz_proj_pair(DK, x[4], x[5], z[4], z[5]);
        return 0.0;
    }
};
template <> struct Prob<4> : ProbOrFullSlack<BZ_D_EITHEROR> {};
template <> struct Prob<5> : ProbOrFullSlack<BZ_D_XOR> {};

template <> struct Prob<6> {   // demo/mpvca.jl:95-100,120-128 mpvca_slack: x = [x1, x2, s1, s2]
    static constexpr int NX = 4;
    static constexpr int NY = 2;
    __device__ static double f_grad(const double (&x)[NX], double (&dfx)[NX]) {
        dfx[0] = 4.0; dfx[1] = 2.0; dfx[2] = 0.0; dfx[3] = 0.0;
        return 4.0 * x[0] + 2.0 * x[1];
    }
    __device__ static void c_eval(const double (&x)[NX], double (&cx)[NY]) {
        cx[0] = x[0] + x[1] - 5.0 * sqrt(2.0) - x[2];
        cx[1] = x[0] + x[1] - 5.0 - x[3];
    }
    __device__ static void c_jt(const double (&)[NX], const double (&v)[NY], double (&jtv)[NX]) {
        jtv[0] = v[0] + v[1];
        jtv[1] = v[0] + v[1];
        jtv[2] = -v[0];
        jtv[3] = -v[1];
    }
    __device__ static void proj(const double (&)[NY], double (&z)[NY]) { z[0] = 0.0; z[1] = 0.0; }
    __device__ static double prox(const double (&x)[NX], double, double (&z)[NX]) {
        bz_proj_pair(BZ_D_VC, x[0], x[2], z[0], z[2]);   // pairs (x1, s1), (x2, s2)
        bz_proj_pair(BZ_D_VC, x[1], x[3], z[1], z[3]);
        return 0.0;
    }
};

template <> struct Prob<7> {   // demo/mpvca.jl:102-108,130-143 mpvca_fullslack: x = [x1, x2, s1..s4]
    static constexpr int NX = 6;
    static constexpr int NY = 4;
    __device__ static double f_grad(const double (&x)[NX], double (&dfx)[NX]) {
        dfx[0] = 4.0; dfx[1] = 2.0; dfx[2] = 0.0; dfx[3] = 0.0; dfx[4] = 0.0; dfx[5] = 0.0;
        return 4.0 * x[0] + 2.0 * x[1];
    }
    __device__ static void c_eval(const double (&x)[NX], double (&cx)[NY]) {
        cx[0] = x[0] - x[2];
        cx[1] = x[0] + x[1] - 5.0 * sqrt(2.0) - x[3];
        cx[2] = x[1] - x[4];
        cx[3] = x[0] + x[1] - 5.0 - x[5];
    }
    __device__ static void c_jt(const double (&)[NX], const double (&v)[NY], double (&jtv)[NX]) {
        jtv[0] = v[0] + v[1] + v[3];
        jtv[1] = v[1] + v[2] + v[3];
        jtv[2] = -v[0]; jtv[3] = -v[1]; jtv[4] = -v[2]; jtv[5] = -v[3];
    }
    __device__ static void proj(const double (&)[NY], double (&z)[NY]) { z[0] = 0.0; z[1] = 0.0; z[2] = 0.0; z[3] = 0.0; }
    __device__ static double prox(const double (&x)[NX], double, double (&z)[NX]) {
        z[0] = fmax(0.0, x[0]);
        z[1] = fmax(0.0, x[1]);
        bz_proj_pair(BZ_D_VC, x[2], x[3], z[2], z[3]);
        bz_proj_pair(BZ_D_VC, x[4], x[5], z[4], z[5]);
        return 0.0;
    }
};

// ---- per-thread augmented-Lagrangian state (auglagfun.jl:11-26; SLACK: auglagfunslack.jl:15-29) -----
// SLACK = 0: AugLagFun + NonsmoothCostFun on x (alps).  SLACK = 1: AugLagFunSlack + NonsmoothCostFunSlack on the
// stacked variable [x; s] (als); NV is the length of the PANOC+ variable.
template <int KIND, int SLACK> struct Al {
    static constexpr int NX = Prob<KIND>::NX;
    static constexpr int NY = Prob<KIND>::NY;
    static constexpr int NV = SLACK ? NX + NY : NX;
    double mu[NY], rmu[NY], y[NY], muy[NY], musqy, cx[NY], s[NY], yupd[NY], fx, gz;   // rmu = 1 ./ mu (see fdiv)
    int ngrad;   // AL-gradient evaluations of this solve so far (the work budget of phase 1, see k_batch_alps)

    __device__ void update() {   // AugLagUpdate!
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < NY; ++i) { muy[i] = mu[i] * y[i]; acc += muy[i] * y[i]; rmu[i] = frecip(mu[i]); }
        musqy = 0.5 * acc;
    }
    __device__ double gradient(const double (&xv)[NV], double (&dlx)[NV]) {   // gradient!(dlx, al, x)
        double dfx[NX], jtv[NX], x[NX], acc = 0.0;
        ++ngrad;
#pragma unroll
        for (int j = 0; j < NX; ++j) x[j] = xv[j];
        if constexpr (SLACK != 0) {   // auglagfunslack.jl:77-97
            fx = Prob<KIND>::f_grad(x, dfx);
            Prob<KIND>::c_eval(x, cx);
#pragma unroll
            for (int i = 0; i < NY; ++i) {
                const double t = (cx[i] + muy[i]) - xv[NX + i];
                acc += fdiv(t * t, mu[i], rmu[i]);
            }
            double Fxs = 0.5 * acc;
            Fxs += fx;
            Fxs -= musqy;
#pragma unroll
            for (int i = 0; i < NY; ++i) yupd[i] = y[i] + fdiv(cx[i] - xv[NX + i], mu[i], rmu[i]);
            Prob<KIND>::c_jt(x, yupd, jtv);
#pragma unroll
            for (int j = 0; j < NX; ++j) dlx[j] = dfx[j] + jtv[j];
#pragma unroll
            for (int i = 0; i < NY; ++i) dlx[NX + i] = -yupd[i];
            return Fxs;
        } else {
        Prob<KIND>::c_eval(x, cx);
#pragma unroll
        for (int i = 0; i < NY; ++i) yupd[i] = cx[i] + muy[i];
        Prob<KIND>::proj(yupd, s);
#pragma unroll
        for (int i = 0; i < NY; ++i) yupd[i] -= s[i];
        // sum((yupd.^2) ./ mu) and yupd ./= mu (auglagfun.jl:78-79) in one sweep.  On an inactive constraint t = +-0 (three
        // times in four): (+0)/mu adds nothing to a sum that is never -0, and t/mu = t for mu > 0 -- one test instead of
        // two divisions (rmu > 0 says mu is positive, finite and not NaN; anything else takes the division).
#pragma unroll
        for (int i = 0; i < NY; ++i) {
            const double t = yupd[i];
            if (t == 0.0 && rmu[i] > 0.0) continue;
            acc += fdiv(t * t, mu[i], rmu[i]);
            yupd[i] = fdiv(t, mu[i], rmu[i]);
        }
        double lx = 0.5 * acc;
        fx = Prob<KIND>::f_grad(x, dfx);
        lx += fx;
        lx -= musqy;
        Prob<KIND>::c_jt(x, yupd, jtv);
#pragma unroll
        for (int j = 0; j < NX; ++j) dlx[j] = dfx[j] + jtv[j];
        return lx;
        }
    }
    // prox!(z, gFun, x, gamma): prox of g on x, and for SLACK the projection onto D of the slack part
    __device__ double prox(const double (&xv)[NV], double gamma, double (&z)[NV]) {
        double x[NX], zx[NX];
#pragma unroll
        for (int j = 0; j < NX; ++j) x[j] = xv[j];
        const double g = Prob<KIND>::prox(x, gamma, zx);
#pragma unroll
        for (int j = 0; j < NX; ++j) z[j] = zx[j];
        if constexpr (SLACK != 0) {   // auglagfunslack.jl:147-164
            double sv[NY], zs[NY];
#pragma unroll
            for (int i = 0; i < NY; ++i) sv[i] = xv[NX + i];
            Prob<KIND>::proj(sv, zs);
#pragma unroll
            for (int i = 0; i < NY; ++i) z[NX + i] = zs[i];
        }
        return g;
    }
};

// left-to-right dot product (NV = 2: a0*b0 + a1*b1, the order of oracle/tiny_oracle.c dotn)
template <int NV>
__device__ __forceinline__ double dotn(const double (&a)[NV], const double (&b)[NV]) {
    double acc = a[0] * b[0];
#pragma unroll
    for (int j = 1; j < NV; ++j) acc += a[j] * b[j];
    return acc;
}
template <int NV>
__device__ __forceinline__ double f_model(double fx, const double (&g)[NV], const double (&res)[NV], double L) {
    const double nr = sqrt(dotn<NV>(res, res));
    return fx - dotn<NV>(g, res) + (L / 2) * (nr * nr);
}

// L-BFGS history kept newest-first and shifted on update, so every index is static (registers);
// the visiting order (loop1 newest->oldest, loop2 oldest->newest) is that of the upstream ring.
template <int MM, int NV>
struct Lbfgs {
    int M, currmem;
    double H, s[MM][NV], y[MM][NV], ys[MM], rys[MM];   // rys = 1 ./ ys (see fdiv)
    __device__ void reset() { currmem = 0; H = 1.0; }
    __device__ void update(const double (&sv)[NV], const double (&yv)[NV]) {
        if (M == 0) return;
        const double ysv = dotn<NV>(sv, yv);
        if (ysv > 0) {
#pragma unroll
            for (int i = MM - 1; i > 0; --i) {
#pragma unroll
                for (int j = 0; j < NV; ++j) { s[i][j] = s[i - 1][j]; y[i][j] = y[i - 1][j]; }
                ys[i] = ys[i - 1];
                rys[i] = rys[i - 1];
            }
#pragma unroll
            for (int j = 0; j < NV; ++j) { s[0][j] = sv[j]; y[0][j] = yv[j]; }
            ys[0] = ysv;
            rys[0] = frecip(ysv);
            currmem = currmem + 1 > M ? M : currmem + 1;
            H = ysv / dotn<NV>(yv, yv);
        }
    }
    __device__ void mul(const double (&v)[NV], double (&d)[NV]) const {
        double al[MM];
#pragma unroll
        for (int j = 0; j < NV; ++j) d[j] = v[j];
#pragma unroll
        for (int i = 0; i < MM; ++i) {
            if (i < currmem) {
                al[i] = fdiv(dotn<NV>(s[i], d), ys[i], rys[i]);
#pragma unroll
                for (int j = 0; j < NV; ++j) d[j] -= al[i] * y[i][j];
            }
        }
#pragma unroll
        for (int j = 0; j < NV; ++j) d[j] = H * d[j];
#pragma unroll
        for (int i = MM - 1; i >= 0; --i) {
            if (i < currmem) {
                const double beta = fdiv(dotn<NV>(y[i], d), ys[i], rys[i]);
                const double coef = al[i] - beta;
#pragma unroll
                for (int j = 0; j < NV; ++j) d[j] += coef * s[i][j];
            }
        }
    }
};

// PANOCplus (Appendix A.2-A.5) on a variable of NV entries; returns k, solution in z
// WARP = false: one thread runs the solve.  WARP = true: the 32 lanes of a warp run the SAME solve in lock-step (identical
// registers), except inside the line search, where lane l evaluates the candidate the sequential loop would reach l
// steps later (see line_search_warp).  budget > 0: give up (return -1) once THIS inner solve has used more than `budget`
// AL-gradients at more than TB_HEAVY per iteration -- phase 1 of the two-phase batch.
template <int KIND, int MM, int SLACK, bool WARP>
__device__ long long panocplus(Al<KIND, SLACK>& al, const double (&x0)[Al<KIND, SLACK>::NV], const TbOpts& o, double tol,
                               double (&z)[Al<KIND, SLACK>::NV], int budget) {
    constexpr int NV = Al<KIND, SLACK>::NV;
    const double a = o.alpha;
    const int ng0 = al.ngrad;
    double x[NV], grad_x[NV], grad_z[NV], tmpg[NV], y[NV], res[NV];
    double x_prev[NV], res_prev[NV], x_d[NV], z_curr[NV], d[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) x[j] = x0[j];
    double f_x = al.gradient(x, grad_x);
    {
        double xe[NV];
#pragma unroll
        for (int j = 0; j < NV; ++j) xe[j] = x[j] + 1.0;
        al.gradient(xe, tmpg);
    }
    double df[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) df[j] = tmpg[j] - grad_x[j];
    const double Lest = sqrt(dotn<NV>(df, df)) / sqrt((double)NV);
    double gamma = a / Lest;
    double rg = frecip(gamma);   // 1/gamma, kept exact through the halvings (see fdiv)
    double aog = a / gamma;      // alpha/gamma for f_model: halving gamma doubles the correctly rounded quotient exactly
                                 // (while it is a normal number: otherwise it is divided afresh)
#pragma unroll
    for (int j = 0; j < NV; ++j) y[j] = x[j] - gamma * grad_x[j];
    double g_z = al.prox(y, gamma, z);
#pragma unroll
    for (int j = 0; j < NV; ++j) res[j] = x[j] - z[j];
    Lbfgs<MM, NV> H;
    H.M = o.mem;
    H.reset();
    double f_up = f_model<NV>(f_x, grad_x, res, aog);
    double f_z = al.gradient(z, grad_z);
    while (f_z > f_up + 10 * BZ_EPS * (1 + fabs(f_z)) && gamma >= o.minimum_gamma) {
        gamma /= 2;
        rg *= 2;
        aog *= 2;
        if (!(fabs(aog) >= 4.5e-308)) aog = a / gamma;
#pragma unroll
        for (int j = 0; j < NV; ++j) y[j] = x[j] - gamma * grad_x[j];
        g_z = al.prox(y, gamma, z);
#pragma unroll
        for (int j = 0; j < NV; ++j) res[j] = x[j] - z[j];
        f_up = f_model<NV>(f_x, grad_x, res, aog);
        f_z = al.gradient(z, grad_z);
    }
    double tau = 1.0;
    long long k = 1;
    bool spec = false;     // WARP mode: start the next line search speculatively (see below)
    (void)spec;
    for (;;) {
        double stopnorm = 0.0;
        bool isnan_ = false;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const double v = (fdiv(res[j], gamma, rg) - grad_x[j]) + grad_z[j];
            if (v != v) isnan_ = true; else stopnorm = fmax(stopnorm, fabs(v));
        }
        if (isnan_) stopnorm = __longlong_as_double(0x7ff8000000000000LL);
        if (k >= o.sub_maxit || stopnorm <= tol) { al.gz = g_z; return k; }
        // phase 1: an inner solve that back-tracks heavily (more than TB_HEAVY AL-gradients per iteration so far) and is past
        // its budget is left to the warp-per-instance kernel, where its line searches run across the lanes.  The test is
        // local to this inner solve: the heavy solves of the demos come late (outer iteration 30 of 32 on the slowest
        // either-or start), after thousands of cheap iterations that would hide them in a whole-solve average.  Solves that
        // merely iterate long (few back-tracks) stay: one lone lane is the fastest place for them.
        if (budget > 0 && al.ngrad - ng0 > budget && (long long)(al.ngrad - ng0) > TB_HEAVY * k) return -1;
#pragma unroll
        for (int j = 0; j < NV; ++j) { x_prev[j] = x[j]; res_prev[j] = res[j]; }
        const double FBE_x = f_model<NV>(f_x, grad_x, res, aog) + g_z;
        double nres[NV];
#pragma unroll
        for (int j = 0; j < NV; ++j) nres[j] = -res[j];
        H.mul(nres, d);
#pragma unroll
        for (int j = 0; j < NV; ++j) x_d[j] = x[j] + d[j];
        f_x = al.gradient(x_d, grad_x);
#pragma unroll
        for (int j = 0; j < NV; ++j) { x[j] = x_d[j]; z_curr[j] = z[j]; }
        tau = 1.0;
        const double sigma = o.beta * fdiv(0.5, gamma, rg) * (1 - a);
        const double nr = sqrt(dotn<NV>(res, res));
        const double thresh = FBE_x - sigma * (nr * nr) + 10 * BZ_EPS * (1 + fabs(FBE_x));
        if constexpr (!WARP) {
            for (int kb = 1; kb <= o.max_backtracks; ++kb) {
#pragma unroll
                for (int j = 0; j < NV; ++j) y[j] = x[j] - gamma * grad_x[j];
                g_z = al.prox(y, gamma, z);
#pragma unroll
                for (int j = 0; j < NV; ++j) res[j] = x[j] - z[j];
                f_up = f_model<NV>(f_x, grad_x, res, aog);
                f_z = al.gradient(z, grad_z);
                if (f_z > f_up + 10 * BZ_EPS * (1 + fabs(f_z)) && gamma >= o.minimum_gamma) {
                    gamma *= 0.5;
                    rg *= 2;
                    aog *= 2;
                    if (!(fabs(aog) >= 4.5e-308)) aog = a / gamma;
                    H.reset();
                    continue;
                }
                if (f_up + g_z <= thresh || kb >= o.max_backtracks) break;
                tau = (kb >= o.max_backtracks - 1) ? 0.0 : tau / 2;
#pragma unroll
                for (int j = 0; j < NV; ++j) x[j] = tau * x_d[j] + (1 - tau) * z_curr[j];
                f_x = al.gradient(x, grad_x);
            }
        } else {
            // ---- the same loop, its iterations evaluated speculatively by the lanes of the warp.  Between two gamma-halvings
            // the candidate of loop index kb + l is a function of (x_d, z_curr, gamma) and of l alone: tau follows the
            // deterministic schedule below, x = tau*x_d + (1 - tau)*z_curr.  Lane l evaluates it with the very statements of
            // the sequential loop; the first lane (in loop order) that either halves gamma or stops is what the sequential
            // loop would have reached, and its registers become everybody's.  A gamma-halving restarts the speculation at the
            // next loop index from that lane's x.  Bit-identical to the sequential loop by construction.
            // Speculation costs a broadcast of the winner's registers, so it is only used where it pays: an iteration starts
            // with the plain sequential step when the previous iteration accepted its first candidate (`spec` false), and
            // switches to speculation as soon as a candidate is rejected.  `carried` says whether lane 0's (x, f_x, grad_x)
            // already belong to loop index kb (true at the start and after a gamma-halving) or must be interpolated.
            const int lane = threadIdx.x & 31;
            int kb = 1;
            bool carried = true, backtracked = false;
            while (kb <= o.max_backtracks) {
                if (!spec && carried) {
                    // ---- loop index kb evaluated by every lane alike (identical registers: the sequential statements)
#pragma unroll
                    for (int j = 0; j < NV; ++j) y[j] = x[j] - gamma * grad_x[j];
                    g_z = al.prox(y, gamma, z);
#pragma unroll
                    for (int j = 0; j < NV; ++j) res[j] = x[j] - z[j];
                    f_up = f_model<NV>(f_x, grad_x, res, aog);
                    f_z = al.gradient(z, grad_z);
                    if (f_z > f_up + 10 * BZ_EPS * (1 + fabs(f_z)) && gamma >= o.minimum_gamma) {
                        gamma *= 0.5;
                        rg *= 2;
                        aog *= 2;
                        if (!(fabs(aog) >= 4.5e-308)) aog = a / gamma;
                        H.reset();
                        ++kb;
                        backtracked = true;
                        continue;
                    }
                    if (f_up + g_z <= thresh || kb >= o.max_backtracks) break;
                    tau = (kb >= o.max_backtracks - 1) ? 0.0 : tau / 2;      // tau of loop index kb + 1
                    ++kb;
                    carried = false;                                          // x of that index is not formed yet
                    spec = true;
                    backtracked = true;
                    continue;
                }
                // tau of loop index kb + lane, from the tau of index kb:  tau <- (k >= max_backtracks - 1) ? 0 : tau / 2 at k = kb .. kb+lane-1
                double tl = tau;
                for (int q = 0; q < lane; ++q) tl = (kb + q >= o.max_backtracks - 1) ? 0.0 : tl / 2;
                const int kl = kb + lane;
                double xl[NV], gxl[NV], fxl = f_x;
#pragma unroll
                for (int j = 0; j < NV; ++j) { xl[j] = x[j]; gxl[j] = grad_x[j]; }
                const bool live = kl <= o.max_backtracks;
                if ((lane > 0 || !carried) && live) {
#pragma unroll
                    for (int j = 0; j < NV; ++j) xl[j] = tl * x_d[j] + (1 - tl) * z_curr[j];
                    fxl = al.gradient(xl, gxl);
                }
                double zl[NV], rl[NV], gzl[NV], gz_l = 0.0, fup_l = 0.0, fz_l = 0.0;
                bool halve = false, stop = false;
                if (live) {
#pragma unroll
                    for (int j = 0; j < NV; ++j) y[j] = xl[j] - gamma * gxl[j];
                    gz_l = al.prox(y, gamma, zl);
#pragma unroll
                    for (int j = 0; j < NV; ++j) rl[j] = xl[j] - zl[j];
                    fup_l = f_model<NV>(fxl, gxl, rl, aog);
                    fz_l = al.gradient(zl, gzl);
                    halve = fz_l > fup_l + 10 * BZ_EPS * (1 + fabs(fz_l)) && gamma >= o.minimum_gamma;
                    stop = !halve && (fup_l + gz_l <= thresh || kl >= o.max_backtracks);
                }
                const unsigned ev = __ballot_sync(0xffffffffu, live && (halve || stop));
                const int src = __ffs(ev) - 1;        // never empty: the lane of loop index max_backtracks halves or stops
                const bool was_halve = __shfl_sync(0xffffffffu, halve ? 1 : 0, src) != 0;
                if (src > 0) backtracked = true;
                // the chosen lane's state becomes the warp's (including the AugLagFun side effects of its last gradient!)
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    x[j] = __shfl_sync(0xffffffffu, xl[j], src);
                    grad_x[j] = __shfl_sync(0xffffffffu, gxl[j], src);
                    z[j] = __shfl_sync(0xffffffffu, zl[j], src);
                    res[j] = __shfl_sync(0xffffffffu, rl[j], src);
                    grad_z[j] = __shfl_sync(0xffffffffu, gzl[j], src);
                }
                f_x = __shfl_sync(0xffffffffu, fxl, src);
                g_z = __shfl_sync(0xffffffffu, gz_l, src);
                f_up = __shfl_sync(0xffffffffu, fup_l, src);
                f_z = __shfl_sync(0xffffffffu, fz_l, src);
                tau = __shfl_sync(0xffffffffu, tl, src);
                al.fx = __shfl_sync(0xffffffffu, al.fx, src);
#pragma unroll
                for (int i = 0; i < Al<KIND, SLACK>::NY; ++i) {
                    al.cx[i] = __shfl_sync(0xffffffffu, al.cx[i], src);
                    al.s[i] = __shfl_sync(0xffffffffu, al.s[i], src);
                    al.yupd[i] = __shfl_sync(0xffffffffu, al.yupd[i], src);
                }
                al.ngrad = __shfl_sync(0xffffffffu, al.ngrad, 0);    // (not used in this mode; kept uniform)
                carried = true;
                if (!was_halve) break;
                gamma *= 0.5;
                rg *= 2;
                aog *= 2;
                if (!(fabs(aog) >= 4.5e-308)) aog = a / gamma;
                H.reset();
                backtracked = true;
                kb += src + 1;
            }
            spec = backtracked;            // the next iteration speculates from its first candidate iff this one back-tracked
        }
        double sv[NV], yv[NV];
#pragma unroll
        for (int j = 0; j < NV; ++j) { sv[j] = x[j] - x_prev[j]; yv[j] = res[j] - res_prev[j]; }
        H.update(sv, yv);
        k += 1;
    }
}

// One complete solve per thread (WARP = false) or per warp (WARP = true): SLACK = 0 `alps` (alps.jl:7-117), SLACK = 1 `als`
// (als.jl:7-120).
//
// Two-phase batch (the loop replaced is demo/mpvca.jl:241-281).  A launch lasts as long as its slowest
// instance, and the slow instances are slow because every PANOC+ iteration exhausts its 20 line-search back-tracks
// (40 dependent AL-gradients per iteration on one lone lane).  Phase 1 runs every instance on its own thread with a
// budget of AL-gradient evaluations per inner solve; an instance whose inner solve exceeds it records the state at the top
// of that outer iteration (TB_CKPT doubles), is appended to `surv` and abandoned.  Phase 2 resumes the survivors from that
// record, one WARP each, with the line search evaluated speculatively across the lanes: the dependent chain of an iteration
// shrinks from up to 41 gradients to 3, bit-identical results (same statements, same order; an outer iteration starts a
// fresh PANOC+ solve, so nothing but the record crosses the hand-over).
// Every thread of phase 1 reports here before it returns; the last one tells the warp-per-instance kernel that the queue gets
// no more entries (queue[2]).  (A separate one-thread kernel behind phase 1 used to do this -- and could not be scheduled
// when the persistent kernel's CTAs had taken every register of every SM: a dead-lock.)  The fence orders a thread's push
// before its count, so whoever sees the flag sees every push.
__device__ __forceinline__ void tb_phase1_exit(int* queue, long long total) {
    __threadfence();
    if (atomicAdd(&queue[5], 1) == (int)(total - 1)) *(volatile int*)&queue[2] = 1;
}

template <int KIND, int MM, int SLACK, bool WARP>
__global__ void __launch_bounds__(TB_THREADS, WARP ? 1 : TB_MINB)
    k_batch_alps(long long batch, const double* __restrict__ x0s, const double* __restrict__ y0s,
                 TbOpts o, double* __restrict__ xs, double* __restrict__ ys, int* __restrict__ tot_its,
                 long long* __restrict__ tot_inners, int* __restrict__ statuses,
                 double* __restrict__ nrps, int budget, int* surv, int* queue, double* ckpt, int drain) {
    // queue[0]: survivors pushed so far (slots reserved), queue[1]: next slot to pop, queue[2]: phase 1 has finished,
    // queue[3]: a persistent warp gave up waiting, queue[5]: threads of phase 1 that have returned.
    // surv[slot] is -1 until the pushing thread has stored the instance index.
    constexpr int NX = Prob<KIND>::NX;
    constexpr int NY = Prob<KIND>::NY;
    constexpr int NV = Al<KIND, SLACK>::NV;
    // Instance -> thread: lane l of warp w takes instance l*nwarps + w, so the 32 lanes of a warp hold instances that are
    // far apart in the batch.  The demos' batches are grids of starting points: neighbours fall into the same basin and are
    // slow together, and in launch order a warp that holds the slowest instance holds several more that take other
    // branches, each serialised against it (r1j, 65,536 MPVC starts: 38.3 ms in launch order, 21.9 ms this way -- the time
    // of the slowest instance alone; either-or 38.7 -> 36.3 ms).  Results do not depend on the mapping.
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long inst;
    if constexpr (!WARP) {
        const long long nwarps = (batch + 31) / 32;
        inst = (t & 31) * nwarps + (t >> 5);
        if ((t >> 5) >= nwarps || inst >= batch) { tb_phase1_exit(queue, (long long)gridDim.x * blockDim.x); return; }
    }
    // WARP: a persistent warp of the concurrent phase-2 kernel; it pops survivors until phase 1 is over and the queue is dry
    for (;;) {
    if constexpr (WARP) {
        int idx = 0;
        if ((threadIdx.x & 31) == 0) {
            // bounded wait: if phase 1 never reports (a failed launch, a device that cannot co-schedule the two kernels),
            // the warp gives up after TB_QUEUE_TIMEOUT_NS, raises queue[3] and the host returns BZ_ERR_CUDA
            unsigned long long t_start = 0;
            for (;;) {
                const int tail = *(volatile int*)&queue[0], head = *(volatile int*)&queue[1];
                if (head < tail) {
                    if (atomicCAS(&queue[1], head, head + 1) == head) { idx = head; break; }
                    continue;
                }
                if (*(volatile int*)&queue[2]) {
                    if (*(volatile int*)&queue[1] >= *(volatile int*)&queue[0]) { idx = -1; break; }
                    continue;
                }
                __nanosleep(400);
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t_start == 0) t_start = now;
                else if (now - t_start > TB_QUEUE_TIMEOUT_NS) { *(volatile int*)&queue[3] = 1; idx = -1; break; }
            }
            if (idx >= 0) { while ((inst = *(volatile int*)&surv[idx]) < 0) __nanosleep(100); __threadfence(); }
        }
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx < 0) return;
        inst = __shfl_sync(0xffffffffu, (int)inst, 0);
    }
    Al<KIND, SLACK> al;
    al.ngrad = 0;
    double x[NX], cx[NY], s[NY];
    double inner_tol = o.inner_tol;
    int tot_it = 0;
    bool have_old = false;
    long long tot_inner = 0;
    double nrp = __longlong_as_double(0x7ff8000000000000LL), nrp_old = nrp;
    double objx = 0.0;
    double dtmp[NX];
    if constexpr (WARP) {
        // resume at the top of the outer iteration phase 1 gave up in: everything the loop carries across iterations
        // (x, s, y, mu, inner_tol, the previous violation, the counters) was recorded by the thread that pushed the instance
        const double* ck = ckpt + (size_t)inst * TB_CKPT;
#pragma unroll
        for (int j = 0; j < NX; ++j) x[j] = __ldcg(ck + j);
#pragma unroll
        for (int i = 0; i < NY; ++i) { s[i] = __ldcg(ck + 6 + i); al.y[i] = __ldcg(ck + 10 + i); al.mu[i] = __ldcg(ck + 14 + i); }
        inner_tol = __ldcg(ck + 18);
        nrp = __ldcg(ck + 19);
        tot_it = (int)__ldcg(ck + 20);
        tot_inner = (long long)__ldcg(ck + 21);
        have_old = __ldcg(ck + 22) != 0.0;
    } else {
    {   // alps.jl:37-45 / als.jl:38-44
        double x0[NX];
#pragma unroll
        for (int j = 0; j < NX; ++j) x0[j] = x0s[inst * NX + j];
        al.gz = Prob<KIND>::prox(x0, BZ_EPS, x);
    }
    objx = Prob<KIND>::f_grad(x, dtmp) + al.gz;
    Prob<KIND>::c_eval(x, cx);
    Prob<KIND>::proj(cx, s);
#pragma unroll
    for (int i = 0; i < NY; ++i) {   // default_penalty_parameter!, safeguards.jl:13-18
        const double df = cx[i] - s[i];
        double m = jmax(1.0, 0.5 * (df * df)) / jmax(1.0, objx);
        m *= 0.1;
        al.mu[i] = jmax(1e-8, (m != m) ? m : fmin(m, 1e8));
        al.y[i] = y0s[inst * NY + i];
    }
    }
    bool solved = false, tired = tot_it >= o.maxit, broken = objx != objx;
    while (!(solved || tired || broken)) {
        tot_it++;
#pragma unroll
        for (int i = 0; i < NY; ++i) al.y[i] = fmax(-1e20, fmin(al.y[i], 1e20));   // safeguards.jl:2-5
        al.update();
        double v0[NV], z[NV];
#pragma unroll
        for (int j = 0; j < NX; ++j) v0[j] = x[j];
        if constexpr (SLACK != 0) {
#pragma unroll
            for (int i = 0; i < NY; ++i) v0[NX + i] = s[i];       // xSlack = [x; s], als.jl:69-70
        }
        // End game of phase 1: once no more than `drain` of its threads are still at work (the persistent kernel has that many
        // warps, nearly all idle by then), a solve moves to a warp of its own at its next outer iteration.  The long solves
        // that are left share warps with each other and run one after the other wherever they disagree; alone they run at
        // the speed of their dependent chain (the 200 longest either-or starts: 13.8 ms together, 9.6 ms for the slowest alone).
        bool move = false;
        if constexpr (!WARP) {
            if (budget > 0 && drain > 0 && tot_it > 1) {
                const long long total = (long long)gridDim.x * blockDim.x;
                move = total - (long long)*(volatile int*)&queue[5] <= (long long)drain;
            }
        }
        const long long sub_it = move ? -1 : panocplus<KIND, MM, SLACK, WARP>(al, v0, o, inner_tol, z, WARP ? 0 : budget);
        if (sub_it < 0) {                                         // a heavy back-tracker over the budget: the warp kernel takes it
            // ... from the top of THIS outer iteration: x, s, mu, inner_tol and the counters have not moved since, y is
            // the clamped y (clamping again changes nothing)
            double* ck = ckpt + (size_t)inst * TB_CKPT;
#pragma unroll
            for (int j = 0; j < NX; ++j) __stcg(ck + j, x[j]);
#pragma unroll
            for (int i = 0; i < NY; ++i) { __stcg(ck + 6 + i, s[i]); __stcg(ck + 10 + i, al.y[i]); __stcg(ck + 14 + i, al.mu[i]); }
            __stcg(ck + 18, inner_tol);
            __stcg(ck + 19, nrp);
            __stcg(ck + 20, (double)(tot_it - 1));
            __stcg(ck + 21, (double)tot_inner);
            __stcg(ck + 22, have_old ? 1.0 : 0.0);
            const int slot = atomicAdd(&queue[0], 1);
            __threadfence();
            *(volatile int*)&surv[slot] = (int)inst;
            tb_phase1_exit(queue, (long long)gridDim.x * blockDim.x);
            return;
        }
#pragma unroll
        for (int j = 0; j < NX; ++j) x[j] = z[j];
        tot_inner += sub_it;
        const bool sub_solved = sub_it < o.subsolver_maxit;
        if constexpr (SLACK != 0) {
#pragma unroll
            for (int i = 0; i < NY; ++i) s[i] = z[NX + i];
            objx = Prob<KIND>::f_grad(x, dtmp) + al.gz;           // f(x) + gSlack.gz, als.jl:78
            Prob<KIND>::c_eval(x, cx);
#pragma unroll
            for (int i = 0; i < NY; ++i) al.y[i] = al.y[i] + fdiv(cx[i] - s[i], al.mu[i], al.rmu[i]);   // als.jl:83
        } else {
            objx = al.fx + al.gz;
            Prob<KIND>::c_eval(x, cx);
            double t[NY];
#pragma unroll
            for (int i = 0; i < NY; ++i) t[i] = cx[i] + al.muy[i];
            Prob<KIND>::proj(t, s);
#pragma unroll
            for (int i = 0; i < NY; ++i) al.y[i] = fdiv(t[i] - s[i], al.mu[i], al.rmu[i]);
        }
        nrp_old = nrp;
        nrp = 0.0;
#pragma unroll
        for (int i = 0; i < NY; ++i) nrp = jmax(nrp, fabs(cx[i] - s[i]));
        solved = (inner_tol <= o.tol_dual && sub_solved) && (nrp <= o.tol_prim);
        tired = tot_it >= o.maxit;
        broken = objx != objx;
        if (!(solved || tired || broken)) {
            if (have_old && nrp > jmax(o.theta_penalty * nrp_old, o.tol_prim)) {
#pragma unroll
                for (int i = 0; i < NY; ++i) al.mu[i] *= o.kappa_penalty;
            }
            inner_tol = jmax(o.kappa_tol * inner_tol, o.tol_dual);
        }
        have_old = true;
    }
    if (!WARP || (threadIdx.x & 31) == 0) {                       // WARP: every lane holds the same result
#pragma unroll
        for (int j = 0; j < NX; ++j) xs[inst * NX + j] = x[j];
#pragma unroll
        for (int i = 0; i < NY; ++i) ys[inst * NY + i] = al.y[i];
        tot_its[inst] = tot_it;
        tot_inners[inst] = tot_inner;
        statuses[inst] = solved ? 0 : tired ? 1 : broken ? 2 : 3;
        nrps[inst] = nrp;
    }
    if constexpr (!WARP) { tb_phase1_exit(queue, (long long)gridDim.x * blockDim.x); return; }
    }   // for (;;): next survivor
}


}  // namespace

static int batch_solve(bz_ctx* ctx, int slack, int kind, int64_t batch, const double* x0, const double* y0,
                       const bz_options* opt, double* x, double* y, int32_t* tot_it,
                       int64_t* tot_inner_it, int32_t* status, double* norm_res_prim,
                       double* kernel_ms) {
    if (!ctx || !x0 || !y0 || !opt || !x || !y || !tot_it || !tot_inner_it || !status || !norm_res_prim)
        return bz_fail(ctx, BZ_ERR_ARG, "bz_batch_solve: NULL argument");
    if (kind < 0 || kind > 7) return bz_fail(ctx, BZ_ERR_UNSUPPORTED, "bz_batch_solve: unknown problem kind");
    if (batch < 0) return bz_fail(ctx, BZ_ERR_ARG, "batch must be nonnegative");
    if (opt->lbfgs_memory < 0 || opt->lbfgs_memory > TB_MAXMEM)
        return bz_fail(ctx, BZ_ERR_ARG, "batched path supports lbfgs_memory in [0, 8]");
    BZ_CUDA(ctx, cudaSetDevice(ctx->device));
    if (kernel_ms) *kernel_ms = 0.0;
    if (batch == 0) return BZ_OK;
    const int nx = kind <= 3 ? 2 : (kind == 6 ? 4 : 6);
    const int ny = (kind == 0 || kind == 6) ? 2 : 4;
    TbOpts o;
    o.tol_prim = opt->tol_prim >= 0 ? opt->tol_prim : opt->tol;
    o.tol_dual = opt->tol_dual >= 0 ? opt->tol_dual : opt->tol;
    o.inner_tol = opt->inner_tol >= 0 ? opt->inner_tol : cbrt(o.tol_dual);
    o.maxit = opt->maxit;
    o.theta_penalty = opt->theta_penalty; o.kappa_penalty = opt->kappa_penalty; o.kappa_tol = opt->kappa_tol;
    o.subsolver_maxit = opt->subsolver_maxit;
    o.alpha = opt->alpha; o.beta = opt->beta; o.minimum_gamma = opt->minimum_gamma;
    o.max_backtracks = opt->max_backtracks; o.mem = opt->lbfgs_memory;
    o.sub_maxit = opt->sub_maxit;

    // one grow-only device arena per context, carved into the eight per-instance arrays: no cudaMalloc / cudaFree (and
    // their implicit synchronisations) on the second and later calls
    const size_t B = (size_t)batch;
    const size_t need = B * (2 * (size_t)nx + 2 * (size_t)ny + 1 + TB_CKPT) * sizeof(double) + B * sizeof(long long) + 3 * B * sizeof(int) + 128;
    auto cleanup = [&]() {};
#define BTRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); \
        return bz_fail(ctx, BZ_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); } } while (0)
    if (ctx->batch_arena_bytes < need) {
        if (ctx->batch_arena) cudaFree(ctx->batch_arena);
        ctx->batch_arena = nullptr;
        ctx->batch_arena_bytes = 0;
        BTRY(cudaMalloc(&ctx->batch_arena, need));
        ctx->batch_arena_bytes = need;
    }
    double* dx0 = reinterpret_cast<double*>(ctx->batch_arena);
    double* dy0 = dx0 + B * nx;
    double* dx = dy0 + B * ny;
    double* dy = dx + B * nx;
    double* dnrp = dy + B * ny;
    double* dck = dnrp + B;        // hand-over records of the instances phase 1 gives up on (TB_CKPT doubles each)
    long long* din = reinterpret_cast<long long*>(dck + B * TB_CKPT);
    int* dit = reinterpret_cast<int*>(din + B);
    int* dst = dit + B;
    int* dsurv = dst + B;          // survivors of phase 1 (instance indices, -1 = not stored yet), then the queue words
    int* dnsurv = dsurv + B;
    // phase-1 budget of AL-gradients per instance (0: single phase).  The speculative line search covers the whole loop
    // with one warp when max_backtracks <= 32.
    const int budget = (ctx->batch_budget > 0 && opt->max_backtracks <= 32 && opt->max_backtracks >= 1) ? ctx->batch_budget : 0;
    BTRY(cudaMemcpyAsync(dx0, x0, batch * nx * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    BTRY(cudaMemcpyAsync(dy0, y0, batch * ny * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    cudaEvent_t e0, e1;
    BTRY(cudaEventCreate(&e0));
    BTRY(cudaEventCreate(&e1));
    const unsigned grid = (unsigned)((batch + TB_THREADS - 1) / TB_THREADS);
    cudaEventRecord(e0, ctx->stream);
    // Two kernels run side by side.  Phase 1 goes first and takes the SMs; the warp-per-instance kernel follows on a second
    // stream: its persistent warps move in as CTAs of phase 1 retire (the median solve is short, so that is early) and pop
    // survivors from the queue as phase 1 produces them.  Phase 1 never waits for it, so nothing can dead-lock; launched the
    // other way round, the persistent kernel held 45 % of every SM's registers from the start and pushed phase 1 into a
    // third wave (either-or grid 27 ms instead of 17), and a phase-1 launch that had to grow the device's local-memory
    // reservation waited for an idle device that never came.
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    const bool concurrent = budget > 0 && ctx->batch_concurrent;     // else: phase 2 after phase 1 on the same stream
    if (budget > 0) BTRY(cudaMemsetAsync(dsurv, 0xff, B * sizeof(int), ctx->stream));
    BTRY(cudaMemsetAsync(dnsurv, 0, 16 * sizeof(int), ctx->stream));      // (words 8..15: scratch queue of the empty launches below)
    if (concurrent) {
        if (!ctx->stream2) BTRY(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
        BTRY(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        BTRY(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
        BTRY(cudaEventRecord(ev_fork, ctx->stream));
        BTRY(cudaStreamWaitEvent(ctx->stream2, ev_fork, 0));
    }
    const int p2ctas = ctx->batch_p2ctas > 0 ? ctx->batch_p2ctas : (concurrent ? 2 : 4);
    const unsigned grid2 = (unsigned)(p2ctas * ctx->num_sms);
    // phase 1 hands its last solves to the persistent warps (one each) once fewer of its threads are left than there are such warps
    const int drain = (concurrent && ctx->batch_drain != 0) ? (ctx->batch_drain > 0 ? ctx->batch_drain : (int)(grid2 * (TB_THREADS / 32))) : 0;
    // L-BFGS history is a register-resident shift register sized at compile time: 5 (the demos') or 8
    // (lazy module loading: the first launch of a kernel may wait for running kernels to finish -- fatal when the running
    // kernel is the persistent one that waits for the kernel being loaded.  Both are loaded before either is launched.)
#define TB_LAUNCH2(K, MMV, SL) do { \
        if (concurrent) { const unsigned bit__ = 1u << ((K) * 4 + ((MMV) > 5 ? 2 : 0) + (SL)); \
            if (!(ctx->batch_presized & bit__)) { \
                /* once per kernel pair and context: an empty launch of each (no instances; a scratch queue that says */ \
                /* "phase 1 is over").  It loads the modules and grows the device's local-memory reservation to the    */ \
                /* larger stack frame now -- a launch that has to grow it waits for an idle device, and with the other  */ \
                /* kernel of the pair already running and waiting for it that is a dead-lock.                         */ \
                cudaMemsetAsync(dnsurv + 8, 1, 4 * sizeof(int), ctx->stream);   /* tail = head, "phase 1 is over" != 0 */ \
                k_batch_alps<K, MMV, SL, false><<<1, TB_THREADS, 0, ctx->stream>>>(0, dx0, dy0, o, dx, dy, dit, din, dst, dnrp, 0, dsurv, dnsurv + 8, dck, 0); \
                k_batch_alps<K, MMV, SL, true><<<1, TB_THREADS, 0, ctx->stream>>>(0, dx0, dy0, o, dx, dy, dit, din, dst, dnrp, 0, dsurv, dnsurv + 8, dck, 0); \
                BTRY(cudaStreamSynchronize(ctx->stream)); \
                ctx->batch_presized |= bit__; ctx->n_launches += 2; } } \
        k_batch_alps<K, MMV, SL, false><<<grid, TB_THREADS, 0, ctx->stream>>>(batch, dx0, dy0, o, dx, dy, dit, din, dst, dnrp, budget, dsurv, dnsurv, dck, drain); \
        if (concurrent) { k_batch_alps<K, MMV, SL, true><<<grid2, TB_THREADS, 0, ctx->stream2>>>(batch, dx0, dy0, o, dx, dy, dit, din, dst, dnrp, 0, dsurv, dnsurv, dck, 0); ctx->n_launches++; } \
        if (budget > 0 && !concurrent) { \
            k_batch_alps<K, MMV, SL, true><<<grid2, TB_THREADS, 0, ctx->stream>>>(batch, dx0, dy0, o, dx, dy, dit, din, dst, dnrp, 0, dsurv, dnsurv, dck, 0); \
            ctx->n_launches++; } \
    } while (0)
#define TB_LAUNCH(K, MMV) do { if (slack) TB_LAUNCH2(K, MMV, 1); else TB_LAUNCH2(K, MMV, 0); } while (0)
#define TB_DISPATCH(K) do { if (o.mem <= 5) TB_LAUNCH(K, 5); else TB_LAUNCH(K, TB_MAXMEM); } while (0)
    switch (kind) {
    case 0: TB_DISPATCH(0); break;
    case 1: TB_DISPATCH(1); break;
    case 2: TB_DISPATCH(2); break;
    case 3: TB_DISPATCH(3); break;
    case 4: TB_DISPATCH(4); break;
    case 5: TB_DISPATCH(5); break;
    case 6: TB_DISPATCH(6); break;
    default: TB_DISPATCH(7); break;
    }
#undef TB_DISPATCH
#undef TB_LAUNCH
#undef TB_LAUNCH2
    if (concurrent) {
        cudaEventRecord(ev_join, ctx->stream2);
        cudaStreamWaitEvent(ctx->stream, ev_join, 0);
    }
    cudaEventRecord(e1, ctx->stream);
    ctx->n_launches++;
    BTRY(cudaGetLastError());
    BTRY(cudaMemcpyAsync(x, dx, batch * nx * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    BTRY(cudaMemcpyAsync(y, dy, batch * ny * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    BTRY(cudaMemcpyAsync(norm_res_prim, dnrp, batch * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    BTRY(cudaMemcpyAsync(tot_it, dit, batch * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    BTRY(cudaMemcpyAsync(status, dst, batch * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    BTRY(cudaMemcpyAsync(tot_inner_it, din, batch * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    int qwords[4] = {0, 0, 0, 0};
    if (budget > 0) BTRY(cudaMemcpyAsync(qwords, dnsurv, sizeof(qwords), cudaMemcpyDeviceToHost, ctx->stream));
    BTRY(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (kernel_ms) *kernel_ms = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ev_fork) { cudaEventDestroy(ev_fork); cudaEventDestroy(ev_join); }
    cleanup();
#undef BTRY
    if (qwords[3]) return bz_fail(ctx, BZ_ERR_CUDA, "bz_batch_solve: the warp-per-instance kernel timed out waiting for phase 1 "
                                  "(the two kernels were not co-scheduled); set BZ_BATCH_CONCURRENT=0");
    return BZ_OK;
}

// ---- fp64 issue-rate microbenchmark ("measure an FMA microbenchmark"): the denominator of the batched
// path's secondary roofline.  Every thread runs 8 independent chains of dependent DFMAs.
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters, double a, double b) {
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (double)(threadIdx.x + k);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fma(v[k], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[k];
    if (s == 123.456) out[0] = s;      // never true: keeps the chains alive
}
extern "C" int bz_fp64_fma_peak(bz_ctx* ctx, double* tflops) {
    if (!ctx || !tflops) return BZ_ERR_ARG;
    BZ_CUDA(ctx, cudaSetDevice(ctx->device));
    double* d = nullptr;
    BZ_CUDA(ctx, cudaMalloc((void**)&d, 64));
    const int iters = 4096, blocks = ctx->num_sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0, ctx->stream);
        k_fp64_peak<<<blocks, threads, 0, ctx->stream>>>(d, iters, 0.999999, 1e-9);
        cudaEventRecord(e1, ctx->stream);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    ctx->n_launches += 4;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    BZ_CUDA(ctx, cudaGetLastError());
    *tflops = 2.0 * 8.0 * iters * (double)blocks * threads / (best * 1e-3) / 1e12;
    return BZ_OK;
}

extern "C" int bz_batch_solve(bz_ctx* ctx, int kind, int64_t batch, const double* x0, const double* y0,
                              const bz_options* opt, double* x, double* y, int32_t* tot_it,
                              int64_t* tot_inner_it, int32_t* status, double* norm_res_prim,
                              double* kernel_ms) {
    return batch_solve(ctx, 0, kind, batch, x0, y0, opt, x, y, tot_it, tot_inner_it, status, norm_res_prim, kernel_ms);
}

extern "C" int bz_batch_solve_als(bz_ctx* ctx, int kind, int64_t batch, const double* x0, const double* y0,
                                  const bz_options* opt, double* x, double* y, int32_t* tot_it,
                                  int64_t* tot_inner_it, int32_t* status, double* norm_res_prim,
                                  double* kernel_ms) {
    return batch_solve(ctx, 1, kind, batch, x0, y0, opt, x, y, tot_it, tot_inner_it, status, norm_res_prim, kernel_ms);
}

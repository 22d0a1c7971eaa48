/*
 * CPU ORACLE (test infrastructure -- NOT a product path): plain-C restatement of one full
 * `Bazinga.alps` solve for the tiny closed-form problems the demos run over grids of starting
 * points.  It follows oracle/bazinga_oracle.py statement by statement (which in
 * turn follows src/algorithms/alps.jl:7-117, src/utilities/auglagfun.jl:73-101,
 * src/utilities/safeguards.jl:2-18, src/projections/orConstraints.jl:7-36,
 * src/projections/vanishingConstraints.jl:27-46 and SURVEY.md Appendix A for the third-party
 * PANOCplus / LBFGS), with sums evaluated left to right and no FMA contraction
 * (build with -ffp-contract=off), so the batched CUDA kernel can be checked against it bit for bit.
 *
 * Problem kinds (same numbering as bz_batch_solve in include/bazinga_b200.h):
 *   0 rosenbrock  demo/rosenbrock.jl:39-80   nx=2 ny=2  f=w(x2+1-(x1+1)^2)^2, g=lambda|x1|, D=EITHEROR
 *   1 either-or   demo/eitheror.jl:58-126    nx=2 ny=4  D = 2 EITHEROR pairs
 *   2 xor         demo/eitheror.jl:128-132   nx=2 ny=4  D = 2 XOR pairs
 *   3 mpvca       demo/mpvca.jl:77-150       nx=2 ny=4  g = max(0,x), D = 2 VC pairs
 *   4 eitheror_fullslack  demo/eitheror.jl:76-82,106-119   nx=6 ny=4  x = [x1,x2,s1..s4], g projects (s1,s2),(s3,s4) onto EITHEROR, D = {0}
 *   5 xor_fullslack       demo/eitheror.jl:84-90,106-119   nx=6 ny=4  the same with XOR
 *   6 mpvca_slack         demo/mpvca.jl:95-100,120-128     nx=4 ny=2  x = [x1,x2,s1,s2], g projects (x1,s1),(x2,s2) onto VC, D = {0}
 *   7 mpvca_fullslack     demo/mpvca.jl:102-108,130-143    nx=6 ny=4  g = max(0,.) on x1,x2 and VC on (s1,s2),(s3,s4), D = {0}
 *
 * tiny_als / tiny_als_batch: the same problems through `Bazinga.als` (src/algorithms/als.jl:7-120) -- PANOCplus
 * on the stacked variable [x; s] with AugLagFunSlack / NonsmoothCostFunSlack (src/utilities/auglagfunslack.jl).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's shared object.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define TNX 6            /* largest nx (the *_fullslack variants) */
#define TNY 4
#define TNV (TNX + TNY)   /* largest PANOCplus variable: [x; s] of als */
#define TMEM 16
#define T_EPS 2.220446049250313e-16

typedef struct {
    double tol, tol_prim, tol_dual, inner_tol;
    int maxit;
    double theta_penalty, kappa_penalty, kappa_tol;
    long long subsolver_maxit;
    double alpha, beta, minimum_gamma;
    int max_backtracks, lbfgs_memory;
    long long sub_maxit;
} tiny_opts;

typedef struct {
    int kind, nx, ny;
    int slack, nv;   /* slack = 1: AugLagFunSlack / NonsmoothCostFunSlack on [x; s] (nv = nx + ny), else nv = nx */
    double mu[TNY], y[TNY], muy[TNY], musqy;
    double cx[TNY], s[TNY], yupd[TNY], fx;
    double gz;
    long long n_grad, n_prox;
} tiny_al;

/* ---- 2-element projections ------------------------------------------------------------------ */
static void proj_eitheror(const double* x, double* z) { /* orConstraints.jl:7-17 */
    z[0] = x[0]; z[1] = x[1];
    if (x[0] < 0 && x[1] < 0) { if (x[0] > x[1]) z[0] = 0.0; else z[1] = 0.0; }
}
static void proj_xor(const double* x, double* z) { /* orConstraints.jl:24-36 */
    z[0] = x[0]; z[1] = x[1];
    if (x[0] * x[1] > 0) {
        if (x[0] > x[1]) { z[0] = fmax(0.0, x[0]); z[1] = fmin(0.0, x[1]); }
        else { z[0] = fmin(0.0, x[0]); z[1] = fmax(0.0, x[1]); }
    }
}
static void proj_vc(const double* x, double* z) { /* vanishingConstraints.jl:27-46 */
    z[0] = 0.0; z[1] = 0.0;
    if (x[0] <= 0) z[1] = x[1];
    else if (x[1] >= 0) { z[0] = x[0]; z[1] = x[1]; }
    else {
        double sm = x[0] + x[1];
        if (sm > 0) z[0] = x[0];
        else if (sm < 0) z[1] = x[1];
        else z[1] = x[1];
    }
}
static void kind_dims(int kind, int* nx, int* ny) {
    *nx = (kind <= 3) ? 2 : (kind == 6 ? 4 : 6);
    *ny = (kind == 0 || kind == 6) ? 2 : 4;
}
static void projD(int kind, const double* v, double* z) {
    switch (kind) {
    case 0: proj_eitheror(v, z); break;                               /* demo/rosenbrock.jl:77-80 */
    case 1: proj_eitheror(v, z); proj_eitheror(v + 2, z + 2); break;  /* demo/eitheror.jl:122-126 */
    case 2: proj_xor(v, z); proj_xor(v + 2, z + 2); break;            /* demo/eitheror.jl:128-132 */
    case 3: proj_vc(v, z); proj_vc(v + 2, z + 2); break;              /* demo/mpvca.jl:146-150 */
    case 6: z[0] = 0.0; z[1] = 0.0; break;                            /* Bazinga.ZeroSet, zeroSet.jl:17-20 */
    default: z[0] = z[1] = z[2] = z[3] = 0.0; break;                  /* 4, 5, 7: ZeroSet */
    }
}

/* ---- oracles --------------------------------------------------------------------------------- */
static double f_grad(int kind, const double* x, double* dfx) {
    int nx, ny;
    kind_dims(kind, &nx, &ny);
    for (int j = 2; j < nx; ++j) dfx[j] = 0.0;   /* dfx .= 0 (eitheror.jl:65, mpvca.jl:84) */
    if (kind == 0) { /* demo/rosenbrock.jl:45-50, w = 10 */
        const double w = 10.0;
        double tmp = x[1] + 1 - (x[0] + 1) * (x[0] + 1);
        dfx[0] = -4 * w * tmp * (x[0] + 1);
        dfx[1] = 2 * w * tmp;
        return w * (tmp * tmp);
    } else if (kind == 1 || kind == 2 || kind == 4 || kind == 5) { /* demo/eitheror.jl:64-68, c = (8,-3): norm(x[1:2]-c)^2 */
        double d0 = x[0] - 8.0, d1 = x[1] - (-3.0);
        dfx[0] = 2 * d0; dfx[1] = 2 * d1;
        double nr = sqrt(d0 * d0 + d1 * d1);
        return nr * nr;
    } else { /* demo/mpvca.jl:83-87, c = (4,2): dot(c, x[1:2]) */
        dfx[0] = 4.0; dfx[1] = 2.0;
        return 4.0 * x[0] + 2.0 * x[1];
    }
}
static double f_value(int kind, const double* x) { double t[TNX]; return f_grad(kind, x, t); }

static void c_eval(int kind, const double* x, double* cx) {
    if (kind == 0) { cx[0] = -x[0] - x[1]; cx[1] = x[1] - x[0]; }       /* rosenbrock.jl:67-70 */
    else if (kind == 1 || kind == 2) {                                   /* eitheror.jl:93-99 */
        cx[0] = 2 * x[1] - x[0] - 4;
        cx[1] = 2 - x[0];
        cx[2] = 4 * x[1] - x[0] * x[0];
        cx[3] = 10 - (x[0] - 3) * (x[0] - 3) - (x[1] - 1) * (x[1] - 1);
    } else if (kind == 4 || kind == 5) {                                 /* eitheror.jl:107-113 ConstraintORfullslack */
        cx[0] = 2 * x[1] - x[0] - 4 - x[2];
        cx[1] = 2 - x[0] - x[3];
        cx[2] = 4 * x[1] - x[0] * x[0] - x[4];
        cx[3] = 10 - (x[0] - 3) * (x[0] - 3) - (x[1] - 1) * (x[1] - 1) - x[5];
    } else if (kind == 3) {                                              /* mpvca.jl:111-114 */
        cx[0] = x[0];
        cx[1] = x[0] + x[1] - 5.0 * sqrt(2.0);
        cx[2] = x[1];
        cx[3] = x[0] + x[1] - 5.0;
    } else if (kind == 6) {                                              /* mpvca.jl:121-124 ConstraintMPVCAslack */
        cx[0] = x[0] + x[1] - 5.0 * sqrt(2.0) - x[2];
        cx[1] = x[0] + x[1] - 5.0 - x[3];
    } else {                                                             /* mpvca.jl:131-139 ConstraintMPVCAfullslack */
        cx[0] = x[0] - x[2];
        cx[1] = x[0] + x[1] - 5.0 * sqrt(2.0) - x[3];
        cx[2] = x[1] - x[4];
        cx[3] = x[0] + x[1] - 5.0 - x[5];
    }
}
static void c_jtprod(int kind, const double* x, const double* v, double* jtv) {
    if (kind == 0) { jtv[0] = -v[0] - v[1]; jtv[1] = v[1] - v[0]; }     /* rosenbrock.jl:71-74 */
    else if (kind == 1 || kind == 2 || kind == 4 || kind == 5) {         /* eitheror.jl:100-104, 114-119 */
        jtv[0] = -v[0] - v[1] - 2 * x[0] * v[2] + 2 * (3 - x[0]) * v[3];
        jtv[1] = 2 * v[0] + 4 * v[2] + 2 * (1 - x[1]) * v[3];
        if (kind >= 4) { jtv[2] = -v[0]; jtv[3] = -v[1]; jtv[4] = -v[2]; jtv[5] = -v[3]; }
    } else if (kind == 6) {                                              /* mpvca.jl:125-128 */
        jtv[0] = v[0] + v[1]; jtv[1] = v[0] + v[1]; jtv[2] = -v[0]; jtv[3] = -v[1];
    } else {                                                             /* mpvca.jl:115-118, 140-143 */
        jtv[0] = v[0] + v[1] + v[3];
        jtv[1] = v[1] + v[2] + v[3];
        if (kind == 7) { jtv[2] = -v[0]; jtv[3] = -v[1]; jtv[4] = -v[2]; jtv[5] = -v[3]; }
    }
}
static double g_prox(int kind, const double* x, double gamma, double* z) {
    if (kind == 0) { /* rosenbrock.jl:55-64, lambda = 1 */
        const double lambda = 1.0;
        double gl = gamma * lambda;
        if (fabs(x[0]) <= gl) z[0] = 0.0;
        else z[0] = (x[0] > 0 ? 1.0 : -1.0) * (fabs(x[0]) - gl);
        z[1] = x[1];
        return lambda * fabs(z[0]);
    } else if (kind == 1 || kind == 2) { z[0] = x[0]; z[1] = x[1]; return 0.0; } /* eitheror.jl:71-74 */
    else if (kind == 4 || kind == 5) {                                             /* eitheror.jl:76-90 */
        z[0] = x[0]; z[1] = x[1];
        if (kind == 4) { proj_eitheror(x + 2, z + 2); proj_eitheror(x + 4, z + 4); }
        else { proj_xor(x + 2, z + 2); proj_xor(x + 4, z + 4); }
        return 0.0;
    } else if (kind == 3) { z[0] = fmax(0.0, x[0]); z[1] = fmax(0.0, x[1]); return 0.0; }   /* mpvca.jl:90-93 */
    else if (kind == 6) {                                                          /* mpvca.jl:95-100: pairs (x1,s1), (x2,s2) */
        double a[2] = {x[0], x[2]}, b[2] = {x[1], x[3]}, za[2], zb[2];
        proj_vc(a, za); proj_vc(b, zb);
        z[0] = za[0]; z[2] = za[1]; z[1] = zb[0]; z[3] = zb[1];
        return 0.0;
    } else {                                                                       /* mpvca.jl:102-108 */
        z[0] = fmax(0.0, x[0]); z[1] = fmax(0.0, x[1]);
        proj_vc(x + 2, z + 2); proj_vc(x + 4, z + 4);
        return 0.0;
    }
}

/* ---- AugLagFun (auglagfun.jl:73-101) ------------------------------------------------------------- */
static void al_update(tiny_al* al) {
    double acc = 0.0;
    for (int i = 0; i < al->ny; ++i) { al->muy[i] = al->mu[i] * al->y[i]; acc += al->muy[i] * al->y[i]; }
    al->musqy = 0.5 * acc;
}
static double al_gradient(tiny_al* al, const double* x, double* dlx) {
    double dfx[TNX], jtv[TNX], acc = 0.0;
    al->n_grad++;
    if (al->slack) { /* gradient!(dFxs, F::AugLagFunSlack, xs), auglagfunslack.jl:77-97 */
        const double* sl = x + al->nx;
        al->fx = f_grad(al->kind, x, dfx);
        c_eval(al->kind, x, al->cx);
        for (int i = 0; i < al->ny; ++i) {
            double t = (al->cx[i] + al->muy[i]) - sl[i];
            acc += (t * t) / al->mu[i];
        }
        double Fxs = 0.5 * acc;
        Fxs += al->fx;
        Fxs -= al->musqy;
        for (int i = 0; i < al->ny; ++i) al->yupd[i] = al->y[i] + (al->cx[i] - sl[i]) / al->mu[i];
        c_jtprod(al->kind, x, al->yupd, jtv);
        for (int j = 0; j < al->nx; ++j) dlx[j] = dfx[j] + jtv[j];
        for (int i = 0; i < al->ny; ++i) dlx[al->nx + i] = -al->yupd[i];
        return Fxs;
    }
    c_eval(al->kind, x, al->cx);
    for (int i = 0; i < al->ny; ++i) al->yupd[i] = al->cx[i] + al->muy[i];
    projD(al->kind, al->yupd, al->s);
    for (int i = 0; i < al->ny; ++i) al->yupd[i] -= al->s[i];
    for (int i = 0; i < al->ny; ++i) acc += (al->yupd[i] * al->yupd[i]) / al->mu[i];
    double lx = 0.5 * acc;
    for (int i = 0; i < al->ny; ++i) al->yupd[i] /= al->mu[i];
    al->fx = f_grad(al->kind, x, dfx);
    lx += al->fx;
    lx -= al->musqy;
    c_jtprod(al->kind, x, al->yupd, jtv);
    for (int j = 0; j < al->nx; ++j) dlx[j] = dfx[j] + jtv[j];
    return lx;
}
static double al_prox(tiny_al* al, const double* x, double gamma, double* z) {
    al->n_prox++;
    al->gz = g_prox(al->kind, x, gamma, z);
    if (al->slack) projD(al->kind, x + al->nx, z + al->nx);   /* NonsmoothCostFunSlack, auglagfunslack.jl:147-164 */
    return al->gz;
}

/* ---- LBFGS + PANOCplus ------------------------------------------------------- */
typedef struct {
    int M, currmem, curridx, nv;
    double H, s[TMEM][TNV], y[TMEM][TNV], ys[TMEM], alphas[TMEM];
} tiny_lbfgs;

/* left-to-right dot product (for nv = 2: a[0]*b[0] + a[1]*b[1]) */
static double dotn(const double* a, const double* b, int n) {
    double acc = a[0] * b[0];
    for (int j = 1; j < n; ++j) acc += a[j] * b[j];
    return acc;
}

static void lbfgs_update(tiny_lbfgs* L, const double* s, const double* yv) {
    if (L->M == 0) return;
    const int n = L->nv;
    double ys = dotn(s, yv, n);
    if (ys > 0) {
        L->curridx += 1;
        if (L->curridx > L->M) L->curridx = 1;
        L->currmem = L->currmem + 1 > L->M ? L->M : L->currmem + 1;
        int i = L->curridx - 1;
        L->ys[i] = ys;
        for (int j = 0; j < n; ++j) { L->s[i][j] = s[j]; L->y[i][j] = yv[j]; }
        L->H = ys / dotn(yv, yv, n);
    }
}
static void lbfgs_mul(tiny_lbfgs* L, const double* v, double* d) {
    const int n = L->nv;
    for (int j = 0; j < n; ++j) d[j] = v[j];
    int idx = L->curridx;
    for (int t = 0; t < L->currmem; ++t) {
        int i = idx - 1;
        L->alphas[i] = dotn(L->s[i], d, n) / L->ys[i];
        for (int j = 0; j < n; ++j) d[j] -= L->alphas[i] * L->y[i][j];
        idx -= 1;
        if (idx == 0) idx = L->M;
    }
    for (int j = 0; j < n; ++j) d[j] = L->H * d[j];
    for (int t = 0; t < L->currmem; ++t) {
        idx += 1;
        if (idx > L->M) idx = 1;
        int i = idx - 1;
        double beta = dotn(L->y[i], d, n) / L->ys[i];
        double coef = L->alphas[i] - beta;
        for (int j = 0; j < n; ++j) d[j] += coef * L->s[i][j];
    }
}
static double f_model(double fx, const double* grad, const double* res, double L, int n) {
    double nr = sqrt(dotn(res, res, n));
    return fx - dotn(grad, res, n) + (L / 2) * (nr * nr);
}

/* returns k; solution in z.  The variable has al->nv entries: x (alps) or [x; s] (als). */
static long long panocplus(tiny_al* al, const double* x0, const tiny_opts* o, double tol,
                           long long maxit, double* z) {
    const double a = o->alpha;
    const int n = al->nv;
    double x[TNV], grad_x[TNV], grad_z[TNV], tmpg[TNV], xe[TNV], df[TNV];
    double y[TNV], res[TNV], x_prev[TNV], res_prev[TNV], x_d[TNV], z_curr[TNV], d[TNV], nres[TNV];
    double sv[TNV], yv[TNV];
    for (int j = 0; j < n; ++j) x[j] = x0[j];
    double f_x = al_gradient(al, x, grad_x);
    for (int j = 0; j < n; ++j) xe[j] = x[j] + 1.0;
    al_gradient(al, xe, tmpg);
    for (int j = 0; j < n; ++j) df[j] = tmpg[j] - grad_x[j];
    double Lest = sqrt(dotn(df, df, n)) / sqrt((double)n);
    double gamma = a / Lest;
    for (int j = 0; j < n; ++j) y[j] = x[j] - gamma * grad_x[j];
    double g_z = al_prox(al, y, gamma, z);
    for (int j = 0; j < n; ++j) res[j] = x[j] - z[j];
    tiny_lbfgs H;
    memset(&H, 0, sizeof H);
    H.M = o->lbfgs_memory; H.H = 1.0; H.nv = n;
    double f_up = f_model(f_x, grad_x, res, a / gamma, n);
    double f_z = al_gradient(al, z, grad_z);
    while (f_z > f_up + 10 * T_EPS * (1 + fabs(f_z)) && gamma >= o->minimum_gamma) {
        gamma /= 2;
        for (int j = 0; j < n; ++j) y[j] = x[j] - gamma * grad_x[j];
        g_z = al_prox(al, y, gamma, z);
        for (int j = 0; j < n; ++j) res[j] = x[j] - z[j];
        f_up = f_model(f_x, grad_x, res, a / gamma, n);
        f_z = al_gradient(al, z, grad_z);
    }
    double tau = 1.0;
    long long k = 1;
    for (;;) {
        double stopnorm = 0.0;
        int isnan_ = 0;
        for (int j = 0; j < n; ++j) {
            double v = (res[j] / gamma - grad_x[j]) + grad_z[j];
            if (v != v) isnan_ = 1; else stopnorm = fmax(stopnorm, fabs(v));
        }
        if (isnan_) stopnorm = NAN;
        if (k >= maxit || stopnorm <= tol) return k;
        for (int j = 0; j < n; ++j) { x_prev[j] = x[j]; res_prev[j] = res[j]; }
        double FBE_x = f_model(f_x, grad_x, res, a / gamma, n) + g_z;
        for (int j = 0; j < n; ++j) nres[j] = -res[j];
        lbfgs_mul(&H, nres, d);
        for (int j = 0; j < n; ++j) x_d[j] = x[j] + d[j];
        f_x = al_gradient(al, x_d, grad_x);
        for (int j = 0; j < n; ++j) { x[j] = x_d[j]; z_curr[j] = z[j]; }
        tau = 1.0;
        double sigma = o->beta * (0.5 / gamma) * (1 - a);
        double nr = sqrt(dotn(res, res, n));
        double thresh = FBE_x - sigma * (nr * nr) + 10 * T_EPS * (1 + fabs(FBE_x));
        for (int kb = 1; kb <= o->max_backtracks; ++kb) {
            for (int j = 0; j < n; ++j) y[j] = x[j] - gamma * grad_x[j];
            g_z = al_prox(al, y, gamma, z);
            for (int j = 0; j < n; ++j) res[j] = x[j] - z[j];
            f_up = f_model(f_x, grad_x, res, a / gamma, n);
            f_z = al_gradient(al, z, grad_z);
            if (f_z > f_up + 10 * T_EPS * (1 + fabs(f_z)) && gamma >= o->minimum_gamma) {
                gamma *= 0.5;
                H.currmem = 0; H.curridx = 0; H.H = 1.0;
                continue;
            }
            if (f_up + g_z <= thresh || kb >= o->max_backtracks) break;
            tau = (kb >= o->max_backtracks - 1) ? 0.0 : tau / 2;
            for (int j = 0; j < n; ++j) x[j] = tau * x_d[j] + (1 - tau) * z_curr[j];
            f_x = al_gradient(al, x, grad_x);
        }
        for (int j = 0; j < n; ++j) { sv[j] = x[j] - x_prev[j]; yv[j] = res[j] - res_prev[j]; }
        lbfgs_update(&H, sv, yv);
        k += 1;
    }
}

static double jl_max(double a, double b) { return (a != a || b != b) ? NAN : fmax(a, b); }

/* one alps solve (alps.jl:7-117); status: 0 first_order, 1 max_iter, 2 exception, 3 unknown */
int tiny_alps(int kind, const double* x0, const double* y0, const tiny_opts* o, double* x_out,
              double* y_out, int* tot_it_out, long long* tot_inner_out, int* status_out,
              double* nrp_out, long long* ngrad_out) {
    int nx, ny;
    kind_dims(kind, &nx, &ny);
    double tol_prim = o->tol_prim >= 0 ? o->tol_prim : o->tol;
    double tol_dual = o->tol_dual >= 0 ? o->tol_dual : o->tol;
    double inner_tol = o->inner_tol >= 0 ? o->inner_tol : cbrt(tol_dual);
    tiny_al al;
    memset(&al, 0, sizeof al);
    al.kind = kind; al.nx = nx; al.ny = ny; al.slack = 0; al.nv = nx;
    double x[TNX], cx[TNY], s[TNY], z[TNX];
    al_prox(&al, x0, T_EPS, x);
    double objx = f_value(kind, x) + al.gz;
    c_eval(kind, x, cx);
    projD(kind, cx, s);
    for (int i = 0; i < ny; ++i) { /* safeguards.jl:13-18 */
        double df = cx[i] - s[i];
        double m = jl_max(1.0, 0.5 * (df * df)) / jl_max(1.0, objx);
        m *= 0.1;
        al.mu[i] = jl_max(1e-8, (m != m) ? NAN : fmin(m, 1e8));
        al.y[i] = y0[i];
    }
    int tot_it = 0, solved = 0, tired = tot_it >= o->maxit, broken = objx != objx, have_old = 0;
    long long tot_inner = 0;
    double nrp = NAN, nrp_old = NAN;
    while (!(solved || tired || broken)) {
        tot_it++;
        for (int i = 0; i < ny; ++i) al.y[i] = fmax(-1e20, fmin(al.y[i], 1e20)); /* safeguards.jl:2-5 */
        al_update(&al);
        long long sub_it = panocplus(&al, x, o, inner_tol, o->sub_maxit, z);
        for (int j = 0; j < nx; ++j) x[j] = z[j];
        objx = al.fx + al.gz;
        tot_inner += sub_it;
        int sub_solved = sub_it < o->subsolver_maxit;
        c_eval(kind, x, cx);
        double t[TNY];
        for (int i = 0; i < ny; ++i) t[i] = cx[i] + al.muy[i];
        projD(kind, t, s);
        for (int i = 0; i < ny; ++i) al.y[i] = (t[i] - s[i]) / al.mu[i];
        nrp_old = nrp;
        nrp = 0.0;
        for (int i = 0; i < ny; ++i) {
            double v = fabs(cx[i] - s[i]);
            nrp = (v != v || nrp != nrp) ? NAN : fmax(nrp, v);
        }
        solved = (inner_tol <= tol_dual && sub_solved) && (nrp <= tol_prim);
        tired = tot_it >= o->maxit;
        broken = objx != objx;
        if (!(solved || tired || broken)) {
            if (have_old && nrp > jl_max(o->theta_penalty * nrp_old, tol_prim))
                for (int i = 0; i < ny; ++i) al.mu[i] *= o->kappa_penalty;
            inner_tol = jl_max(o->kappa_tol * inner_tol, tol_dual);
        }
        have_old = 1;
    }
    for (int j = 0; j < nx; ++j) x_out[j] = x[j];
    for (int i = 0; i < ny; ++i) y_out[i] = al.y[i];
    *tot_it_out = tot_it;
    *tot_inner_out = tot_inner;
    *status_out = solved ? 0 : tired ? 1 : broken ? 2 : 3;
    *nrp_out = nrp;
    if (ngrad_out) *ngrad_out = al.n_grad;
    return 0;
}

/* the demos' `for i = 1:ntests` loop; OpenMP over instances when built with -fopenmp */
int tiny_alps_batch(int kind, long long batch, const double* x0, const double* y0,
                    const tiny_opts* o, double* x, double* y, int* tot_it, long long* tot_inner,
                    int* status, double* nrp) {
    int nx, ny;
    kind_dims(kind, &nx, &ny);
#pragma omp parallel for schedule(dynamic, 64)
    for (long long i = 0; i < batch; ++i)
        tiny_alps(kind, x0 + i * nx, y0 + i * ny, o, x + i * nx, y + i * ny, tot_it + i,
                  tot_inner + i, status + i, nrp + i, 0);
    return 0;
}

/* one als solve (src/algorithms/als.jl:7-120): PANOCplus on [x; s]; same outputs and status codes as tiny_alps */
int tiny_als(int kind, const double* x0, const double* y0, const tiny_opts* o, double* x_out,
             double* y_out, int* tot_it_out, long long* tot_inner_out, int* status_out,
             double* nrp_out, long long* ngrad_out) {
    int nx, ny;
    kind_dims(kind, &nx, &ny);
    double tol_prim = o->tol_prim >= 0 ? o->tol_prim : o->tol;
    double tol_dual = o->tol_dual >= 0 ? o->tol_dual : o->tol;
    double inner_tol = o->inner_tol >= 0 ? o->inner_tol : cbrt(tol_dual);
    tiny_al al;
    memset(&al, 0, sizeof al);
    al.kind = kind; al.nx = nx; al.ny = ny; al.slack = 0; al.nv = nx;
    double x[TNX], cx[TNY], s[TNY], xs[TNV], zs[TNV];
    al_prox(&al, x0, T_EPS, x);                       /* als.jl:38-39 (plain NonsmoothCostFun) */
    double objx = f_value(kind, x) + al.gz;
    c_eval(kind, x, cx);
    projD(kind, cx, s);
    for (int i = 0; i < ny; ++i) { /* safeguards.jl:13-18 */
        double df = cx[i] - s[i];
        double m = jl_max(1.0, 0.5 * (df * df)) / jl_max(1.0, objx);
        m *= 0.1;
        al.mu[i] = jl_max(1e-8, (m != m) ? NAN : fmin(m, 1e8));
        al.y[i] = y0[i];
    }
    al.slack = 1; al.nv = nx + ny;                   /* fSlack / gSlack from here on (als.jl:46-47) */
    int tot_it = 0, solved = 0, tired = tot_it >= o->maxit, broken = objx != objx, have_old = 0;
    long long tot_inner = 0;
    double nrp = NAN, nrp_old = NAN;
    while (!(solved || tired || broken)) {
        tot_it++;
        for (int i = 0; i < ny; ++i) al.y[i] = fmax(-1e20, fmin(al.y[i], 1e20)); /* safeguards.jl:2-5 */
        al_update(&al);
        for (int j = 0; j < nx; ++j) xs[j] = x[j];
        for (int i = 0; i < ny; ++i) xs[nx + i] = s[i];
        long long sub_it = panocplus(&al, xs, o, inner_tol, o->sub_maxit, zs);
        for (int j = 0; j < nx; ++j) x[j] = zs[j];
        for (int i = 0; i < ny; ++i) s[i] = zs[nx + i];
        objx = f_value(kind, x) + al.gz;              /* als.jl:78 */
        tot_inner += sub_it;
        int sub_solved = sub_it < o->subsolver_maxit;
        c_eval(kind, x, cx);
        for (int i = 0; i < ny; ++i) al.y[i] = al.y[i] + (cx[i] - s[i]) / al.mu[i];   /* als.jl:83 */
        nrp_old = nrp;
        nrp = 0.0;
        for (int i = 0; i < ny; ++i) {
            double v = fabs(cx[i] - s[i]);
            nrp = (v != v || nrp != nrp) ? NAN : fmax(nrp, v);
        }
        solved = (inner_tol <= tol_dual && sub_solved) && (nrp <= tol_prim);
        tired = tot_it >= o->maxit;
        broken = objx != objx;
        if (!(solved || tired || broken)) {
            if (have_old && nrp > jl_max(o->theta_penalty * nrp_old, tol_prim))
                for (int i = 0; i < ny; ++i) al.mu[i] *= o->kappa_penalty;
            inner_tol = jl_max(o->kappa_tol * inner_tol, tol_dual);
        }
        have_old = 1;
    }
    for (int j = 0; j < nx; ++j) x_out[j] = x[j];
    for (int i = 0; i < ny; ++i) y_out[i] = al.y[i];
    *tot_it_out = tot_it;
    *tot_inner_out = tot_inner;
    *status_out = solved ? 0 : tired ? 1 : broken ? 2 : 3;
    *nrp_out = nrp;
    if (ngrad_out) *ngrad_out = al.n_grad;
    return 0;
}

int tiny_als_batch(int kind, long long batch, const double* x0, const double* y0,
                   const tiny_opts* o, double* x, double* y, int* tot_it, long long* tot_inner,
                   int* status, double* nrp) {
    int nx, ny;
    kind_dims(kind, &nx, &ny);
#pragma omp parallel for schedule(dynamic, 64)
    for (long long i = 0; i < batch; ++i)
        tiny_als(kind, x0 + i * nx, y0 + i * ny, o, x + i * nx, y + i * ny, tot_it + i,
                 tot_inner + i, status + i, nrp + i, 0);
    return 0;
}

/* ---- synthetic data generator (C form of bazinga_oracle.bz_randn_like; Philox-4x32-10, counter =
 * element index, key = (seed, stream)); used by bench.py's CPU legs to build the 4096x16384 instance
 * quickly.  tests/test_oracle_pins.py checks it against the numpy definition. ------------------------- */
static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1, uint32_t* out) {
    uint32_t c2 = 0, c3 = 0;
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
int oracle_randn_fill(uint64_t seed, uint64_t stream, uint64_t start, long long count, double* out) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < count; ++i) {
        uint32_t u[4];
        uint64_t idx = start + (uint64_t)i;
        philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)seed, (uint32_t)stream, u);
        int64_t ssum = (int64_t)((uint64_t)u[0] + u[1] + u[2] + u[3]) - (int64_t)(2ull * 0xFFFFFFFFull);
        out[i] = (double)ssum * 0x1.bb67ae8584caap-32;
    }
    return 0;
}

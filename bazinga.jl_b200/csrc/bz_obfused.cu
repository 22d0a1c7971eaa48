// X kernel: one fused pass for a whole forward-backward evaluation of the obstacle problems (config C4;
// demo/obstacle.jl obstacle_l1 / obstacle_l2, 1-D operator or the 2-D 5-point grid):
//
//     grad_x = gradient!(., alFun, x)            auglagfun.jl:73-86 around demo/obstacle.jl:97-113  (stencil, twice)
//     z = prox_g(x - gamma*grad_x); res = x - z  nonsmoothcostfun.jl:17-22, demo/obstacle.jl:67-81
//     grad_z = gradient!(., alFun, z)            the second AL-gradient of a PANOC+ iteration
//     f(x), f(z), <grad_x,res>, ||res||^2, g(z), ||res/gamma - grad_x + grad_z||_inf, then the decisions of
//     SURVEY Appendix A.4/A.5 by the last CTA
//
// i.e. the kernels [G, F, G, D] of the generic cycle in ONE read of x, mu, mu.*y (5 N-vectors) and ONE write of z and
// res (6 N-vectors): 11 N-vectors instead of 53 (two 13-vector stencil gradients, 15 for F, 12 for D).  Nothing of
// grad_x / grad_z / c(x) / yupd touches HBM -- every value is a pointwise function of a radius-4 neighbourhood of x:
//
//     w_x  = (x1 + A x2 - x3 + mu y)/mu          needs x2 at distance 1
//     grad_x = [f1 + w_x; (x2 - 1) + A w_x; -w_x]   needs w_x at distance 1
//     z    = prox(x - gamma grad_x)              pointwise
//     w_z  = (z1 + A z2 - z3 + mu y)/mu          needs z2 at distance 1
//     grad_z = [f1z + w_z; (z2 - 1) + A w_z; -w_z]  needs w_z at distance 1
//
// Layout of the computation (no shared memory, no block barrier in the row loop): a WARP owns 32 consecutive grid
// columns and walks down a chunk of rows; lane l holds column c0 - 4 + l.  Vertical neighbours are the lane's own
// registers from earlier rows (a software pipeline: while raw row j is consumed, w_x is formed for row j-1, grad_x / z
// for row j-2, w_z for row j-3 and grad_z for row j-4), horizontal neighbours come from the adjacent lanes by shuffle.
// Each of the four stencil applications invalidates one more lane at either end of the warp, so lanes 4..27 own the 24
// output columns and adjacent warps overlap by 8 columns (re-read through L1/L2, not DRAM).  Points outside the grid
// hold exact zeros -- subtracting 0.0 is what "no neighbour" means, bit for bit.  A chunk starts 4 rows early and ends 4
// rows late (warm-up of the pipeline, ~2.5 % of the rows at C4's chunk height).
//
// The arithmetic per entry is the arithmetic of k_obstacle_c / k_obstacle_jt / k_fstep / k_dstep in the same order, so
// z, res, grad_x and grad_z are bit-identical to the unfused kernels (and to the numpy oracle); the reductions are
// order-fixed (lane rows ascending -> warp butterfly -> warps -> CTAs in index order).  Compiled with -fmad=false; the
// divisions by gamma use the three-FMA quotient of bz_fdiv.cuh (the same bits as the IEEE division).
#include "bz_vecdev.cuh"
#include "bz_fdiv.cuh"
#include <cstdlib>

#define OF_THREADS 128
#define OF_WARPS (OF_THREADS / 32)
#define OF_HALO 4
#define OF_OUT (32 - 2 * OF_HALO)   // 24 output columns per warp
#define OF_PF 1  /* (documentation only) */                      // raw rows in flight ahead of the row being consumed (one row-time of a warp is several memory latencies)
#define OF_NSUM 9

struct ObFusedGeom { int ncg, nrc, rows_per_chunk, ctas; };

__device__ __forceinline__ double of_shl(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }     // value of lane - 1
__device__ __forceinline__ double of_shr(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }   // value of lane + 1

// diag*v - left - right - up - down, the operation order of bz_stencil_rc (missing neighbours are exact zeros)
__device__ __forceinline__ double of_stencil(double diag, double c, double left, double right, double up, double dn) {
    double out = diag * c;
    out -= left;
    out -= right;
    out -= up;
    out -= dn;
    return out;
}

// The software pipeline of one lane.  Every delay line is a circular buffer of THREE rows with STATIC indices: the row
// loop is unrolled three times and step P of a group finds the row of age k in slot (P-dependent constant + k) mod 3, so
// nothing is ever shifted (a shifting version spent 29 % of its instructions on register moves; a six-row version did
// not fit the instruction cache).  Raw rows are loaded one row ahead into staging registers.  Registers are the scarce
// resource (a first version spilled its loop bounds and stalled on local-memory loads, ncu r2d): indices are 32-bit,
// the stop norm keeps res/gamma - grad_x (3 values per row) instead of res and grad_x (6), the debug outputs and the
// rarely needed second stop norm are compile-time variants.
template <bool NEED2>
struct OfPipe {
    double X1[3], X2[3], X3[3], MU[3], RM[3], MY[3];   // raw rows jc, jc-1, jc-2 (x, mu, 1/mu, mu*y)
    double S1, S2, S3, SMU, SRM, SMY;                  // staging: row jc+1, in flight
    double WX[3];                                       // w_x of rows jc-1, jc-2, jc-3
    double Z1[3], Z2[3], Z3[3];                         // z of rows jc-2, jc-3, jc-4
    double T1[3], T2[3], T3[3];                         // res/gamma - grad_x of the same rows
    double U1[NEED2 ? 3 : 1], U2[NEED2 ? 3 : 1], U3[NEED2 ? 3 : 1];   // res/(gamma/2) - grad_x (NEED2 only)
    double WZ[3];                                       // w_z of rows jc-3, jc-4, jc-5
};
struct OfCtx {
    const double* __restrict__ x; const double* __restrict__ mu; const double* __restrict__ rmu; const double* __restrict__ muy;
    double* __restrict__ zp; double* __restrict__ rp; double* __restrict__ dbg_gx; double* __restrict__ dbg_gz;
    int N, GX, GY, col, rb, re;
    // slab mode (grid rows split over the ranks, bz_internal.cuh): rows ylo .. yhi-1 exist -- 4 rows beyond the slab on a side
    // with a neighbour, whose raw rows (x, mu, 1/mu, mu*y at the same offsets of ITS slab) are read through the peer mapping
    int ylo, yhi;
    const double* x_up; const double* mu_up; const double* rmu_up; const double* muy_up;
    const double* x_dn; const double* mu_dn; const double* rmu_dn; const double* muy_dn;
    double gamma, gamma_h, rg, rgh, diag;
    bool colok, own_lane;
    double a[OF_NSUM];
    double m1, m2, nan1, nan2;     // running maxima (NaN-blind) and NaN detectors: sums of |v|, NaN iff some v was NaN
};

// fdiv of bz_fdiv.cuh with the range test first: the common quotient costs three FMAs and one integer test
__device__ __forceinline__ double of_div(double a, double b, double rb) {
    const double q = a * rb;
    const double q2 = fma(fma(-b, q, a), rb, q);
    if (fmid<512>(q2)) return q2;
    if (a != a || (a == 0.0 && q == 0.0)) return q;
    return fdiv_rare(a, b, rb);
}

// One row step.  MASKED = false is the steady state of an interior warp: every lane's column is inside the grid, rows
// jc-4 .. jc+1 are grid rows, and all four stage rows belong to this warp's chunk -- no validity selects, no range
// tests; the halo lanes accumulate garbage that is discarded at the end (c.own_lane).  MASKED = true is the general step.
// L2 = obstacle_l2 (f adds 0.5||x1||^2: dfx block 1 = x1; g without the threshold on block 1 and with value 0), else
// obstacle_l1.  NEED2: also the stop norm at gamma/2.  DEBUG: store grad_x / grad_z (parity entry point).
template <int P, bool MASKED, bool L2, bool NEED2, bool DEBUG>
__device__ __forceinline__ void of_step(OfPipe<NEED2>& s, OfCtx& c, const int jc) {
    constexpr bool l2f = L2, l1g = !L2;
    constexpr int n0 = (3 - P % 3) % 3;
#define SL(age) ((n0 + (age)) % 3)
    const int base = jc * c.GX + c.col;                // element index of (row jc, this lane's column)
    // ---- the row evicted from the raw lines is row jc-3: stage 3 wants its mu, 1/mu, mu*y
    const double ev_mu = s.MU[SL(0)], ev_rm = s.RM[SL(0)], ev_my = s.MY[SL(0)];
    s.X1[SL(0)] = s.S1; s.X2[SL(0)] = s.S2; s.X3[SL(0)] = s.S3;            // row jc arrives from staging
    s.MU[SL(0)] = s.SMU; s.RM[SL(0)] = s.SRM; s.MY[SL(0)] = s.SMY;
    // ---- all horizontal neighbours of this step at once: the four stencil centres (x2 of row jc-1, w_x of row jc-2, z2 of
    // row jc-3, w_z of row jc-4) were produced in EARLIER steps, so their eight shuffles are issued back to back here
    const double cx2 = s.X2[SL(1)], cwx = s.WX[SL(1)], cz2 = s.Z2[SL(1)], cwz = s.WZ[SL(1)];
    const double lx2 = of_shl(cx2), rx2 = of_shr(cx2), lwx = of_shl(cwx), rwx = of_shr(cwx);
    const double lz2 = of_shl(cz2), rz2 = of_shr(cz2), lwz = of_shl(cwz), rwz = of_shr(cwz);
    {   // ---- raw row jc + 1 into staging (zeros outside the grid)
        const bool ok = MASKED ? (c.colok && jc + 1 >= c.ylo && jc + 1 < c.yhi) : true;
        const int i = base + c.GX;
        if (MASKED && ok && (jc + 1 < 0 || jc + 1 >= c.GY)) {
            // a row of the neighbouring rank's slab (after the cross-rank barrier at the top of the kernel); uncached loads
            const bool up = jc + 1 < 0;
            const int ih = (up ? jc + 1 + c.GY : jc + 1 - c.GY) * c.GX + c.col;
            const double* hx = up ? c.x_up : c.x_dn;
            s.S1 = __ldcv(hx + ih);
            s.S2 = __ldcv(hx + c.N + ih);
            s.S3 = __ldcv(hx + 2 * c.N + ih);
            s.SMU = __ldcv((up ? c.mu_up : c.mu_dn) + ih);
            s.SRM = __ldcv((up ? c.rmu_up : c.rmu_dn) + ih);
            s.SMY = __ldcv((up ? c.muy_up : c.muy_dn) + ih);
        } else {
        s.S1 = ok ? __ldg(c.x + i) : 0.0;
        s.S2 = ok ? __ldg(c.x + c.N + i) : 0.0;
        s.S3 = ok ? __ldg(c.x + 2 * c.N + i) : 0.0;
        s.SMU = ok ? __ldg(c.mu + i) : 1.0;
        s.SRM = ok ? __ldg(c.rmu + i) : 1.0;      // 1 ./ mu, kept up to date by k_al_update (of_div: the bits of a / mu)
        s.SMY = ok ? __ldg(c.muy + i) : 0.0;
        }
    }
    {   // ---- stage 1: w_x on row r1 = jc - 1 (x2 rows jc-2, jc-1, jc)                      k_obstacle_c
        const int r1 = jc - 1;
        const double x1 = s.X1[SL(1)], x2 = s.X2[SL(1)], x3 = s.X3[SL(1)];
        const double mua = s.MU[SL(1)], rma = s.RM[SL(1)], muya = s.MY[SL(1)];
        const double cva = (x1 + of_stencil(c.diag, x2, lx2, rx2, s.X2[SL(2)], s.X2[SL(0)])) - x3;
        const double t = cva + muya;
        const double u = t - 0.0;                      // SetObstacle: proj = 0 (demo/obstacle.jl:116-119)
        if (MASKED) {
            const bool ok = c.colok && r1 >= c.ylo && r1 < c.yhi;
            s.WX[SL(0)] = ok ? of_div(u, mua, rma) : 0.0;
            if (c.own_lane && r1 >= c.rb && r1 < c.re) {
                c.a[0] += of_div(u * u, mua, rma);
                if (l2f) c.a[1] += x1 * x1;
                c.a[1] += x2 * x2;
                c.a[2] += x2;
            }
        } else {
            s.WX[SL(0)] = of_div(u, mua, rma);
            c.a[0] += of_div(u * u, mua, rma);
            if (l2f) c.a[1] += x1 * x1;
            c.a[1] += x2 * x2;
            c.a[2] += x2;
        }
    }
    {   // ---- stage 2: grad_x, z, res on row r2 = jc - 2 (w_x rows jc-3, jc-2, jc-1)          k_obstacle_jt + k_fstep
        const int r2 = jc - 2;
        const double x1 = s.X1[SL(2)], x2 = s.X2[SL(2)], x3 = s.X3[SL(2)];
        const double w = s.WX[SL(1)];
        const double aw = of_stencil(c.diag, w, lwx, rwx, s.WX[SL(2)], s.WX[SL(0)]);
        const double g1 = (l2f ? x1 : 0.0) + w;
        const double g2 = (x2 - 1.0) + aw;
        const double g3 = 0.0 + (-w);
        double z1 = l1g ? fmax(0.0, (x1 - c.gamma * g1) - c.gamma) : fmax(0.0, x1 - c.gamma * g1);
        double z2 = fmax(0.0, x2 - c.gamma * g2);
        double z3 = fmax(0.0, x3 - c.gamma * g3);
        if (z2 > z3) z3 = 0.0; else z2 = 0.0;
        bool st_ok = c.own_lane;
        if (MASKED) {
            const bool ok = c.colok && r2 >= c.ylo && r2 < c.yhi;
            if (!ok) { z1 = 0.0; z2 = 0.0; z3 = 0.0; }
            st_ok = c.own_lane && r2 >= c.rb && r2 < c.re;
        }
        const double q1 = x1 - z1, q2 = x2 - z2, q3 = x3 - z3;
        s.Z1[SL(0)] = z1; s.Z2[SL(0)] = z2; s.Z3[SL(0)] = z3;
        // what the stop norm needs of this row: res/gamma - grad_x  (k_dstep: (r/gamma - gx) + gz)
        s.T1[SL(0)] = of_div(q1, c.gamma, c.rg) - g1;
        s.T2[SL(0)] = of_div(q2, c.gamma, c.rg) - g2;
        s.T3[SL(0)] = of_div(q3, c.gamma, c.rg) - g3;
        if (NEED2) {
            s.U1[SL(0) % (NEED2 ? 3 : 1)] = of_div(q1, c.gamma_h, c.rgh) - g1;
            s.U2[SL(0) % (NEED2 ? 3 : 1)] = of_div(q2, c.gamma_h, c.rgh) - g2;
            s.U3[SL(0) % (NEED2 ? 3 : 1)] = of_div(q3, c.gamma_h, c.rgh) - g3;
        }
        if (st_ok) {
            const int i = base - 2 * c.GX;
            c.zp[i] = z1; c.zp[c.N + i] = z2; c.zp[2 * c.N + i] = z3;
            c.rp[i] = q1; c.rp[c.N + i] = q2; c.rp[2 * c.N + i] = q3;
            if (DEBUG) { c.dbg_gx[i] = g1; c.dbg_gx[c.N + i] = g2; c.dbg_gx[2 * c.N + i] = g3; }
        }
        if (!MASKED || st_ok) {
            if (l1g) c.a[5] += z1;
            c.a[3] += g1 * q1; c.a[4] += q1 * q1;
            c.a[3] += g2 * q2; c.a[4] += q2 * q2;
            c.a[3] += g3 * q3; c.a[4] += q3 * q3;
        }
    }
    {   // ---- stage 3: w_z on row r3 = jc - 3 (z2 rows jc-4, jc-3, jc-2)
        const int r3 = jc - 3;
        const double z1 = s.Z1[SL(1)], z2 = s.Z2[SL(1)], z3 = s.Z3[SL(1)];
        const double cva = (z1 + of_stencil(c.diag, z2, lz2, rz2, s.Z2[SL(2)], s.Z2[SL(0)])) - z3;
        const double t = cva + ev_my;
        const double u = t - 0.0;
        if (MASKED) {
            const bool ok = c.colok && r3 >= c.ylo && r3 < c.yhi;
            s.WZ[SL(0)] = ok ? of_div(u, ev_mu, ev_rm) : 0.0;
            if (c.own_lane && r3 >= c.rb && r3 < c.re) {
                c.a[6] += of_div(u * u, ev_mu, ev_rm);
                if (l2f) c.a[7] += z1 * z1;
                c.a[7] += z2 * z2;
                c.a[8] += z2;
            }
        } else {
            s.WZ[SL(0)] = of_div(u, ev_mu, ev_rm);
            c.a[6] += of_div(u * u, ev_mu, ev_rm);
            if (l2f) c.a[7] += z1 * z1;
            c.a[7] += z2 * z2;
            c.a[8] += z2;
        }
    }
    {   // ---- stage 4: grad_z and the stop norm on row r4 = jc - 4 (w_z rows jc-5, jc-4, jc-3)   k_dstep
        const int r4 = jc - 4;
        const double w = s.WZ[SL(1)];
        const double aw = of_stencil(c.diag, w, lwz, rwz, s.WZ[SL(2)], s.WZ[SL(0)]);
        if (!MASKED || (c.own_lane && r4 >= c.rb && r4 < c.re)) {
            const double h1 = (l2f ? s.Z1[SL(2)] : 0.0) + w;
            const double h2 = (s.Z2[SL(2)] - 1.0) + aw;
            const double h3 = 0.0 + (-w);
            if (DEBUG) { if (c.own_lane) { const int i = base - 4 * c.GX; c.dbg_gz[i] = h1; c.dbg_gz[c.N + i] = h2; c.dbg_gz[2 * c.N + i] = h3; } }
            // NaN-propagating max (Julia's norm(., Inf)): the running maximum ignores NaN; a sum of |v| (never Inf - Inf)
            // is NaN exactly when some v was
            const double v1 = fabs(s.T1[SL(2)] + h1), v2 = fabs(s.T2[SL(2)] + h2), v3 = fabs(s.T3[SL(2)] + h3);
            c.m1 = fmax(c.m1, fmax(fmax(v1, v2), v3));
            c.nan1 += (v1 + v2) + v3;
            if (NEED2) {   // the norm at gamma/2 is only read when this back-track is the last one allowed (bz_d_decide)
                const double u1 = fabs(s.U1[SL(2) % (NEED2 ? 3 : 1)] + h1), u2 = fabs(s.U2[SL(2) % (NEED2 ? 3 : 1)] + h2),
                             u3 = fabs(s.U3[SL(2) % (NEED2 ? 3 : 1)] + h3);
                c.m2 = fmax(c.m2, fmax(fmax(u1, u2), u3));
                c.nan2 += (u1 + u2) + u3;
            }
        }
    }
#undef SL
}

template <int P, bool L2, bool NEED2, bool DEBUG>
__device__ __forceinline__ void of_step_any(OfPipe<NEED2>& s, OfCtx& c, int jc, int s_lo, int s_hi) {
    if (jc >= s_lo && jc < s_hi) of_step<P, false, L2, NEED2, DEBUG>(s, c, jc);
    else of_step<P, true, L2, NEED2, DEBUG>(s, c, jc);
}

// force: 0 = member of the cycle (predicated on st->next_op == OP_X, runs the decisions), 1 = parity entry point
// (evaluates at v.x with st->gamma, leaves the nine scalars in st->out, takes no decision)
template <bool L2, bool NEED2, bool DEBUG>
__device__ __forceinline__ void of_kernel_body(BzState* st, const BzVecs& v, const BzObstacle& ob, const BzGParams& g,
                                               const double* __restrict__ mu, const double* __restrict__ rmu,
                                               const double* __restrict__ muy, double* part, unsigned int* counter,
                                               const ObFusedGeom& geo, int force, double* __restrict__ dbg_gx,
                                               double* __restrict__ dbg_gz, double* scratch) {
    // buffer rotation (bz_lbfgs.cu): the cooperative L kernel swaps the roles of (x, x_prev), (res, res_prev), (z, z_curr)
    // instead of copying them; st->rot says which physical buffer plays which role right now
    const int rot = force ? 0 : __ldcg(&st->rot);
    OfCtx c;
    c.gamma = __ldcg(&st->gamma);
    c.gamma_h = c.gamma * 0.5;
    c.rg = frecip(c.gamma); c.rgh = frecip(c.gamma_h);
    c.N = (int)ob.N;
    c.GX = (int)(ob.gx == 0 ? ob.N : ob.gx); c.GY = (int)(ob.gx == 0 ? 1 : ob.gy);
    c.diag = ob.gx == 0 ? 2.0 : 4.0;
    c.x = (rot & 1) ? v.x_prev : v.x;
    c.mu = mu; c.rmu = rmu; c.muy = muy;
    c.zp = (rot & 4) ? v.z_curr : v.z;
    c.rp = (rot & 2) ? v.res_prev : v.res;
    c.dbg_gx = dbg_gx; c.dbg_gz = dbg_gz;
    c.ylo = 0; c.yhi = c.GY;
    c.x_up = c.x_dn = c.mu_up = c.mu_dn = c.rmu_up = c.rmu_dn = c.muy_up = c.muy_dn = nullptr;
    const BzXr& xr = st->xr;
    const bool slab = !force && xr.nranks > 1;
    unsigned long long xseq = 0ULL;
    if (slab) {
        // after this barrier the L kernel (or the outer-step kernels) of EVERY rank is complete: the neighbours' x, mu,
        // 1/mu, mu*y rows may be read.  The reductions at the end are the second rendezvous: no rank leaves this kernel --
        // and overwrites x in its next L step -- before every rank has finished reading.
        xseq = *xr.seq;
        bz_xr_barrier_all(st, xseq);
        const double* base = xr.slab[xr.rank];
        if (xr.rank > 0) {
            const double* nb = xr.slab[xr.rank - 1];
            c.x_up = nb + (c.x - base); c.mu_up = nb + (mu - base); c.rmu_up = nb + (rmu - base); c.muy_up = nb + (muy - base);
            c.ylo = -OF_HALO;
        }
        if (xr.rank + 1 < xr.nranks) {
            const double* nb = xr.slab[xr.rank + 1];
            c.x_dn = nb + (c.x - base); c.mu_dn = nb + (mu - base); c.rmu_dn = nb + (rmu - base); c.muy_dn = nb + (muy - base);
            c.yhi = c.GY + OF_HALO;
        }
    }

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wg = (int)blockIdx.x * OF_WARPS + warp;
    const int cg = wg % geo.ncg, rc = wg / geo.ncg;
    c.col = cg * OF_OUT - OF_HALO + lane;
    c.colok = c.col >= 0 && c.col < c.GX;
    c.own_lane = lane >= OF_HALO && lane < 32 - OF_HALO && c.colok;
    c.rb = rc * geo.rows_per_chunk;
    c.re = c.rb + geo.rows_per_chunk;
    if (c.re > c.GY) c.re = c.GY;
    const bool active = rc < geo.nrc && c.rb < c.GY;
#pragma unroll
    for (int k = 0; k < OF_NSUM; ++k) c.a[k] = 0.0;
    c.m1 = 0.0; c.m2 = 0.0;
    c.nan1 = 0.0; c.nan2 = 0.0;

    if (active) {
        OfPipe<NEED2> s;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            s.X1[k] = 0.0; s.X2[k] = 0.0; s.X3[k] = 0.0; s.MU[k] = 1.0; s.RM[k] = 1.0; s.MY[k] = 0.0;
            s.WX[k] = s.Z1[k] = s.Z2[k] = s.Z3[k] = s.T1[k] = s.T2[k] = s.T3[k] = s.WZ[k] = 0.0;
        }
#pragma unroll
        for (int k = 0; k < (NEED2 ? 3 : 1); ++k) { s.U1[k] = 0.0; s.U2[k] = 0.0; s.U3[k] = 0.0; }
        s.S1 = s.S2 = s.S3 = 0.0; s.SMU = 1.0; s.SRM = 1.0; s.SMY = 0.0;
        // steps jc = rb-5 .. re+3: the first one only loads row rb-4 into staging (its stages see zeros and own nothing)
        const int j_begin = c.rb - 5, j_end = c.re + 4;
        // unmasked range: interior warp, stage rows jc-4 .. jc-1 inside [rb, re), loaded row jc+1 a grid row
        const bool interior = (cg * OF_OUT - OF_HALO >= 0) && (cg * OF_OUT - OF_HALO + 31 < c.GX);
        int s_lo = c.rb + 4, s_hi = c.re + 1;
        if (s_hi > c.GY - 1) s_hi = c.GY - 1;
        if (!interior) s_hi = s_lo;
        for (int jb = j_begin; jb < j_end; jb += 3) {
            of_step_any<0, L2, NEED2, DEBUG>(s, c, jb, s_lo, s_hi);
            if (jb + 1 < j_end) of_step_any<1, L2, NEED2, DEBUG>(s, c, jb + 1, s_lo, s_hi);
            if (jb + 2 < j_end) of_step_any<2, L2, NEED2, DEBUG>(s, c, jb + 2, s_lo, s_hi);
        }
        if (!c.own_lane) {           // halo lanes accumulated garbage in the unmasked steps
#pragma unroll
            for (int k = 0; k < OF_NSUM; ++k) c.a[k] = 0.0;
            c.m1 = 0.0; c.m2 = 0.0;
            c.nan1 = 0.0; c.nan2 = 0.0;
        }
    }
    double a[OF_NSUM];
#pragma unroll
    for (int k = 0; k < OF_NSUM; ++k) a[k] = c.a[k];
    const double m1 = (c.nan1 != c.nan1) ? NANQ : c.m1, m2 = (c.nan2 != c.nan2) ? NANQ : c.m2;
    const double gamma = c.gamma;

    // ---- per-CTA partials (fixed order), then the last CTA sums them and runs the scalar logic
    bz_block_sum<OF_NSUM>(a, scratch);
    double mx[2] = {m1 != m1 ? 0.0 : m1, m2 != m2 ? 0.0 : m2};
    double fl[2] = {m1 != m1 ? 1.0 : 0.0, m2 != m2 ? 1.0 : 0.0};
    bz_block_max<2>(mx, scratch);
    bz_block_max<2>(fl, scratch);
    const int nb = gridDim.x;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < OF_NSUM; ++k) part[(size_t)k * nb + blockIdx.x] = a[k];
        part[(size_t)OF_NSUM * nb + blockIdx.x] = fl[0] > 0.0 ? NANQ : mx[0];
        part[(size_t)(OF_NSUM + 1) * nb + blockIdx.x] = fl[1] > 0.0 ? NANQ : mx[1];
    }
    if (!bz_last_cta(counter, nb)) return;
    double sums[OF_NSUM];
#pragma unroll
    for (int k = 0; k < OF_NSUM; ++k) {
        double t = 0.0;
        for (int i = threadIdx.x; i < nb; i += blockDim.x) t += __ldcg(part + (size_t)k * nb + i);
        sums[k] = t;
    }
    bz_block_sum<OF_NSUM>(sums, scratch);
    const double stop1 = bz_final_nanmax(part + (size_t)OF_NSUM * nb, nb, scratch);
    const double stop2 = bz_final_nanmax(part + (size_t)(OF_NSUM + 1) * nb, nb, scratch);
    if (threadIdx.x != 0) return;
    double stop1x = stop1, stop2x = stop2;
    if (slab) {      // sums and stop norms over the ranks in ONE exchange, added / compared in rank order: the same bits everywhere
        *xr.seq = xseq + 1ULL;                 // (the barrier at the top was exchange xseq)
        double all[OF_NSUM + 2];
#pragma unroll
        for (int k = 0; k < OF_NSUM; ++k) all[k] = sums[k];
        all[OF_NSUM] = stop1; all[OF_NSUM + 1] = stop2;
        // (no system-scope fence: z and res, all this kernel wrote, are read by no peer before its next fenced barrier)
        bz_xr_allreduce(st, all, OF_NSUM + 2, (1u << OF_NSUM) | (1u << (OF_NSUM + 1)), false);   // the last two: NaN-propagating max
#pragma unroll
        for (int k = 0; k < OF_NSUM; ++k) sums[k] = all[k];
        stop1x = all[OF_NSUM]; stop2x = all[OF_NSUM + 1];
    }
    // AL values (auglagfun.jl:78-82): lx = 0.5*sum((t-s)^2/mu) + f - 0.5*sum(mu y^2); f = 0.5*fa - fb (demo/obstacle.jl:24-26,40-42)
    const double musqy = st->musqy;
    double fx_x = 0.5 * sums[1]; fx_x = fx_x + (-1.0) * sums[2];
    double lx_x = 0.5 * sums[0]; lx_x += fx_x; lx_x -= musqy;
    double fx_z = 0.5 * sums[7]; fx_z = fx_z + (-1.0) * sums[8];
    double lx_z = 0.5 * sums[6]; lx_z += fx_z; lx_z -= musqy;
    if (force) {
        st->out[0] = lx_x; st->out[1] = fx_x; st->out[2] = sums[3]; st->out[3] = sums[4];
        st->out[4] = bz_g_value_scale(g) * sums[5]; st->out[5] = lx_z; st->out[6] = fx_z; st->out[7] = stop1; st->out[8] = stop2;
        __threadfence();
        return;
    }
    const int fresh = st->grad_x_fresh;
    bz_f_finish(st, g, gamma, fresh, sums[3], sums[4], sums[5], lx_x, fx_x);
    bz_d_decide(st, gamma, lx_z, fx_z, stop1x, stop2x);
    bz_toc(st, 2);
    __threadfence();
}

// MINB resident CTAs per SM: 3 (168 registers) or 2 (255 registers, nothing spilled); chosen at launch (BZ_OF_OCC)
template <bool L2, int MINB>
__global__ void __launch_bounds__(OF_THREADS, MINB)
    k_obstacle_fused(BzState* st, BzVecs v, BzObstacle ob, BzGParams g, const double* __restrict__ mu,
                     const double* __restrict__ rmu, const double* __restrict__ muy, double* part, unsigned int* counter,
                     ObFusedGeom geo, int force, double* __restrict__ dbg_gx, double* __restrict__ dbg_gz) {
    __shared__ double scratch[OF_NSUM * 32];
    bz_pdl_wait();
    bz_pdl_launch_dependents();
    if (force) {      // parity entry point: both stop norms, grad_x / grad_z written out
        of_kernel_body<L2, true, true>(st, v, ob, g, mu, rmu, muy, part, counter, geo, 1, dbg_gx, dbg_gz, scratch);
        return;
    }
    if (__ldcg(&st->next_op) != OP_X) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) bz_tic(st, 2);
    // stop norm at gamma/2: read by bz_d_decide only when a gamma-halving at this back-track index runs off the loop
    const bool need2 = __ldcg(&st->mode) == MODE_ITER && __ldcg(&st->k_bt) + 1 > __ldcg(&st->max_backtracks);
    if (need2) of_kernel_body<L2, true, false>(st, v, ob, g, mu, rmu, muy, part, counter, geo, 0, nullptr, nullptr, scratch);
    else of_kernel_body<L2, false, false>(st, v, ob, g, mu, rmu, muy, part, counter, geo, 0, nullptr, nullptr, scratch);
}

// chunks of rows x groups of 24 columns -> warps; the grid is sized to ONE wave of resident CTAs when the problem is big
static ObFusedGeom of_geometry(const bz_ctx* ctx, const BzObstacle& ob, int resident_per_sm) {
    const long long GX = ob.gx == 0 ? ob.N : ob.gx, GY = ob.gx == 0 ? 1 : ob.gy;
    ObFusedGeom g;
    const long long ncg = (GX + OF_OUT - 1) / OF_OUT;
    const long long max_warps = (long long)ctx->num_sms * resident_per_sm * OF_WARPS;
    long long nrc = max_warps / ncg;               // row chunks that fit one wave
    if (nrc < 1) nrc = 1;
    long long rpc = (GY + nrc - 1) / nrc;
    if (rpc < 64 && GY > 64) rpc = 64;             // keep the 8 warm-up rows a small fraction (more than one wave then)
    if (rpc < 1) rpc = 1;
    nrc = (GY + rpc - 1) / rpc;
    g.ncg = (int)ncg; g.nrc = (int)nrc; g.rows_per_chunk = (int)rpc;
    g.ctas = (int)((ncg * nrc + OF_WARPS - 1) / OF_WARPS);
    return g;
}

static int of_occ() {
    static int occ = 0;
    if (occ == 0) { const char* e = getenv("BZ_OF_OCC"); occ = (e && atoi(e) == 2) ? 2 : 3; }
    return occ;
}
static int of_resident(bz_ctx* ctx) {
    static int per_sm[64] = {0};
    int& r = per_sm[ctx->device & 63];
    if (r == 0) {
        cudaFuncSetAttribute(k_obstacle_fused<false, 3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_obstacle_fused<true, 3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_obstacle_fused<false, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_obstacle_fused<true, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaError_t e = of_occ() == 2 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r, k_obstacle_fused<false, 2>, OF_THREADS, 0)
                                      : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r, k_obstacle_fused<false, 3>, OF_THREADS, 0);
        if (e != cudaSuccess || r < 1) r = 1;
    }
    return r;
}

int bz_obfused_partials(bz_ctx* ctx, const BzObstacle& ob) {
    return of_geometry(ctx, ob, of_resident(ctx)).ctas * (OF_NSUM + 2);
}

int bz_launch_obstacle_fused(bz_ctx* ctx, BzState* st, const BzVecs& v, const BzObstacle& ob, const BzGParams& g,
                             const double* mu, const double* rmu, const double* muy, double* part, unsigned int* counter,
                             int force, double* dbg_gx, double* dbg_gz) {
    const ObFusedGeom geo = of_geometry(ctx, ob, of_resident(ctx));
    // obstacle_l1 = (SmoothCostObstacleL1, NonsmoothCostObstacleL1), obstacle_l2 = (..L2, ..L2): the catalogue check of
    // bz_problem_create admits no other pairing
#define OF_LAUNCH(L2V, OCC) bz_launch_pdl(ctx, k_obstacle_fused<L2V, OCC>, dim3(geo.ctas), dim3(OF_THREADS), 0, 0, st, v, ob, g, mu, rmu, \
                                          muy, part, counter, geo, force, dbg_gx, dbg_gz)
    if (of_occ() == 2) return ob.l2 ? OF_LAUNCH(true, 2) : OF_LAUNCH(false, 2);
    return ob.l2 ? OF_LAUNCH(true, 3) : OF_LAUNCH(false, 3);
#undef OF_LAUNCH
}

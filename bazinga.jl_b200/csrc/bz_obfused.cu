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

#define OF_THREADS 128
#define OF_WARPS (OF_THREADS / 32)
#define OF_HALO 4
#define OF_OUT (32 - 2 * OF_HALO)   // 24 output columns per warp
#define OF_PF 2                      // raw rows in flight ahead of the row being consumed
#define OF_NSUM 9

struct ObFusedGeom { int ncg, nrc, rows_per_chunk, ctas; };

__device__ __forceinline__ double of_shl(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }     // value of lane - 1
__device__ __forceinline__ double of_shr(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }   // value of lane + 1

// diag*v - left - right - up - down, the operation order of bz_stencil_rc (missing neighbours are exact zeros)
__device__ __forceinline__ double of_stencil(double diag, double c, double up, double dn) {
    double out = diag * c;
    out -= of_shl(c);
    out -= of_shr(c);
    out -= up;
    out -= dn;
    return out;
}

template <int Q>
__device__ __forceinline__ void of_shift(double (&q)[Q]) {
#pragma unroll
    for (int k = Q - 1; k > 0; --k) q[k] = q[k - 1];
}

// force: 0 = member of the cycle (predicated on st->next_op == OP_X, runs the decisions), 1 = parity entry point
// (evaluates at v.x with st->gamma, leaves the eleven scalars in st->out, takes no decision)
__global__ void __launch_bounds__(OF_THREADS, 3)
    k_obstacle_fused(BzState* st, BzVecs v, BzObstacle ob, BzGParams g, const double* __restrict__ mu,
                     const double* __restrict__ muy, double* part, unsigned int* counter, ObFusedGeom geo, int force,
                     double* __restrict__ dbg_gx, double* __restrict__ dbg_gz) {
    __shared__ double scratch[OF_NSUM * 32];
    bz_pdl_wait();
    bz_pdl_launch_dependents();
    if (!force) {
        if (__ldcg(&st->next_op) != OP_X) return;
        if (blockIdx.x == 0 && threadIdx.x == 0) bz_tic(st, 2);
    }
    const double gamma = __ldcg(&st->gamma);
    const double gamma_h = gamma * 0.5;
    const double rg = frecip(gamma), rgh = frecip(gamma_h);
    const long long N = ob.N;
    const long long GX = ob.gx == 0 ? N : ob.gx, GY = ob.gx == 0 ? 1 : ob.gy;
    const double diag = ob.gx == 0 ? 2.0 : 4.0;
    const bool l2f = ob.l2 != 0;                       // SmoothCostObstacleL2: f adds 0.5||x1||^2, dfx block 1 = x1
    const bool l1g = g.kind == BZ_G_OBSTACLE_L1;       // NonsmoothCostObstacleL1: threshold gamma on block 1, value sum(z1)
    const double* __restrict__ x1p = v.x;
    const double* __restrict__ x2p = v.x + N;
    const double* __restrict__ x3p = v.x + 2 * N;
    double* __restrict__ zp = v.z;
    double* __restrict__ rp = v.res;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long wg = (long long)blockIdx.x * OF_WARPS + warp;
    const long long cg = wg % geo.ncg, rc = wg / geo.ncg;
    const long long c = cg * OF_OUT - OF_HALO + lane;
    const bool colok = c >= 0 && c < GX;
    const bool own_lane = lane >= OF_HALO && lane < 32 - OF_HALO && colok;
    const long long rb = rc * geo.rows_per_chunk;
    long long re = rb + geo.rows_per_chunk;
    if (re > GY) re = GY;
    const bool active = rc < geo.nrc && rb < GY;

    // sums: 0 lx(x) 1 fa(x) 2 fb(x) 3 <grad_x,res> 4 ||res||^2 5 g 6 lx(z) 7 fa(z) 8 fb(z); maxima: stop norm at gamma, gamma/2
    double a[OF_NSUM];
#pragma unroll
    for (int k = 0; k < OF_NSUM; ++k) a[k] = 0.0;
    double m1 = 0.0, m2 = 0.0;

    if (active) {
        // delay lines, index 0 = newest.  Raw rows: [0] = row jc + PF (just loaded) ... [PF] = row jc.
        double X1[OF_PF + 3], X2[OF_PF + 3], X3[OF_PF + 3], MU[OF_PF + 4], MY[OF_PF + 4];
        double WX[3], Z1[3], Z2[3], Z3[2], G1[3], G2[3], G3[3], R1[3], R2[3], R3[3], WZ[3];
#pragma unroll
        for (int k = 0; k < OF_PF + 3; ++k) { X1[k] = 0.0; X2[k] = 0.0; X3[k] = 0.0; }
#pragma unroll
        for (int k = 0; k < OF_PF + 4; ++k) { MU[k] = 1.0; MY[k] = 0.0; }
#pragma unroll
        for (int k = 0; k < 3; ++k) { WX[k] = Z1[k] = Z2[k] = G1[k] = G2[k] = G3[k] = R1[k] = R2[k] = R3[k] = WZ[k] = 0.0; }
        Z3[0] = Z3[1] = 0.0;

        auto load_row = [&](long long j) {            // raw row j into slot 0 (zeros outside the grid)
            const bool ok = colok && j >= 0 && j < GY;
            const long long i = j * GX + c;
            X1[0] = ok ? __ldg(x1p + i) : 0.0;
            X2[0] = ok ? __ldg(x2p + i) : 0.0;
            X3[0] = ok ? __ldg(x3p + i) : 0.0;
            MU[0] = ok ? __ldg(mu + i) : 1.0;
            MY[0] = ok ? __ldg(muy + i) : 0.0;
        };
        // prologue: rows rb-4 .. rb-4+PF-1 into the queue
#pragma unroll
        for (int k = 0; k < OF_PF; ++k) {
            of_shift(X1); of_shift(X2); of_shift(X3); of_shift(MU); of_shift(MY);
            load_row(rb - 4 + k);
        }
        for (long long jc = rb - 4; jc < re + 4; ++jc) {
            of_shift(X1); of_shift(X2); of_shift(X3); of_shift(MU); of_shift(MY);
            load_row(jc + OF_PF);
            of_shift(WX); of_shift(Z1); of_shift(Z2); of_shift(G1); of_shift(G2); of_shift(G3);
            of_shift(R1); of_shift(R2); of_shift(R3); of_shift(WZ);
            Z3[1] = Z3[0];

            // ---- stage 1: w_x on row r1 = jc - 1 (x2 rows jc-2, jc-1, jc)            k_obstacle_c
            {
                const long long r1 = jc - 1;
                const bool ok = colok && r1 >= 0 && r1 < GY;
                const double x1 = X1[OF_PF + 1], x2 = X2[OF_PF + 1], x3 = X3[OF_PF + 1];
                const double mua = MU[OF_PF + 1], muya = MY[OF_PF + 1];
                const double cva = (x1 + of_stencil(diag, x2, X2[OF_PF + 2], X2[OF_PF])) - x3;
                const double t = cva + muya;
                const double u = t - 0.0;                      // SetObstacle: proj = 0 (demo/obstacle.jl:116-119)
                WX[0] = ok ? u / mua : 0.0;
                if (own_lane && r1 >= rb && r1 < re) {
                    a[0] += (u * u) / mua;
                    if (l2f) a[1] += x1 * x1;
                    a[1] += x2 * x2;
                    a[2] += x2;
                }
            }
            // ---- stage 2: grad_x, z, res on row r2 = jc - 2 (w_x rows jc-3, jc-2, jc-1)  k_obstacle_jt + k_fstep
            {
                const long long r2 = jc - 2;
                const bool ok = colok && r2 >= 0 && r2 < GY;
                const double x1 = X1[OF_PF + 2], x2 = X2[OF_PF + 2], x3 = X3[OF_PF + 2];
                const double w = WX[1];
                const double aw = of_stencil(diag, w, WX[2], WX[0]);
                const double g1 = (l2f ? x1 : 0.0) + w;
                const double g2 = (x2 - 1.0) + aw;
                const double g3 = 0.0 + (-w);
                double z1 = l1g ? fmax(0.0, (x1 - gamma * g1) - gamma) : fmax(0.0, x1 - gamma * g1);
                double z2 = fmax(0.0, x2 - gamma * g2);
                double z3 = fmax(0.0, x3 - gamma * g3);
                if (z2 > z3) z3 = 0.0; else z2 = 0.0;
                if (!ok) { z1 = 0.0; z2 = 0.0; z3 = 0.0; }
                const double q1 = x1 - z1, q2 = x2 - z2, q3 = x3 - z3;
                Z1[0] = z1; Z2[0] = z2; Z3[0] = z3;
                G1[0] = g1; G2[0] = g2; G3[0] = g3;
                R1[0] = q1; R2[0] = q2; R3[0] = q3;
                if (own_lane && r2 >= rb && r2 < re) {
                    const long long i = r2 * GX + c;
                    zp[i] = z1; zp[N + i] = z2; zp[2 * N + i] = z3;
                    rp[i] = q1; rp[N + i] = q2; rp[2 * N + i] = q3;
                    if (dbg_gx) { dbg_gx[i] = g1; dbg_gx[N + i] = g2; dbg_gx[2 * N + i] = g3; }
                    if (l1g) a[5] += z1;
                    a[3] += g1 * q1; a[4] += q1 * q1;
                    a[3] += g2 * q2; a[4] += q2 * q2;
                    a[3] += g3 * q3; a[4] += q3 * q3;
                }
            }
            // ---- stage 3: w_z on row r3 = jc - 3 (z2 rows jc-4, jc-3, jc-2)
            {
                const long long r3 = jc - 3;
                const bool ok = colok && r3 >= 0 && r3 < GY;
                const double z1 = Z1[1], z2 = Z2[1], z3 = Z3[1];
                const double mua = MU[OF_PF + 3], muya = MY[OF_PF + 3];
                const double cva = (z1 + of_stencil(diag, z2, Z2[2], Z2[0])) - z3;
                const double t = cva + muya;
                const double u = t - 0.0;
                WZ[0] = ok ? u / mua : 0.0;
                if (own_lane && r3 >= rb && r3 < re) {
                    a[6] += (u * u) / mua;
                    if (l2f) a[7] += z1 * z1;
                    a[7] += z2 * z2;
                    a[8] += z2;
                }
            }
            // ---- stage 4: grad_z and the stop norm on row r4 = jc - 4 (w_z rows jc-5, jc-4, jc-3)   k_dstep
            {
                const long long r4 = jc - 4;
                const double w = WZ[1];
                const double aw = of_stencil(diag, w, WZ[2], WZ[0]);
                if (own_lane && r4 >= rb && r4 < re) {
                    const double h1 = (l2f ? Z1[2] : 0.0) + w;
                    const double h2 = (Z2[2] - 1.0) + aw;
                    const double h3 = 0.0 + (-w);
                    if (dbg_gz) { const long long i = r4 * GX + c; dbg_gz[i] = h1; dbg_gz[N + i] = h2; dbg_gz[2 * N + i] = h3; }
                    m1 = bz_nanmax(m1, (fdiv(R1[2], gamma, rg) - G1[2]) + h1);
                    m2 = bz_nanmax(m2, (fdiv(R1[2], gamma_h, rgh) - G1[2]) + h1);
                    m1 = bz_nanmax(m1, (fdiv(R2[2], gamma, rg) - G2[2]) + h2);
                    m2 = bz_nanmax(m2, (fdiv(R2[2], gamma_h, rgh) - G2[2]) + h2);
                    m1 = bz_nanmax(m1, (fdiv(R3[2], gamma, rg) - G3[2]) + h3);
                    m2 = bz_nanmax(m2, (fdiv(R3[2], gamma_h, rgh) - G3[2]) + h3);
                }
            }
        }
    }

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
    {
        const double* ptrs[OF_NSUM];
        int counts[OF_NSUM];
#pragma unroll
        for (int k = 0; k < OF_NSUM; ++k) { ptrs[k] = part + (size_t)k * nb; counts[k] = nb; }
#pragma unroll
        for (int k = 0; k < OF_NSUM; ++k) {
            double s = 0.0;
            for (int i = threadIdx.x; i < counts[k]; i += blockDim.x) s += __ldcg(ptrs[k] + i);
            sums[k] = s;
        }
        bz_block_sum<OF_NSUM>(sums, scratch);
    }
    const double stop1 = bz_final_nanmax(part + (size_t)OF_NSUM * nb, nb, scratch);
    const double stop2 = bz_final_nanmax(part + (size_t)(OF_NSUM + 1) * nb, nb, scratch);
    if (threadIdx.x != 0) return;
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
    bz_d_decide(st, gamma, lx_z, fx_z, stop1, stop2);
    bz_toc(st, 2);
    __threadfence();
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

static int of_resident(bz_ctx* ctx) {
    static int per_sm[64] = {0};
    int& r = per_sm[ctx->device & 63];
    if (r == 0) {
        cudaFuncSetAttribute(k_obstacle_fused, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r, k_obstacle_fused, OF_THREADS, 0) != cudaSuccess || r < 1) r = 1;
    }
    return r;
}

int bz_obfused_partials(bz_ctx* ctx, const BzObstacle& ob) {
    return of_geometry(ctx, ob, of_resident(ctx)).ctas * (OF_NSUM + 2);
}

int bz_launch_obstacle_fused(bz_ctx* ctx, BzState* st, const BzVecs& v, const BzObstacle& ob, const BzGParams& g,
                             const double* mu, const double* muy, double* part, unsigned int* counter, int force,
                             double* dbg_gx, double* dbg_gz) {
    const ObFusedGeom geo = of_geometry(ctx, ob, of_resident(ctx));
    return bz_launch_pdl(ctx, k_obstacle_fused, dim3(geo.ctas), dim3(OF_THREADS), 0, 0, st, v, ob, g, mu, muy, part, counter,
                         geo, force, dbg_gx, dbg_gz);
}

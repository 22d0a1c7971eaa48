// F and D kernels of the PANOC+ state machine (see bz_vec.cu for the cycle [L, G, F, G, D]).
//   F  k_fstep   assemble grad_x, forward-backward step  y = x - gamma*grad; z = prox_g(y); res = x - z,
//                fused with <grad,res>, ||res||^2 and g(z)      (+ the two init modes of Appendix A.4)
//   D  k_dstep   assemble grad_z and f(z); stop norm; gamma-halving / sufficient-decrease decisions
// Both finish with the "last CTA done" pattern: per-CTA partials in fixed slots, the last CTA sums
// them in fixed order and runs the scalar logic -- deterministic, no float atomics.
// When the provider left per-cluster partial rows (dense passes) the assembly is the blocked one of
// bz_vecdev.cuh (64 columns per CTA pass, 4 partial groups); otherwise plain grid-stride.
// Compiled with -fmad=false (the reference's Julia broadcasts are not contracted).
#include "bz_vecdev.cuh"

// The dense passes run with the maximum shared-memory carve-out; asking for the same carve-out here
// keeps the SMs from being reconfigured at every kernel boundary of the cycle.
template <typename Kern>
static void prefer_max_smem(Kern kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

static int fbd_grid(const bz_ctx* ctx, const BzGradSrc& gs, long long items) {
    if (gs.Pf || gs.Pc) {
        long long nb = (items + BZ_ASM_COLS - 1) / BZ_ASM_COLS;
        const long long cap = (long long)ctx->num_sms * 8;
        if (nb < 1) nb = 1;
        return (int)(nb > cap ? cap : nb);
    }
    return bz_vec_grid(ctx, items);
}

// ---------------------------------------------------------------------------------------------
// F
// ---------------------------------------------------------------------------------------------
// MINB = 4 caps the kernel at 64 registers: twice the resident threads for long vectors (C4: 651 -> 366 us); the
// latency-bound short-vector case (C2) is 1 us faster with the registers the compiler wants
template <bool HEAVY, int MINB>
__global__ void __launch_bounds__(BZ_VEC_THREADS, MINB)
    k_fstep(BzState* st, BzVecs v, BzGradSrc gs, BzGParams g, double* part, unsigned int* counter,
            int force_mode) {
    __shared__ double scratch[6 * 32];
    __shared__ double asm_sm[2 * BZ_VEC_THREADS];
    bz_pdl_wait();
    bz_pdl_launch_dependents();
    int mode = force_mode;
    if (mode < 0) {
        if (__ldcg(&st->next_op) != OP_F) return;
        mode = __ldcg(&st->mode);
        if (blockIdx.x == 0 && threadIdx.x == 0) bz_tic(st, 2);
    }
    const long long n = v.n;
    bz_gs_sync_counts(gs, st);
    const bool blocked = bz_gs_blocked(gs);
    const double gamma = __ldcg(&st->gamma);
    const int fresh = (mode >= MODE_INITFB) ? __ldcg(&st->grad_x_fresh) : 0;
    double a[3] = {0.0, 0.0, 0.0};

    // elementwise work for entry j given the assembled gradient value (or the stored grad_x).  The vectors never
    // alias: every load is issued before the first store so that the loads of an entry overlap.
    const double* __restrict__ vx = v.x;
    double* __restrict__ vgx = v.grad_x;
    double* __restrict__ vz = v.z;
    double* __restrict__ vres = v.res;
    // A speculated L step that D confirmed (bz_lspec.cuh, bz_d_decide) is still to be copied into place: the new x waits
    // in the shadow buffer (the dense pass before this kernel read it there), the candidate pair in the spare history
    // slot.  Entry j is moved by the thread that then steps it, before any restrict-qualified access to the same arrays.
    const bool commit = (mode >= MODE_INITFB) && force_mode < 0 && v.xs != nullptr && __ldcg(&st->spec_commit) != 0;
    const int c_pair = commit ? __ldcg(&st->sp_pair) : 0;
    const int c_idx = commit ? __ldcg(&st->curridx) : 0;      // (= sp_curridx: D has put the scalars in place)
    auto commit_elem = [&](long long j) -> double {
        const double xo = v.x[j], ro = v.res[j], zo = v.z[j], xn = __ldcg(v.xs + j);
        double sj = 0.0, yj = 0.0;
        if (c_pair) { sj = __ldcg(v.sM + (size_t)BZ_MAXMEM * n + j); yj = __ldcg(v.yM + (size_t)BZ_MAXMEM * n + j); }
        v.x_prev[j] = xo;
        v.res_prev[j] = ro;
        v.x_d[j] = xn;
        v.x[j] = xn;
        v.z_curr[j] = zo;
        if (c_pair) { v.sM[(size_t)(c_idx > 0 ? c_idx - 1 : 0) * n + j] = sj; v.yM[(size_t)(c_idx > 0 ? c_idx - 1 : 0) * n + j] = yj; }
        asm volatile("" ::: "memory");
        return xn;
    };
    auto elem = [&](long long j, double gr) {
        if (mode == MODE_INIT0) {
            const double xj = vx[j];
            vgx[j] = gr;
            v.xeps[j] = xj + 1.0;                // lower_bound_smoothness_constant: x .+ 1
        } else if (mode == MODE_INIT1) {
            const double df = gr - vgx[j];
            a[0] += df * df;
        } else {
            const double xj = commit ? commit_elem(j) : vx[j];
            if (fresh) vgx[j] = gr;
            const double y = xj - gamma * gr;
            const double z = (g.snx > 0 && j >= g.snx) ? bz_proj_elem(g.sdk, g.slo, g.shi, y, j - g.snx)   // als: proj_D on the slack part
                                                      : bz_prox_elem<HEAVY>(g, y, gamma, j, a[2]);
            const double r = xj - z;
            vz[j] = z;
            vres[j] = r;
            a[0] += gr * r;
            a[1] += r * r;
        }
    };

    const bool need_assembly = (mode == MODE_INIT0) || (mode == MODE_INIT1) || fresh;
    if (mode >= MODE_INITFB && (g.kind == BZ_G_OBSTACLE_L1 || g.kind == BZ_G_OBSTACLE_L2)) {
        // demo/obstacle.jl:67-81: soft threshold block 1 by gamma, max(0,.) on blocks 2-3, then
        // complementarity zeroing between blocks 2 and 3 (tie zeroes the block-2 entry)
        const long long N = g.N;
        const bool l1 = g.kind == BZ_G_OBSTACLE_L1;
        const long long stride = (long long)gridDim.x * blockDim.x;
        const double* __restrict__ Ec = gs.Ec;
        const bool plain = fresh && gs.Ec && !gs.Pf && !gs.Pc && !gs.Ef;   // the stencil provider: grad = Ec
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
            double gr[3], xv[3], zv[3];
#pragma unroll
            for (int b = 0; b < 3; ++b) {        // all six loads first (no store in between: they overlap)
                const long long j = i + b * N;
                gr[b] = plain ? __ldcg(Ec + j) : (fresh ? bz_assemble_grad(gs, j, n) : vgx[j]);
                xv[b] = vx[j];
            }
            zv[0] = l1 ? fmax(0.0, (xv[0] - gamma * gr[0]) - gamma) : fmax(0.0, xv[0] - gamma * gr[0]);   // L2: no threshold
            zv[1] = fmax(0.0, xv[1] - gamma * gr[1]);
            zv[2] = fmax(0.0, xv[2] - gamma * gr[2]);
            if (zv[1] > zv[2]) zv[2] = 0.0; else zv[1] = 0.0;
            if (l1) a[2] += zv[0];               // NonsmoothCostObstacleL2 returns zero (:61)
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                const long long j = i + b * N;
                const double r = xv[b] - zv[b];
                if (fresh) vgx[j] = gr[b];
                vz[j] = zv[b];
                vres[j] = r;
                a[0] += gr[b] * r;
                a[1] += r * r;
            }
        }
    } else if (blocked) {
        for (long long jb = (long long)blockIdx.x * BZ_ASM_COLS; jb < n; jb += (long long)gridDim.x * BZ_ASM_COLS) {
            const double gr = need_assembly ? bz_assemble_blocked(gs, jb, n, asm_sm) : 0.0;
            const long long j = jb + threadIdx.x;
            if (threadIdx.x < BZ_ASM_COLS && j < n) elem(j, need_assembly ? gr : v.grad_x[j]);
        }
    } else {
        const long long stride = (long long)gridDim.x * blockDim.x;
        for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
            elem(j, need_assembly ? bz_assemble_grad(gs, j, n) : v.grad_x[j]);
    }
    bz_block_sum<3>(a, scratch);
    const int nb = gridDim.x;
    if (threadIdx.x == 0) {
        part[blockIdx.x] = a[0];
        part[nb + blockIdx.x] = a[1];
        part[2 * nb + blockIdx.x] = a[2];
    }
    if (!bz_last_cta(counter, nb)) return;

    // ---- last CTA: scalar bookkeeping
    const bool need_value = (mode == MODE_INIT0) || (mode >= MODE_INITFB && fresh);
    const double* const ptrs[6] = {part, part + nb, part + 2 * nb, need_value ? gs.lxpart : nullptr,
                                   need_value ? gs.fxpart : nullptr, need_value ? gs.fxpart2 : nullptr};
    const int counts[6] = {nb, nb, nb, gs.nlx, gs.nfx, gs.nfx2};
    double sums[6];
    bz_final_sums<6>(ptrs, counts, sums, scratch);
    if (threadIdx.x != 0) return;
    bz_xr_allreduce(st, sums, 6, 0u);         // slab mode: sums over ranks
    const double s0 = sums[0], s1 = sums[1], s2 = sums[2];
    double fx = 0.0, lx = 0.0;
    if (need_value) lx = bz_value_from_sums(gs, st->musqy, sums[3], sums[4], sums[5], fx);
    if (mode == MODE_INIT0) {
        st->f_x = lx; st->al_fx = fx; st->al_lx = lx;
        st->n_grad += 1;
        st->grad_target = TGT_XEPS;
        st->after_g = OP_F;
        st->mode = MODE_INIT1;
        st->next_op = OP_G;
    } else if (mode == MODE_INIT1) {
        const double Lest = sqrt(s0) / sqrt((double)(st->n_global > 0 ? st->n_global : n));
        st->Lest = Lest;
        st->gamma = st->alpha / Lest;     // L = 0 -> gamma = Inf, IEEE semantics as upstream
        st->n_grad += 1;
        st->mode = MODE_INITFB;
        st->grad_x_fresh = 0;
        st->next_op = st->fused ? OP_X : OP_F;   // fused problems recompute grad_x inside X
    } else {
        bz_f_finish(st, g, gamma, fresh, s0, s1, s2, lx, fx);
        st->spec_commit = 0;                 // (every CTA read it before it got here)
        st->grad_target = TGT_Z;
        st->after_g = OP_D;
        st->next_op = OP_G;
    }
    if (force_mode < 0) bz_toc(st, 2);
    __threadfence();
}

int bz_launch_fstep(bz_ctx* ctx, BzState* st, const BzVecs& v, const BzGradSrc& gs,
                    const BzGParams& g, double* part, unsigned int* counter, int force_mode) {
    const long long items = (g.kind == BZ_G_OBSTACLE_L1 || g.kind == BZ_G_OBSTACLE_L2) ? g.N : v.n;
    static BzOncePerDevice once;
    if (once.first(ctx->device)) { prefer_max_smem(k_fstep<false, 1>); prefer_max_smem(k_fstep<false, 4>); prefer_max_smem(k_fstep<true, 1>); }
    if (bz_g_is_heavy(g.kind))
        return bz_launch_pdl(ctx, k_fstep<true, 1>, dim3(fbd_grid(ctx, gs, items)), dim3(BZ_VEC_THREADS), 0, 0, st, v, gs, g,
                             part, counter, force_mode);
    if (items > (1LL << 20))
        return bz_launch_pdl(ctx, k_fstep<false, 4>, dim3(fbd_grid(ctx, gs, items)), dim3(BZ_VEC_THREADS), 0, 0, st, v, gs, g,
                             part, counter, force_mode);
    return bz_launch_pdl(ctx, k_fstep<false, 1>, dim3(fbd_grid(ctx, gs, items)), dim3(BZ_VEC_THREADS), 0, 0, st, v, gs, g,
                         part, counter, force_mode);
}

// ---------------------------------------------------------------------------------------------
// D: gradient at z is in; decide (A.4 backtrack_stepsize!, A.5 loop body, A.3 stop rule)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BZ_VEC_THREADS)
    k_dstep(BzState* st, BzVecs v, BzGradSrc gs, double* part, unsigned int* counter) {
    __shared__ double scratch[6 * 32];
    __shared__ double asm_sm[2 * BZ_VEC_THREADS];
    bz_pdl_wait();
    bz_pdl_launch_dependents();
    if (__ldcg(&st->next_op) != OP_D) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) bz_tic(st, 3);
    const long long n = v.n;
    bz_gs_sync_counts(gs, st);
    const double gamma = __ldcg(&st->gamma);
    const double gamma_h = gamma * 0.5;
    double m1 = 0.0, m2 = 0.0;
    const double* __restrict__ vres = v.res;
    const double* __restrict__ vgx = v.grad_x;
    double* __restrict__ vgz = v.grad_z;
    auto elem = [&](long long j, double gz) {
        const double r = vres[j], gx = vgx[j];          // loads before the store: they overlap
        vgz[j] = gz;
        m1 = bz_nanmax(m1, (r / gamma - gx) + gz);      // norm(res/gamma - grad_x + grad_z, Inf)
        m2 = bz_nanmax(m2, (r / gamma_h - gx) + gz);    // same with the halved gamma (edge case below)
    };
    if (bz_gs_blocked(gs)) {
        for (long long jb = (long long)blockIdx.x * BZ_ASM_COLS; jb < n; jb += (long long)gridDim.x * BZ_ASM_COLS) {
            const double gz = bz_assemble_blocked(gs, jb, n, asm_sm);
            const long long j = jb + threadIdx.x;
            if (threadIdx.x < BZ_ASM_COLS && j < n) elem(j, gz);
        }
    } else {
        const long long stride = (long long)gridDim.x * blockDim.x;
        long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        for (; j + stride < n; j += 2 * stride) {       // two entries per trip: six loads in flight
            const double g0 = bz_assemble_grad(gs, j, n), g1 = bz_assemble_grad(gs, j + stride, n);
            const double r0 = vres[j], x0 = vgx[j], r1 = vres[j + stride], x1 = vgx[j + stride];
            vgz[j] = g0; vgz[j + stride] = g1;
            m1 = bz_nanmax(m1, (r0 / gamma - x0) + g0);
            m2 = bz_nanmax(m2, (r0 / gamma_h - x0) + g0);
            m1 = bz_nanmax(m1, (r1 / gamma - x1) + g1);
            m2 = bz_nanmax(m2, (r1 / gamma_h - x1) + g1);
        }
        if (j < n) elem(j, bz_assemble_grad(gs, j, n));
    }
    // NaN-propagating block max: the NaN flag travels in a second lane
    double a[2] = {m1 != m1 ? 0.0 : m1, m2 != m2 ? 0.0 : m2};
    double f[2] = {m1 != m1 ? 1.0 : 0.0, m2 != m2 ? 1.0 : 0.0};
    bz_block_max<2>(a, scratch);
    bz_block_max<2>(f, scratch);
    const int nb = gridDim.x;
    if (threadIdx.x == 0) {
        part[blockIdx.x] = f[0] > 0.0 ? NANQ : a[0];
        part[nb + blockIdx.x] = f[1] > 0.0 ? NANQ : a[1];
    }
    if (!bz_last_cta(counter, nb)) return;

    double stop1 = bz_final_nanmax(part, nb, scratch);
    double stop2 = bz_final_nanmax(part + nb, nb, scratch);
    const double* const ptrs[3] = {gs.lxpart, gs.fxpart, gs.fxpart2};
    const int counts[3] = {gs.nlx, gs.nfx, gs.nfx2};
    double sums[3];
    bz_final_sums<3>(ptrs, counts, sums, scratch);
    if (threadIdx.x != 0) return;
    if (st->xr.nranks > 1) {                  // slab mode: max / sums over ranks
        double v5[5] = {stop1, stop2, sums[0], sums[1], sums[2]};
        bz_xr_allreduce(st, v5, 5, 3u);
        stop1 = v5[0]; stop2 = v5[1]; sums[0] = v5[2]; sums[1] = v5[3]; sums[2] = v5[4];
    }
    double fx;
    const double f_z = bz_value_from_sums(gs, st->musqy, sums[0], sums[1], sums[2], fx);

    bz_d_decide(st, gamma, f_z, fx, stop1, stop2);
    bz_toc(st, 3);
    __threadfence();
}

int bz_launch_dstep(bz_ctx* ctx, BzState* st, const BzVecs& v, const BzGradSrc& gs, double* part,
                    unsigned int* counter) {
    static BzOncePerDevice once;
    if (once.first(ctx->device)) prefer_max_smem(k_dstep);
    return bz_launch_pdl(ctx, k_dstep, dim3(fbd_grid(ctx, gs, v.n)), dim3(BZ_VEC_THREADS), 0, 0, st, v, gs, part,
                         counter);
}

// assemble a gradient + value for the parity entry point bz_auglag_grad
__global__ void __launch_bounds__(BZ_VEC_THREADS)
    k_assemble(BzState* st, BzGradSrc gs, long long n, double* out, double* part,
               unsigned int* counter) {
    __shared__ double scratch[4 * 32];
    __shared__ double asm_sm[2 * BZ_VEC_THREADS];
    bz_gs_sync_counts(gs, st);
    if (bz_gs_blocked(gs)) {
        for (long long jb = (long long)blockIdx.x * BZ_ASM_COLS; jb < n; jb += (long long)gridDim.x * BZ_ASM_COLS) {
            const double gr = bz_assemble_blocked(gs, jb, n, asm_sm);
            const long long j = jb + threadIdx.x;
            if (threadIdx.x < BZ_ASM_COLS && j < n) out[j] = gr;
        }
    } else {
        const long long stride = (long long)gridDim.x * blockDim.x;
        for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
            out[j] = bz_assemble_grad(gs, j, n);
    }
    if (!bz_last_cta(counter, gridDim.x)) return;
    double fx;
    double lx = bz_assemble_value(gs, st, scratch, fx);
    if (threadIdx.x == 0) {
        if (st->xr.nranks > 1) {              // slab mode: lx = sum_r (0.5*slx_r + fx_r) - musqy (musqy is global already)
            double v2[2] = {lx + st->musqy, fx};
            bz_xr_allreduce(st, v2, 2, 0u);
            lx = v2[0] - st->musqy; fx = v2[1];
        }
        st->al_fx = fx; st->al_lx = lx;
        st->out[0] = lx; st->out[1] = fx;
        __threadfence();
    }
}
int bz_launch_assemble(bz_ctx* ctx, BzState* st, const BzGradSrc& gs, long long n, double* out,
                       double* part, unsigned int* counter) {
    k_assemble<<<fbd_grid(ctx, gs, n), BZ_VEC_THREADS, 0, ctx->stream>>>(st, gs, n, out, part, counter);
    ctx->n_launches++;
    BZ_CUDA(ctx, cudaGetLastError());
    return BZ_OK;
}

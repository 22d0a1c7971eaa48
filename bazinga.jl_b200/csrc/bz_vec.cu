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
#include "bz_vec.cuh"
#include <cooperative_groups.h>

#define NANQ __longlong_as_double(0x7ff8000000000000LL)

__device__ __forceinline__ double bz_nanmax(double m, double v) {
    v = fabs(v);
    return (v != v || m != m) ? NANQ : fmax(m, v);
}
__device__ __forceinline__ double jl_max(double a, double b) {   // Julia max propagates NaN
    return (a != a || b != b) ? NANQ : fmax(a, b);
}

int bz_vec_grid(const bz_ctx* ctx, long long n) {
    long long nb = (n + BZ_VEC_THREADS - 1) / BZ_VEC_THREADS;
    long long cap = (long long)ctx->num_sms * 8;
    if (nb < 1) nb = 1;
    if (nb > cap) nb = cap;
    return (int)nb;
}
int bz_l_grid(const bz_ctx* ctx, long long n) {
    long long nb = (n + 4 * BZ_L_THREADS - 1) / (4 * BZ_L_THREADS);
    if (nb < 1) nb = 1;
    if (nb > ctx->num_sms) nb = ctx->num_sms;
    return (int)nb;
}

// ---------------------------------------------------------------------------------------------
// gradient / value assembly from the provider outputs (auglagfun.jl:78-84)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double bz_assemble_grad(const BzGradSrc& gs, long long j, long long n) {
    double dfx = 0.0, jtv = 0.0;
    if (gs.Pf) {
        double s = 0.0;
        for (int p = 0; p < gs.npf; ++p) s += __ldcg(gs.Pf + (size_t)p * n + j);
        dfx = s;
    } else if (gs.Ef) {
        dfx = __ldcg(gs.Ef + j);
    }
    if (gs.Pc) {
        double s = 0.0;
        for (int p = 0; p < gs.npc; ++p) s += __ldcg(gs.Pc + (size_t)p * n + j);
        jtv = s;
    } else if (gs.Ec) {
        jtv = __ldcg(gs.Ec + j);
    }
    return dfx + jtv;   // dlx .= al.dfx .+ al.jtv
}

// fixed-order parallel sum of a global partial array by one whole CTA; result in every thread
__device__ inline double bz_final_sum(const double* part, int n, double* scratch) {
    double v[1] = {0.0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) v[0] += __ldcg(part + i);
    bz_block_sum<1>(v, scratch);
    return v[0];
}
__device__ inline double bz_final_nanmax(const double* part, int n, double* scratch) {
    double m = 0.0;
    int isnan_ = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double v = __ldcg(part + i);
        if (v != v) isnan_ = 1; else m = fmax(m, v);
    }
    double a[1] = {m};
    bz_block_max<1>(a, scratch);
    double b[1] = {(double)isnan_};
    bz_block_max<1>(b, scratch);
    return b[0] > 0.0 ? NANQ : a[0];
}

// AL value from the provider partials:  lx = 0.5*sum(..) ; lx += fx ; lx -= musqy   (:78-82)
__device__ inline double bz_assemble_value(const BzGradSrc& gs, BzState* st, double* scratch,
                                           double& fx_out) {
    const double slx = gs.lxpart ? bz_final_sum(gs.lxpart, gs.nlx, scratch) : 0.0;
    const double sfx = gs.fxpart ? bz_final_sum(gs.fxpart, gs.nfx, scratch) : 0.0;
    const double sfx2 = gs.fxpart2 ? bz_final_sum(gs.fxpart2, gs.nfx2, scratch) : 0.0;
    double lx = 0.5 * slx;
    double fx = gs.fxpart ? gs.fx_scale * sfx : 0.0;
    if (gs.fxpart2) fx = fx + gs.fx_scale2 * sfx2;
    lx += fx;
    lx -= st->musqy;
    fx_out = fx;
    return lx;
}

// ---------------------------------------------------------------------------------------------
// F: forward-backward step (also the two init modes of Appendix A.4)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BZ_VEC_THREADS)
    k_fstep(BzState* st, BzVecs v, BzGradSrc gs, BzGParams g, double* part, unsigned int* counter,
            int force_mode) {
    __shared__ double scratch[4 * 32];
    int mode = force_mode;
    if (mode < 0) {
        if (__ldcg(&st->next_op) != OP_F) return;
        mode = __ldcg(&st->mode);
    }
    const long long n = v.n;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long j0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double a[3] = {0.0, 0.0, 0.0};
    int fresh = 0;

    if (mode == MODE_INIT0) {
        for (long long j = j0; j < n; j += stride) {
            const double gr = bz_assemble_grad(gs, j, n);
            v.grad_x[j] = gr;
            v.xeps[j] = v.x[j] + 1.0;   // lower_bound_smoothness_constant: x .+ 1
        }
    } else if (mode == MODE_INIT1) {
        for (long long j = j0; j < n; j += stride) {
            const double df = bz_assemble_grad(gs, j, n) - v.grad_x[j];
            a[0] += df * df;
        }
    } else {
        const double gamma = __ldcg(&st->gamma);
        fresh = __ldcg(&st->grad_x_fresh);
        if (g.kind == BZ_G_OBSTACLE_L1) {
            // demo/obstacle.jl:67-81: soft threshold block 1 by gamma, max(0,.) on blocks 2-3, then
            // complementarity zeroing between blocks 2 and 3 (tie zeroes the block-2 entry)
            const long long N = g.N;
            for (long long i = j0; i < N; i += stride) {
                double gr[3], xv[3], zv[3];
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    const long long j = i + b * N;
                    gr[b] = fresh ? bz_assemble_grad(gs, j, n) : v.grad_x[j];
                    if (fresh) v.grad_x[j] = gr[b];
                    xv[b] = v.x[j];
                }
                zv[0] = fmax(0.0, (xv[0] - gamma * gr[0]) - gamma);
                zv[1] = fmax(0.0, xv[1] - gamma * gr[1]);
                zv[2] = fmax(0.0, xv[2] - gamma * gr[2]);
                if (zv[1] > zv[2]) zv[2] = 0.0; else zv[1] = 0.0;
                a[2] += zv[0];
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    const long long j = i + b * N;
                    const double r = xv[b] - zv[b];
                    v.z[j] = zv[b];
                    v.res[j] = r;
                    a[0] += gr[b] * r;
                    a[1] += r * r;
                }
            }
        } else {
            for (long long j = j0; j < n; j += stride) {
                const double gr = fresh ? bz_assemble_grad(gs, j, n) : v.grad_x[j];
                if (fresh) v.grad_x[j] = gr;
                const double xj = v.x[j];
                const double y = xj - gamma * gr;
                const double z = bz_prox_elem(g, y, gamma, j, a[2]);
                const double r = xj - z;
                v.z[j] = z;
                v.res[j] = r;
                a[0] += gr * r;
                a[1] += r * r;
            }
        }
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
    const double s0 = bz_final_sum(part, nb, scratch);
    const double s1 = bz_final_sum(part + nb, nb, scratch);
    const double s2 = bz_final_sum(part + 2 * nb, nb, scratch);
    double fx = 0.0, lx = 0.0;
    const bool need_value = (mode == MODE_INIT0) || (mode >= MODE_INITFB && fresh);
    if (need_value) lx = bz_assemble_value(gs, st, scratch, fx);
    if (threadIdx.x != 0) return;
    if (mode == MODE_INIT0) {
        st->f_x = lx; st->al_fx = fx; st->al_lx = lx;
        st->n_grad += 1;
        st->grad_target = TGT_XEPS;
        st->after_g = OP_F;
        st->mode = MODE_INIT1;
        st->next_op = OP_G;
    } else if (mode == MODE_INIT1) {
        const double Lest = sqrt(s0) / sqrt((double)n);
        st->Lest = Lest;
        st->gamma = st->alpha / Lest;     // L = 0 -> gamma = Inf, IEEE semantics as upstream
        st->n_grad += 1;
        st->mode = MODE_INITFB;
        st->grad_x_fresh = 0;
        st->next_op = OP_F;
    } else {
        if (fresh) {
            st->f_x = lx; st->al_fx = fx; st->al_lx = lx;
            st->n_grad += 1;
        }
        const double gamma = st->gamma;
        const double nrm = sqrt(s1);
        const double nrm2 = nrm * nrm;                       // norm(res)^2
        st->dot_gr = s0;
        st->res_nrm2sq = nrm2;
        st->f_up = st->f_x - s0 + ((st->alpha / gamma) / 2) * nrm2;   // f_model
        st->g_z = bz_g_value_scale(g) * s2;
        st->n_prox += 1;
        st->grad_x_fresh = 0;
        st->grad_target = TGT_Z;
        st->after_g = OP_D;
        st->next_op = OP_G;
        st->out[0] = st->g_z; st->out[1] = s0; st->out[2] = s1;
    }
    __threadfence();
}

int bz_launch_fstep(bz_ctx* ctx, BzState* st, const BzVecs& v, const BzGradSrc& gs,
                    const BzGParams& g, double* part, unsigned int* counter, int force_mode) {
    const long long items = (g.kind == BZ_G_OBSTACLE_L1) ? g.N : v.n;
    k_fstep<<<bz_vec_grid(ctx, items), BZ_VEC_THREADS, 0, ctx->stream>>>(st, v, gs, g, part, counter,
                                                                         force_mode);
    ctx->n_launches++;
    BZ_CUDA(ctx, cudaGetLastError());
    return BZ_OK;
}

// ---------------------------------------------------------------------------------------------
// D: gradient at z is in; decide (A.4 backtrack_stepsize!, A.5 loop body, A.3 stop rule)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BZ_VEC_THREADS)
    k_dstep(BzState* st, BzVecs v, BzGradSrc gs, double* part, unsigned int* counter) {
    __shared__ double scratch[4 * 32];
    if (__ldcg(&st->next_op) != OP_D) return;
    const long long n = v.n;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const double gamma = __ldcg(&st->gamma);
    const double gamma_h = gamma * 0.5;
    double m1 = 0.0, m2 = 0.0;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const double gz = bz_assemble_grad(gs, j, n);
        v.grad_z[j] = gz;
        const double r = v.res[j], gx = v.grad_x[j];
        m1 = bz_nanmax(m1, (r / gamma - gx) + gz);      // norm(res/gamma - grad_x + grad_z, Inf)
        m2 = bz_nanmax(m2, (r / gamma_h - gx) + gz);
    }
    // NaN-propagating block max: encode NaN as +inf marker in a second lane
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

    const double stop1 = bz_final_nanmax(part, nb, scratch);
    const double stop2 = bz_final_nanmax(part + nb, nb, scratch);
    double fx;
    const double f_z = bz_assemble_value(gs, st, scratch, fx);
    if (threadIdx.x != 0) return;

    st->f_z = f_z; st->al_fx = fx; st->al_lx = f_z;
    st->n_grad += 1;
    const int mode = st->mode;
    const bool cond = (f_z > st->f_up + 10 * BZ_EPS * (1 + fabs(f_z))) && (gamma >= st->gamma_min);
    bool end_iter = false;
    double stopn = stop1;
    if (mode == MODE_INITFB) {          // backtrack_stepsize! (A.4)
        if (cond) {
            st->gamma = gamma / 2;
            st->grad_x_fresh = 0;
            st->next_op = OP_F;
        } else {
            end_iter = true;
            st->skip_update = 1;
            st->mode = MODE_ITER;
        }
    } else {                            // A.5 loop body at backtrack index k
        const int k = st->k_bt;
        if (cond) {
            st->gamma = gamma * 0.5;
            st->currmem = 0; st->curridx = 0; st->H = 1.0;       // reset_direction_state!
            if (k + 1 > st->max_backtracks) {                    // `continue` ran off the loop
                end_iter = true;
                stopn = stop2;
            } else {
                st->k_bt = k + 1;
                st->grad_x_fresh = 0;
                st->next_op = OP_F;
            }
        } else if (st->f_up + st->g_z <= st->thresh || k >= st->max_backtracks) {
            end_iter = true;
        } else {
            st->tau = (k >= st->max_backtracks - 1) ? 0.0 : st->tau / 2;
            st->k_bt = k + 1;
            st->l_role = LROLE_INTERP;
            st->next_op = OP_L;
        }
        if (end_iter) st->skip_update = 0;
    }
    if (end_iter) {
        st->iter += 1;
        st->stopnorm = stopn;
        if (st->iter >= st->maxit || stopn <= st->tol) {
            st->inner_status = (stopn <= st->tol) ? IST_CONVERGED : IST_MAXIT;
            st->next_op = OP_DONE;
        } else {
            st->l_role = LROLE_NEWITER;
            st->next_op = OP_L;
        }
    }
    __threadfence();
}

int bz_launch_dstep(bz_ctx* ctx, BzState* st, const BzVecs& v, const BzGradSrc& gs, double* part,
                    unsigned int* counter) {
    k_dstep<<<bz_vec_grid(ctx, v.n), BZ_VEC_THREADS, 0, ctx->stream>>>(st, v, gs, part, counter);
    ctx->n_launches++;
    BZ_CUDA(ctx, cudaGetLastError());
    return BZ_OK;
}

// assemble a gradient + value for the parity entry point bz_auglag_grad
__global__ void __launch_bounds__(BZ_VEC_THREADS)
    k_assemble(BzState* st, BzGradSrc gs, long long n, double* out, double* part,
               unsigned int* counter) {
    __shared__ double scratch[4 * 32];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
        out[j] = bz_assemble_grad(gs, j, n);
    if (!bz_last_cta(counter, gridDim.x)) return;
    double fx;
    const double lx = bz_assemble_value(gs, st, scratch, fx);
    if (threadIdx.x == 0) {
        st->al_fx = fx; st->al_lx = lx;
        st->out[0] = lx; st->out[1] = fx;
        __threadfence();
    }
}
int bz_launch_assemble(bz_ctx* ctx, BzState* st, const BzGradSrc& gs, long long n, double* out,
                       double* part, unsigned int* counter) {
    k_assemble<<<bz_vec_grid(ctx, n), BZ_VEC_THREADS, 0, ctx->stream>>>(st, gs, n, out, part, counter);
    ctx->n_launches++;
    BZ_CUDA(ctx, cudaGetLastError());
    return BZ_OK;
}

// ---------------------------------------------------------------------------------------------
// L: L-BFGS update + two-loop + x_d (cooperative launch; all-CTA reductions in fixed order)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

struct TeamCtx {
    BzTeam tm;
    int idx;
    double* scratch;   // >= 2*32 + 2 doubles of shared memory
};

template <int K>
__device__ inline void team_sum(TeamCtx& tc, double (&v)[K]) {
    bz_block_sum<K>(v, tc.scratch);
    const int nb = gridDim.x;
    if (nb == 1) return;
    const int j = tc.idx++;
    unsigned int* bar = tc.tm.bars + tc.tm.parity * BZ_MAXBAR + j;
    double* sl = tc.tm.slots + ((size_t)(tc.tm.parity * BZ_MAXBAR + j) * tc.tm.max_ctas) * 2;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) sl[blockIdx.x * 2 + k] = v[k];
        __threadfence();
        atomicAdd(bar, 1u);
        while (ld_acquire_u32(bar) < (unsigned int)nb) { }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double s = 0.0;
            for (int c = threadIdx.x; c < nb; c += 32) s += __ldcg(sl + c * 2 + k);
            s = bz_warp_sum(s);
            if (threadIdx.x == 0) tc.scratch[64 + k] = s;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = tc.scratch[64 + k];
    __syncthreads();
}

__global__ void __launch_bounds__(BZ_L_THREADS)
    k_lstep(BzState* st, BzVecs v, BzTeam tm, unsigned int* counter, int force_role) {
    __shared__ double scratch[2 * 32 + 4];
    // the other parity's barrier words are free again: reset them for the launch after this one
    if (blockIdx.x == 0 && threadIdx.x < BZ_MAXBAR)
        tm.bars[(1 - tm.parity) * BZ_MAXBAR + threadIdx.x] = 0u;
    int role = force_role;
    if (role < 0) {
        if (__ldcg(&st->next_op) != OP_L) return;
        role = __ldcg(&st->l_role);
    }
    const long long n = v.n;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long j0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;

    if (role == LROLE_INTERP) {      // x = tau*x_d + (1-tau)*z_curr   (A.5)
        const double tau = __ldcg(&st->tau);
        for (long long j = j0; j < n; j += stride)
            v.x[j] = tau * v.x_d[j] + (1 - tau) * v.z_curr[j];
        if (bz_last_cta(counter, gridDim.x) && threadIdx.x == 0) {
            st->grad_target = TGT_X; st->after_g = OP_F; st->grad_x_fresh = 1;
            st->next_op = OP_G;
            __threadfence();
        }
        return;
    }

    TeamCtx tc{tm, 0, scratch};
    const int M = __ldcg(&st->mem);
    int currmem = __ldcg(&st->currmem), curridx = __ldcg(&st->curridx);
    double H = __ldcg(&st->H);
    const int skip = __ldcg(&st->skip_update);
    double ys[BZ_MAXMEM], al[BZ_MAXMEM];
    for (int i = 0; i < BZ_MAXMEM; ++i) { ys[i] = __ldcg(&st->ys[i]); al[i] = 0.0; }

    // ---- update!(H, x - x_prev, res - res_prev)
    if (!skip && M > 0) {
        double a[2] = {0.0, 0.0};
        for (long long j = j0; j < n; j += stride) {
            const double s = v.x[j] - v.x_prev[j], yv = v.res[j] - v.res_prev[j];
            a[0] += s * yv;
            a[1] += yv * yv;
        }
        team_sum<2>(tc, a);
        if (a[0] > 0) {
            curridx += 1;
            if (curridx > M) curridx = 1;
            currmem = currmem + 1 > M ? M : currmem + 1;
            ys[curridx - 1] = a[0];
            double* sm = v.sM + (size_t)(curridx - 1) * n;
            double* ym = v.yM + (size_t)(curridx - 1) * n;
            for (long long j = j0; j < n; j += stride) {
                sm[j] = v.x[j] - v.x_prev[j];
                ym[j] = v.res[j] - v.res_prev[j];
            }
            H = a[0] / a[1];
        }
    }

    // ---- d = H * (-res): two-loop recursion, each axpy fused with the next dot product
    int idx = curridx;
    {
        double a[1] = {0.0};
        if (currmem > 0) {
            const double* sm = v.sM + (size_t)(idx - 1) * n;
            for (long long j = j0; j < n; j += stride) {
                const double dj = -v.res[j];
                v.d[j] = dj;
                a[0] += sm[j] * dj;
            }
        } else {
            for (long long j = j0; j < n; j += stride) v.d[j] = H * (-v.res[j]);
        }
        for (int t = 0; t < currmem; ++t) {       // loop1!: newest -> oldest
            team_sum<1>(tc, a);
            const int i = idx - 1;
            const double alpha_i = a[0] / ys[i];
            al[i] = alpha_i;
            int idx2 = idx - 1;
            if (idx2 == 0) idx2 = M;
            const double* ym = v.yM + (size_t)i * n;
            a[0] = 0.0;
            if (t < currmem - 1) {
                const double* sm2 = v.sM + (size_t)(idx2 - 1) * n;
                for (long long j = j0; j < n; j += stride) {
                    const double dj = v.d[j] - alpha_i * ym[j];
                    v.d[j] = dj;
                    a[0] += sm2[j] * dj;
                }
            } else {                               // last: also d .= H .* d, then first dot of loop2!
                int idn = idx2 + 1;
                if (idn > M) idn = 1;
                const double* ymn = v.yM + (size_t)(idn - 1) * n;
                for (long long j = j0; j < n; j += stride) {
                    double dj = v.d[j] - alpha_i * ym[j];
                    dj = H * dj;
                    v.d[j] = dj;
                    a[0] += ymn[j] * dj;
                }
            }
            idx = idx2;
        }
        for (int t = 0; t < currmem; ++t) {       // loop2!: oldest -> newest
            idx += 1;
            if (idx > M) idx = 1;
            team_sum<1>(tc, a);
            const int i = idx - 1;
            const double beta = a[0] / ys[i];
            const double coef = al[i] - beta;
            const double* sm = v.sM + (size_t)i * n;
            a[0] = 0.0;
            if (t < currmem - 1) {
                int idn = idx + 1;
                if (idn > M) idn = 1;
                const double* ymn = v.yM + (size_t)(idn - 1) * n;
                for (long long j = j0; j < n; j += stride) {
                    const double dj = v.d[j] + coef * sm[j];
                    v.d[j] = dj;
                    a[0] += ymn[j] * dj;
                }
            } else {
                for (long long j = j0; j < n; j += stride) v.d[j] = v.d[j] + coef * sm[j];
            }
        }
    }
    // ---- x_prev <- x; res_prev <- res; x_d = x + d; x <- x_d; z_curr <- z
    for (long long j = j0; j < n; j += stride) {
        const double xj = v.x[j];
        v.x_prev[j] = xj;
        v.res_prev[j] = v.res[j];
        const double xd = xj + v.d[j];
        v.x_d[j] = xd;
        v.x[j] = xd;
        v.z_curr[j] = v.z[j];
    }
    if (bz_last_cta(counter, gridDim.x) && threadIdx.x == 0) {
        const double gamma = st->gamma;
        const double FBE_x =
            (st->f_x - st->dot_gr + ((st->alpha / gamma) / 2) * st->res_nrm2sq) + st->g_z;
        const double sigma = st->beta * (0.5 / gamma) * (1 - st->alpha);
        st->FBE_x = FBE_x;
        st->sigma = sigma;
        st->thresh = FBE_x - sigma * st->res_nrm2sq + 10 * BZ_EPS * (1 + fabs(FBE_x));
        st->tau = 1.0;
        st->k_bt = 1;
        st->currmem = currmem; st->curridx = curridx; st->H = H;
        for (int i = 0; i < BZ_MAXMEM; ++i) { st->ys[i] = ys[i]; st->alphas[i] = al[i]; }
        st->skip_update = 0;
        st->grad_target = TGT_X; st->after_g = OP_F; st->grad_x_fresh = 1;
        st->next_op = OP_G;
        __threadfence();
    }
}

int bz_launch_lstep(bz_ctx* ctx, BzState* st, const BzVecs& v, const BzTeam& tm,
                    unsigned int* counter, int force_role) {
    const int nb = bz_l_grid(ctx, v.n);
    BzVecs vv = v;
    BzTeam tt = tm;
    void* args[] = {(void*)&st, (void*)&vv, (void*)&tt, (void*)&counter, (void*)&force_role};
    BZ_CUDA(ctx, cudaLaunchCooperativeKernel((void*)k_lstep, dim3(nb), dim3(BZ_L_THREADS), args, 0,
                                             ctx->stream));
    ctx->n_launches++;
    return BZ_OK;
}

// ---------------------------------------------------------------------------------------------
// AL-gradient providers for c = Identity and c = Obstacle
// ---------------------------------------------------------------------------------------------
// cx = x; t = cx + mu*y; s = proj_D(t); yupd = (t - s)/mu; jtv = yupd     (identityFunction.jl)
__global__ void __launch_bounds__(BZ_VEC_THREADS)
    k_identity_c(BzState* st, int force_target, const double* x0, const double* x1, const double* x2,
                 long long n, const double* __restrict__ mu, const double* __restrict__ muy,
                 BzDParams D, double* cx, double* s, double* yupd, double* lxpart,
                 unsigned int* counter, int set_after) {
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
            const double t = c + muy[j];
            const double sv = bz_proj_elem(D.kind, D.lo, D.hi, t);
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
                         int set_after) {
    k_identity_c<<<bz_vec_grid(ctx, n), BZ_VEC_THREADS, 0, ctx->stream>>>(
        st, force_target, src[0], src[1], src[2], n, mu, muy, D, cx, s, yupd, lxpart, counter,
        set_after);
    ctx->n_launches++;
    BZ_CUDA(ctx, cudaGetLastError());
    return BZ_OK;
}

// A*v for the obstacle stencil at grid index i (1-D: SymTridiagonal(2,-1); 2-D: 5-point, Dirichlet)
__device__ __forceinline__ double bz_stencil(const double* __restrict__ v, long long i,
                                             const BzObstacle& ob) {
    if (ob.gx == 0) {
        double out = 2.0 * v[i];
        if (i > 0) out -= v[i - 1];
        if (i + 1 < ob.N) out -= v[i + 1];
        return out;
    }
    const long long r = i / ob.gx, c = i % ob.gx;
    double out = 4.0 * v[i];
    if (c > 0) out -= v[i - 1];
    if (c + 1 < ob.gx) out -= v[i + 1];
    if (r > 0) out -= v[i - ob.gx];
    if (r + 1 < ob.gy) out -= v[i + ob.gx];
    return out;
}

// pass 1: cx = x1 + A x2 - x3; yupd = (cx + mu*y - 0)/mu; partials of lx and of f(x)
__global__ void __launch_bounds__(BZ_VEC_THREADS)
    k_obstacle_c(BzState* st, int force_target, const double* x0, const double* x1p,
                 const double* x2p, BzObstacle ob, const double* __restrict__ mu,
                 const double* __restrict__ muy, double* cx, double* s, double* yupd, double* lxpart,
                 double* fxpart) {
    __shared__ double scratch[3 * 32];
    int target = force_target;
    if (target < 0) {
        if (__ldcg(&st->next_op) != OP_G) return;
        target = __ldcg(&st->grad_target);
    }
    const double* x = target == 0 ? x0 : (target == 1 ? x1p : x2p);
    const long long N = ob.N;
    const long long stride = (long long)gridDim.x * blockDim.x;
    double a[3] = {0.0, 0.0, 0.0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        const double x2 = x[N + i];
        const double c = (x[i] + bz_stencil(x + N, i, ob)) - x[2 * N + i];
        const double t = c + muy[i];
        const double u = t - 0.0;            // SetObstacle: z .= 0 (demo/obstacle.jl:116-119)
        a[0] += (u * u) / mu[i];
        cx[i] = c; s[i] = 0.0; yupd[i] = u / mu[i];
        a[1] += x2 * x2;                     // f = 0.5*sum(x2.^2) - sum(x2)  (:40-42)
        a[2] += x2;
    }
    bz_block_sum<3>(a, scratch);
    if (threadIdx.x == 0) {
        lxpart[blockIdx.x] = a[0];
        fxpart[blockIdx.x] = a[1];
        fxpart[gridDim.x + blockIdx.x] = a[2];
    }
}
// pass 2: grad = dfx + jtv with dfx = [0; x2-1; 0], jtv = [w; A w; -w]
__global__ void __launch_bounds__(BZ_VEC_THREADS)
    k_obstacle_jt(BzState* st, int force_target, const double* x0, const double* x1p,
                  const double* x2p, BzObstacle ob, const double* __restrict__ yupd, double* Ec,
                  unsigned int* counter, int set_after) {
    int target = force_target;
    if (target < 0) {
        if (__ldcg(&st->next_op) != OP_G) return;
        target = __ldcg(&st->grad_target);
    }
    const double* x = target == 0 ? x0 : (target == 1 ? x1p : x2p);
    const long long N = ob.N;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        const double w = yupd[i];
        Ec[i] = 0.0 + w;
        Ec[N + i] = (x[N + i] - 1.0) + bz_stencil(yupd, i, ob);
        Ec[2 * N + i] = 0.0 + (-w);
    }
    if (bz_last_cta(counter, gridDim.x) && threadIdx.x == 0 && set_after) {
        st->next_op = st->after_g;
        __threadfence();
    }
}

int bz_launch_obstacle(bz_ctx* ctx, BzState* st, int force_target, const double* const* src,
                       const BzObstacle& ob, const double* mu, const double* muy, double* cx,
                       double* s, double* yupd, double* Ec, double* lxpart, double* fxpart,
                       unsigned int* counter, int set_after) {
    const int nb = bz_vec_grid(ctx, ob.N);
    k_obstacle_c<<<nb, BZ_VEC_THREADS, 0, ctx->stream>>>(st, force_target, src[0], src[1], src[2], ob,
                                                          mu, muy, cx, s, yupd, lxpart, fxpart);
    k_obstacle_jt<<<nb, BZ_VEC_THREADS, 0, ctx->stream>>>(st, force_target, src[0], src[1], src[2],
                                                           ob, yupd, Ec, counter, set_after);
    ctx->n_launches += 2;
    BZ_CUDA(ctx, cudaGetLastError());
    return BZ_OK;
}

// ---------------------------------------------------------------------------------------------
// standalone prox / proj, and the ALPS outer-step kernels
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BZ_VEC_THREADS)
    k_prox(BzState* st, const double* in, double gamma, double* out, long long n, BzGParams g,
           double* part, unsigned int* counter) {
    __shared__ double scratch[32];
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long j0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double a[1] = {0.0};
    if (g.kind == BZ_G_OBSTACLE_L1) {
        const long long N = g.N;
        for (long long i = j0; i < N; i += stride) {
            double z0 = fmax(0.0, in[i] - gamma);
            double z1 = fmax(0.0, in[N + i]), z2 = fmax(0.0, in[2 * N + i]);
            if (z1 > z2) z2 = 0.0; else z1 = 0.0;
            out[i] = z0; out[N + i] = z1; out[2 * N + i] = z2;
            a[0] += z0;
        }
    } else {
        for (long long j = j0; j < n; j += stride) out[j] = bz_prox_elem(g, in[j], gamma, j, a[0]);
    }
    bz_block_sum<1>(a, scratch);
    if (threadIdx.x == 0) part[blockIdx.x] = a[0];
    if (!bz_last_cta(counter, gridDim.x)) return;
    const double s = bz_final_sum(part, gridDim.x, scratch);
    if (threadIdx.x == 0) {
        st->g_z = bz_g_value_scale(g) * s;
        st->out[0] = st->g_z;
        __threadfence();
    }
}
int bz_launch_prox(bz_ctx* ctx, BzState* st, const double* in, double gamma, double* out, long long n,
                   const BzGParams& g, double* part, unsigned int* counter) {
    const long long items = (g.kind == BZ_G_OBSTACLE_L1) ? g.N : n;
    k_prox<<<bz_vec_grid(ctx, items), BZ_VEC_THREADS, 0, ctx->stream>>>(st, in, gamma, out, n, g, part,
                                                                        counter);
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
        for (long long j = j0; j < m; j += stride) out[j] = bz_proj_elem(D.kind, D.lo, D.hi, in[j]);
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
                int count_outer, double* part, unsigned int* counter) {
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
        a[0] += my * yj;
    }
    bz_block_sum<1>(a, scratch);
    bz_block_max<1>(bad, scratch);
    if (threadIdx.x == 0) { part[blockIdx.x] = a[0]; part[gridDim.x + blockIdx.x] = bad[0]; }
    if (!bz_last_cta(counter, gridDim.x)) return;
    const double s = bz_final_sum(part, gridDim.x, scratch);
    const double b = bz_final_sum(part + gridDim.x, gridDim.x, scratch);
    if (threadIdx.x == 0) {
        st->musqy = 0.5 * s;
        if (b > 0.0) st->err_flag = 1;
        if (count_outer) st->tot_it += 1;
        __threadfence();
    }
}
int bz_launch_al_update(bz_ctx* ctx, BzState* st, double* mu, double* y, double* muy, long long m,
                        int safeguard, int count_outer, double* part, unsigned int* counter) {
    k_al_update<<<bz_vec_grid(ctx, m), BZ_VEC_THREADS, 0, ctx->stream>>>(st, mu, y, muy, m, safeguard,
                                                                         count_outer, part, counter);
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
            const double sv = bz_proj_elem(D.kind, D.lo, D.hi, t);  // proj!(s, D, y)
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
    const double nrp = bz_final_nanmax(part, gridDim.x, scratch);
    if (threadIdx.x != 0) return;
    st->out[0] = nrp;
    if (outer_logic) bz_outer_logic(st, nrp);
    __threadfence();
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
            if (compute_s) { sv = bz_proj_elem(D.kind, D.lo, D.hi, cx[j]); s[j] = sv; }
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
                              double gamma_min, int max_backtracks, int mem) {
    st->tol = tol; st->maxit = maxit; st->alpha = alpha; st->beta = beta;
    st->gamma_min = gamma_min; st->max_backtracks = max_backtracks; st->mem = mem;
    st->iter = 0; st->inner_status = IST_RUNNING;
    st->currmem = 0; st->curridx = 0; st->H = 1.0;
    st->tau = 1.0; st->k_bt = 1; st->skip_update = 1;
    st->mode = MODE_INIT0; st->grad_target = TGT_X; st->after_g = OP_F; st->grad_x_fresh = 1;
    st->l_role = LROLE_NEWITER;
    st->next_op = OP_G;
}
int bz_launch_panoc_begin(bz_ctx* ctx, BzState* st, double tol, long long maxit, double alpha,
                          double beta, double gamma_min, int max_backtracks, int mem) {
    k_panoc_begin<<<1, 1, 0, ctx->stream>>>(st, tol, maxit, alpha, beta, gamma_min, max_backtracks, mem);
    ctx->n_launches++;
    BZ_CUDA(ctx, cudaGetLastError());
    return BZ_OK;
}

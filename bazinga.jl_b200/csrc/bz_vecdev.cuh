// Device helpers shared by the vector kernels (bz_vec.cu, bz_fbd.cu, bz_lbfgs.cu).
#pragma once
#include "bz_vec.cuh"

#define NANQ __longlong_as_double(0x7ff8000000000000LL)

__device__ __forceinline__ double bz_nanmax(double m, double v) {
    v = fabs(v);
    return (v != v || m != m) ? NANQ : fmax(m, v);
}
__device__ __forceinline__ double jl_max(double a, double b) {   // Julia max propagates NaN
    return (a != a || b != b) ? NANQ : fmax(a, b);
}

// ---------------------------------------------------------------------------------------------
// gradient / value assembly from the provider outputs (auglagfun.jl:78-84)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double bz_assemble_grad(const BzGradSrc& gs, long long j, long long n) {
    if (gs.nx > 0) {                      // als: [dfx + jtv; -yupd]
        if (j >= gs.nx) return -__ldcg(gs.tail + (j - gs.nx));
        n = gs.nx;                        // row stride of the partial arrays
    }
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

// several fixed-order partial-array sums with ONE block reduction (last-CTA epilogues)
template <int K>
__device__ inline void bz_final_sums(const double* const (&ptrs)[K], const int (&counts)[K],
                                     double (&out)[K], double* scratch) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double s = 0.0;
        if (ptrs[k])
            for (int i = threadIdx.x; i < counts[k]; i += blockDim.x) s += __ldcg(ptrs[k] + i);
        out[k] = s;
    }
    bz_block_sum<K>(out, scratch);
}
// AL value given the three provider sums (auglagfun.jl:78-82)
__device__ __forceinline__ double bz_value_from_sums(const BzGradSrc& gs, double musqy, double slx,
                                                     double sfx, double sfx2, double& fx_out) {
    double lx = 0.5 * slx;
    double fx = gs.fxpart ? gs.fx_scale * sfx : 0.0;
    if (gs.fxpart2) fx = fx + gs.fx_scale2 * sfx2;
    lx += fx;
    lx -= musqy;
    fx_out = fx;
    return lx;
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
// Blocked assembly for providers that leave per-cluster partial rows (the dense passes): all
// BZ_VEC_THREADS threads of the CTA cooperate on BZ_ASM_COLS consecutive columns -- thread
// (grp, col) sums a contiguous quarter of the partial rows with its loads batched 8 deep, the four
// group sums are combined in fixed order.  The value is returned in threads tid < BZ_ASM_COLS (for
// column jb + tid); other threads get 0.  Deterministic, no atomics.  `sm` needs 2*256 doubles.
// ---------------------------------------------------------------------------------------------
#define BZ_ASM_COLS 64
#define BZ_ASM_GROUPS (BZ_VEC_THREADS / BZ_ASM_COLS)

__device__ __forceinline__ double bz_sum_group(const double* __restrict__ P, int np, int grp,
                                               long long j, long long n) {
    const int cnt = (np + BZ_ASM_GROUPS - 1) / BZ_ASM_GROUPS;
    int p = grp * cnt;
    int p1 = p + cnt;
    if (p1 > np) p1 = np;
    double s = 0.0;
    const double* q = P + (size_t)p * n + j;
    // all loads of a 24-row batch are issued before the (order-fixed) adds: one L2 round trip
    for (; p < p1; p += 24, q += 24 * (size_t)n) {
        double v[24];
#pragma unroll
        for (int k = 0; k < 24; ++k) v[k] = (p + k < p1) ? __ldcg(q + (size_t)k * n) : 0.0;
#pragma unroll
        for (int k = 0; k < 24; ++k)
            if (p + k < p1) s += v[k];
    }
    return s;
}

__device__ __forceinline__ bool bz_gs_blocked(const BzGradSrc& gs) { return gs.Pf || gs.Pc; }
// the dense constraint pass records how many partial rows it wrote (it runs on fewer row clusters when the speculative
// L step rides along, bz_lspec.cuh): consumers take the count from the state
__device__ __forceinline__ void bz_gs_sync_counts(BzGradSrc& gs, const BzState* st) {
    const int npc = __ldcg(&st->npc_c);
    if (gs.Pc && npc > 0) { gs.npc = npc; gs.nlx = npc; }
}

__device__ inline double bz_assemble_blocked(const BzGradSrc& gs, long long jb, long long n, double* sm) {
    const int col = threadIdx.x & (BZ_ASM_COLS - 1), grp = threadIdx.x / BZ_ASM_COLS;
    const long long j = jb + col;
    const long long nhead = gs.nx > 0 ? gs.nx : n;     // als: entries >= nx are -yupd (no partial rows there)
    double sf = 0.0, sc = 0.0;
    if (j < nhead) {
        if (gs.Pf) sf = bz_sum_group(gs.Pf, gs.npf, grp, j, nhead);
        if (gs.Pc) sc = bz_sum_group(gs.Pc, gs.npc, grp, j, nhead);
    }
    sm[grp * BZ_ASM_COLS + col] = sf;
    sm[BZ_VEC_THREADS + grp * BZ_ASM_COLS + col] = sc;
    __syncthreads();
    double out = 0.0;
    if (grp == 0 && j >= nhead && j < n) out = -__ldcg(gs.tail + (j - nhead));
    if (grp == 0 && j < nhead) {
        double dfx = 0.0, jtv = 0.0;
        if (gs.Pf) dfx = (sm[col] + sm[BZ_ASM_COLS + col]) + (sm[2 * BZ_ASM_COLS + col] + sm[3 * BZ_ASM_COLS + col]);
        else if (gs.Ef) dfx = __ldcg(gs.Ef + j);
        const double* sc4 = sm + BZ_VEC_THREADS;
        if (gs.Pc) jtv = (sc4[col] + sc4[BZ_ASM_COLS + col]) + (sc4[2 * BZ_ASM_COLS + col] + sc4[3 * BZ_ASM_COLS + col]);
        else if (gs.Ec) jtv = __ldcg(gs.Ec + j);
        out = dfx + jtv;   // dlx .= al.dfx .+ al.jtv
    }
    __syncthreads();
    return out;
}

// ---------------------------------------------------------------------------------------------
// Scalar logic of the state machine, run by ONE thread once the sums of a step are known.  Shared by the stand-alone
// F / D kernels (bz_fbd.cu) and the fused obstacle kernel (bz_obfused.cu), so both take the same decisions bit for bit.
// `fused` problems go L -> X -> (X | L): where the plain cycle asks for a gradient at x followed by F, they ask for X.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bz_goto_fb(BzState* st) {     // "evaluate the AL-gradient at x, then the forward-backward step"
    if (st->fused) { st->next_op = OP_X; return; }
    st->grad_target = TGT_X; st->after_g = OP_F;
    st->next_op = OP_G;
}
// forward-backward step done at gamma with s0 = <grad_x, res>, s1 = ||res||^2 (before the sqrt round trip), s2 = g partial sum
__device__ __forceinline__ void bz_f_finish(BzState* st, const BzGParams& g, double gamma, int fresh, double s0, double s1,
                                            double s2, double lx, double fx) {
    if (fresh) {
        st->f_x = lx; st->al_fx = fx; st->al_lx = lx;
        st->n_grad += 1;
    }
    const double nrm = sqrt(s1);
    const double nrm2 = nrm * nrm;                       // norm(res)^2
    st->dot_gr = s0;
    st->res_nrm2sq = nrm2;
    st->f_up = st->f_x - s0 + ((st->alpha / gamma) / 2) * nrm2;   // f_model
    st->g_z = bz_g_value_scale(g) * s2;
    st->n_prox += 1;
    st->grad_x_fresh = 0;
    st->out[0] = st->g_z; st->out[1] = s0; st->out[2] = s1;
}
// gradient at z is in: backtrack_stepsize! (A.4), the loop body of A.5, the stop rule A.3
__device__ __forceinline__ void bz_d_decide(BzState* st, double gamma, double f_z, double fx, double stop1, double stop2) {
    const int refb = st->fused ? OP_X : OP_F;            // "redo the forward-backward step with the halved gamma"
    const int spec = st->spec_ok;                        // an L step was speculated beside this G(z) (bz_lspec.cuh)
    st->spec_ok = 0; st->spec_commit = 0;
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
            st->next_op = refb;
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
                st->next_op = refb;
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
        } else if (spec && !cond) {
            // The L step of the next iteration was computed beside this G(z) on exactly this verdict (bz_lspec.cuh): its
            // scalars go into place here, the L kernel is skipped, G reads x + d from the shadow buffer and the F step
            // behind it copies the vectors into place (spec_commit stays set until then).  (cond: the line search ran off
            // its loop on a gamma-halving, which resets the history the speculation used -- the plain L kernel runs.)
            const double gm = st->gamma;
            const double FBE_x = (st->f_x - st->dot_gr + ((st->alpha / gm) / 2) * st->res_nrm2sq) + st->g_z;
            const double sigma = st->beta * (0.5 / gm) * (1 - st->alpha);
            st->FBE_x = FBE_x;
            st->sigma = sigma;
            st->thresh = FBE_x - sigma * st->res_nrm2sq + 10 * BZ_EPS * (1 + fabs(FBE_x));
            st->tau = 1.0;
            st->k_bt = 1;
            st->currmem = st->sp_currmem; st->curridx = st->sp_curridx; st->H = st->sp_H;
            for (int i = 0; i < BZ_MAXMEM; ++i) { st->ys[i] = st->sp_ys[i]; st->alphas[i] = st->sp_alphas[i]; }
            st->skip_update = 0;
            st->spec_commit = 1;
            st->grad_x_fresh = 1;
            bz_goto_fb(st);
        } else {
            st->l_role = LROLE_NEWITER;
            st->next_op = OP_L;
        }
    }
}

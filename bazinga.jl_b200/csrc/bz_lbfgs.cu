// L kernels of the PANOC+ state machine: L-BFGS update!(H, x - x_prev, res - res_prev), the two-loop
// direction d = H*(-res) (ProximalAlgorithms LBFGS operator), x_d = x + d,
// and the tau-interpolation x = tau*x_d + (1-tau)*z_curr.  Two implementations with the same
// arithmetic per element and the same order of the 2*currmem (+1) sequential dot products:
//   k_lstep_cluster  n <= 8 CTAs x 256 threads x 16: ONE thread-block cluster; d and the vectors
//                    of the current step live in registers, every dot product is a warp-shuffle +
//                    shared-memory block sum followed by an all-to-all through distributed shared
//                    memory and one hardware cluster barrier (north_star: "warp-shuffle and
//                    cluster-level reductions").
//   k_lstep          any n: cooperative launch over up to one CTA per SM with a software grid
//                    barrier; each axpy pass is fused with the next dot product.
// All sums are order-fixed (no float atomics).  Compiled with -fmad=false.
#include "bz_vecdev.cuh"
#include <cooperative_groups.h>

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

struct TeamCtx {
    BzTeam tm;
    int idx;
    double* scratch;   // >= 2*32 + 2 doubles of shared memory
    BzState* st;
    unsigned long long seq0;   // slab mode: cross-rank exchange sequence number at kernel start
    int xidx;                  // ... and the number of exchanges this launch has done
};

template <int K>
__device__ inline void team_sum(TeamCtx& tc, double (&v)[K]) {
    bz_block_sum<K>(v, tc.scratch);
    const int nb = gridDim.x;
    if (nb > 1) {
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
    if (tc.st->xr.nranks > 1) {
        // slab mode: the local totals of all ranks, added in rank order by every CTA (identical bits
        // everywhere).  CTA 0 posts; the local grid barrier of the next team_sum keeps any CTA from
        // running more than one exchange ahead of the slowest one.
        const unsigned long long s = tc.seq0 + (unsigned long long)(tc.xidx++);
        if (threadIdx.x == 0) {
            if (blockIdx.x == 0) bz_xr_post(tc.st->xr, s, v, K);
            double g[XR_MAXR * XR_K];
            bz_xr_gather(tc.st, s, K, g);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                double acc = g[k];
                for (int r = 1; r < tc.st->xr.nranks; ++r) acc += g[r * XR_K + k];
                tc.scratch[64 + k] = acc;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = tc.scratch[64 + k];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(BZ_L_THREADS, BZ_L_CTAS_PER_SM)
    k_lstep(BzState* st, BzVecs v, BzTeam tm, unsigned int* counter, int force_role) {
    __shared__ double scratch[2 * 32 + 4];
    // the other parity's barrier words are free again: reset them for the launch after this one
    if (blockIdx.x == 0 && threadIdx.x < BZ_MAXBAR)
        tm.bars[(1 - tm.parity) * BZ_MAXBAR + threadIdx.x] = 0u;
    int role = force_role;
    if (role < 0) {
        if (__ldcg(&st->next_op) != OP_L) return;
        role = __ldcg(&st->l_role);
        if (blockIdx.x == 0 && threadIdx.x == 0) bz_tic(st, 1);
    }
    const long long n = v.n;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long j0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;

    if (role == LROLE_INTERP) {      // x = tau*x_d + (1-tau)*z_curr   (A.5)
        const double tau = __ldcg(&st->tau);
        _Pragma("unroll 4") for (long long j = j0; j < n; j += stride)
            v.x[j] = tau * v.x_d[j] + (1 - tau) * v.z_curr[j];
        if (bz_last_cta(counter, gridDim.x) && threadIdx.x == 0) {
            st->grad_target = TGT_X; st->after_g = OP_F; st->grad_x_fresh = 1;
            st->next_op = OP_G;
            if (force_role < 0) bz_toc(st, 1);
            __threadfence();
        }
        return;
    }

    TeamCtx tc{tm, 0, scratch, st, st->xr.nranks > 1 ? *st->xr.seq : 0ULL, 0};
    const int M = __ldcg(&st->mem);
    int currmem = __ldcg(&st->currmem), curridx = __ldcg(&st->curridx);
    double H = __ldcg(&st->H);
    const int skip = __ldcg(&st->skip_update);
    double ys[BZ_MAXMEM], al[BZ_MAXMEM];
    for (int i = 0; i < BZ_MAXMEM; ++i) { ys[i] = __ldcg(&st->ys[i]); al[i] = 0.0; }

    // ---- update!(H, x - x_prev, res - res_prev)
    if (!skip && M > 0) {
        double a[2] = {0.0, 0.0};
        _Pragma("unroll 4") for (long long j = j0; j < n; j += stride) {
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
            _Pragma("unroll 4") for (long long j = j0; j < n; j += stride) {
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
            _Pragma("unroll 4") for (long long j = j0; j < n; j += stride) {
                const double dj = -v.res[j];
                v.d[j] = dj;
                a[0] += sm[j] * dj;
            }
        } else {
            _Pragma("unroll 4") for (long long j = j0; j < n; j += stride) v.d[j] = H * (-v.res[j]);
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
                _Pragma("unroll 4") for (long long j = j0; j < n; j += stride) {
                    const double dj = v.d[j] - alpha_i * ym[j];
                    v.d[j] = dj;
                    a[0] += sm2[j] * dj;
                }
            } else {                               // last: also d .= H .* d, then first dot of loop2!
                int idn = idx2 + 1;
                if (idn > M) idn = 1;
                const double* ymn = v.yM + (size_t)(idn - 1) * n;
                _Pragma("unroll 4") for (long long j = j0; j < n; j += stride) {
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
                _Pragma("unroll 4") for (long long j = j0; j < n; j += stride) {
                    const double dj = v.d[j] + coef * sm[j];
                    v.d[j] = dj;
                    a[0] += ymn[j] * dj;
                }
            } else {
                _Pragma("unroll 4") for (long long j = j0; j < n; j += stride) v.d[j] = v.d[j] + coef * sm[j];
            }
        }
    }
    // ---- x_prev <- x; res_prev <- res; x_d = x + d; x <- x_d; z_curr <- z
    _Pragma("unroll 4") for (long long j = j0; j < n; j += stride) {
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
        if (st->xr.nranks > 1) *st->xr.seq = tc.seq0 + (unsigned long long)tc.xidx;
        if (force_role < 0) bz_toc(st, 1);
        __threadfence();
    }
}


// ---------------------------------------------------------------------------------------------
// cluster version
// ---------------------------------------------------------------------------------------------
#define LC_THREADS 256
#define LC_EPT 16          // max elements per thread
#define LC_MAXCS 8         // portable cluster size

__device__ __forceinline__ uint32_t lc_smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint32_t lc_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t lc_cluster_size() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void lc_cluster_sync_relaxed() {   // pairs with fence.mbarrier_init, no membar
    asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ void lc_st_remote(void* local_ptr, uint32_t rank, double v) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(lc_smem_u32(local_ptr)), "r"(rank));
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(remote), "d"(v) : "memory");
}

__device__ __forceinline__ void lc_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(lc_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void lc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(lc_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void lc_mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    const uint32_t addr = lc_smem_u32(bar);
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
// one double into CTA `rank`'s shared memory, completing 8 bytes of tx on that CTA's mbarrier
__device__ __forceinline__ void lc_st_async(void* local_data, uint64_t* local_bar, uint32_t rank, double v) {
    uint32_t rdata, rbar;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rdata) : "r"(lc_smem_u32(local_data)), "r"(rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(lc_smem_u32(local_bar)), "r"(rank));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(rdata),
                 "d"(v), "r"(rbar)
                 : "memory");
}

struct ClusterRed {
    double* wpart;     // [2][8] warp partials (K <= 2)
    double* slots;     // [2 parities][LC_MAXCS][2]
    uint64_t* bars;    // [2 parities]: all senders' stores of a round complete their tx here
    uint32_t rank, cs;
    int round;
};

// All-cluster sum of K values per thread; result in every thread of every CTA, identical bits.
// Warp shuffles -> shared-memory block sum -> all-to-all through distributed shared memory:
// every CTA st.async's its block sums into every CTA's slot row, each store completing on the
// receiver's mbarrier, so there is no cluster barrier (and no membar) per reduction.  Two slot
// sets alternate; a CTA can only be one round ahead of the slowest CTA (it needs everybody's
// value of round r+1 before it can send round r+2), so a set is never overwritten while in use.
template <int K>
__device__ inline void cluster_sum(ClusterRed& cr, double (&v)[K]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = bz_warp_sum(v[k]);
    const int par = cr.round & 1;
    const uint32_t phase = (uint32_t)((cr.round >> 1) & 1);
    cr.round++;
    double* wp = cr.wpart + par * 16;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) wp[k * 8 + warp] = v[k];
    }
    __syncthreads();
    double* sl = cr.slots + par * (LC_MAXCS * 2);
    uint64_t* bar = cr.bars + par;
    if (threadIdx.x < cr.cs) {     // thread t sends this CTA's block sums to CTA t
        if (threadIdx.x == 0) lc_mbar_expect_tx(bar, cr.cs * K * 8u);
        double bs[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const double* w = wp + k * 8;
            bs[k] = ((w[0] + w[1]) + (w[2] + w[3])) + ((w[4] + w[5]) + (w[6] + w[7]));
        }
#pragma unroll
        for (int k = 0; k < K; ++k) lc_st_async(&sl[cr.rank * 2 + k], bar, threadIdx.x, bs[k]);
    }
    lc_mbar_wait_cluster(bar, phase);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double s = 0.0;
        for (uint32_t c = 0; c < cr.cs; ++c) s += sl[c * 2 + k];
        v[k] = s;
    }
}

template <int EPT>
__global__ void __launch_bounds__(LC_THREADS)
    k_lstep_cluster(BzState* st, BzVecs v, int force_role) {
    __shared__ double wpart[2 * 2 * 8];
    __shared__ double slots[2 * LC_MAXCS * 2];
    __shared__ uint64_t bars[2];
    bz_pdl_wait();
    bz_pdl_launch_dependents();
    int role = force_role;
    if (role < 0) {
        if (__ldcg(&st->next_op) != OP_L) return;     // uniform over the cluster
        role = __ldcg(&st->l_role);
        if (blockIdx.x == 0 && threadIdx.x == 0) bz_tic(st, 1);
    }
    ClusterRed cr{wpart, slots, bars, lc_cluster_rank(), lc_cluster_size(), 0};
    if (threadIdx.x == 0) {
        lc_mbar_init(&bars[0], 1);
        lc_mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    lc_cluster_sync_relaxed();   // every CTA's barriers exist before anything is sent to them
    const long long n = v.n;
    const int stride = (int)(cr.cs * LC_THREADS);
    const int j0 = (int)(cr.rank * LC_THREADS + threadIdx.x);
    // element k of this thread is j0 + k*stride; valid iff < n
#define LC_FOR(k) _Pragma("unroll") for (int k = 0; k < EPT; ++k)
#define LC_J(k) ((long long)j0 + (long long)(k) * stride)

    if (role == LROLE_INTERP) {      // x = tau*x_d + (1-tau)*z_curr   (A.5)
        const double tau = __ldcg(&st->tau);
        LC_FOR(k) {
            const long long j = LC_J(k);
            if (j < n) v.x[j] = tau * v.x_d[j] + (1 - tau) * v.z_curr[j];
        }
        // (the consumers of this state are later kernels: they see it only after the whole grid is done)
        if (cr.rank == 0 && threadIdx.x == 0) {
            st->grad_target = TGT_X; st->after_g = OP_F; st->grad_x_fresh = 1;
            st->next_op = OP_G;
            if (force_role < 0) bz_toc(st, 1);
            __threadfence();
        }
        return;
    }

    const int M = __ldcg(&st->mem);
    int currmem = __ldcg(&st->currmem), curridx = __ldcg(&st->curridx);
    double H = __ldcg(&st->H);
    const int skip = __ldcg(&st->skip_update);
    double ys[BZ_MAXMEM], al[BZ_MAXMEM];
    for (int i = 0; i < BZ_MAXMEM; ++i) { ys[i] = __ldcg(&st->ys[i]); al[i] = 0.0; }

    double xv[EPT], rv[EPT], d[EPT];
    LC_FOR(k) {
        const long long j = LC_J(k);
        xv[k] = j < n ? v.x[j] : 0.0;
        rv[k] = j < n ? v.res[j] : 0.0;
    }
    // ---- update!(H, x - x_prev, res - res_prev)
    if (!skip && M > 0) {
        double sv[EPT], yv[EPT];
        double a[2] = {0.0, 0.0};
        LC_FOR(k) {
            const long long j = LC_J(k);
            sv[k] = j < n ? xv[k] - v.x_prev[j] : 0.0;
            yv[k] = j < n ? rv[k] - v.res_prev[j] : 0.0;
        }
        LC_FOR(k) { a[0] += sv[k] * yv[k]; a[1] += yv[k] * yv[k]; }
        cluster_sum<2>(cr, a);
        if (a[0] > 0) {
            curridx += 1;
            if (curridx > M) curridx = 1;
            currmem = currmem + 1 > M ? M : currmem + 1;
            ys[curridx - 1] = a[0];
            double* sm = v.sM + (size_t)(curridx - 1) * n;
            double* ym = v.yM + (size_t)(curridx - 1) * n;
            LC_FOR(k) {
                const long long j = LC_J(k);
                if (j < n) { sm[j] = sv[k]; ym[j] = yv[k]; }
            }
            H = a[0] / a[1];
        }
    }
    // ---- d = H * (-res): two-loop recursion, d in registers.  The 2*currmem steps are run as one
    // sequence (loop1!: newest -> oldest with (dot, axpy) vectors (s_i, y_i); loop2!: oldest -> newest
    // with (y_i, s_i)); the vectors of step k+1 are loaded BEFORE the cluster reduction of step k so
    // their L2 latency hides behind it.
    LC_FOR(k) d[k] = -rv[k];
    const int nsteps = 2 * currmem;
    auto hist_index = [&](int kk) {      // ring slot visited by step kk
        const int k1 = kk < currmem ? kk : nsteps - 1 - kk;
        return ((curridx - 1 - k1) % M + M) % M;
    };
    auto load_step = [&](int kk, double (&A)[EPT], double (&B)[EPT]) {
        if (kk >= nsteps) return;
        const int i = hist_index(kk);
        const double* pa = (kk < currmem ? v.sM : v.yM) + (size_t)i * n;
        const double* pb = (kk < currmem ? v.yM : v.sM) + (size_t)i * n;
        LC_FOR(e) {
            const long long j = LC_J(e);
            A[e] = j < n ? __ldcg(pa + j) : 0.0;
            B[e] = j < n ? __ldcg(pb + j) : 0.0;
        }
    };
    double ra[EPT], rb[EPT];
    load_step(0, ra, rb);
    for (int kk = 0; kk < nsteps; ++kk) {
        if (kk == currmem) { LC_FOR(e) d[e] = H * d[e]; }      // d .= L.H .* d between the loops
        double a[1] = {0.0};
        LC_FOR(e) a[0] += ra[e] * d[e];
        double na[EPT], nb[EPT];
        LC_FOR(e) { na[e] = 0.0; nb[e] = 0.0; }
        load_step(kk + 1, na, nb);
        cluster_sum<1>(cr, a);
        const int i = hist_index(kk);
        if (kk < currmem) {
            const double alpha_i = a[0] / ys[i];
            al[i] = alpha_i;
            LC_FOR(e) d[e] = d[e] - alpha_i * rb[e];
        } else {
            const double beta = a[0] / ys[i];
            const double coef = al[i] - beta;
            LC_FOR(e) d[e] = d[e] + coef * rb[e];
        }
        LC_FOR(e) { ra[e] = na[e]; rb[e] = nb[e]; }
    }
    if (currmem == 0) { LC_FOR(e) d[e] = H * d[e]; }
    // ---- x_prev <- x; res_prev <- res; x_d = x + d; x <- x_d; z_curr <- z
    LC_FOR(k) {
        const long long j = LC_J(k);
        if (j < n) {
            v.d[j] = d[k];
            v.x_prev[j] = xv[k];
            v.res_prev[j] = rv[k];
            const double xd = xv[k] + d[k];
            v.x_d[j] = xd;
            v.x[j] = xd;
            v.z_curr[j] = v.z[j];
        }
    }
    if (cr.rank == 0 && threadIdx.x == 0) {
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
        if (force_role < 0) bz_toc(st, 1);
        __threadfence();
    }
#undef LC_FOR
#undef LC_J
}

int bz_launch_lstep(bz_ctx* ctx, BzState* st, const BzVecs& v, const BzTeam& tm,
                    unsigned int* counter, int force_role, bool slab) {
    const long long per_cta = (long long)LC_THREADS * LC_EPT;
    static BzOncePerDevice once;
    if (once.first(ctx->device)) {   // same carve-out as the dense passes: no SM reconfiguration inside the cycle
        cudaFuncSetAttribute(k_lstep_cluster<8>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_lstep_cluster<LC_EPT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_lstep, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    if (v.n <= per_cta * LC_MAXCS && !slab) {   // (slab mode: only the cooperative kernel exchanges across ranks)
        // smallest power-of-two cluster that keeps <= 8 elements per thread when possible
        int cs = 1;
        while (cs < LC_MAXCS && (long long)cs * LC_THREADS * 8 < v.n) cs *= 2;
        if ((long long)cs * LC_THREADS * 8 >= v.n)
            return bz_launch_pdl(ctx, k_lstep_cluster<8>, dim3(cs), dim3(LC_THREADS), 0, cs, st, v, force_role);
        return bz_launch_pdl(ctx, k_lstep_cluster<LC_EPT>, dim3(cs), dim3(LC_THREADS), 0, cs, st, v, force_role);
    }
    const int nb = bz_l_grid(ctx, v.n);
    BzVecs vv = v;
    BzTeam tt = tm;
    void* args[] = {(void*)&st, (void*)&vv, (void*)&tt, (void*)&counter, (void*)&force_role};
    BZ_CUDA(ctx, cudaLaunchCooperativeKernel((void*)k_lstep, dim3(nb), dim3(BZ_L_THREADS), args, 0,
                                             ctx->stream));
    ctx->n_launches++;
    return BZ_OK;
}

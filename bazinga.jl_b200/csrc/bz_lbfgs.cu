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
    double* scratch;   // >= 3*32 + 3 doubles of shared memory (block sums of K <= 3 values, then the team result)
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
        double* sl = tc.tm.slots + ((size_t)(tc.tm.parity * BZ_MAXBAR + j) * tc.tm.max_ctas) * 4;
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) sl[blockIdx.x * 4 + k] = v[k];
            __threadfence();
            atomicAdd(bar, 1u);
            while (ld_acquire_u32(bar) < (unsigned int)nb) { }
        }
        __syncthreads();
        if (threadIdx.x < 32) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                double s = 0.0;
                for (int c = threadIdx.x; c < nb; c += 32) s += __ldcg(sl + c * 4 + k);
                s = bz_warp_sum(s);
                if (threadIdx.x == 0) tc.scratch[96 + k] = s;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = tc.scratch[96 + k];
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
                tc.scratch[96 + k] = acc;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = tc.scratch[96 + k];
        __syncthreads();
    }
}

// V entries per access: double2 when n is even (every vector of the state is 256-byte aligned and the history slots
// are n entries apart), plain doubles otherwise
template <int V> struct Pk { double v[V]; };
template <int V> __device__ __forceinline__ Pk<V> ldv(const double* p, long long j);
template <> __device__ __forceinline__ Pk<1> ldv<1>(const double* p, long long j) { Pk<1> r; r.v[0] = p[j]; return r; }
template <> __device__ __forceinline__ Pk<2> ldv<2>(const double* p, long long j) {
    const double2 t = reinterpret_cast<const double2*>(p)[j];
    Pk<2> r; r.v[0] = t.x; r.v[1] = t.y; return r;
}
template <int V> __device__ __forceinline__ void stv(double* p, long long j, const Pk<V>& a);
template <> __device__ __forceinline__ void stv<1>(double* p, long long j, const Pk<1>& a) { p[j] = a.v[0]; }
template <> __device__ __forceinline__ void stv<2>(double* p, long long j, const Pk<2>& a) {
    reinterpret_cast<double2*>(p)[j] = make_double2(a.v[0], a.v[1]);
}

// Passes over the n-vectors per new iterate with memory M (all 2M+1 dot products stay sequential, in upstream's order):
//   update      read x, x_prev, res, res_prev; write the candidate pair (s, y) into the SPARE history slot; the same
//               pass forms <s,y>, <y,y> and <s,-res> -- the first dot product of loop1 if the pair is accepted      6
//   loop1, 1st  d = -res - alpha*y_new is formed from res (d is never initialised in memory), dot with the next s   4
//   2M-2 steps  read d, the axpy vector, the next dot vector; write d                                               4 each
//   last step   d + coef*s fused with  x_prev <- x, res_prev <- res, x_d = x <- x + d, z_curr <- z                 10
// = 52 at M = 5 (61 before: the pair was recomputed in a second pass, d = -res took a pass of its own, and the last
// axpy was separate from the end-of-iteration copies).  A rejected pair (<s,y> <= 0) leaves the ring untouched: the
// physical slots are addressed through BzState::slot[], the spare one is swapped in only on acceptance.
// rot != -1: buffer rotation (fused problems, bz_obfused.cu).  The roles (x, x_prev), (res, res_prev), (z, z_curr) are
// swapped instead of copied at the end of the step: the last pass writes only x_d and the new x (5 vector passes instead
// of 10, 47 per step at M = 5); BzState::rot records which physical buffer plays which role.
template <int V>
__device__ __forceinline__ void lstep_newiter(BzState* st, const BzVecs& v, TeamCtx& tc, unsigned int* counter,
                                               int force_role, int rot) {
    const bool rotate = rot >= 0;
    const int rb_ = rotate ? rot : 0;
    const long long n = v.n, nv = n / V;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long j0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int M = __ldcg(&st->mem);
    int currmem = __ldcg(&st->currmem), curridx = __ldcg(&st->curridx);
    double H = __ldcg(&st->H);
    const int skip = __ldcg(&st->skip_update);
    double ys[BZ_MAXMEM], al[BZ_MAXMEM];
    int slot[BZ_MAXMEM];
    for (int i = 0; i < BZ_MAXMEM; ++i) { ys[i] = __ldcg(&st->ys[i]); al[i] = 0.0; slot[i] = __ldcg(&st->slot[i]); }
    int spare = __ldcg(&st->spare);
    double* const x = (rb_ & 1) ? v.x_prev : v.x;
    double* const x_prev_buf = (rb_ & 1) ? v.x : v.x_prev;
    const double* __restrict__ res = (rb_ & 2) ? v.res_prev : v.res;
    const double* const res_prev_buf = (rb_ & 2) ? v.res : v.res_prev;
    const double* __restrict__ zc = (rb_ & 4) ? v.z_curr : v.z;
    double* __restrict__ d = v.d;

    // ---- update!(H, x - x_prev, res - res_prev)
    bool have_dot0 = false;
    double dot0 = 0.0;
    if (!skip && M > 0) {
        const double* __restrict__ xp = x_prev_buf;
        const double* __restrict__ rp = res_prev_buf;
        double* __restrict__ sn = v.sM + (size_t)spare * n;
        double* __restrict__ yn = v.yM + (size_t)spare * n;
        double a[3] = {0.0, 0.0, 0.0};
        for (long long j = j0; j < nv; j += stride) {
            const Pk<V> xa = ldv<V>(x, j), xb = ldv<V>(xp, j), ra = ldv<V>(res, j), rb = ldv<V>(rp, j);
            Pk<V> sv, yv;
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const double s_ = xa.v[e] - xb.v[e], y_ = ra.v[e] - rb.v[e];
                sv.v[e] = s_; yv.v[e] = y_;
                a[0] += s_ * y_;
                a[1] += y_ * y_;
                a[2] += s_ * (-ra.v[e]);
            }
            stv<V>(sn, j, sv);
            stv<V>(yn, j, yv);
        }
        team_sum<3>(tc, a);
        if (a[0] > 0) {
            curridx += 1;
            if (curridx > M) curridx = 1;
            currmem = currmem + 1 > M ? M : currmem + 1;
            ys[curridx - 1] = a[0];
            const int old = slot[curridx - 1];
            slot[curridx - 1] = spare;
            spare = old;
            H = a[0] / a[1];
            dot0 = a[2];
            have_dot0 = true;
        }
    }
    auto SM = [&](int ring1) { return v.sM + (size_t)slot[ring1 - 1] * n; };   // 1-based ring position -> history vector
    auto YM = [&](int ring1) { return v.yM + (size_t)slot[ring1 - 1] * n; };
    const bool write_d = force_role >= 0;   // the parity entry point bz_lbfgs_apply reads d; the cycle only needs x_d

    // last pass: d_final = dsrc + coef*sadd (or H*(-res) without history), fused with the end-of-iteration copies
    auto final_pass = [&](const double* __restrict__ dsrc, const double* __restrict__ sadd, double coef, bool from_res) {
        double* __restrict__ xdw = v.x_d;
        if (rotate) {      // the new x goes where x_prev was; x, res, z stay in place and change roles
            double* __restrict__ xnew = x_prev_buf;
            for (long long j = j0; j < nv; j += stride) {
                const Pk<V> xa = ldv<V>(x, j);
                Pk<V> da, sa, ra;
                if (!from_res) { da = ldv<V>(dsrc, j); sa = ldv<V>(sadd, j); } else ra = ldv<V>(res, j);
                Pk<V> xd, dn;
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    dn.v[e] = from_res ? H * (-ra.v[e]) : da.v[e] + coef * sa.v[e];
                    xd.v[e] = xa.v[e] + dn.v[e];
                }
                stv<V>(xdw, j, xd);
                stv<V>(xnew, j, xd);
                if (write_d) stv<V>(d, j, dn);
            }
            return;
        }
        double* __restrict__ xpw = v.x_prev;
        double* __restrict__ rpw = v.res_prev;
        double* __restrict__ zcw = v.z_curr;
        for (long long j = j0; j < nv; j += stride) {
            const Pk<V> xa = ldv<V>(x, j), ra = ldv<V>(res, j), za = ldv<V>(zc, j);
            Pk<V> da, sa;
            if (!from_res) { da = ldv<V>(dsrc, j); sa = ldv<V>(sadd, j); }
            Pk<V> xd, dn;
#pragma unroll
            for (int e = 0; e < V; ++e) {
                dn.v[e] = from_res ? H * (-ra.v[e]) : da.v[e] + coef * sa.v[e];
                xd.v[e] = xa.v[e] + dn.v[e];
            }
            stv<V>(xpw, j, xa);
            stv<V>(rpw, j, ra);
            stv<V>(xdw, j, xd);
            stv<V>(x, j, xd);
            stv<V>(zcw, j, za);
            if (write_d) stv<V>(d, j, dn);
        }
    };

    if (currmem == 0) {
        final_pass(nullptr, nullptr, 0.0, true);
    } else {
        // ---- d = H * (-res): two-loop recursion, each axpy fused with the next dot product
        int idx = curridx;
        double a[1] = {0.0};
        if (!have_dot0) {                     // no (accepted) update in this call: <s_newest, -res> needs its own pass
            const double* __restrict__ sm = SM(idx);
            for (long long j = j0; j < nv; j += stride) {
                const Pk<V> ra = ldv<V>(res, j), sa = ldv<V>(sm, j);
#pragma unroll
                for (int e = 0; e < V; ++e) a[0] += sa.v[e] * (-ra.v[e]);
            }
        }
        for (int t = 0; t < currmem; ++t) {       // loop1!: newest -> oldest
            if (t == 0 && have_dot0) a[0] = dot0; else team_sum<1>(tc, a);
            const int i = idx - 1;
            const double alpha_i = a[0] / ys[i];
            al[i] = alpha_i;
            int idx2 = idx - 1;
            if (idx2 == 0) idx2 = M;
            const double* __restrict__ ym = YM(idx);
            const bool last = t == currmem - 1;
            int idn = idx2 + 1;
            if (idn > M) idn = 1;
            const double* __restrict__ nxt = last ? YM(idn) : SM(idx2);   // last: d .= H .* d, then the first dot of loop2!
            const double* __restrict__ dsrc = t == 0 ? res : d;            // t = 0: d = -res, read straight from res
            a[0] = 0.0;
            // two trips per iteration: six loads in flight per thread (ncu r2i: the passes wait on global loads with too few
            // bytes in flight -- 0.81 of the copy peak); the accumulation order (j ascending) is unchanged
            long long j = j0;
            for (; j + stride < nv; j += 2 * stride) {
                const Pk<V> da = ldv<V>(dsrc, j), ya = ldv<V>(ym, j), na = ldv<V>(nxt, j);
                const Pk<V> db = ldv<V>(dsrc, j + stride), yb = ldv<V>(ym, j + stride), nb = ldv<V>(nxt, j + stride);
                Pk<V> dn, dm;
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    double dj = (t == 0 ? -da.v[e] : da.v[e]) - alpha_i * ya.v[e];
                    if (last) dj = H * dj;
                    dn.v[e] = dj;
                    a[0] += na.v[e] * dj;
                }
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    double dj = (t == 0 ? -db.v[e] : db.v[e]) - alpha_i * yb.v[e];
                    if (last) dj = H * dj;
                    dm.v[e] = dj;
                    a[0] += nb.v[e] * dj;
                }
                stv<V>(d, j, dn);
                stv<V>(d, j + stride, dm);
            }
            for (; j < nv; j += stride) {
                const Pk<V> da = ldv<V>(dsrc, j), ya = ldv<V>(ym, j), na = ldv<V>(nxt, j);
                Pk<V> dn;
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    double dj = (t == 0 ? -da.v[e] : da.v[e]) - alpha_i * ya.v[e];
                    if (last) dj = H * dj;
                    dn.v[e] = dj;
                    a[0] += na.v[e] * dj;
                }
                stv<V>(d, j, dn);
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
            const double* __restrict__ sm = SM(idx);
            a[0] = 0.0;
            if (t < currmem - 1) {
                int idn = idx + 1;
                if (idn > M) idn = 1;
                const double* __restrict__ ymn = YM(idn);
                long long j = j0;
                for (; j + stride < nv; j += 2 * stride) {
                    const Pk<V> da = ldv<V>(d, j), sa = ldv<V>(sm, j), na = ldv<V>(ymn, j);
                    const Pk<V> db = ldv<V>(d, j + stride), sb = ldv<V>(sm, j + stride), nb = ldv<V>(ymn, j + stride);
                    Pk<V> dn, dm;
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        const double dj = da.v[e] + coef * sa.v[e];
                        dn.v[e] = dj;
                        a[0] += na.v[e] * dj;
                    }
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        const double dj = db.v[e] + coef * sb.v[e];
                        dm.v[e] = dj;
                        a[0] += nb.v[e] * dj;
                    }
                    stv<V>(d, j, dn);
                    stv<V>(d, j + stride, dm);
                }
                for (; j < nv; j += stride) {
                    const Pk<V> da = ldv<V>(d, j), sa = ldv<V>(sm, j), na = ldv<V>(ymn, j);
                    Pk<V> dn;
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        const double dj = da.v[e] + coef * sa.v[e];
                        dn.v[e] = dj;
                        a[0] += na.v[e] * dj;
                    }
                    stv<V>(d, j, dn);
                }
            } else {
                final_pass(d, sm, coef, false);
            }
        }
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
        for (int i = 0; i < BZ_MAXMEM; ++i) { st->ys[i] = ys[i]; st->alphas[i] = al[i]; st->slot[i] = slot[i]; }
        st->spare = spare;
        if (rotate) st->rot = rb_ ^ 7;          // x <-> x_prev, res <-> res_prev, z <-> z_curr
        st->skip_update = 0;
        st->grad_x_fresh = 1;
        bz_goto_fb(st);
        if (st->xr.nranks > 1) *st->xr.seq = tc.seq0 + (unsigned long long)tc.xidx;
        if (force_role < 0) bz_toc(st, 1);
        __threadfence();
    }
}

__global__ void __launch_bounds__(BZ_L_THREADS, BZ_L_CTAS_PER_SM)
    k_lstep(BzState* st, BzVecs v, BzTeam tm, unsigned int* counter, int force_role, int rotate) {
    __shared__ double scratch[3 * 32 + 4];
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
    const bool vec2 = (n % 2) == 0;

    if (role == LROLE_INTERP) {      // x = tau*x_d + (1-tau)*z_curr   (A.5)
        const double tau = __ldcg(&st->tau);
        const long long stride = (long long)gridDim.x * blockDim.x;
        const long long j0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        const int rot = rotate ? __ldcg(&st->rot) : 0;
        const double* __restrict__ xd = v.x_d;
        const double* __restrict__ zc = (rot & 4) ? v.z : v.z_curr;
        double* __restrict__ x = (rot & 1) ? v.x_prev : v.x;
        if (vec2) {
            for (long long j = j0; j < n / 2; j += stride) {
                const Pk<2> a = ldv<2>(xd, j), b = ldv<2>(zc, j);
                Pk<2> o;
                o.v[0] = tau * a.v[0] + (1 - tau) * b.v[0];
                o.v[1] = tau * a.v[1] + (1 - tau) * b.v[1];
                stv<2>(x, j, o);
            }
        } else {
            for (long long j = j0; j < n; j += stride) x[j] = tau * xd[j] + (1 - tau) * zc[j];
        }
        if (bz_last_cta(counter, gridDim.x) && threadIdx.x == 0) {
            st->grad_x_fresh = 1;
            bz_goto_fb(st);
            if (force_role < 0) bz_toc(st, 1);
            __threadfence();
        }
        return;
    }
    TeamCtx tc{tm, 0, scratch, st, st->xr.nranks > 1 ? *st->xr.seq : 0ULL, 0};
    const int rot = rotate ? __ldcg(&st->rot) : -1;
    if (vec2) lstep_newiter<2>(st, v, tc, counter, force_role, rot);
    else lstep_newiter<1>(st, v, tc, counter, force_role, rot);
}


// ---------------------------------------------------------------------------------------------
// cluster version
// ---------------------------------------------------------------------------------------------
#define LC_THREADS 512     // (the speculative copy of this kernel, bz_lspec.cuh, must use the same numbers)
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
    double* wpart;     // [2 parities][K <= 2][16] warp partials
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
    double* wp = cr.wpart + par * 32;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) wp[k * 16 + warp] = v[k];
    }
    __syncthreads();
    double* sl = cr.slots + par * (LC_MAXCS * 2);
    uint64_t* bar = cr.bars + par;
    if (threadIdx.x < cr.cs) {     // thread t sends this CTA's block sums to CTA t
        if (threadIdx.x == 0) lc_mbar_expect_tx(bar, cr.cs * K * 8u);
        double bs[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const double* w = wp + k * 16;
            bs[k] = (((w[0] + w[1]) + (w[2] + w[3])) + ((w[4] + w[5]) + (w[6] + w[7]))) +
                    (((w[8] + w[9]) + (w[10] + w[11])) + ((w[12] + w[13]) + (w[14] + w[15])));
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
    __shared__ double wpart[2 * 2 * 16];
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
            st->grad_x_fresh = 1;
            bz_goto_fb(st);
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
        st->grad_x_fresh = 1;
        bz_goto_fb(st);
        if (force_role < 0) bz_toc(st, 1);
        __threadfence();
    }
#undef LC_FOR
#undef LC_J
}

int bz_launch_lstep(bz_ctx* ctx, BzState* st, const BzVecs& v, const BzTeam& tm,
                    unsigned int* counter, int force_role, bool slab, bool rotate) {
    const long long per_cta = (long long)LC_THREADS * LC_EPT;
    static BzOncePerDevice once;
    if (once.first(ctx->device)) {   // same carve-out as the dense passes: no SM reconfiguration inside the cycle
        cudaFuncSetAttribute(k_lstep_cluster<8>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_lstep_cluster<LC_EPT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_lstep, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    if (v.n <= per_cta * LC_MAXCS && !slab && !rotate) {   // (slab mode: only the cooperative kernel exchanges across ranks)
        // smallest power-of-two cluster that keeps <= 8 elements per thread when possible
        int cs = 1;
        while (cs < LC_MAXCS && (long long)cs * LC_THREADS * 8 < v.n) cs *= 2;      // n = 16384: four CTAs of 512 threads
        if ((long long)cs * LC_THREADS * 8 >= v.n)
            return bz_launch_pdl(ctx, k_lstep_cluster<8>, dim3(cs), dim3(LC_THREADS), 0, cs, st, v, force_role);
        return bz_launch_pdl(ctx, k_lstep_cluster<LC_EPT>, dim3(cs), dim3(LC_THREADS), 0, cs, st, v, force_role);
    }
    const int nb = bz_l_grid(ctx, v.n);
    BzVecs vv = v;
    BzTeam tt = tm;
    int rot_arg = rotate ? 1 : 0;
    void* args[] = {(void*)&st, (void*)&vv, (void*)&tt, (void*)&counter, (void*)&force_role, (void*)&rot_arg};
    BZ_CUDA(ctx, cudaLaunchCooperativeKernel((void*)k_lstep, dim3(nb), dim3(BZ_L_THREADS), args, 0,
                                             ctx->stream));
    ctx->n_launches++;
    return BZ_OK;
}

// Speculative L step, hidden behind the dense AL-gradient at z (config C2: 23.5 of every 242 us were the L kernel).
//
// A PANOC+ iteration is the chain  L -> G(x) -> F -> G(z) -> D  and the next L needs nothing that G(z) or D compute
// except D's verdict: its inputs (x, x_prev, res, res_prev, the L-BFGS history, gamma and the scalars F left) are all
// final once F has run.  So while 72 row clusters of k_dense_algrad stream A for G(z), two spare clusters of the
// same launch (4 CTAs at n = 16384) run the L step of the NEXT iteration on the
// assumption that D will accept the iterate, writing into shadow storage only:
//     x + d              -> BzLSpec::xs
//     (s, y) candidate   -> the spare history slot BZ_MAXMEM
//     currmem, curridx, H, ys[], alphas[], pair accepted or not -> BzState::sp_*
// D sets spec_commit when it ends the iteration on this very evaluation (99.8 % of C2's iterations); the L kernel then
// only copies the shadow into place.  Any other verdict (gamma halving, tau back-track, done) drops the shadow and the
// plain L kernel runs.
//
// The arithmetic is that of k_lstep_cluster statement by statement -- same element -> thread mapping (cs CTAs x 256
// threads x EPT elements), same reduction trees, explicit _rn intrinsics because this header is compiled into
// bz_dense.cu, where contraction is allowed -- so an iteration gives the same bits whether it was speculated or not.
// The cs CTAs exchange their block sums through global memory (tagged, self-validating words: no counters to reset).
#pragma once
#include "bz_internal.cuh"
#include "bz_vec.cuh"

#define LS_THREADS 512     // = LC_THREADS of bz_lbfgs.cu
#define LS_MAXCS 8

struct BzLSpec {
    int enabled;               // this launch carries the speculative L step (G(z) of a dense, n <= 16384 problem)
    int dense_clusters;        // row clusters of the dense pass; clusters beyond them run the L step
    int cs;                    // CTAs of the L step (the cluster size k_lstep_cluster would use for this n)
    BzVecs v;
    double* xs;                // shadow x + d
    unsigned long long* xbuf;  // exchange: [32 rounds][LS_MAXCS][2 values][2 words]
};

__device__ __forceinline__ void ls_post(unsigned long long* slot, unsigned long long tag, double v) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"(tag | (bits & 0xffffffffULL)),
                 "l"(tag | (bits >> 32)) : "memory");
}
// the wait for a sibling CTA is bounded (a device shared with another tenant may never make the whole grid resident):
// after ~4 s the speculation is abandoned, err_flag = 4 stops the state machine and the host reports BZ_ERR_CUDA
__device__ __forceinline__ double ls_take(BzState* st, const unsigned long long* slot, unsigned long long tag) {
    unsigned long long w0, w1;
    unsigned int spins = 0;
    const unsigned long long t0 = bz_globaltimer();
    for (;;) {
        asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(slot) : "memory");
        if ((((w0 ^ tag) | (w1 ^ tag)) >> 32) == 0) break;
        if ((++spins & 1023u) == 0u && (bz_globaltimer() - t0 > 4000000000ULL || *(volatile int*)&st->err_flag == 4)) {
            st->err_flag = 4;
            return __longlong_as_double(0x7ff8000000000000LL);
        }
    }
    return __longlong_as_double((long long)((w0 & 0xffffffffULL) | (w1 << 32)));
}

// sum of K values per thread over the cs CTAs; the trees of cluster_sum (bz_lbfgs.cu): warp butterfly, eight warp sums
// pairwise, CTA sums in rank order
template <int K>
__device__ inline void ls_sum(BzState* st, double (&v)[K], double* wpart, unsigned long long* xbuf, unsigned long long seq,
                              int& round, int rank, int cs) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = bz_warp_sum(v[k]);
    double* wp = wpart + (round & 1) * 32;
    if (threadIdx.x < LS_THREADS && lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) wp[k * 16 + warp] = v[k];
    }
    __syncthreads();
    const unsigned long long tag = ((seq + (unsigned long long)round + 1ULL) & 0xffffffffULL) << 32;
    unsigned long long* base = xbuf + (size_t)(round & 31) * (LS_MAXCS * 2 * 2);
    if (threadIdx.x < K) {
        const double* w = wp + threadIdx.x * 16;
        const double bs = (((w[0] + w[1]) + (w[2] + w[3])) + ((w[4] + w[5]) + (w[6] + w[7]))) +
                          (((w[8] + w[9]) + (w[10] + w[11])) + ((w[12] + w[13]) + (w[14] + w[15])));
        ls_post(base + ((size_t)rank * 2 + threadIdx.x) * 2, tag, bs);
    }
    double s[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double acc = 0.0;
        for (int c = 0; c < cs; ++c) acc += ls_take(st, base + ((size_t)c * 2 + k) * 2, tag);
        s[k] = acc;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = s[k];
    round++;
    __syncthreads();     // wpart[(round) & 1] is reused two reductions later
}

// the body run by CTA `rank` of the cs speculating CTAs (blockDim may exceed LS_THREADS: the extra threads only take
// part in the barriers).  Mirrors k_lstep_cluster<EPT> in its NEWITER role.
template <int EPT>
__device__ inline void ls_body(BzState* st, const BzLSpec& ls, int rank, int cs, double* wpart) {
    const BzVecs& v = ls.v;
    const long long n = v.n;
    const bool worker = threadIdx.x < LS_THREADS;
    const int stride = cs * LS_THREADS;
    const int j0 = rank * LS_THREADS + (int)threadIdx.x;
    const unsigned long long seq = __ldcg(&st->lspec_seq);
    int round = 0;
#define LS_FOR(k) _Pragma("unroll") for (int k = 0; k < EPT; ++k)
#define LS_J(k) ((long long)j0 + (long long)(k) * stride)
#define LS_OK(k) (worker && LS_J(k) < n)
    const int M = __ldcg(&st->mem);
    int currmem = __ldcg(&st->currmem), curridx = __ldcg(&st->curridx);
    double H = __ldcg(&st->H);
    const int skip = (__ldcg(&st->mode) == MODE_INITFB) ? 1 : 0;   // the first L of an inner solve makes no update (A.4 -> A.5)
    double ys[BZ_MAXMEM], al[BZ_MAXMEM];
    for (int i = 0; i < BZ_MAXMEM; ++i) { ys[i] = __ldcg(&st->ys[i]); al[i] = 0.0; }

    double xv[EPT], rv[EPT], d[EPT];
    LS_FOR(k) {
        const long long j = LS_J(k);
        xv[k] = LS_OK(k) ? v.x[j] : 0.0;
        rv[k] = LS_OK(k) ? v.res[j] : 0.0;
    }
    // ---- update!(H, x - x_prev, res - res_prev): the candidate pair goes to the spare slot; on acceptance the ring
    // slot `pair_slot` is what the plain kernel would have written -- here it is only READ through `cand` below
    int pair_accepted = 0, pair_slot = -1;
    double* sc = v.sM + (size_t)BZ_MAXMEM * n;
    double* yc = v.yM + (size_t)BZ_MAXMEM * n;
    if (!skip && M > 0) {
        double sv[EPT], yv[EPT];
        double a[2] = {0.0, 0.0};
        LS_FOR(k) {
            const long long j = LS_J(k);
            sv[k] = LS_OK(k) ? __dsub_rn(xv[k], v.x_prev[j]) : 0.0;
            yv[k] = LS_OK(k) ? __dsub_rn(rv[k], v.res_prev[j]) : 0.0;
        }
        LS_FOR(k) { a[0] = __dadd_rn(a[0], __dmul_rn(sv[k], yv[k])); a[1] = __dadd_rn(a[1], __dmul_rn(yv[k], yv[k])); }
        ls_sum<2>(st, a, wpart, ls.xbuf, seq, round, rank, cs);
        if (a[0] > 0) {
            curridx += 1;
            if (curridx > M) curridx = 1;
            currmem = currmem + 1 > M ? M : currmem + 1;
            ys[curridx - 1] = a[0];
            pair_accepted = 1;
            pair_slot = curridx - 1;
            LS_FOR(k) {
                const long long j = LS_J(k);
                if (LS_OK(k)) { sc[j] = sv[k]; yc[j] = yv[k]; }
            }
            H = __ddiv_rn(a[0], a[1]);
        }
        // the two-loop below re-reads the candidate through global memory like any other pair: make it visible CTA-wide
        __syncthreads();
    }
    // ---- d = H * (-res): two-loop recursion, d in registers (see k_lstep_cluster)
    LS_FOR(k) d[k] = -rv[k];
    const int nsteps = 2 * currmem;
    auto hist_index = [&](int kk) {
        const int k1 = kk < currmem ? kk : nsteps - 1 - kk;
        return ((curridx - 1 - k1) % M + M) % M;
    };
    auto load_step = [&](int kk, double (&A)[EPT], double (&B)[EPT]) {
        if (kk >= nsteps) return;
        const int i = hist_index(kk);
        const bool cand = (i == pair_slot);           // the pair accepted a moment ago lives in the spare slot
        const double* sbase = cand ? sc : v.sM + (size_t)i * n;
        const double* ybase = cand ? yc : v.yM + (size_t)i * n;
        const double* pa = kk < currmem ? sbase : ybase;
        const double* pb = kk < currmem ? ybase : sbase;
        LS_FOR(e) {
            const long long j = LS_J(e);
            A[e] = LS_OK(e) ? __ldcg(pa + j) : 0.0;
            B[e] = LS_OK(e) ? __ldcg(pb + j) : 0.0;
        }
    };
    double ra[EPT], rb[EPT];
    load_step(0, ra, rb);
    for (int kk = 0; kk < nsteps; ++kk) {
        if (kk == currmem) { LS_FOR(e) d[e] = __dmul_rn(H, d[e]); }
        double a[1] = {0.0};
        LS_FOR(e) a[0] = __dadd_rn(a[0], __dmul_rn(ra[e], d[e]));
        double na[EPT], nb[EPT];
        LS_FOR(e) { na[e] = 0.0; nb[e] = 0.0; }
        load_step(kk + 1, na, nb);
        ls_sum<1>(st, a, wpart, ls.xbuf, seq, round, rank, cs);
        const int i = hist_index(kk);
        if (kk < currmem) {
            const double alpha_i = __ddiv_rn(a[0], ys[i]);
            al[i] = alpha_i;
            LS_FOR(e) d[e] = __dsub_rn(d[e], __dmul_rn(alpha_i, rb[e]));
        } else {
            const double beta = __ddiv_rn(a[0], ys[i]);
            const double coef = __dsub_rn(al[i], beta);
            LS_FOR(e) d[e] = __dadd_rn(d[e], __dmul_rn(coef, rb[e]));
        }
        LS_FOR(e) { ra[e] = na[e]; rb[e] = nb[e]; }
    }
    if (currmem == 0) { LS_FOR(e) d[e] = __dmul_rn(H, d[e]); }
    LS_FOR(k) {
        const long long j = LS_J(k);
        if (LS_OK(k)) ls.xs[j] = __dadd_rn(xv[k], d[k]);
    }
    if (rank == 0 && threadIdx.x == 0) {
        st->sp_currmem = currmem; st->sp_curridx = curridx; st->sp_H = H;
        st->sp_pair = pair_accepted;
        for (int i = 0; i < BZ_MAXMEM; ++i) { st->sp_ys[i] = ys[i]; st->sp_alphas[i] = al[i]; }
        st->spec_ok = (*(volatile int*)&st->err_flag == 4) ? 0 : 1;
    }
#undef LS_FOR
#undef LS_J
#undef LS_OK
}

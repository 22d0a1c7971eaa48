// Single-pass fused AL-gradient for WIDE dense A (nx > 16384; config C5, 32768 x 262144 row-sharded):
//     r = A x - b;  t = r + mu.*y;  s = proj_D(t);  w = (t - s)./mu;  g = A' w      (auglagfun.jl:73-86
//     around demo/basispursuit.jl:38-49), A read from HBM exactly ONCE.
//
// A row of C5 is 2 MB: no cluster can hold it, but the whole GPU can hold ~16 of them in shared
// memory.  So the split is by COLUMNS over the whole grid: CTA c (one per SM, all co-resident) owns
// columns [c*W, c*W + W) of every row and streams its slice of each row (W*8 bytes, one cp.async.bulk
// per row) through a ring of shared-memory slots.  For row j every CTA forms its slice dot product,
// the 148 slice dots are exchanged through an L2-resident ring in global memory, every CTA adds them
// in the same fixed order (bitwise identical w everywhere, no float atomics), and the row -- still
// sitting in shared memory -- is accumulated into the CTA's column sums.  The ring (16 rows per SM at
// C5's width, ~5 us of HBM time) is what hides the grid-wide exchange (~2.3 us).
//
// Warp roles; everything hands over through mbarriers, there is no __syncthreads in the row loop:
//   CH dot warps       slice dot of each row with the x slice in registers, as soon as the row lands
//   CH update warps    acc += w_j * A[j, slice] once w_j is published, then release the slot; they
//                      are decoupled from the dot warps, so a late w never delays the posting of
//                      later rows (a fused dot+update loop with a fixed skew ran at 4.7 TB/s, this
//                      split at 6.1 TB/s)
//   XW exchange warps  row j is handled by exchange warp j % XW: sums the CH warp partials, posts the
//                      CTA's slice dot, gathers all G slice dots, runs the constraint-space
//                      elementwise stage and publishes w_j
//   1 producer warp    one lane issues the bulk copies as slots are released
//
// Exchange format: D[(seq0 + j) % GR_DRING][cta] is two 64-bit words {tag:32 | value bits 31..0},
// {tag:32 | value bits 63..32}; tag = low 32 bits of a sequence number that is unique per (launch, row).
// Each aligned 8-byte word is single-copy atomic and validates itself, so writer and readers need
// only relaxed gpu-scope accesses: no fences on the critical path.  Sharded per-position arrival
// counters (red.relaxed, monotonic over the life of the problem) tell the gathering warps when a
// round of loads is worth issuing; they are a hint, not a synchronisation.
//
// Co-residency: the grid is <= the number of SMs at 1 CTA/SM and CTAs spin on global memory only after
// griddepcontrol.wait has returned (the predecessor grid has completed), so every CTA of this grid can
// be scheduled; dependents are released only once every CTA has started.  This holds on a device the process
// has to itself (the stated requirement, include/bazinga_b200.h); if another tenant (a second context, MPS)
// holds SMs, part of a team may never start: every cross-CTA wait is bounded by GR_TIMEOUT_NS, after which
// the launch drains with BzState::err_flag = 4 and the solve returns BZ_ERR_CUDA instead of hanging.
#include "bz_internal.cuh"
#include "bz_dense.cuh"
#include "bz_ptx.cuh"

#define GR_MAXSLOTS 24
#define GR_PRING 32         // ring of per-row warp-partial sets / published w values (>= GR_MAXSLOTS)
#define GR_DRING 256        // rows of the global exchange ring
#define GR_RING_DOUBLES 28624   // 223.6 KB of row slices (16 slots of the C5 slice width 1772)
#define GR_TIMEOUT_NS 4000000000ULL   // bound on every cross-CTA wait of the exchange (see the exchange warps)
#define GR_SHARDS 16        // arrival-counter shards per ring position (same-address L2 atomics serialise at ~27 cycles each)
#define GR_SMEM_BYTES (GR_RING_DOUBLES * 8 + 2 * GR_MAXSLOTS * 8 + 2 * GR_PRING * 8 + GR_PRING * 8 * 8 + GR_PRING * 8 + 16 * 8)

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void st_relaxed_v2(unsigned long long* p, unsigned long long a,
                                              unsigned long long b) {
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void ld_relaxed_v2(const unsigned long long* p, unsigned long long& a,
                                              unsigned long long& b) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}


// CH "dot" warps + CH "update" warps (thread t of either group owns columns {64 CH c + 2t, +1}, c < NCH),
// XW exchange warps, one producer warp; RB rows per loop trip of the dot / update warps
template <int CH, int NCH, int XW, int RB>
__global__ void __launch_bounds__((2 * CH + XW + 1) * 32, 1)
    k_grid_algrad(BzDenseArgs a, int W, int nslots, int GS, int nshards) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* ring = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + GR_RING_DOUBLES);   // [GR_MAXSLOTS] TMA landed
    uint64_t* empty = full + GR_MAXSLOTS;                                    // [GR_MAXSLOTS] column update done
    uint64_t* pbar = empty + GR_MAXSLOTS;                                    // [GR_PRING] CH warp partials in
    uint64_t* wbar = pbar + GR_PRING;                                        // [GR_PRING] w published
    double* wsum = reinterpret_cast<double*>(wbar + GR_PRING);               // [GR_PRING][8]
    double* wring = wsum + GR_PRING * 8;                                     // [GR_PRING]
    double* lxs = wring + GR_PRING;                                          // [16] lx of the exchange warps

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    bz_pdl_wait();
    bz_pdl_launch_dependents();
    int target = a.force_target;
    if (target < 0) {
        if (__ldcg(&a.st->next_op) != OP_G) return;   // uniform over the grid
        target = __ldcg(&a.st->grad_target);
        if (blockIdx.x == 0 && tid == 0 && a.set_after) bz_tic(a.st, 0);
    }
    const double* __restrict__ xsrc = a.src[target];

    // team geometry: GS CTAs share the rows [r0, r0 + m) and split their columns; G = GS, cta = rank in team
    const int G = GS, team = blockIdx.x / GS, cta = blockIdx.x % GS;
    // the team's row sequence base: every team consumes consecutive sequence numbers (its own row count per launch),
    // so ring positions, tags and arrival counts stay in step even when the last team has fewer rows
    const unsigned long long seq0 = __ldcg(a.gseq + team);
    const long long n = a.n;
    const long long r0 = (long long)team * a.rows_per_cluster;
    const long long mrem = a.m - r0;
    const int m = (int)(mrem < 0 ? 0 : (mrem < a.rows_per_cluster ? mrem : a.rows_per_cluster));
    const long long col0 = (long long)cta * W;
    const int Wl = (int)((n - col0) < W ? (n - col0) : W);    // local slice width (even, > 0)
    const int stride = W;                                      // slot stride in doubles (even)
    const uint32_t row_bytes = (uint32_t)Wl * 8u;
    unsigned long long* const Dteam = a.D + (size_t)team * ((size_t)GR_DRING * GS * 2);   // this team's exchange ring
    // this team's arrival counters: GR_DRING positions x nshards counters, 256 B apart
    unsigned long long* const Dcnt = a.D + (size_t)GR_DRING * 160 * 2 + (size_t)team * ((size_t)GR_DRING * nshards * 32);

    if (tid == 0) {
        for (int s = 0; s < GR_MAXSLOTS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CH); }
        for (int s = 0; s < GR_PRING; ++s) { mbar_init(&pbar[s], CH); mbar_init(&wbar[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp < CH) {
        // ------------------------------------------------------------------ dot warps ("phase A")
        // slice dot of every row with the x slice held in registers, as soon as the row has landed
        const int lc0 = 2 * tid;
        double2 xr[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int lc = c * (64 * CH) + lc0;
            xr[c] = (lc < Wl) ? *reinterpret_cast<const double2*>(xsrc + col0 + lc) : make_double2(0.0, 0.0);
        }
        int slot = 0, par = 0;
        for (int k = 0; k < m; k += RB) {
            double pa[RB];
            if (k + RB <= m) {   // full batch: RB independent latency chains
                const double2* rowv[RB];
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    mbar_wait(&full[slot], (uint32_t)par);
                    rowv[r] = reinterpret_cast<const double2*>(ring + (size_t)slot * stride) + tid;
                    if (++slot == nslots) { slot = 0; par ^= 1; }
                }
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    double qa = 0.0, qb = 0.0;
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        if (c * (64 * CH) + lc0 < Wl) {
                            const double2 av = rowv[r][c * (32 * CH)];
                            qa = fma(av.x, xr[c].x, qa);
                            qb = fma(av.y, xr[c].y, qb);
                        }
                    }
                    pa[r] = qa + qb;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                    for (int r = 0; r < RB; ++r) pa[r] += __shfl_xor_sync(0xffffffffu, pa[r], o);
                }
                if (lane == 0) {
#pragma unroll
                    for (int r = 0; r < RB; ++r) {
                        const int pr = (k + r) % GR_PRING;
                        wsum[pr * 8 + warp] = pa[r];
                        mbar_arrive(&pbar[pr]);
                    }
                }
            } else {             // ragged tail, one row at a time
                for (int r = 0; k + r < m; ++r) {
                    mbar_wait(&full[slot], (uint32_t)par);
                    const double2* rv = reinterpret_cast<const double2*>(ring + (size_t)slot * stride) + tid;
                    if (++slot == nslots) { slot = 0; par ^= 1; }
                    double qa = 0.0, qb = 0.0;
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        if (c * (64 * CH) + lc0 < Wl) {
                            const double2 av = rv[c * (32 * CH)];
                            qa = fma(av.x, xr[c].x, qa);
                            qb = fma(av.y, xr[c].y, qb);
                        }
                    }
                    const double part = bz_warp_sum(qa + qb);
                    if (lane == 0) {
                        const int pr = (k + r) % GR_PRING;
                        wsum[pr * 8 + warp] = part;
                        mbar_arrive(&pbar[pr]);
                    }
                }
            }
        }
    } else if (warp < 2 * CH) {
        // ------------------------------------------------------------------ update warps ("phase B")
        // acc += w_j * A[j, slice] from the same shared-memory slot once w_j has been published; decoupled
        // from the dot warps, so a late w never delays the posting of later rows
        const int t2 = tid - CH * 32;
        const int lc0 = 2 * t2;
        double2 acc[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) acc[c] = make_double2(0.0, 0.0);
        int slot = 0;
        for (int j = 0; j < m; j += RB) {
            if (j + RB <= m) {
                double w[RB];
                const double2* rowv[RB];
                int sl[RB];
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    const int pr = (j + r) % GR_PRING;
                    mbar_wait(&wbar[pr], (uint32_t)(((j + r) / GR_PRING) & 1));
                    w[r] = wring[pr];
                    sl[r] = slot;
                    rowv[r] = reinterpret_cast<const double2*>(ring + (size_t)slot * stride) + t2;
                    if (++slot == nslots) slot = 0;
                }
#pragma unroll
                for (int r = 0; r < RB; ++r) {
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        if (c * (64 * CH) + lc0 < Wl) {
                            const double2 av = rowv[r][c * (32 * CH)];
                            acc[c].x = fma(av.x, w[r], acc[c].x);
                            acc[c].y = fma(av.y, w[r], acc[c].y);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) {
#pragma unroll
                    for (int r = 0; r < RB; ++r) mbar_arrive(&empty[sl[r]]);
                }
            } else {
                for (int r = 0; j + r < m; ++r) {
                    const int pr = (j + r) % GR_PRING;
                    mbar_wait(&wbar[pr], (uint32_t)(((j + r) / GR_PRING) & 1));
                    const double w = wring[pr];
                    const double2* rv = reinterpret_cast<const double2*>(ring + (size_t)slot * stride) + t2;
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        if (c * (64 * CH) + lc0 < Wl) {
                            const double2 av = rv[c * (32 * CH)];
                            acc[c].x = fma(av.x, w, acc[c].x);
                            acc[c].y = fma(av.y, w, acc[c].y);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[slot]);
                    if (++slot == nslots) slot = 0;
                }
            }
        }
        // within a team each column is owned by exactly one CTA: P[team] is the team's partial A'w
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int lc = c * (64 * CH) + lc0;
            if (lc < Wl) *reinterpret_cast<double2*>(a.P + (size_t)team * n + col0 + lc) = acc[c];
        }
    } else if (warp < 2 * CH + XW) {
        // ------------------------------------------------------------------ exchange warps
        const int e = warp - 2 * CH;
        double lx = 0.0;
        bool dead = false;
        const unsigned long long t_start = bz_globaltimer();
        for (int j = e; j < m; j += XW) {
            const double b_j = __ldg(a.b + r0 + j), mu_j = __ldg(a.mu + r0 + j), muy_j = __ldg(a.muy + r0 + j);
            const int pr = j % GR_PRING;
            mbar_wait(&pbar[pr], (uint32_t)((j / GR_PRING) & 1));
            // ring position, tag and arrival counts are all derived from the global row sequence number
            const unsigned long long sq = seq0 + (unsigned long long)j;
            const unsigned long long tag = ((sq + 1ULL) & 0xffffffffULL) << 32;
            const unsigned int pos = (unsigned int)(sq % GR_DRING);
            unsigned long long* Drow = Dteam + (size_t)pos * (size_t)G * 2;
            {   // this CTA's slice dot: fixed tree over the CH warp partials, posted by lane 0
                double v = (lane < CH) ? wsum[pr * 8 + lane] : 0.0;
                if (CH > 4) v += __shfl_xor_sync(0xffffffffu, v, 4);
                if (CH > 2) v += __shfl_xor_sync(0xffffffffu, v, 2);
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                if (lane == 0) {
                    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
                    st_relaxed_v2(Drow + 2 * cta, tag | (bits & 0xffffffffULL), tag | (bits >> 32));
                    if (nshards > 0) {
                        unsigned long long* mycnt = Dcnt + ((size_t)pos * nshards + (cta % nshards)) * 32;
                        asm volatile("red.relaxed.gpu.global.add.u64 [%0], 1;" ::"l"(mycnt) : "memory");
                    }
                }
            }
            // The arrival counters are only a HINT that keeps 148 x XW warps from hammering L2 with gather
            // rounds: correctness rests on the self-validating tags below.  CTA c counts on shard c % nshards
            // (same-address L2 atomics from 148 CTAs serialise at ~27 cycles each: 4000 cycles unsharded).
            // Every wait on another CTA is bounded (GR_TIMEOUT_NS): a team that is not fully resident -- another tenant
            // holds SMs of this device -- sets err_flag = 4 and the launch drains instead of spinning forever
            // (the host then returns BZ_ERR_CUDA).  Once a warp has timed out it stops waiting altogether.
            if (lane < nshards && !dead) {
                const unsigned long long* cnt = Dcnt + ((size_t)pos * nshards + lane) * 32;
                const unsigned long long want = (sq / GR_DRING + 1ULL) * (unsigned long long)((G - lane + nshards - 1) / nshards);
                unsigned long long have;
                unsigned int spins = 0;
                do {
                    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(have) : "l"(cnt) : "memory");
                    if (have < want && (++spins & 255u) == 0u && bz_globaltimer() - t_start > GR_TIMEOUT_NS) { dead = true; break; }
                } while (have < want);
            }
            dead = __any_sync(0xffffffffu, dead);
            // gather the G slice dots: lane l owns CTAs l, l+32, ...; all loads of a round are issued
            // before the first is inspected; only entries that failed validation are re-read
            constexpr int NQ = 5;   // G <= 160
            unsigned long long w0[NQ], w1[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) { w0[q] = ~tag; w1[q] = ~tag; }
            bool pending = !dead;
            unsigned int rounds = 0;
            while (pending) {
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    const int c = lane + 32 * q;
                    if (c >= G) { w0[q] = tag; w1[q] = tag; }
                    else if ((((w0[q] ^ tag) | (w1[q] ^ tag)) >> 32) != 0) ld_relaxed_v2(Drow + 2 * c, w0[q], w1[q]);
                }
                pending = false;
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                    pending |= (((w0[q] ^ tag) | (w1[q] ^ tag)) >> 32) != 0;
                if (pending && (++rounds & 255u) == 0u && bz_globaltimer() - t_start > GR_TIMEOUT_NS) { dead = true; break; }
            }
            dead = __any_sync(0xffffffffu, dead);
            if (dead && lane == 0) a.st->err_flag = 4;
            double sum = 0.0;
#pragma unroll
            for (int q = 0; q < NQ; ++q)
                sum += __longlong_as_double((long long)((w0[q] & 0xffffffffULL) | (w1[q] << 32)));
            __syncwarp();
            const double tot = bz_warp_sum(sum);
            const double cxv = tot - b_j;                                // cx = A*x - b
            double sv, w;
            if (a.slack) {                                               // als, auglagfunslack.jl:88-92
                sv = __ldg(xsrc + n + r0 + j);
                const double tt = (cxv + muy_j) - sv;
                lx += (tt * tt) / mu_j;
                w = __ldg(a.y + r0 + j) + (cxv - sv) / mu_j;
            } else {
                const double t = cxv + muy_j;                            // auglagfun.jl:75
                sv = bz_proj_elem(a.dkind, a.dlo, a.dhi, t, r0 + j);     // :76
                const double u = t - sv;                                 // :77
                lx += (u * u) / mu_j;                                    // :78 (0.5 applied later)
                w = u / mu_j;                                            // :79
            }
            if (lane == 0) {
                wring[pr] = w;
                mbar_arrive(&wbar[pr]);
                if (cta == 0) {
                    if (a.cx) a.cx[r0 + j] = cxv;
                    if (a.s) a.s[r0 + j] = sv;
                    if (a.yupd) a.yupd[r0 + j] = w;
                }
            }
        }
        if (lane == 0) lxs[e] = lx;
    } else if (lane == 0) {
        // ------------------------------------------------------------------ producer
        const uint64_t l2pol = l2_evict_first_policy();
        const double* src = a.A + r0 * a.lda + col0;
        for (int r = 0; r < m; ++r) {
            const int slot = r % nslots;
            if (r >= nslots) mbar_wait(&empty[slot], (uint32_t)(((r / nslots) - 1) & 1));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&full[slot], row_bytes);
            if (a.evict_first)
                bulk_g2s_hint(ring + (size_t)slot * stride, src + (long long)r * a.lda, row_bytes, &full[slot], l2pol);
            else
                bulk_g2s(ring + (size_t)slot * stride, src + (long long)r * a.lda, row_bytes, &full[slot]);
        }
    }

    __syncthreads();
    if (cta == 0 && tid == 0) {   // the team's lx partial: exchange warps in fixed order
        double lx = 0.0;
        for (int e = 0; e < XW; ++e) lx += (e < m) ? lxs[e] : 0.0;
        a.lxpart[team] = lx;
    }
    if (bz_last_cta(a.counter, gridDim.x)) {
        const int T = gridDim.x / GS;
        for (int t = tid; t < T; t += blockDim.x) {   // next launch: fresh tags for every team
            const long long rem = a.m - (long long)t * a.rows_per_cluster;
            a.gseq[t] += (unsigned long long)(rem < a.rows_per_cluster ? rem : a.rows_per_cluster);
        }
        if (tid == 0) {
            if (a.set_after) {
                a.st->next_op = a.st->after_g;
                if (a.force_target < 0) bz_toc(a.st, 0);
            }
            if (a.st->err_flag == 4) {   // a bounded wait expired: stop the state machine, the host reports the error
                a.st->next_op = OP_DONE;
                a.st->inner_status = IST_MAXIT;
            }
            __threadfence();
        }
    }
}

// Geometry of the pass.  The SMs are divided into T teams of GS CTAs; team t owns the rows
// [t*rpt, (t+1)*rpt) and its CTAs split the columns into slices of W.  Wanted: every SM busy, and slices of
// 1024..4096 columns (below ~1024 a row takes less HBM time than one exchange can sustain; above 4096 the x slice
// and the accumulators no longer fit in registers).  C5 (n = 262144): one team of 148, W = 1772; n = 65536: 4 teams
// of 37; n = 32768: 9 teams of 16.  Returns false when the shape cannot take it (odd n, tiny problems, ...): the cluster
// kernel / the two-pass kernels of bz_dense.cu run instead.
bool bz_grid_geometry(const bz_ctx* ctx, const double* A, long long m, long long n, long long lda, BzGridGeom* g) {
    if (ctx->wide_two_pass) return false;
    if (n % 2 || lda % 2 || (reinterpret_cast<uintptr_t>(A) & 15)) return false;
    if (m < 1 || m > 0x7fffff00LL) return false;
    const int sms = ctx->num_sms < 160 ? ctx->num_sms : 160;
    const bool forced = ctx->force_wide != 0;
    // n <= 16384 stays with the 2-CTA cluster kernel of bz_dense.cu: measured on C2 (4096 x 16384) it needs 90 us per
    // pass, this kernel 103 us (37 teams of 4) / 106 us (18 teams of 8) -- a team's rows are few (111) and each pays the
    // exchange latency the cluster kernel's DSMEM hand-over does not have
    if (!forced && n <= DA_MAX_N) return false;
    long long gs_min = (n + 4095) / 4096;
    if (gs_min > sms) return false;
    int best_gs = 0, best_w = 0, best_t = 0;
    double best_u = -1.0;
    for (long long gs0 = gs_min; gs0 <= sms; ++gs0) {
        long long w = (n + gs0 - 1) / gs0;
        w = (w + 1) & ~1LL;
        if (w < 2) w = 2;
        if (w > 4096) continue;
        if (w < 1024 && !forced && gs0 > gs_min) break;         // slices only get narrower from here
        const long long gs = (n + w - 1) / w;                    // CTAs that really get columns
        long long t = sms / gs;
        const long long tmax = (m + 7) / 8;                      // at least ~8 rows per team
        if (t > tmax) t = tmax;
        if (t < 1) t = 1;
        if (ctx->grid_team > 0 && gs != ctx->grid_team) continue;
        const double u = (double)(gs * t) / sms;
        if (u > best_u + 1e-9 || (u > best_u - 1e-9 && gs > best_gs)) { best_u = u; best_gs = (int)gs; best_w = (int)w; best_t = (int)t; }
    }
    if (best_gs == 0) return false;
    if (!forced && best_w < 512) return false;   // rows too short for one exchange per row
    g->GS = best_gs; g->W = best_w; g->T = best_t;
    g->rpt = (int)((m + best_t - 1) / best_t);
    g->T = (int)((m + g->rpt - 1) / g->rpt);
    return true;
}

template <int CH, int NCH, int XW, int RB>
static int launch_variant(bz_ctx* ctx, const BzDenseArgs& a, const BzGridGeom& g, int nslots) {
    static_assert(CH >= 2 && CH <= 8 && (CH & (CH - 1)) == 0, "wsum holds 8 warp partials per row");
    static_assert(XW <= 16, "lxs holds 16 partials");
    if (g.W > NCH * 64 * CH) return bz_fail(ctx, BZ_ERR_UNSUPPORTED, "grid-wide dense pass: slice wider than the variant");
    static BzOncePerDevice once;
    static int resident_per_sm = 0;
    if (once.first(ctx->device)) {
        BZ_CUDA(ctx, cudaFuncSetAttribute(k_grid_algrad<CH, NCH, XW, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, GR_SMEM_BYTES));
        BZ_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident_per_sm, k_grid_algrad<CH, NCH, XW, RB>,
                                                                   (2 * CH + XW + 1) * 32, GR_SMEM_BYTES));
    }
    // the CTAs of a team wait for each other: the whole grid has to be resident at once
    if ((long long)g.GS * g.T > (long long)resident_per_sm * ctx->num_sms)
        return bz_fail(ctx, BZ_ERR_CUDA, "grid-wide dense pass: the grid cannot be co-resident on this device");
    // arrival-counter shards per ring position: ~8 CTAs per counter, at most 16 (GR_SHARDS); teams of <= 8 CTAs
    // validate their few entries directly (measured on 37 teams of 4: 103 us per pass without the hint, 490 us with it)
    int nshards = g.GS <= 8 ? 0 : g.GS / 8;
    if (nshards > GR_SHARDS) nshards = GR_SHARDS;
    return bz_launch_pdl(ctx, k_grid_algrad<CH, NCH, XW, RB>, dim3(g.GS * g.T), dim3((2 * CH + XW + 1) * 32), GR_SMEM_BYTES, 0, a,
                         g.W, nslots, g.GS, nshards);
}

int bz_launch_grid_algrad(bz_ctx* ctx, const BzDenseArgs& args) {
    BzGridGeom g;
    if (!bz_grid_geometry(ctx, args.A, args.m, args.n, args.lda, &g))
        return bz_fail(ctx, BZ_ERR_UNSUPPORTED, "grid-wide dense pass: unsupported shape");
    int nslots = GR_RING_DOUBLES / g.W;          // >= 6 (W <= 4096)
    if (nslots > GR_MAXSLOTS) nslots = GR_MAXSLOTS;
    BzDenseArgs a2 = args;
    a2.rows_per_cluster = g.rpt;
    if (g.W > 2048) return launch_variant<8, 8, 7, 2>(ctx, a2, g, nslots);
    // measured on 8192 x 262144 (us per pass): <4,8,15,2> 2818, <4,8,11,2> 2858, <4,8,15,4> 2962, <4,8,15,1> 3150,
    // <8,4,15,2> 3157, <2,16,15,2> 4352; BZ_GRID_VARIANT=1 keeps the 8+8-warp split reachable for the tests
    if (ctx->grid_variant == 1) return launch_variant<8, 4, 15, 2>(ctx, a2, g, nslots);
    return launch_variant<4, 8, 15, 2>(ctx, a2, g, nslots);
}

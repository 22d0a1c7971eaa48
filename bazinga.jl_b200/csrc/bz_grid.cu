// Single-pass fused AL-gradient for WIDE dense A (nx > 16384; config C5, 32768 x 262144 row-sharded):
//     r = A x - b;  t = r + mu.*y;  s = proj_D(t);  w = (t - s)./mu;  g = A' w      (auglagfun.jl:73-86
//     around demo/basispursuit.jl:38-49), A read from HBM exactly ONCE.
//
// A row of C5 is 2 MB: no cluster can hold it, but the whole GPU can hold ~16 of them in shared
// memory.  So the split is by COLUMNS over the whole grid: CTA c (one per SM, all co-resident) owns
// columns [c*W, c*W + W) of every row, keeps its x slice and its A'w accumulators in registers, and
// streams its slice of each row (W*8 bytes, one cp.async.bulk per row) through a ring of shared-memory
// slots.  For row j every CTA forms its slice dot product ("phase A"), the 148 slice dots are exchanged
// through an L2-resident ring in global memory, every CTA adds them in the same fixed order (bitwise
// identical w everywhere, no float atomics), and the row -- still sitting in shared memory -- is
// accumulated into the column sums ("phase B").  Phase A runs `depth` rows ahead of phase B, which is
// what hides the grid-wide exchange (~2 us against 0.3 us of HBM time per row).
//
// Warp roles (no __syncthreads in the row loop, everything hands over through mbarriers):
//   warps 0..15        compute: phase A of row k, then phase B of row k - depth
//   warps 16..16+XW-1  exchange: row j is handled by exchange warp j % XW: sums the 16 warp partials,
//                      posts the CTA's slice dot, gathers all G slice dots, runs the constraint-space
//                      elementwise stage and publishes w_j to the compute warps
//   last warp          producer: one lane issues the bulk copies as slots are released
//
// Exchange format: D[(seq0 + j) % GR_DRING][cta] is two 64-bit words {tag:32 | value bits 31..0},
// {tag:32 | value bits 63..32}; tag = low 32 bits of a sequence number that is unique per (launch, row).
// Each aligned 8-byte word is single-copy atomic and validates itself, so writer and readers need
// only relaxed gpu-scope accesses: no fences on the critical path.  A per-position arrival counter
// (red.relaxed, monotonic over the life of the problem) tells the gathering warps when a round of
// loads is worth issuing; it is a hint, not a synchronisation.
//
// Co-residency: the grid is <= the number of SMs at 1 CTA/SM and CTAs spin on global memory only after
// griddepcontrol.wait has returned (the predecessor grid has completed), so every CTA of this grid can
// be scheduled; dependents are released only once every CTA has started.
#include "bz_internal.cuh"
#include "bz_dense.cuh"
#include "bz_ptx.cuh"

#define GR_MAXSLOTS 24
#define GR_PRING 16         // ring of per-row warp-partial sets / published w values
#define GR_DRING 256        // rows of the global exchange ring
#define GR_RING_DOUBLES 28672   // 224 KB of row slices
#define GR_SHARDS 16        // arrival-counter shards per ring position (same-address L2 atomics serialise at ~27 cycles each)
#define GR_SMEM_BYTES (GR_RING_DOUBLES * 8 + 2 * GR_MAXSLOTS * 8 + 2 * GR_PRING * 8 + GR_PRING * 16 * 8 + GR_PRING * 8)

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// spin on test_wait (no hardware suspend window): lower wake-up latency than try_wait for the two
// hand-overs that sit on the exchange's critical path
__device__ __forceinline__ void mbar_spin(uint64_t* bar, uint32_t parity, bool spin) {
    if (!spin) { mbar_wait(bar, parity); return; }
    uint32_t done;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void st_relaxed_v2(unsigned long long* p, unsigned long long a,
                                              unsigned long long b) {
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void ld_relaxed_v2(const unsigned long long* p, unsigned long long& a,
                                              unsigned long long& b) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}


// CW compute warps (thread t owns columns {64 CW c + 2t, +1}, c < NCH), XW exchange warps, RB rows per
// compute-warp loop trip
template <int CW, int NCH, int XW, int RB>
__global__ void __launch_bounds__((CW + XW + 1) * 32, 1)
    k_grid_algrad(BzDenseArgs a, int W, int nslots, int depth) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* ring = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + GR_RING_DOUBLES);   // [GR_MAXSLOTS] TMA landed
    uint64_t* empty = full + GR_MAXSLOTS;                                    // [GR_MAXSLOTS] phase B done
    uint64_t* pbar = empty + GR_MAXSLOTS;                                    // [GR_PRING] 16 warp partials in
    uint64_t* wbar = pbar + GR_PRING;                                        // [GR_PRING] w published
    double* wsum = reinterpret_cast<double*>(wbar + GR_PRING);               // [GR_PRING][16]
    double* wring = wsum + GR_PRING * 16;                                    // [GR_PRING]

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
    const unsigned long long seq0 = __ldcg(a.gseq);

    const int G = gridDim.x, cta = blockIdx.x;
    const bool spin = (a.xflags & 2) != 0, spin_w = (a.xflags & 4) != 0;
    const long long n = a.n;
    const int m = (int)a.m;
    const long long col0 = (long long)cta * W;
    const int Wl = (int)((n - col0) < W ? (n - col0) : W);    // local slice width (even, > 0)
    const int stride = W;                                      // slot stride in doubles (even)
    const uint32_t row_bytes = (uint32_t)Wl * 8u;

    if (tid == 0) {
        for (int s = 0; s < GR_MAXSLOTS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CW); }
        for (int s = 0; s < GR_PRING; ++s) { mbar_init(&pbar[s], CW); mbar_init(&wbar[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp < CW) {
        // ------------------------------------------------------------------ compute warps
        const int lc0 = 2 * tid;
        double2 xr[NCH], acc[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int lc = c * (64 * CW) + lc0;
            xr[c] = (lc < Wl) ? *reinterpret_cast<const double2*>(xsrc + col0 + lc) : make_double2(0.0, 0.0);
            acc[c] = make_double2(0.0, 0.0);
        }
        // RB rows per loop trip: the RB slice dots (each a latency chain of shared-memory loads, fp64 FMAs
        // and a 5-step shuffle tree) and the RB column updates are independent, so their latencies overlap
        int slotA = 0, parA = 0, slotB = 0;
        for (int kA = 0; kA < m + depth; kA += RB) {   // depth is a multiple of RB
            if (kA < m) {   // phase A of rows kA .. kA+RB-1
                int sl[RB];
                double pa[RB];
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    sl[r] = slotA;
                    if (kA + r < m) {
                        mbar_wait(&full[slotA], (uint32_t)parA);
                        if (++slotA == nslots) { slotA = 0; parA ^= 1; }
                    }
                }
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    const double2* rowv = reinterpret_cast<const double2*>(ring + (size_t)sl[r] * stride);
                    double qa = 0.0, qb = 0.0;
                    if (kA + r < m) {
#pragma unroll
                        for (int c = 0; c < NCH; ++c) {
                            if (c * (64 * CW) + lc0 < Wl) {
                                const double2 av = rowv[c * (32 * CW) + tid];
                                qa = fma(av.x, xr[c].x, qa);
                                qb = fma(av.y, xr[c].y, qb);
                            }
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
                        if (kA + r < m) {
                            const int pr = (kA + r) % GR_PRING;
                            wsum[pr * 16 + warp] = pa[r];
                            mbar_arrive(&pbar[pr]);
                        }
                    }
                }
            }
            const int jB = kA - depth;
            if (jB >= 0) {  // phase B of rows jB .. jB+RB-1
                double w[RB];
                int sl[RB];
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    const int j = jB + r;
                    sl[r] = slotB;
                    w[r] = 0.0;
                    if (j < m) {
                        mbar_spin(&wbar[j % GR_PRING], (uint32_t)((j / GR_PRING) & 1), spin_w);
                        w[r] = wring[j % GR_PRING];
                        if (++slotB == nslots) slotB = 0;
                    }
                }
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    if (jB + r < m) {
                        const double2* rowv = reinterpret_cast<const double2*>(ring + (size_t)sl[r] * stride);
#pragma unroll
                        for (int c = 0; c < NCH; ++c) {
                            if (c * (64 * CW) + lc0 < Wl) {
                                const double2 av = rowv[c * (32 * CW) + tid];
                                acc[c].x = fma(av.x, w[r], acc[c].x);
                                acc[c].y = fma(av.y, w[r], acc[c].y);
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) {
#pragma unroll
                    for (int r = 0; r < RB; ++r)
                        if (jB + r < m) mbar_arrive(&empty[sl[r]]);
                }
            }
        }
        // each column is owned by exactly one CTA: P is the finished A'w (one "partial" row)
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int lc = c * (64 * CW) + lc0;
            if (lc < Wl) *reinterpret_cast<double2*>(a.P + col0 + lc) = acc[c];
        }
    } else if (warp < CW + XW) {
        // ------------------------------------------------------------------ exchange warps
        const int e = warp - CW;
        double lx = 0.0;
        for (int j = e; j < m; j += XW) {
            const double b_j = __ldg(a.b + j), mu_j = __ldg(a.mu + j), muy_j = __ldg(a.muy + j);
            const int pr = j % GR_PRING;
            mbar_spin(&pbar[pr], (uint32_t)((j / GR_PRING) & 1), spin);
            // ring position, tag and arrival counts are all derived from the global row sequence number
            const unsigned long long sq = seq0 + (unsigned long long)j;
            const unsigned long long tag = ((sq + 1ULL) & 0xffffffffULL) << 32;
            const unsigned int pos = (unsigned int)(sq % GR_DRING);
            const unsigned long long* Drow = a.D + (size_t)pos * (size_t)G * 2;
            {   // this CTA's slice dot: fixed tree over the 16 warp partials, posted by lane 0
                double v = (lane < CW) ? wsum[pr * 16 + lane] : 0.0;
                if (CW > 8) v += __shfl_xor_sync(0xffffffffu, v, 8);
                if (CW > 4) v += __shfl_xor_sync(0xffffffffu, v, 4);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                if (lane == 0) {
                    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
                    st_relaxed_v2(a.D + ((size_t)pos * (size_t)G + cta) * 2, tag | (bits & 0xffffffffULL), tag | (bits >> 32));
                    unsigned long long* mycnt = a.D + (size_t)GR_DRING * 160 * 2 + ((size_t)pos * GR_SHARDS + (cta % GR_SHARDS)) * 32;
                    asm volatile("red.relaxed.gpu.global.add.u64 [%0], 1;" ::"l"(mycnt) : "memory");
                }
            }
            // The arrival counters are only a HINT that keeps 148 x XW warps from hammering L2 with gather
            // rounds: correctness rests on the self-validating tags below.  CTA c counts on shard c % 16.
            if (lane < GR_SHARDS && lane < G) {
                const unsigned long long* cnt = a.D + (size_t)GR_DRING * 160 * 2 + ((size_t)pos * GR_SHARDS + lane) * 32;
                const unsigned long long want = (sq / GR_DRING + 1ULL) * (unsigned long long)((G - lane + GR_SHARDS - 1) / GR_SHARDS);
                unsigned long long have;
                do {
                    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(have) : "l"(cnt) : "memory");
                } while (have < want);
            }
            __syncwarp();
            // gather the G slice dots: lane l owns CTAs l, l+32, ...; all loads of a round are issued
            // before the first is inspected; only entries that failed validation are re-read
            constexpr int NQ = 5;   // G <= 160
            unsigned long long w0[NQ], w1[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) { w0[q] = ~tag; w1[q] = ~tag; }
            bool pending = true;
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
            }
            double sum = 0.0;
#pragma unroll
            for (int q = 0; q < NQ; ++q)
                sum += __longlong_as_double((long long)((w0[q] & 0xffffffffULL) | (w1[q] << 32)));
            __syncwarp();
            const double tot = bz_warp_sum(sum);
            const double cxv = tot - b_j;                                // cx = A*x - b
            const double t = cxv + muy_j;                                // auglagfun.jl:75
            const double sv = bz_proj_elem(a.dkind, a.dlo, a.dhi, t);    // :76
            const double u = t - sv;                                     // :77
            lx += (u * u) / mu_j;                                        // :78 (0.5 applied later)
            const double w = u / mu_j;                                   // :79
            if (lane == 0) {
                wring[pr] = w;
                mbar_arrive(&wbar[pr]);
                if (cta == 0) {
                    if (a.cx) a.cx[j] = cxv;
                    if (a.s) a.s[j] = sv;
                    if (a.yupd) a.yupd[j] = w;
                }
            }
        }
        if (cta == 0 && lane == 0) {
            a.lxpart[e] = lx;
            if (e == 0) for (int q = XW; q < 16; ++q) a.lxpart[q] = 0.0;   // the consumers add 16 partials
        }
    } else if (lane == 0) {
        // ------------------------------------------------------------------ producer
        const uint64_t l2pol = l2_evict_first_policy();
        const double* src = a.A + col0;
        const int pf = a.l2_prefetch_rows;
        for (int r = 0; r < pf && r < m; ++r) l2_prefetch(src + (long long)r * a.lda, row_bytes);
        for (int r = 0; r < m; ++r) {
            const int slot = r % nslots;
            if (r >= nslots) mbar_wait(&empty[slot], (uint32_t)(((r / nslots) - 1) & 1));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&full[slot], row_bytes);
            if (a.evict_first)
                bulk_g2s_hint(ring + (size_t)slot * stride, src + (long long)r * a.lda, row_bytes, &full[slot], l2pol);
            else
                bulk_g2s(ring + (size_t)slot * stride, src + (long long)r * a.lda, row_bytes, &full[slot]);
            // the ring cannot hold enough rows to cover HBM latency AND the exchange: pull the rows behind
            // the ring into L2 early, so that the copy into shared memory only pays L2 latency
            if (pf > 0 && r + pf < m) l2_prefetch(src + (long long)(r + pf) * a.lda, row_bytes);
        }
    }

    if (bz_last_cta(a.counter, gridDim.x)) {
        if (tid == 0) {
            *a.gseq = seq0 + (unsigned long long)m;   // next launch: fresh tags
            if (a.set_after) {
                a.st->next_op = a.st->after_g;
                if (a.force_target < 0) bz_toc(a.st, 0);
            }
            __threadfence();
        }
    }
}

// geometry of the grid-wide pass; returns false when the shape cannot take it (then the two-pass
// kernels of bz_dense.cu run instead)
bool bz_grid_geometry(const bz_ctx* ctx, const double* A, long long m, long long n, long long lda, int* G,
                      int* W, int* nlx) {
    if (ctx->wide_two_pass) return false;
    if (n % 2 || lda % 2 || (reinterpret_cast<uintptr_t>(A) & 15)) return false;
    if (m > 0x7fffff00LL) return false;
    const int sms = ctx->num_sms < 160 ? ctx->num_sms : 160;
    long long w = (n + sms - 1) / sms;
    w = (w + 1) & ~1LL;
    if (w < 2) w = 2;
    if (w > 4096) return false;
    *W = (int)w;
    *G = (int)((n + w - 1) / w);
    *nlx = 16;   // lx partials: one per exchange warp (<= 15), the rest stay zero
    return true;
}

template <int CW, int NCH, int XW, int RB>
static int launch_variant(bz_ctx* ctx, const BzDenseArgs& a, int G, int W, int nslots, int inflight) {
    static bool attr_set = false;
    if (!attr_set) {
        BZ_CUDA(ctx, cudaFuncSetAttribute(k_grid_algrad<CW, NCH, XW, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, GR_SMEM_BYTES));
        attr_set = true;
    }
    // phase A runs `depth` rows ahead of phase B (a multiple of the compute warps' row batch RB); the
    // ring holds depth + RB rows awaiting their w plus the loads in flight
    int depth = ((nslots - inflight - RB) / RB) * RB;
    if (depth > GR_PRING - RB) depth = GR_PRING - RB;
    if (depth < RB) depth = RB;
    return bz_launch_pdl(ctx, k_grid_algrad<CW, NCH, XW, RB>, dim3(G), dim3((CW + XW + 1) * 32), GR_SMEM_BYTES, 0, a, W, nslots, depth);
}

int bz_launch_grid_algrad(bz_ctx* ctx, const BzDenseArgs& args) {
    int G, W, nlx;
    if (!bz_grid_geometry(ctx, args.A, args.m, args.n, args.lda, &G, &W, &nlx))
        return bz_fail(ctx, BZ_ERR_UNSUPPORTED, "grid-wide dense pass: unsupported shape");
    int nslots = GR_RING_DOUBLES / W;
    if (nslots > GR_MAXSLOTS) nslots = GR_MAXSLOTS;
    int inflight = nslots / 4;
    if (inflight < 1) inflight = 1;
    if (inflight > 4) inflight = 4;
    if (ctx->grid_inflight > 0 && ctx->grid_inflight < nslots) inflight = ctx->grid_inflight;
    BzDenseArgs a2 = args;
    a2.xflags = ctx->grid_xflags;
    a2.l2_prefetch_rows = ctx->grid_prefetch;
    if (W > 2048) return launch_variant<8, 8, 7, 2>(ctx, a2, G, W, nslots, inflight);     // nslots >= 7
    switch (ctx->grid_variant) {                                                           // nslots >= 14
    case 1: return launch_variant<16, 2, 15, 4>(ctx, a2, G, W, nslots, inflight);
    case 2: return launch_variant<4, 8, 11, 4>(ctx, a2, G, W, nslots, inflight);
    case 3: return launch_variant<8, 4, 15, 2>(ctx, a2, G, W, nslots, inflight);
    default: return launch_variant<8, 4, 15, 4>(ctx, a2, G, W, nslots, inflight);
    }
}

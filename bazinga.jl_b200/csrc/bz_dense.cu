// Dense kernels: the fused single-pass augmented-Lagrangian gradient for c(x) = A x - b
// (demo/basispursuit.jl:38-49 inside src/utilities/auglagfun.jl:73-86), the same pass reused for
// f = LeastSquares(A,b), a plain GEMV for f = Quadratic(Q,q), plus layout / generator helpers.
//
// Fused pass (k_dense_algrad): A is row-major m x n and is read from HBM exactly ONCE per
// evaluation.  A cluster of two CTAs owns a block of rows; CTA `rank` owns the column half
// [rank*Wh, ...).  Row slices stream through a shared-memory ring filled by cp.async.bulk (TMA
// bulk copies, mbarrier complete_tx).  For row i: phase A forms the slice dot product with x
// (x lives in registers), the two halves are exchanged through distributed shared memory and one
// split-phase cluster barrier, the constraint-space elementwise stage
//     t = r_i - b_i + mu_i*y_i ;  s = proj_D(t) ;  w = (t - s)/mu_i ;  lx += (t-s)^2/mu_i
// is evaluated redundantly by every thread, and phase B accumulates w * A[i, :] into per-thread
// column accumulators from the SAME shared-memory slots.  Per-cluster partial gradients go to
// P[cluster][n]; the consumer kernels (bz_vec.cu) sum them in fixed order, so the result is
// deterministic (no float atomics).
#include "bz_internal.cuh"
#include "bz_dense.cuh"
#include "bz_rng.cuh"

#include "bz_ptx.cuh"


// ---------------------------------------------------------------------------------------------
// fused AL-gradient pass
// ---------------------------------------------------------------------------------------------

// SLACK: the als formulation (AugLagFunSlack) -- a separate instantiation, so that the alps kernel keeps its registers
template <bool BULK, bool SLACK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(DA_THREADS, 1)
    k_dense_algrad(BzDenseArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* ring = reinterpret_cast<double*>(smem_raw);                      // DA_RING_DOUBLES
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + DA_RING_DOUBLES);    // [DA_MAXSLOTS]
    double* wsum = reinterpret_cast<double*>(full + DA_MAXSLOTS);            // [2][16] warp partials
    double* xchg = wsum + 32;                                                 // [2]: peer's row dots
    uint64_t* xbar = reinterpret_cast<uint64_t*>(xchg + 2);                   // [2]: their arrival barriers

    const int tid = threadIdx.x;
    bz_pdl_launch_dependents();
    // Everything up to bz_pdl_wait() below overlaps the tail of the predecessor kernel (programmatic
    // dependent launch): barrier setup and -- because no kernel of the cycle ever writes A -- the
    // bulk copies that fill the ring.  Only x and the state machine have to wait.

    const uint32_t rank = cluster_ctarank();
    const int q = blockIdx.x >> 1;
    const long long n = a.n;
    const long long Wh = (((n + 1) >> 1) + 1) & ~1LL;   // even split point
    const long long col0 = rank ? Wh : 0;
    long long Wl = rank ? (n - Wh) : (Wh < n ? Wh : n);
    if (Wl < 0) Wl = 0;
    const int W = (int)Wl;                               // slice width <= DA_MAXCH*DA_CHUNK
    const int nch = (W + DA_CHUNK - 1) / DA_CHUNK;
    const long long r0 = (long long)q * a.rows_per_cluster;
    long long r1 = r0 + a.rows_per_cluster;
    if (r1 > a.m) r1 = a.m;
    const int nrows = r1 > r0 ? (int)(r1 - r0) : 0;
    const bool is_spec = a.ls.enabled && q >= a.ls.dense_clusters;   // a cluster of the speculative L step: owns no rows

    // ring of whole row slices: slot stride rounded to an even number of doubles (16 B)
    const int stride = ((W + 1) & ~1) > 0 ? ((W + 1) & ~1) : 2;
    int nslots = DA_RING_DOUBLES / stride;
    if (nslots > DA_MAXSLOTS) nslots = DA_MAXSLOTS;
    const uint32_t row_bytes = (uint32_t)W * 8u;

    if (tid == 0) {
        if (BULK)
            for (int s = 0; s < DA_MAXSLOTS; ++s) mbar_init(&full[s], 1);
        mbar_init(&xbar[0], 1);
        mbar_init(&xbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster_arrive_relaxed();   // the peer's barriers exist before anything is sent to them
    cluster_wait();

    const uint64_t l2pol = l2_evict_first_policy();
    // bring local row `row` into ring slot `slot`: one bulk copy + one mbarrier phase (BULK), or a
    // cooperative synchronous fill (ragged / odd shapes)
    auto issue_row = [&](int row, int slot) {
        const double* src = a.A + (r0 + row) * a.lda + col0;
        if (BULK) {
            if (tid == 0 && W > 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&full[slot], row_bytes);
                if (a.evict_first)
                    bulk_g2s_hint(ring + (size_t)slot * stride, src, row_bytes, &full[slot], l2pol);
                else
                    bulk_g2s(ring + (size_t)slot * stride, src, row_bytes, &full[slot]);
                const int pf = row + a.l2_prefetch_rows;
                if (a.l2_prefetch_rows > 0 && pf < nrows)
                    l2_prefetch(a.A + (r0 + pf) * a.lda + col0, row_bytes);
            }
            __syncwarp();
        } else {
            double* dst = ring + (size_t)slot * stride;
            for (int k = tid; k < W; k += DA_THREADS) dst[k] = __ldg(src + k);
            __syncthreads();
        }
    };

    const int npre = nrows < nslots ? nrows : nslots;
    if (BULK)
        for (int r = 0; r < npre; ++r) issue_row(r, r);

    bz_pdl_wait();
    int target = a.force_target;
    if (target < 0) {
        if (__ldcg(&a.st->next_op) != OP_G) {          // both CTAs of a cluster take the same branch
            if (BULK && W > 0)                           // the copies in flight must land before the CTA may exit
                for (int r = 0; r < npre; ++r) mbar_wait(&full[r], 0u);
            return;
        }
        target = __ldcg(&a.st->grad_target);
        if (blockIdx.x == 0 && tid == 0 && a.set_after) bz_tic(a.st, 0);
    }
    if (is_spec) {
        // ---- speculative L step of the NEXT iteration beside G(z) (bz_lspec.cuh); nothing here touches the dense pass
        if (a.force_target < 0 && target == TGT_Z) {
            const int mode = __ldcg(&a.st->mode);
            const int rank_ls = (int)blockIdx.x - 2 * a.ls.dense_clusters;
            const int cs = a.ls.cs;
            if ((mode == MODE_INITFB || mode == MODE_ITER) && rank_ls < cs) {
                // (the ring is unused in these CTAs: its first doubles are the reduction scratch)
                if ((long long)cs * LS_THREADS * 8 >= a.ls.v.n) ls_body<8>(a.st, a.ls, rank_ls, cs, ring);
                else ls_body<16>(a.st, a.ls, rank_ls, cs, ring);
            }
        }
    } else {
    // (a speculated L step that D has confirmed leaves the new x in the shadow buffer until the F step that follows this
    // launch copies it into place, bz_lspec.cuh)
    const double* __restrict__ xsrc = (a.x_shadow && a.force_target < 0 && target == TGT_X && __ldcg(&a.st->spec_commit))
                                          ? a.x_shadow : a.src[target];
    const double* __restrict__ sslack = xsrc + a.n;   // als: the slack part of the stacked input
    if (!BULK)
        for (int r = 0; r < npre; ++r) issue_row(r, r);

    // x slice and accumulators in registers; lane validity is per (chunk, element)
    const int lc0 = 2 * tid;
    double2 xr[DA_MAXCH], acc[DA_MAXCH];
#pragma unroll
    for (int c = 0; c < DA_MAXCH; ++c) {
        const int lc = c * DA_CHUNK + lc0;
        xr[c].x = (lc < W) ? xsrc[col0 + lc] : 0.0;
        xr[c].y = (lc + 1 < W) ? xsrc[col0 + lc + 1] : 0.0;
        acc[c] = make_double2(0.0, 0.0);
    }

    double lx = 0.0;
    double own_prev = 0.0;           // this CTA's slice dot of the previous row
    double b_prev = 0.0, mu_prev = 1.0, muy_prev = 0.0;
    int slot = 0, par = 0, slot_prev = 0;

    // constraint-space stage + phase B for local row `rowp`, given both slice dots
    auto finish_row = [&](int rowp, int slotp, double own, double peer, double bi, double mui,
                          double muyi) {
        const double p0 = rank ? peer : own, p1 = rank ? own : peer;
        const double cxv = (p0 + p1) - bi;                           // cx = A*x - b
        double sv, w;
        if constexpr (SLACK) {                                       // auglagfunslack.jl:88-92
            sv = __ldg(sslack + r0 + rowp);
            const double tt = (cxv + muyi) - sv;
            lx += (tt * tt) / mui;
            w = __ldg(a.y + r0 + rowp) + (cxv - sv) / mui;
        } else {
            const double t = cxv + muyi;                             // auglagfun.jl:75
            sv = bz_proj_elem(a.dkind, a.dlo, a.dhi, t, r0 + rowp);  // :76
            const double u = t - sv;                                 // :77
            lx += (u * u) / mui;                                     // :78 (0.5 applied later)
            w = u / mui;                                             // :79
        }
        if (rank == 0 && tid == 0) {
            const long long gi = r0 + rowp;
            if (a.cx) a.cx[gi] = cxv;
            if (a.s) a.s[gi] = sv;
            if (a.yupd) a.yupd[gi] = w;
        }
        const double2* rowv = reinterpret_cast<const double2*>(ring + (size_t)slotp * stride);
#pragma unroll
        for (int c = 0; c < DA_MAXCH; ++c) {
            if (c < nch) {
                const int lc = c * DA_CHUNK + lc0;
                const double2 av = rowv[c * (DA_CHUNK / 2) + tid];
                if (lc < W) acc[c].x = fma(av.x, w, acc[c].x);
                if (lc + 1 < W) acc[c].y = fma(av.y, w, acc[c].y);
            }
        }
    };

    for (int i = 0; i < nrows; ++i) {
        const long long gi = r0 + i;
        const double b_i = __ldg(a.b + gi), mu_i = __ldg(a.mu + gi), muy_i = __ldg(a.muy + gi);
        // ---- phase A: slice dot of row i
        if (BULK && W > 0) mbar_wait(&full[slot], (uint32_t)par);
        const double2* rowv = reinterpret_cast<const double2*>(ring + (size_t)slot * stride);
        double pa = 0.0, pb = 0.0;
#pragma unroll
        for (int c = 0; c < DA_MAXCH; ++c) {
            if (c < nch) {
                const int lc = c * DA_CHUNK + lc0;
                const double2 av = rowv[c * (DA_CHUNK / 2) + tid];
                if (lc < W) pa = fma(av.x, xr[c].x, pa);
                if (lc + 1 < W) pb = fma(av.y, xr[c].y, pb);
            }
        }
        const double part = bz_warp_sum(pa + pb);
        // ---- finish row i-1 (its peer dot arrived with the previous cluster-barrier phase)
        if (i > 0) {
            mbar_wait_cluster(&xbar[(i - 1) & 1], (uint32_t)(((i - 1) >> 1) & 1));
            const double peer = xchg[(i - 1) & 1];
            finish_row(i - 1, slot_prev, own_prev, peer, b_prev, mu_prev, muy_prev);
        }
        double* ws = wsum + (i & 1) * 16;
        if ((tid & 31) == 0) ws[tid >> 5] = part;
        __syncthreads();   // row i-1's slot is free; warp sums of row i are visible
        if (i > 0 && i - 1 + nslots < nrows) issue_row(i - 1 + nslots, slot_prev);
        // fixed-order pairwise tree over the 16 warp partials (identical in every thread / both CTAs)
        double t8[8];
#pragma unroll
        for (int w = 0; w < 8; ++w) t8[w] = ws[2 * w] + ws[2 * w + 1];
        const double own = ((t8[0] + t8[1]) + (t8[2] + t8[3])) + ((t8[4] + t8[5]) + (t8[6] + t8[7]));
        if (tid == 0) {
            mbar_expect_tx(&xbar[i & 1], 8);                                  // arm my receive of row i
            st_async_peer_f64(&xchg[i & 1], &xbar[i & 1], rank ^ 1u, own);   // send my half to the peer
        }
        __syncwarp();
        own_prev = own; b_prev = b_i; mu_prev = mu_i; muy_prev = muy_i;
        slot_prev = slot;
        if (++slot == nslots) { slot = 0; par ^= 1; }
    }
    if (nrows > 0) {
        mbar_wait_cluster(&xbar[(nrows - 1) & 1], (uint32_t)(((nrows - 1) >> 1) & 1));
        const double peer = xchg[(nrows - 1) & 1];
        finish_row(nrows - 1, slot_prev, own_prev, peer, b_prev, mu_prev, muy_prev);
    }

    // ---- write this cluster's partial gradient slice and its lx partial
    double* Pq = a.P + (size_t)q * n;
#pragma unroll
    for (int c = 0; c < DA_MAXCH; ++c) {
        const int lc = c * DA_CHUNK + lc0;
        if (lc < W) Pq[col0 + lc] = acc[c].x;
        if (lc + 1 < W) Pq[col0 + lc + 1] = acc[c].y;
    }
    if (rank == 0 && tid == 0) a.lxpart[q] = lx;
    // (no closing cluster barrier: every store into this CTA's shared memory has been waited for
    // above, and the last store this CTA issued targets a peer that is still waiting for it)
    }   // !is_spec

    if (bz_last_cta(a.counter, gridDim.x)) {
        if (tid == 0) {
            if (a.is_cpass) a.st->npc_c = a.ls.enabled ? a.ls.dense_clusters : (int)(gridDim.x >> 1);
            if (a.ls.enabled) a.st->lspec_seq += 32ULL;
            if (a.set_after) {
                a.st->next_op = a.st->after_g;
                if (a.force_target < 0) bz_toc(a.st, 0);
            }
            if (a.st->err_flag == 4) {   // a bounded wait of the speculative L step expired: stop, the host reports the error
                a.st->next_op = OP_DONE;
                a.st->inner_status = IST_MAXIT;
            }
            __threadfence();
        }
    }
}

int bz_launch_dense_algrad(bz_ctx* ctx, const BzDenseArgs& args, int nclusters) {
    static BzOncePerDevice once;
    const size_t smem = (size_t)DA_RING_DOUBLES * 8 + DA_MAXSLOTS * 8 + 32 * 8 + 4 * 8 + 2 * 8;
    if (once.first(ctx->device)) {
        BZ_CUDA(ctx, cudaFuncSetAttribute(k_dense_algrad<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        BZ_CUDA(ctx, cudaFuncSetAttribute(k_dense_algrad<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        BZ_CUDA(ctx, cudaFuncSetAttribute(k_dense_algrad<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        BZ_CUDA(ctx, cudaFuncSetAttribute(k_dense_algrad<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const bool bulk = (args.n % 2 == 0) && (args.lda % 2 == 0) &&
                      ((reinterpret_cast<uintptr_t>(args.A) & 15) == 0);
    // with a speculative L step on board: `nclusters` row clusters + ceil(cs / 2) clusters for the L step
    const dim3 grid(2 * (nclusters + (args.ls.enabled ? (args.ls.cs + 1) / 2 : 0))), block(DA_THREADS);
    if (args.slack)
        return bulk ? bz_launch_pdl(ctx, k_dense_algrad<true, true>, grid, block, smem, 0, args)
                    : bz_launch_pdl(ctx, k_dense_algrad<false, true>, grid, block, smem, 0, args);
    return bulk ? bz_launch_pdl(ctx, k_dense_algrad<true, false>, grid, block, smem, 0, args)
                : bz_launch_pdl(ctx, k_dense_algrad<false, false>, grid, block, smem, 0, args);
}


// ---------------------------------------------------------------------------------------------
// f = Quadratic(Q, q): dfx = Q x + q, fx = 0.5 x'Qx + q'x.  One warp per row, rows strided over the
// grid; per-CTA fx partials (fixed order).  Only used by the reference's small QP tests.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_quad_grad(const double* __restrict__ Q,
                                                   const double* __restrict__ qv, BzState* st,
                                                   int force_target, const double* x0,
                                                   const double* x1, const double* x2, long long n,
                                                   double* __restrict__ Ef, double* fxpart) {
    __shared__ double scratch[32];
    int target = force_target;
    if (target < 0) {
        if (__ldcg(&st->next_op) != OP_G) return;
        target = __ldcg(&st->grad_target);
    }
    const double* x = target == 0 ? x0 : (target == 1 ? x1 : x2);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    double v[1] = {0.0};
    for (long long i = (long long)blockIdx.x * wpb + warp; i < n; i += (long long)gridDim.x * wpb) {
        double sacc = 0.0;
        for (long long j = lane; j < n; j += 32) sacc = fma(Q[i * n + j], x[j], sacc);
        sacc = bz_warp_sum(sacc);
        if (lane == 0) {
            Ef[i] = sacc + qv[i];
            v[0] += 0.5 * x[i] * sacc + qv[i] * x[i];
        }
    }
    bz_block_sum<1>(v, scratch);
    if (threadIdx.x == 0) fxpart[blockIdx.x] = v[0];
}

int bz_launch_quad_grad(bz_ctx* ctx, const double* Q, const double* q, BzState* st, int force_target,
                        const double* const* src, long long n, double* Ef, double* fxpart,
                        int nblocks) {
    k_quad_grad<<<nblocks, 256, 0, ctx->stream>>>(Q, q, st, force_target, src[0], src[1], src[2], n,
                                                   Ef, fxpart);
    ctx->n_launches++;
    BZ_CUDA(ctx, cudaGetLastError());
    return BZ_OK;
}

// ---------------------------------------------------------------------------------------------
// layout / generator helpers
// ---------------------------------------------------------------------------------------------
__global__ void k_transpose(const double* __restrict__ in, double* __restrict__ out, long long rows,
                            long long cols) {
    // in: column-major rows x cols (Julia) == row-major cols x rows; out: row-major rows x cols
    __shared__ double tile[32][33];
    long long bx = (long long)blockIdx.x * 32, by = (long long)blockIdx.y * 32;
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        long long r = bx + threadIdx.x, c = by + k;   // in(r, c) at in[c*rows + r]
        tile[k][threadIdx.x] = (r < rows && c < cols) ? in[c * rows + r] : 0.0;
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        long long r = bx + k, c = by + threadIdx.x;
        if (r < rows && c < cols) out[r * cols + c] = tile[threadIdx.x][k];
    }
}

int bz_launch_transpose(bz_ctx* ctx, const double* in_colmajor, double* out_rowmajor, long long rows,
                        long long cols) {
    dim3 block(32, 8), grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32));
    k_transpose<<<grid, block, 0, ctx->stream>>>(in_colmajor, out_rowmajor, rows, cols);
    ctx->n_launches++;
    BZ_CUDA(ctx, cudaGetLastError());
    return BZ_OK;
}

__global__ void k_randn_fill(unsigned long long seed, unsigned long long stream,
                             unsigned long long start, long long count, double* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (long long)gridDim.x * blockDim.x)
        out[i] = bz_randn_at(seed, stream, start + (unsigned long long)i);
}

int bz_launch_randn_fill(bz_ctx* ctx, uint64_t seed, uint64_t stream, uint64_t start, long long count,
                         double* out_dev) {
    if (count <= 0) return BZ_OK;
    long long blocks = (count + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    k_randn_fill<<<(unsigned)blocks, 256, 0, ctx->stream>>>(seed, stream, start, count, out_dev);
    ctx->n_launches++;
    BZ_CUDA(ctx, cudaGetLastError());
    return BZ_OK;
}

// ---------------------------------------------------------------------------------------------
// Wide path (nx > 16384, e.g. config C5 32768 x 262144 row-sharded): TWO passes over A per
// AL-gradient, i.e. at most half of the single-pass roofline by construction -- the x slice and the
// A'w accumulators of the fused kernel no longer fit in one cluster's registers.  Both passes are
// plain streaming kernels with order-fixed reductions:
//   k_wide_rows   R = 8 rows per CTA pass, thread t owns columns {1024c + 2t, +1}: x values are
//                 loaded once per column chunk and reused for the 8 rows (8 independent 16 B loads in
//                 flight per thread), block sum of the 8 row dots, elementwise stage, lx partial
//   k_wide_cols   thread owns 2 columns, CTA (colblock, rowchunk) accumulates w[row]*A[row, cols] over
//                 its row chunk (w staged in shared memory), one partial row per row chunk
// ---------------------------------------------------------------------------------------------
#define WR_ROWS 8
#define WR_THREADS 512

__global__ void __launch_bounds__(WR_THREADS)
    k_wide_rows(BzDenseArgs a) {
    __shared__ double scratch[WR_ROWS * 32];
    bz_pdl_wait();
    bz_pdl_launch_dependents();
    const int tid = threadIdx.x;
    int target = a.force_target;
    if (target < 0) {
        if (__ldcg(&a.st->next_op) != OP_G) return;
        target = __ldcg(&a.st->grad_target);
        if (blockIdx.x == 0 && tid == 0 && a.set_after) bz_tic(a.st, 0);
    }
    const double* __restrict__ x = a.src[target];
    const long long n = a.n, m = a.m;
    const long long ngroups = (m + WR_ROWS - 1) / WR_ROWS;
    const bool vec = (n % 2 == 0) && (a.lda % 2 == 0) && ((reinterpret_cast<uintptr_t>(a.A) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    double lx = 0.0;
    for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const long long row0 = grp * WR_ROWS;
        const int nr = (int)((m - row0) < WR_ROWS ? (m - row0) : WR_ROWS);
        double acc[WR_ROWS];
#pragma unroll
        for (int r = 0; r < WR_ROWS; ++r) acc[r] = 0.0;
        if (vec) {
            for (long long c = 2LL * tid; c < n; c += 2LL * WR_THREADS) {
                const double2 xv = *reinterpret_cast<const double2*>(x + c);
                double2 av[WR_ROWS];
#pragma unroll
                for (int r = 0; r < WR_ROWS; ++r)
                    av[r] = (r < nr) ? __ldcs(reinterpret_cast<const double2*>(a.A + (row0 + r) * a.lda + c))
                                     : make_double2(0.0, 0.0);
#pragma unroll
                for (int r = 0; r < WR_ROWS; ++r) acc[r] = fma(av[r].y, xv.y, fma(av[r].x, xv.x, acc[r]));
            }
        } else {
            for (long long c = tid; c < n; c += WR_THREADS) {
                const double xv = x[c];
#pragma unroll
                for (int r = 0; r < WR_ROWS; ++r)
                    if (r < nr) acc[r] = fma(__ldg(a.A + (row0 + r) * a.lda + c), xv, acc[r]);
            }
        }
        bz_block_sum<WR_ROWS>(acc, scratch);
        if (tid == 0) {
            for (int r = 0; r < nr; ++r) {
                const long long gi = row0 + r;
                const double cxv = acc[r] - a.b[gi];
                const double mui = a.mu[gi];
                double sv, w;
                if (a.slack) {                                   // auglagfunslack.jl:88-92
                    sv = x[n + gi];
                    const double tt = (cxv + a.muy[gi]) - sv;
                    lx += (tt * tt) / mui;
                    w = a.y[gi] + (cxv - sv) / mui;
                } else {
                    const double t = cxv + a.muy[gi];
                    sv = bz_proj_elem(a.dkind, a.dlo, a.dhi, t, gi);
                    const double u = t - sv;
                    lx += (u * u) / mui;
                    w = u / mui;
                }
                if (a.cx) a.cx[gi] = cxv;
                if (a.s) a.s[gi] = sv;
                a.yupd[gi] = w;
            }
        }
        __syncthreads();
    }
    if (tid == 0) a.lxpart[blockIdx.x] = lx;
}

#define WC_THREADS 256
#define WC_COLS (2 * WC_THREADS)
#define WC_WCHUNK 1024      // rows of w staged in shared memory at a time

__global__ void __launch_bounds__(WC_THREADS)
    k_wide_cols(BzDenseArgs a, int rows_per_chunk) {
    __shared__ double wsm[WC_WCHUNK];
    bz_pdl_wait();
    bz_pdl_launch_dependents();
    const int tid = threadIdx.x;
    if (a.force_target < 0 && __ldcg(&a.st->next_op) != OP_G) return;
    const long long n = a.n, m = a.m;
    const long long c0 = (long long)blockIdx.x * WC_COLS + 2LL * tid;
    const long long r0 = (long long)blockIdx.y * rows_per_chunk;
    long long r1 = r0 + rows_per_chunk;
    if (r1 > m) r1 = m;
    const bool vec = (n % 2 == 0) && (a.lda % 2 == 0) && ((reinterpret_cast<uintptr_t>(a.A) & 15) == 0);
    double2 acc = make_double2(0.0, 0.0);
    for (long long rb = r0; rb < r1; rb += WC_WCHUNK) {
        const int nr = (int)((r1 - rb) < WC_WCHUNK ? (r1 - rb) : WC_WCHUNK);
        __syncthreads();
        for (int k = tid; k < nr; k += WC_THREADS) wsm[k] = a.yupd[rb + k];
        __syncthreads();
        if (vec && c0 + 1 < n) {
            const double* base = a.A + rb * a.lda + c0;
            int r = 0;
            for (; r + 8 <= nr; r += 8) {
                double2 av[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) av[k] = __ldcs(reinterpret_cast<const double2*>(base + (long long)(r + k) * a.lda));
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    acc.x = fma(av[k].x, wsm[r + k], acc.x);
                    acc.y = fma(av[k].y, wsm[r + k], acc.y);
                }
            }
            for (; r < nr; ++r) {
                const double2 av = __ldcs(reinterpret_cast<const double2*>(base + (long long)r * a.lda));
                acc.x = fma(av.x, wsm[r], acc.x);
                acc.y = fma(av.y, wsm[r], acc.y);
            }
        } else {
            for (int r = 0; r < nr; ++r) {
                const double* row = a.A + (rb + r) * a.lda;
                if (c0 < n) acc.x = fma(__ldg(row + c0), wsm[r], acc.x);
                if (c0 + 1 < n) acc.y = fma(__ldg(row + c0 + 1), wsm[r], acc.y);
            }
        }
    }
    double* Pq = a.P + (size_t)blockIdx.y * n;
    if (c0 < n) Pq[c0] = acc.x;
    if (c0 + 1 < n) Pq[c0 + 1] = acc.y;
    if (bz_last_cta(a.counter, gridDim.x * gridDim.y)) {
        if (tid == 0 && a.set_after) {
            a.st->next_op = a.st->after_g;
            if (a.force_target < 0) bz_toc(a.st, 0);
            __threadfence();
        }
    }
}

// geometry of the wide path: number of k_wide_rows CTAs (= lx partials) and of row chunks (= partial rows)
void bz_wide_geometry(const bz_ctx* ctx, long long m, long long n, int* row_ctas, int* row_chunks,
                      int* rows_per_chunk) {
    const long long groups = (m + WR_ROWS - 1) / WR_ROWS;
    long long g1 = (long long)ctx->num_sms * 4;
    if (g1 > groups) g1 = groups;
    if (g1 < 1) g1 = 1;
    *row_ctas = (int)g1;
    const long long colblocks = (n + WC_COLS - 1) / WC_COLS;
    long long want = ((long long)ctx->num_sms * 8 + colblocks - 1) / colblocks;   // ~8 CTAs per SM in total
    if (want > 64) want = 64;
    if (want < 1) want = 1;
    long long rpc = (m + want - 1) / want;
    if (rpc < 1) rpc = 1;
    *rows_per_chunk = (int)rpc;
    *row_chunks = (int)((m + rpc - 1) / rpc);
}

int bz_launch_dense_wide(bz_ctx* ctx, const BzDenseArgs& args) {
    int row_ctas, row_chunks, rpc;
    bz_wide_geometry(ctx, args.m, args.n, &row_ctas, &row_chunks, &rpc);
    BZ_TRY_RC(bz_launch_pdl(ctx, k_wide_rows, dim3(row_ctas), dim3(WR_THREADS), 0, 0, args));
    const unsigned colblocks = (unsigned)((args.n + WC_COLS - 1) / WC_COLS);
    return bz_launch_pdl(ctx, k_wide_cols, dim3(colblocks, row_chunks), dim3(WC_THREADS), 0, 0, args, rpc);
}

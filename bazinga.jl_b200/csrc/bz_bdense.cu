// Batched shared-A mode of the dense AL-gradient (BASELINE north_star job 2: "tensor-core GEMMs only in the batched
// shared-A mode where they really are dense contractions"): B <= 8 basis-pursuit instances that share the matrix A
// (demo/basispursuit.jl:38-49 -- cx .= A*x - b, jtv .= A'*v -- evaluated for a batch of x, b, mu, y):
//
//     R = A X          (m x n)(n x 8)      pass 1, DMMA m8n8k4 (fp64 tensor cores; tcgen05 has no fp64 kind)
//     W = (R - b + mu.*y - proj_D(.)) ./ mu     elementwise, in the epilogue of pass 1 (auglagfun.jl:74-79)
//     G = A' W         (n x m)(m x 8)      pass 2, DMMA m8n8k4
//
// A single instance is a GEMV (0.25 flop/B, bz_dense.cu); eight instances that share A are a skinny GEMM whose cost is
// still the read of A -- so eight AL-gradients cost what one costs.  This first version reads A once per pass (twice per
// batched gradient; the single-pass form needs the row-panel exchange of bz_grid.cu widened to 8 x 8 blocks and is the
// next step): 2 x 537 MB for 8 gradients at 4096 x 16384 instead of 8 x 537 MB.
//
// Pass 1: a CTA owns 16 rows; chunks of 128 columns of those rows (16 x 1 KB, cp.async.bulk per row) and of X (8 x 1 KB)
// stream through a 4-stage shared-memory ring; warp w multiplies the 8-row tile (w & 1) with the 32-column quarter (w >> 1)
// of the chunk: 8 DMMAs per chunk on two independent accumulator pairs.  Pass 2: a CTA owns 64 columns; chunks of 32 rows
// x 64 columns stream through the ring, warp w owns 8 columns: 8 DMMAs per chunk (M = columns of A, K = rows, N = batch).
// Two CTAs per SM in both passes.
// Row strides in shared memory are = 4 (mod 16) doubles, which makes every fragment load conflict-free.  No partial
// results cross CTAs: every output element is produced by one warp in a fixed order (deterministic, no atomics).
#include "bz_internal.cuh"
#include "bz_ptx.cuh"
#include "bz_dense.cuh"
#include "bz_vec.cuh"

#include <vector>

#define BD_B 8            // batch width = N of m8n8k4
#define BD_THREADS 256
#define BD_STAGES 4
#define BD_RP 16          // pass 1: rows per CTA (two 8-row tiles; 4 x 25 KB of ring -> two CTAs per SM, 256 CTAs at m = 4096)
#define BD_KC 128         // pass 1: columns per chunk
#ifndef BD_CP
#define BD_CP 64          // pass 2: columns per CTA (eight 8-column groups; two CTAs per SM, 256 CTAs at n = 16384).  -DBD_CP=128 (sixteen
#endif                    // groups, one CTA per SM, six stages, 1 KB row pieces) was measured: 292 us per batched gradient against 291
#define BD2_THREADS (BD_CP * 4)
#define BD2_MINB (BD_CP > 64 ? 1 : 2)
#define BD2_STAGES (BD_CP > 64 ? 6 : 4)
#define BD_RC 32          // pass 2: rows per chunk
#define BD_PAD 4
#define BD_SMEM1 (BD_STAGES * ((BD_RP + BD_B) * (BD_KC + BD_PAD)) * 8 + 64 + (3 * BD_RP * BD_B + 2 * BD_B) * 8)   /* ring, 8 barriers, reduction scratch */
#define BD_SMEM2 (BD2_STAGES * (BD_RC * (BD_CP + BD_PAD) + BD_B * (BD_RC + BD_PAD)) * 8 + 64)

struct bz_bdense {
    bz_ctx* ctx;
    long long m, n, ldm;      // ldm = m rounded up to even: leading dimension of the batch x m device arrays
    double* A;                 // row-major m x n
    double *X, *G;             // [8][n]
    double *bv, *mu, *y, *cx, *yupd, *W;   // [8][m]
    double* lxpart;            // [ctas1][8]
    int ctas1, ctas2;
};

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void bd_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Both passes: 8 consumer warps + 1 producer warp.  The producer's 32 lanes each issue one row segment of the next A
// tile (and lanes 0..7 one row of the X / W chunk) as soon as the stage's `empty` barrier says all consumers are done
// with it; consumers wait on `full`, multiply, and arrive on `empty`.  No block-wide barrier inside the loop (a version
// with one __syncthreads per chunk and a single issuing thread spent two thirds of its cycles at that barrier: ncu r2d).

// ---- pass 1: R = A X for the CTA's 32 rows, then the elementwise AL stage ------------------------------------------------
// ldm = m rounded up to even: leading dimension of the batch x m arrays (16-byte aligned rows for the bulk copies)
__global__ void __launch_bounds__(BD_THREADS + 32, 2)
    k_bd_ax(const double* __restrict__ A, long long m, long long n, long long ldm, const double* __restrict__ X,
            const double* __restrict__ bv, const double* __restrict__ mu, const double* __restrict__ yv, int dkind, double dlo,
            double dhi, double* __restrict__ cx, double* __restrict__ yupd, double* __restrict__ W, double* __restrict__ lxpart) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int LD = BD_KC + BD_PAD;                       // 132 doubles: = 4 (mod 16)
    constexpr int STAGE = (BD_RP + BD_B) * LD;               // A tile (32 rows) then X chunk (8 rows)
    double* ring = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + BD_STAGES * STAGE);
    uint64_t* empty = full + BD_STAGES;
    double* red = reinterpret_cast<double*>(full + 8);       // [3][16][8] k-quarter partial tiles, then [2][8] lx partials
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long r0 = (long long)blockIdx.x * BD_RP;
    const int nchunks = (int)((n + BD_KC - 1) / BD_KC);

    for (int i = tid; i < BD_STAGES * STAGE; i += blockDim.x) ring[i] = 0.0;   // tails must hold finite values
    if (tid == 0) {
        for (int s = 0; s < BD_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], BD_THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // two independent accumulator pairs per warp (even / odd k-steps): one dependent chain of DMMAs per chunk left the
    // tensor pipe 85 % idle (ncu r2d); the pairs are added in fixed order at the end
    double ca[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    double c0 = 0.0, c1 = 0.0;
    const int mt = warp & 1, kq = (warp >> 1) & 3;
    if (warp == BD_THREADS / 32) {
        // ------------------------------------------------------------------ producer warp
        const uint64_t pol = l2_evict_first_policy();
        long long row = r0 + (lane & (BD_RP - 1));
        if (row >= m) row = m - 1;                            // rows past the end: a valid row again, its results are dropped
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % BD_STAGES;
            if (c >= BD_STAGES) mbar_wait(&empty[s], (uint32_t)(((c / BD_STAGES) - 1) & 1));
            const long long cc0 = (long long)c * BD_KC;
            const int valid = (int)((n - cc0) < BD_KC ? (n - cc0) : BD_KC);
            double* dst = ring + (size_t)s * STAGE;
            if (valid < BD_KC) {     // a partial last chunk: stale X beyond `valid` must read as zero (stale A x 0 = 0; A is finite)
                for (int i = lane; i < BD_B * (BD_KC - valid); i += 32)
                    dst[(size_t)(BD_RP + i / (BD_KC - valid)) * LD + valid + i % (BD_KC - valid)] = 0.0;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_expect_tx(&full[s], (uint32_t)((BD_RP + BD_B) * valid * 8));
            __syncwarp();
            if (lane < BD_RP) bulk_g2s_hint(dst + (size_t)lane * LD, A + row * n + cc0, (uint32_t)valid * 8u, &full[s], pol);
            else if (lane < BD_RP + BD_B)
                bulk_g2s(dst + (size_t)lane * LD, X + (size_t)(lane - BD_RP) * n + cc0, (uint32_t)valid * 8u, &full[s]);
        }
    } else {
        // ------------------------------------------------------------------ consumer warps: tile (mt) x k-quarter (kq)
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % BD_STAGES;
            mbar_wait(&full[s], (uint32_t)((c / BD_STAGES) & 1));
            const double* at = ring + (size_t)s * STAGE + (size_t)(8 * mt + (lane >> 2)) * LD + kq * 32 + (lane & 3);
            const double* xt = ring + (size_t)s * STAGE + (size_t)(BD_RP + (lane >> 2)) * LD + kq * 32 + (lane & 3);
            double af[8], xf[8];
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) { af[kk] = at[4 * kk]; xf[kk] = xt[4 * kk]; }     // all fragment loads first
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) dmma_m8n8k4(ca[kk & 1][0], ca[kk & 1][1], af[kk], xf[kk]);
            __syncwarp();
            if (lane == 0) bd_arrive(&empty[s]);
        }
        c0 = ca[0][0] + ca[1][0];
        c1 = ca[0][1] + ca[1][1];
    }
    // ---- k-quarters: the warps of quarters 1..3 hand their tiles to the warps of quarter 0 (fixed order 0 + 1 + 2 + 3)
    const int row_l = 8 * mt + (lane >> 2), bcol = 2 * (lane & 3);
    const bool consumer = warp < BD_THREADS / 32;
    if (consumer && kq > 0) {
        red[((kq - 1) * BD_RP + row_l) * BD_B + bcol] = c0;
        red[((kq - 1) * BD_RP + row_l) * BD_B + bcol + 1] = c1;
    }
    __syncthreads();
    double lx0 = 0.0, lx1 = 0.0;
    if (consumer && kq == 0) {
        const double rr[2] = {((c0 + red[row_l * BD_B + bcol]) + red[(BD_RP + row_l) * BD_B + bcol]) + red[(2 * BD_RP + row_l) * BD_B + bcol],
                              ((c1 + red[row_l * BD_B + bcol + 1]) + red[(BD_RP + row_l) * BD_B + bcol + 1]) +
                                  red[(2 * BD_RP + row_l) * BD_B + bcol + 1]};
        const long long row = r0 + row_l;
        if (row < m) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {                     // auglagfun.jl:74-79 for instance k = bcol + e, uncontracted
                const size_t idx = (size_t)(bcol + e) * ldm + row;
                const double mu_j = mu[idx];
                const double cxv = __dsub_rn(rr[e], bv[idx]);                    // cx = A*x - b
                const double t = __dadd_rn(cxv, __dmul_rn(mu_j, yv[idx]));       // yupd = cx + mu.*y
                const double sv = bz_proj_elem(dkind, dlo, dhi, t, row);         // proj!(s, D, yupd)
                const double u = __dsub_rn(t, sv);
                const double w = __ddiv_rn(u, mu_j);
                (e == 0 ? lx0 : lx1) = __ddiv_rn(__dmul_rn(u, u), mu_j);
                cx[idx] = cxv; yupd[idx] = w; W[idx] = w;
            }
        }
        // sum over the 8 rows of the tile (lanes with equal lane & 3), fixed butterfly order
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            lx0 += __shfl_xor_sync(0xffffffffu, lx0, o);
            lx1 += __shfl_xor_sync(0xffffffffu, lx1, o);
        }
    }
    __syncthreads();
    if (consumer && kq == 0 && lane < 4) { red[3 * BD_RP * BD_B + mt * BD_B + 2 * lane] = lx0; red[3 * BD_RP * BD_B + mt * BD_B + 2 * lane + 1] = lx1; }
    __syncthreads();
    if (tid < BD_B) {
        const double* q = red + 3 * BD_RP * BD_B;
        lxpart[(size_t)blockIdx.x * BD_B + tid] = q[tid] + q[BD_B + tid];
    }
}

// ---- pass 2: G = A' W for the CTA's 128 columns ----------------------------------------------------------------------------
__global__ void __launch_bounds__(BD2_THREADS + 32, BD2_MINB)
    k_bd_atw(const double* __restrict__ A, long long m, long long n, long long ldm, const double* __restrict__ W,
             double* __restrict__ G) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int LDA = BD_CP + BD_PAD;                      // 132
    constexpr int LDW = BD_RC + BD_PAD;                      // 36 = 4 (mod 16)
    constexpr int STAGE = BD_RC * LDA + BD_B * LDW;
    double* ring = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + BD2_STAGES * STAGE);
    uint64_t* empty = full + BD2_STAGES;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long c0 = (long long)blockIdx.x * BD_CP;
    const int cvalid = (int)((n - c0) < BD_CP ? (n - c0) : BD_CP);
    const int nchunks = (int)((m + BD_RC - 1) / BD_RC);

    for (int i = tid; i < BD2_STAGES * STAGE; i += blockDim.x) ring[i] = 0.0;
    if (tid == 0) {
        for (int s = 0; s < BD2_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], BD2_THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};   // [even / odd k-step][2]: the warp's 8-column group
    if (warp == BD2_THREADS / 32) {
        const uint64_t pol = l2_evict_first_policy();
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % BD2_STAGES;
            if (c >= BD2_STAGES) mbar_wait(&empty[s], (uint32_t)(((c / BD2_STAGES) - 1) & 1));
            const long long rr0 = (long long)c * BD_RC;
            const int rvalid = (int)((m - rr0) < BD_RC ? (m - rr0) : BD_RC);
            const int rcopy = (rvalid + 1) & ~1;              // whole 16-byte units: W rows are padded to ldm (the pad entry is zero)
            double* dst = ring + (size_t)s * STAGE;
            if (rcopy < BD_RC) {                              // rows past the end contribute A(m-1, :) * 0
                for (int i = lane; i < BD_B * (BD_RC - rcopy); i += 32)
                    dst[(size_t)BD_RC * LDA + (size_t)(i / (BD_RC - rcopy)) * LDW + rcopy + i % (BD_RC - rcopy)] = 0.0;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_expect_tx(&full[s], (uint32_t)(BD_RC * cvalid * 8 + BD_B * rcopy * 8));
            __syncwarp();
            long long row = rr0 + lane;
            if (row >= m) row = m - 1;                        // finite filler; its W entries are zero
            bulk_g2s_hint(dst + (size_t)lane * LDA, A + row * n + c0, (uint32_t)cvalid * 8u, &full[s], pol);
            if (lane < BD_B)
                bulk_g2s(dst + (size_t)BD_RC * LDA + (size_t)lane * LDW, W + (size_t)lane * ldm + rr0, (uint32_t)rcopy * 8u, &full[s]);
        }
    } else {
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % BD2_STAGES;
            mbar_wait(&full[s], (uint32_t)((c / BD2_STAGES) & 1));
            const double* tile = ring + (size_t)s * STAGE;
            const double* wt = tile + (size_t)BD_RC * LDA + (size_t)(lane >> 2) * LDW + (lane & 3);
            double wf[BD_RC / 4], af[BD_RC / 4];
#pragma unroll
            for (int kk = 0; kk < BD_RC / 4; ++kk) {
                wf[kk] = wt[4 * kk];
                af[kk] = tile[(size_t)(4 * kk + (lane & 3)) * LDA + 8 * warp + (lane >> 2)];
            }
#pragma unroll
            for (int kk = 0; kk < BD_RC / 4; ++kk) dmma_m8n8k4(acc[kk & 1][0], acc[kk & 1][1], af[kk], wf[kk]);
            __syncwarp();
            if (lane == 0) bd_arrive(&empty[s]);
        }
        const long long col = c0 + 8 * warp + (lane >> 2);
        if (col < n) {
            G[(size_t)(2 * (lane & 3)) * n + col] = acc[0][0] + acc[1][0];
            G[(size_t)(2 * (lane & 3) + 1) * n + col] = acc[0][1] + acc[1][1];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
extern "C" int bz_bdense_create(bz_ctx* ctx, int64_t m, int64_t n, const double* A, int32_t layout, int32_t gen,
                                uint64_t gen_seed, bz_bdense** out) {
    if (!ctx || !out || m < 1 || n < 2 || (n % 2)) return bz_fail(ctx, BZ_ERR_ARG, "bz_bdense_create: need m >= 1 and an even n >= 2");
    if (!gen && !A) return bz_fail(ctx, BZ_ERR_ARG, "bz_bdense_create: A is NULL");
    if (!gen && layout != BZ_ROW_MAJOR) return bz_fail(ctx, BZ_ERR_UNSUPPORTED, "bz_bdense_create: row-major A only");
    *out = nullptr;
    BZ_CUDA(ctx, cudaSetDevice(ctx->device));
    bz_bdense* h = new bz_bdense();
    h->ctx = ctx; h->m = m; h->n = n; h->ldm = (m + 1) & ~1LL;
    h->ctas1 = (int)((m + BD_RP - 1) / BD_RP);
    h->ctas2 = (int)((n + BD_CP - 1) / BD_CP);
    auto fail = [&](cudaError_t e) { bz_bdense_destroy(h); return bz_fail(ctx, BZ_ERR_CUDA, cudaGetErrorString(e)); };
    cudaError_t e;
    if ((e = cudaMalloc((void**)&h->A, (size_t)m * n * 8)) != cudaSuccess) return fail(e);
    double** mv[] = {&h->bv, &h->mu, &h->y, &h->cx, &h->yupd, &h->W};
    for (double** q : mv) {
        if ((e = cudaMalloc((void**)q, (size_t)BD_B * h->ldm * 8)) != cudaSuccess) return fail(e);
        cudaMemsetAsync(*q, 0, (size_t)BD_B * h->ldm * 8, ctx->stream);
    }
    double** nv[] = {&h->X, &h->G};
    for (double** q : nv) if ((e = cudaMalloc((void**)q, (size_t)BD_B * n * 8)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc((void**)&h->lxpart, (size_t)h->ctas1 * BD_B * 8)) != cudaSuccess) return fail(e);
    if (gen) {
        int rc = bz_launch_randn_fill(ctx, gen_seed, 0, 0, m * n, h->A);
        if (rc != BZ_OK) { bz_bdense_destroy(h); return rc; }
    } else if ((e = cudaMemcpyAsync(h->A, A, (size_t)m * n * 8, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) return fail(e);
    if ((e = cudaFuncSetAttribute(k_bd_ax, cudaFuncAttributeMaxDynamicSharedMemorySize, BD_SMEM1)) != cudaSuccess) return fail(e);
    if ((e = cudaFuncSetAttribute(k_bd_atw, cudaFuncAttributeMaxDynamicSharedMemorySize, BD_SMEM2)) != cudaSuccess) return fail(e);
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return fail(e);
    *out = h;
    return BZ_OK;
}

extern "C" int bz_bdense_destroy(bz_bdense* h) {
    if (!h) return BZ_OK;
    cudaSetDevice(h->ctx->device);
    cudaStreamSynchronize(h->ctx->stream);
    double* fr[] = {h->A, h->X, h->G, h->bv, h->mu, h->y, h->cx, h->yupd, h->W, h->lxpart};
    for (double* q : fr) if (q) cudaFree(q);
    delete h;
    return BZ_OK;
}

extern "C" int bz_bdense_auglag_grad(bz_bdense* h, int32_t batch, const double* X, const double* b, const double* mu,
                                     const double* y, int32_t d_kind, double d_lo, double d_hi, double* dlx, double* lx,
                                     double* cx, double* yupd, int32_t reps, double* us_per_call) {
    if (!h || !X || !b || !mu || !y || batch < 1 || batch > BD_B) return BZ_ERR_ARG;
    bz_ctx* ctx = h->ctx;
    BZ_CUDA(ctx, cudaSetDevice(ctx->device));
    if (d_kind != BZ_D_ZERO && d_kind != BZ_D_FREE && d_kind != BZ_D_BOX)
        return bz_fail(ctx, BZ_ERR_UNSUPPORTED, "bz_bdense_auglag_grad: elementwise sets only (ZeroSet, FreeSet, box)");
    const long long m = h->m, n = h->n, ldm = h->ldm;
    // instances beyond `batch` are padded: x = 0, b = 0, mu = 1, y = 0
    BZ_CUDA(ctx, cudaMemsetAsync(h->X, 0, (size_t)BD_B * n * 8, ctx->stream));
    BZ_CUDA(ctx, cudaMemsetAsync(h->bv, 0, (size_t)BD_B * ldm * 8, ctx->stream));
    BZ_CUDA(ctx, cudaMemsetAsync(h->y, 0, (size_t)BD_B * ldm * 8, ctx->stream));
    BZ_CUDA(ctx, cudaMemsetAsync(h->W, 0, (size_t)BD_B * ldm * 8, ctx->stream));
    BZ_TRY_RC(bz_launch_fill(ctx, h->mu, BD_B * ldm, 1.0));
    BZ_CUDA(ctx, cudaMemcpyAsync(h->X, X, (size_t)batch * n * 8, cudaMemcpyHostToDevice, ctx->stream));
    BZ_CUDA(ctx, cudaMemcpy2DAsync(h->bv, ldm * 8, b, m * 8, m * 8, batch, cudaMemcpyHostToDevice, ctx->stream));
    BZ_CUDA(ctx, cudaMemcpy2DAsync(h->mu, ldm * 8, mu, m * 8, m * 8, batch, cudaMemcpyHostToDevice, ctx->stream));
    BZ_CUDA(ctx, cudaMemcpy2DAsync(h->y, ldm * 8, y, m * 8, m * 8, batch, cudaMemcpyHostToDevice, ctx->stream));
    if (reps < 1) reps = 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int r = 0; r < reps + 1; ++r) {            // the first launch is a warm-up when reps > 1
        if (r == 1 || reps == 1) cudaEventRecord(e0, ctx->stream);
        k_bd_ax<<<h->ctas1, BD_THREADS + 32, BD_SMEM1, ctx->stream>>>(h->A, m, n, ldm, h->X, h->bv, h->mu, h->y, d_kind, d_lo, d_hi,
                                                                      h->cx, h->yupd, h->W, h->lxpart);
        k_bd_atw<<<h->ctas2, BD2_THREADS + 32, BD_SMEM2, ctx->stream>>>(h->A, m, n, ldm, h->W, h->G);
        ctx->n_launches += 2;
        if (reps == 1) break;
    }
    cudaEventRecord(e1, ctx->stream);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (e != cudaSuccess) return bz_fail(ctx, BZ_ERR_CUDA, cudaGetErrorString(e));
    if (us_per_call) *us_per_call = 1e3 * (double)ms / reps;
    if (dlx) BZ_CUDA(ctx, cudaMemcpyAsync(dlx, h->G, (size_t)batch * n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (cx) BZ_CUDA(ctx, cudaMemcpy2DAsync(cx, m * 8, h->cx, ldm * 8, m * 8, batch, cudaMemcpyDeviceToHost, ctx->stream));
    if (yupd) BZ_CUDA(ctx, cudaMemcpy2DAsync(yupd, m * 8, h->yupd, ldm * 8, m * 8, batch, cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<double> part((size_t)h->ctas1 * BD_B);
    BZ_CUDA(ctx, cudaMemcpyAsync(part.data(), h->lxpart, part.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    BZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (lx) {
        // lx_k = 0.5*sum((t-s)^2 ./ mu) + f(x) - 0.5*sum(mu.*y.^2) with f = Zero (auglagfun.jl:78-82); partials in CTA order
        for (int k = 0; k < batch; ++k) {
            double s = 0.0;
            for (int c = 0; c < h->ctas1; ++c) s += part[(size_t)c * BD_B + k];
            double musqy = 0.0;
            for (long long i = 0; i < m; ++i) musqy += (mu[(size_t)k * m + i] * y[(size_t)k * m + i]) * y[(size_t)k * m + i];
            lx[k] = 0.5 * s - 0.5 * musqy;
        }
    }
    return BZ_OK;
}

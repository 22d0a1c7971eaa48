// Matrix prox operators of src/proxoperators: Rank (rank.jl:28-33), NuclearNorm (nuclearNorm.jl:29-34) and
// SchattenNormLpPower (schattenNormLp.jl:36-42 with the scalar subproblem of :60-97).  All three are
//     Y = U * diag(h(sigma)) * V'   with   X = U * diag(sigma) * V'
// for an elementwise h, so no ordered SVD is needed: a one-sided Jacobi (Hestenes) iteration rotates the columns of X until
// they are mutually orthogonal, A = X*V with A'A diagonal; then sigma_k = ||a_k||, u_k = a_k / sigma_k and
//     Y = sum_k (h(sigma_k) / sigma_k) * a_k * v_k'.
// The reference calls LAPACK through `svd(X)` for every prox (demo/distance_matrix_completion.jl works on 20 x 20 matrices);
// here one CTA owns one matrix of a batch, A and V live in shared memory, the ceil(n/2) disjoint column pairs of a
// round-robin step are rotated by one warp each (lanes = rows), and sweeps repeat until no pair needed a rotation.
// One-sided Jacobi computes singular values to high relative accuracy; results agree with LAPACK's to rounding
// (tests/test_gpu_matrix_prox.py: 1e-12 relative on sigma, 1e-10 on Y where `pow` is involved).
// Compiled with -fmad=false like the other prox code.
#include "bz_internal.cuh"

#define MP_MAXDIM 64                 // rows, cols <= 64: A (64 x 65) + V (64 x 65) doubles = 65 KB of shared memory
#define MP_LD (MP_MAXDIM + 1)
#define MP_MAXSWEEPS 60

namespace {

// schattenNormLp.jl:60-97 solve_lp_quasi_norm_subproblem(x, p, a, gamma): min_z>=0 a z^p + (z - x)^2 / (2 gamma)
__device__ double mp_lp_subproblem(double x, double p, double a, double gamma) {
    const double alpha = a * gamma;
    if (x <= 0.0) return 0.0;
    const double zbar = pow(1 / (alpha * p * (1 - p)), 1 / (p - 2));
    const double psi_zbar = zbar + alpha * p * pow(zbar, p - 1);
    if (psi_zbar >= x) return 0.0;
    double z = zbar + 0.1;                       // perturbation to the right
    for (int it = 0; it < 1000;) {
        const double dphi = z - x + alpha * p * pow(z, p - 1);
        if (fabs(dphi) <= 1e-14) break;
        ++it;
        const double ddphi = 1 + alpha * p * (p - 1) * pow(z, p - 2);
        z -= dphi / ddphi;
    }
    const double phi_0 = 0.5 * (x * x);
    const double dz = z - x;
    const double phi_z = 0.5 * (dz * dz) + alpha * pow(z, p);
    if (phi_0 <= phi_z) return 0.0;
    return z;
}

__device__ __forceinline__ double mp_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// X, Y: batch matrices, each m x n with contiguous columns (Julia layout; a C-ordered (n, m) array is the same buffer and
// the prox commutes with the transpose).  gz[b] = g(Y_b) as the reference's prox! returns it; sigma (optional): the n
// singular values of each matrix in descending order (zeros beyond min(m, n)); sweeps (optional): Jacobi sweeps used.
__global__ void __launch_bounds__(1024)
    k_matprox(int kind, double lambda, double p, double gamma, int m, int n, long long batch,
              const double* __restrict__ X, double* __restrict__ Y, double* __restrict__ gz,
              double* __restrict__ sigma_out, int* __restrict__ sweeps_out) {
    extern __shared__ double smem[];
    double* A = smem;                          // m x n, leading dimension MP_LD
    double* V = smem + MP_MAXDIM * MP_LD;      // n x n
    __shared__ double sig[MP_MAXDIM], w[MP_MAXDIM];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int ne = n + (n & 1);                // players of the round-robin tournament (one dummy when n is odd)
    const int npairs = ne / 2;
    for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
        const double* Xb = X + (size_t)b * m * n;
        for (int e = tid; e < m * n; e += blockDim.x) A[(e / m) * MP_LD + (e % m)] = Xb[e];
        for (int e = tid; e < n * n; e += blockDim.x) V[(e / n) * MP_LD + (e % n)] = (e / n == e % n) ? 1.0 : 0.0;
        __syncthreads();
        int sweeps = 0;
        for (; sweeps < MP_MAXSWEEPS; ++sweeps) {
            int rotated = 0;
            for (int step = 0; step < ne - 1; ++step) {
                for (int k = warp; k < npairs; k += nwarps) {
                    // circle method: player ne-1 stays, the others rotate
                    int i = (k == 0) ? ne - 1 : (step + k) % (ne - 1);
                    int j = (k == 0) ? step : (step - k + (ne - 1)) % (ne - 1);
                    if (i > j) { const int t = i; i = j; j = t; }
                    if (j >= n) continue;      // the dummy player
                    double* ai = A + i * MP_LD;
                    double* aj = A + j * MP_LD;
                    const double x0 = lane < m ? ai[lane] : 0.0, x1 = lane + 32 < m ? ai[lane + 32] : 0.0;
                    const double y0 = lane < m ? aj[lane] : 0.0, y1 = lane + 32 < m ? aj[lane + 32] : 0.0;
                    const double alpha = mp_warp_sum(x0 * x0 + x1 * x1);
                    const double beta = mp_warp_sum(y0 * y0 + y1 * y1);
                    const double gam = mp_warp_sum(x0 * y0 + x1 * y1);
                    // columns count as orthogonal once |a_i . a_j| <= 4 eps ||a_i|| ||a_j||
                    if (gam == 0.0 || !(fabs(gam) > 4 * BZ_EPS * sqrt(alpha) * sqrt(beta))) continue;
                    rotated = 1;
                    const double zeta = (beta - alpha) / (2 * gam);
                    const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1 + zeta * zeta));
                    const double c = 1 / sqrt(1 + t * t), s = c * t;
                    if (lane < m) { ai[lane] = c * x0 - s * y0; aj[lane] = s * x0 + c * y0; }
                    if (lane + 32 < m) { ai[lane + 32] = c * x1 - s * y1; aj[lane + 32] = s * x1 + c * y1; }
                    double* vi = V + i * MP_LD;
                    double* vj = V + j * MP_LD;
                    for (int r = lane; r < n; r += 32) {
                        const double p0 = vi[r], q0 = vj[r];
                        vi[r] = c * p0 - s * q0;
                        vj[r] = s * p0 + c * q0;
                    }
                }
                __syncthreads();
            }
            if (!__syncthreads_or(rotated)) break;
        }
        // singular values = column norms; the elementwise map h and the weights h(sigma) / sigma
        for (int k = warp; k < n; k += nwarps) {
            const double* ak = A + k * MP_LD;
            const double x0 = lane < m ? ak[lane] : 0.0, x1 = lane + 32 < m ? ak[lane + 32] : 0.0;
            const double sg = sqrt(mp_warp_sum(x0 * x0 + x1 * x1));
            if (lane == 0) {
                double h;
                if (kind == BZ_MATPROX_RANK) h = (sg > sqrt(2 * gamma * lambda)) ? sg : 0.0;          // rank.jl:30-31
                else if (kind == BZ_MATPROX_NUCLEAR) h = fmax(0.0, sg - lambda * gamma);              // nuclearNorm.jl:31
                else h = mp_lp_subproblem(sg, p, lambda, gamma);                                      // schattenNormLp.jl:39
                sig[k] = sg;
                w[k] = sg > 0.0 ? h / sg : 0.0;
            }
        }
        __syncthreads();
        // Y = sum_k w_k a_k v_k'
        double* Yb = Y + (size_t)b * m * n;
        for (int e = tid; e < m * n; e += blockDim.x) {
            const int r = e % m, cidx = e / m;
            double acc = 0.0;
            for (int k = 0; k < n; ++k) acc += (w[k] * A[k * MP_LD + r]) * V[k * MP_LD + cidx];
            Yb[e] = acc;
        }
        __syncthreads();
        if (tid == 0) {
            // the value prox! returns, summed over the singular values in descending order like F.S
            for (int a = 1; a < n; ++a) {        // insertion sort of (sig, w), n <= 64
                const double sa = sig[a], wa = w[a];
                int q = a - 1;
                while (q >= 0 && sig[q] < sa) { sig[q + 1] = sig[q]; w[q + 1] = w[q]; --q; }
                sig[q + 1] = sa; w[q + 1] = wa;
            }
            double acc = 0.0;
            const int nsv = m < n ? m : n;       // svd(X) has min(m, n) singular values
            for (int k = 0; k < nsv; ++k) {
                const double hs = w[k] * sig[k];
                if (kind == BZ_MATPROX_RANK) acc += (sig[k] > sqrt(2 * gamma * lambda)) ? 1.0 : 0.0;
                else if (kind == BZ_MATPROX_NUCLEAR) acc += fmax(0.0, sig[k] - lambda * gamma);
                else acc += pow(hs, p);
            }
            gz[b] = lambda * acc;
            if (sigma_out) for (int k = 0; k < n; ++k) sigma_out[(size_t)b * n + k] = sig[k];
            if (sweeps_out) sweeps_out[b] = sweeps;
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int bz_prox_matrix(bz_ctx* ctx, int kind, double lambda, double p, double gamma, int rows, int cols,
                              int64_t batch, const double* X, double* Y, double* gz, double* sigma, int32_t* sweeps) {
    if (!ctx || !X || !Y || !gz) return bz_fail(ctx, BZ_ERR_ARG, "bz_prox_matrix: NULL argument");
    if (kind != BZ_MATPROX_RANK && kind != BZ_MATPROX_NUCLEAR && kind != BZ_MATPROX_SCHATTEN_LP)
        return bz_fail(ctx, BZ_ERR_UNSUPPORTED, "bz_prox_matrix: unknown matrix prox kind");
    if (lambda < 0) return bz_fail(ctx, BZ_ERR_ARG, kind == BZ_MATPROX_SCHATTEN_LP ? "alpha must be nonnegative" : "parameter λ must be nonnegative");
    if (kind == BZ_MATPROX_SCHATTEN_LP && p <= 0) return bz_fail(ctx, BZ_ERR_ARG, "p must be positive");
    if (kind == BZ_MATPROX_SCHATTEN_LP && p >= 1) return bz_fail(ctx, BZ_ERR_ARG, "p must be smaller than one");
    if (rows < 1 || cols < 1 || rows > MP_MAXDIM || cols > MP_MAXDIM)
        return bz_fail(ctx, BZ_ERR_UNSUPPORTED, "bz_prox_matrix: matrices of 1..64 rows and columns");
    if (batch < 0) return bz_fail(ctx, BZ_ERR_ARG, "batch must be nonnegative");
    if (batch == 0) return BZ_OK;
    BZ_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t nel = (size_t)batch * rows * cols;
    double *dX = nullptr, *dY = nullptr, *dg = nullptr, *ds = nullptr;
    int* dsw = nullptr;
    auto cleanup = [&]() { cudaFree(dX); cudaFree(dY); cudaFree(dg); cudaFree(ds); cudaFree(dsw); };
#define MTRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); \
        return bz_fail(ctx, BZ_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); } } while (0)
    MTRY(cudaMalloc(&dX, nel * sizeof(double)));
    MTRY(cudaMalloc(&dY, nel * sizeof(double)));
    MTRY(cudaMalloc(&dg, (size_t)batch * sizeof(double)));
    if (sigma) MTRY(cudaMalloc(&ds, (size_t)batch * cols * sizeof(double)));
    if (sweeps) MTRY(cudaMalloc(&dsw, (size_t)batch * sizeof(int)));
    MTRY(cudaMemcpyAsync(dX, X, nel * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const int npairs = (cols + 1) / 2;
    const int threads = 32 * (npairs < 1 ? 1 : npairs);
    const size_t smem = (size_t)2 * MP_MAXDIM * MP_LD * sizeof(double);
    MTRY(cudaFuncSetAttribute(k_matprox, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = batch < (long long)ctx->num_sms * 2 ? batch : (long long)ctx->num_sms * 2;
    k_matprox<<<(unsigned)grid, threads, smem, ctx->stream>>>(kind, lambda, p, gamma, rows, cols, batch, dX, dY, dg, ds, dsw);
    ctx->n_launches++;
    MTRY(cudaGetLastError());
    MTRY(cudaMemcpyAsync(Y, dY, nel * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MTRY(cudaMemcpyAsync(gz, dg, (size_t)batch * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (sigma) MTRY(cudaMemcpyAsync(sigma, ds, (size_t)batch * cols * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (sweeps) MTRY(cudaMemcpyAsync(sweeps, dsw, (size_t)batch * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MTRY(cudaStreamSynchronize(ctx->stream));
#undef MTRY
    cleanup();
    return BZ_OK;
}

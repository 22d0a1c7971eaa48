// Launch interfaces of bz_dense.cu
#pragma once
#include "bz_internal.cuh"
#include "bz_lspec.cuh"

#define DA_THREADS 512
#define DA_CHUNK 1024   // doubles per ring slot (8 KB)
#define DA_MAXCH 8      // ring slots per row slice -> slice width <= 8192, n <= 16384
#define DA_RING_DOUBLES (26 * 1024)   // 208 KB ring of whole row slices
#define DA_MAXSLOTS 16
#define DA_MAX_N (2 * DA_MAXCH * DA_CHUNK)

struct BzDenseArgs {
    const double* A;       // row-major m x n (local rows)
    long long lda;
    const double* b;       // m
    const double* mu;      // m (ones for the LeastSquares pass)
    const double* muy;     // m (zeros for the LeastSquares pass)
    const double* src[3];  // candidate input vectors: x, z, xeps
    BzState* st;
    int force_target;      // >= 0: ignore the state machine, read src[force_target]
    long long m, n;
    int rows_per_cluster;
    int dkind;
    double dlo, dhi;
    double* cx;            // outputs over rows (may be null)
    double* s;
    double* yupd;
    double* P;             // nclusters x n partial rows of A' w
    double* lxpart;        // nclusters partial sums of (t-s)^2/mu
    unsigned int* counter;
    int set_after;         // last CTA sets st->next_op = st->after_g
    int evict_first;       // stream A with an L2 evict-first policy
    int l2_prefetch_rows;  // >0: cp.async.bulk.prefetch.L2 this many rows ahead of each issue
    // als (AugLagFunSlack, auglagfunslack.jl:77-97): the input vector is [x; s] (n + m entries) and the slack s replaces
    // proj_D:  lx += (cx + mu*y - s)^2/mu,  yupd = y + (cx - s)/mu
    int slack;
    const double* y;       // m multipliers (slack mode only)
    int is_cpass;          // the constraint pass (not the LeastSquares pass): records its partial-row count in BzState::npc_c
    const double* x_shadow; // problems that speculate the L step: where x + d waits while BzState::spec_commit is set (else null)
    BzLSpec ls;            // speculative L step riding on this launch (k_dense_algrad only; ls.enabled = 0: none)
    // grid-wide single-pass kernel (bz_grid.cu) only
    unsigned long long* D;     // exchange ring: GR_DRING x 160 x 2 words, then GR_DRING arrival counters
    unsigned long long* gseq;  // [teams] sequence base of the exchange tags, advanced by the team's row count per launch
};

int bz_launch_dense_algrad(bz_ctx* ctx, const BzDenseArgs& args, int nclusters);
int bz_launch_dense_wide(bz_ctx* ctx, const BzDenseArgs& args);
// grid-wide single-pass kernel (bz_grid.cu): teams of CTAs split the columns of a row block
#define BZ_GRID_EXCHANGE_WORDS (256 * 160 * 2 + 160 * 256 * 32)   // entries + arrival counters (256 B apart): teams x shards <= 160
struct BzGridGeom { int GS, T, W, rpt; };   // team size, teams (= partial rows), slice width, rows per team
bool bz_grid_geometry(const bz_ctx* ctx, const double* A, long long m, long long n, long long lda, BzGridGeom* g);
int bz_launch_grid_algrad(bz_ctx* ctx, const BzDenseArgs& args);
void bz_wide_geometry(const bz_ctx* ctx, long long m, long long n, int* row_ctas, int* row_chunks,
                      int* rows_per_chunk);
int bz_launch_quad_grad(bz_ctx* ctx, const double* Q, const double* q, BzState* st, int force_target,
                        const double* const* src, long long n, double* Ef, double* fxpart,
                        int nblocks);
int bz_launch_transpose(bz_ctx* ctx, const double* in_colmajor, double* out_rowmajor, long long rows,
                        long long cols);
int bz_launch_randn_fill(bz_ctx* ctx, uint64_t seed, uint64_t stream, uint64_t start, long long count,
                         double* out_dev);

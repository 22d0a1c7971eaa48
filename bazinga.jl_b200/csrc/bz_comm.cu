// See bz_comm.cuh.
#include "bz_comm.cuh"
#include "bz_vec.cuh"
#include <dlfcn.h>
#include <nccl.h>

namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};
NcclApi g_nccl;

bool nccl_load() {
    if (g_nccl.lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) { g_nccl.err = std::string("dlopen libnccl.so.2 failed: ") + dlerror(); return false; }
#define LOAD(field, sym)                                                            \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(g_nccl.lib, sym)); \
    if (!g_nccl.field) { g_nccl.err = std::string("dlsym failed: ") + sym; g_nccl.lib = nullptr; return false; }
    LOAD(GetUniqueId, "ncclGetUniqueId")
    LOAD(CommInitRank, "ncclCommInitRank")
    LOAD(CommDestroy, "ncclCommDestroy")
    LOAD(AllReduce, "ncclAllReduce")
    LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
    return true;
}
}  // namespace

extern "C" int bz_comm_unique_id(void* id128) {
    if (!id128) return BZ_ERR_ARG;
    if (!nccl_load()) return bz_fail(nullptr, BZ_ERR_COMM, g_nccl.err);
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) return bz_fail(nullptr, BZ_ERR_COMM, g_nccl.GetErrorString(r));
    memcpy(id128, &id, 128);
    return BZ_OK;
}

extern "C" int bz_comm_init(bz_ctx* ctx, const void* id128, int nranks, int rank) {
    if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return BZ_ERR_ARG;
    if (!nccl_load()) return bz_fail(ctx, BZ_ERR_COMM, g_nccl.err);
    BZ_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm;
    ncclResult_t r = g_nccl.CommInitRank(&comm, nranks, id, rank);
    if (r != ncclSuccess) return bz_fail(ctx, BZ_ERR_COMM, g_nccl.GetErrorString(r));
    ctx->nccl_comm = comm;
    ctx->nranks = nranks;
    ctx->rank = rank;
    return BZ_OK;
}

extern "C" int bz_comm_destroy(bz_ctx* ctx) {
    if (!ctx) return BZ_OK;
    if (ctx->nccl_comm && g_nccl.lib) {
        g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    ctx->nranks = 1;
    ctx->rank = 0;
    return BZ_OK;
}

int bz_comm_allreduce(bz_ctx* ctx, double* dev_buf, long long count, int op) {
    if (ctx->nranks <= 1) return BZ_OK;
    if (!ctx->nccl_comm) return bz_fail(ctx, BZ_ERR_COMM, "communicator not initialised");
    ncclResult_t r = g_nccl.AllReduce(dev_buf, dev_buf, (size_t)count, ncclFloat64, (ncclRedOp_t)op,
                                      (ncclComm_t)ctx->nccl_comm, ctx->stream);
    ctx->n_launches++;
    if (r != ncclSuccess) return bz_fail(ctx, BZ_ERR_COMM, g_nccl.GetErrorString(r));
    return BZ_OK;
}

extern "C" int bz_comm_bench_allreduce(bz_ctx* ctx, int64_t count, int reps, double* us_per_call) {
    if (!ctx || !us_per_call || count < 1 || reps < 1) return BZ_ERR_ARG;
    BZ_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->nccl_comm) return bz_fail(ctx, BZ_ERR_COMM, "communicator not initialised");
    double* buf = nullptr;
    BZ_CUDA(ctx, cudaMalloc((void**)&buf, (size_t)count * sizeof(double)));
    cudaMemsetAsync(buf, 0, (size_t)count * sizeof(double), ctx->stream);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int rc = BZ_OK;
    for (int i = 0; i < 3 && rc == BZ_OK; ++i) rc = bz_comm_allreduce(ctx, buf, count, BZ_COMM_SUM);   // warm-up
    cudaEventRecord(e0, ctx->stream);
    for (int i = 0; i < reps && rc == BZ_OK; ++i) rc = bz_comm_allreduce(ctx, buf, count, BZ_COMM_SUM);
    cudaEventRecord(e1, ctx->stream);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(buf);
    if (rc != BZ_OK) return rc;
    if (e != cudaSuccess) return bz_fail(ctx, BZ_ERR_CUDA, cudaGetErrorString(e));
    *us_per_call = 1e3 * (double)ms / reps;
    return BZ_OK;
}

int bz_shard_alloc(bz_ctx* ctx, BzShard* sh, long long n) {
    sh->n = n;
    BZ_CUDA(ctx, cudaMalloc((void**)&sh->reduced, (size_t)(n + 8) * sizeof(double)));
    BZ_CUDA(ctx, cudaMemsetAsync(sh->reduced, 0, (size_t)(n + 8) * sizeof(double), ctx->stream));
    return BZ_OK;
}
void bz_shard_free(bz_ctx* ctx, BzShard* sh) {
    (void)ctx;
    if (sh->reduced) cudaFree(sh->reduced);
    sh->reduced = nullptr;
}

// out[j] = sum_p P[p][j] (fixed order); out[n] = sum_p lxpart[p]
__global__ void __launch_bounds__(256) k_local_reduce(const double* __restrict__ P, int np,
                                                      const double* __restrict__ lxpart, int nlx,
                                                      long long n, double* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        double s = 0.0;
        for (int p = 0; p < np; ++p) s += __ldcg(P + (size_t)p * n + j);
        out[j] = s;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double s = 0.0;
        for (int p = 0; p < nlx; ++p) s += __ldcg(lxpart + p);
        out[n] = s;
    }
}

int bz_shard_allreduce(bz_ctx* ctx, BzShard* sh, const double* P, int np, const double* lxpart, int nlx,
                       long long n) {
    k_local_reduce<<<bz_vec_grid(ctx, n), 256, 0, ctx->stream>>>(P, np, lxpart, nlx, n, sh->reduced);
    ctx->n_launches++;
    BZ_CUDA(ctx, cudaGetLastError());
    return bz_comm_allreduce(ctx, sh->reduced, n + 8, BZ_COMM_SUM);
}

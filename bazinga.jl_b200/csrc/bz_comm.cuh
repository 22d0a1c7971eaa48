// Row-sharded multi-GPU support: one process per GPU, NCCL over NVLink for the single
// exchange step of the path -- the allreduce of the n-vector A_g' w_g with the scalar partial appended.
// NCCL is loaded lazily with dlopen so that the library has no link-time dependency on it.
#pragma once
#include "bz_internal.cuh"

struct BzShard {
    double* reduced = nullptr;   // n + 8 doubles: [A' w (n) | sum (t-s)^2/mu | pad]
    long long n = 0;
};

enum BzCommOp { BZ_COMM_SUM = 0, BZ_COMM_MAX = 2 };

int bz_shard_alloc(bz_ctx* ctx, BzShard* sh, long long n);
void bz_shard_free(bz_ctx* ctx, BzShard* sh);
// local fixed-order reduction of the per-cluster partial rows, then allreduce(sum) over ranks
int bz_shard_allreduce(bz_ctx* ctx, BzShard* sh, const double* P, int np, const double* lxpart, int nlx,
                       long long n);
int bz_comm_allreduce(bz_ctx* ctx, double* dev_buf, long long count, int op);

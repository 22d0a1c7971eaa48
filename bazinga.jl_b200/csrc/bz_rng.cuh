// Counter-based synthetic data generator, bit-for-bit identical to oracle/bazinga_oracle.py
// (philox4x32_10 / bz_randn_like): Philox-4x32-10 keyed by (seed, stream), counter = element index,
// value = (u0+u1+u2+u3 - 2*(2^32-1)) * (sqrt(3) * 2^-32).  The integer sum is exact and the single
// multiply is correctly rounded on both sides, so host and device agree exactly.
#pragma once
#include <stdint.h>

__host__ __device__ inline void bz_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1,
                                                 uint32_t out[4]) {
    uint32_t c2 = 0, c3 = 0;
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__host__ __device__ inline double bz_randn_at(uint64_t seed, uint64_t stream, uint64_t idx) {
    uint32_t u[4];
    bz_philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)seed, (uint32_t)stream, u);
    const int64_t ssum = (int64_t)((uint64_t)u[0] + u[1] + u[2] + u[3]) - (int64_t)(2ull * 0xFFFFFFFFull);
    return (double)ssum * 0x1.bb67ae8584caap-32;
}

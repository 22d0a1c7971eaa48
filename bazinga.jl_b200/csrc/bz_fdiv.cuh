// Division by a reused divisor in three dependent FMAs (Markstein), shared by the batched tiny-instance kernel
// (bz_batched.cu) and the fused obstacle kernel (bz_obfused.cu).  Same bits as a / b.
#pragma once
#include <cuda_runtime.h>

// a / b given rb = frecip(b), for divisors that are reused many times (mu_i over an inner solve, ys_i over the life of
// an L-BFGS pair, gamma between halvings).  q = RN(a*rb), the exact remainder r = a - b*q (one FMA) and q' = RN(q + r*rb)
// give the correctly rounded quotient (Markstein's division-by-FMA theorem; tests/test_host_logic.py runs this very
// routine, restated in C, against the hardware division on 2e8 operand pairs, adversarial significands and special
// values included), i.e. the very bits of a / b, in 3 dependent fp64 operations and ONE range test.  A batch runs as long
// as its slowest instance -- a lone thread of a lone warp that pays every instruction's full latency -- so what this buys
// is instructions per division, not throughput (r1j: slowest mpvca instance alone 51.1 -> 22.0 ms, either-or 65.0 ->
// 29.7 ms).  The range test is on q' alone: frecip() poisons rb (NaN) unless 2^-256 <= |b| < 2^256, so a q' with
// 2^-512 <= |q'| < 2^512 proves 2^-769 < |a| < 2^769: q, the remainder (whose last bit is >= 2^-875) and q' are all
// normal, and the product r*rb is never rounded on its own inside the FMA.  Ahead of it, the two numerators that are
// common and need no arithmetic: +-0 (three in four of the AL gradient's (c(x) + mu y - s)_i vanish on an inactive
// constraint) and NaN (a diverged instance keeps iterating to its caps and is the slowest of its batch; the IEEE
// routine's special-value path is its longest).  Everything else is rare and out of line.
template <int E>
__device__ __forceinline__ bool fmid(double v) {          // 2^-E <= |v| < 2^E (false for 0, Inf, NaN); E a power of two
    return (unsigned)((__double2hiint(v) & 0x7fffffff) - ((1023 - E) << 20)) < ((2u * E) << 20);
}
__device__ __forceinline__ double frecip(double b) {
    return fmid<256>(b) ? 1.0 / b : __longlong_as_double(0x7ff8000000000000LL);
}
static __device__ __noinline__ double fdiv_rare(double a, double b, double rb) {
    const double aa = fabs(a);
    if (rb == rb ? !(aa > 0.0 && aa < __longlong_as_double(0x7ff0000000000000LL)) : (b != b))
        return a * rb;                                    // +-Inf over a good b; a NaN divisor
    return a / b;                                         // tiny or huge operands: the IEEE routine
}
__device__ __forceinline__ double fdiv(double a, double b, double rb) {
    const double q = a * rb;
    if (a != a || (a == 0.0 && q == 0.0)) return q;       // NaN; +-0 over a good b (q != 0 there means rb is poisoned)
    const double q2 = fma(fma(-b, q, a), rb, q);
    if (fmid<512>(q2)) return q2;
    return fdiv_rare(a, b, rb);
}


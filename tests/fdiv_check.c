/* Test infrastructure: C restatement of fdiv()/frecip()/fmid() from bazinga.jl_b200/csrc/bz_batched.cu, checked bit for
 * bit against the hardware division.  Usage: fdiv_check <pairs-per-thread>; prints "pairs N bad B".
 * Build: gcc -O2 -ffp-contract=off [-mfma] -fopenmp fdiv_check.c -lm */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t mix(uint64_t* s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline uint64_t bits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static inline double from_bits(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }

static inline int fmid(double v, uint32_t E) {
    uint32_t hi = (uint32_t)(bits(v) >> 32);
    return (uint32_t)((hi & 0x7fffffffu) - ((1023u - E) << 20)) < ((2u * E) << 20);
}
static inline double frecip(double b) { return fmid(b, 256) ? 1.0 / b : from_bits(0x7ff8000000000000ULL); }
static inline double fdiv_rare(double a, double b, double rb) {
    const double aa = fabs(a);
    if (rb == rb ? !(aa > 0.0 && aa < INFINITY) : (b != b)) return a * rb;
    return a / b;
}
static inline double fdiv(double a, double b, double rb, int* fast) {
    const double q = a * rb;
    if (a != a || (a == 0.0 && q == 0.0)) return q;
    const double q2 = fma(fma(-b, q, a), rb, q);
    if (fmid(q2, 512)) { *fast += 1; return q2; }
    return fdiv_rare(a, b, rb);
}
static inline int same(double x, double y) { return (x != x && y != y) || bits(x) == bits(y); }

int main(int argc, char** argv) {
    const long long per = argc > 1 ? atoll(argv[1]) : 10000000LL;
    long long bad = 0, tot = 0, fastn = 0;
    static const double special[] = {0.0, -0.0, INFINITY, -INFINITY, NAN, 1.0, -1.0, 5e-324, 1.7976931348623157e308,
                                     2.2250738585072014e-308, 0x1p-256, 0x1p256, 0x1.fffffffffffffp255, 0x1p-257,
                                     0x1p-512, 0x1p512, 0x1p-768, 0x1p767, 0x1.fffffffffffffp767, 0x1p-769, 0x1p-1000};
    const int ns = (int)(sizeof(special) / sizeof(special[0]));
    for (int i = 0; i < ns; ++i)
        for (int j = 0; j < ns; ++j) {
            int f = 0;
            const double a = special[i], b = special[j];
            if (!same(fdiv(a, b, frecip(b), &f), a / b)) { ++bad; printf("BAD special a=%a b=%a\n", a, b); }
            ++tot;
        }
#pragma omp parallel reduction(+ : bad, tot, fastn)
    {
        uint64_t s = 0x1234567ULL;
#ifdef _OPENMP
        extern int omp_get_thread_num(void);
        s += 7919ULL * (uint64_t)omp_get_thread_num();
#endif
        for (long long i = 0; i < per; ++i) {
            uint64_t ma = mix(&s), mb = mix(&s), e = mix(&s);
            const int mode = (int)(i & 15);
            if (mode == 1) mb |= 0xFFFFFFFF00000ULL;            /* divisor significand near all ones */
            if (mode == 2) mb &= 0xFFFFFULL;                    /* ... near 1.0 */
            if (mode == 3) ma |= 0xFFFFFFFFF0000ULL;
            if (mode == 4) ma &= 0xFFFFULL;
            if (mode == 5) mb = 0xFFFFFFFFFFFFFULL - (mb & 0xFF);
            int ea, eb;
            if (mode < 10) { ea = (int)(e % 81) - 40; eb = (int)((e >> 20) % 81) - 40; }          /* everyday magnitudes */
            else if (mode < 13) { ea = (int)(e % 1601) - 800; eb = (int)((e >> 20) % 561) - 280; } /* across the range tests */
            else { ea = (int)(e % 2046) - 1022; eb = (int)((e >> 20) % 2046) - 1022; }            /* the whole format */
            double a = from_bits(((uint64_t)(1023 + ea) << 52) | (ma & 0xFFFFFFFFFFFFFULL) | (ma & 0x8000000000000000ULL));
            double b = from_bits(((uint64_t)(1023 + eb) << 52) | (mb & 0xFFFFFFFFFFFFFULL) | (mb & 0x8000000000000000ULL));
            if (mode == 15 && (e & (1ULL << 50))) a = from_bits(ma & 0x800FFFFFFFFFFFFFULL);       /* subnormal numerator */
            int f = 0;
            const double got = fdiv(a, b, frecip(b), &f), want = a / b;
            if (!same(got, want)) { if (++bad < 5) printf("BAD a=%a b=%a got %a want %a\n", a, b, got, want); }
            ++tot; fastn += f;
        }
    }
    printf("pairs %lld bad %lld fast %lld\n", tot, bad, fastn);
    return bad != 0;
}

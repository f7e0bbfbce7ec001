/* oracle/hostgen.c -- TEST / BENCH INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Host-side twin of the device generator of synthetic "generic data" (bioen_b200/csrc/bioen_b200.cu, k_generate;
 * NumPy restatement: tests/util_rng.py):  yTilde[i][j] = a_i + b * z(seed, i, col0 + j),  z ~ N(0,1) from a
 * counter-based hash (splitmix64 finaliser) and Box-Muller.  bench.py's reference arm (--impl reference) uses it
 * to build, without a GPU, the same N = 1e6 x M = 1e3 problem the B200 arm generates on the device, so the two
 * arms minimise the same function.  Entries agree with the device's to the last 1-2 ulp (libm cos/log vs CUDA's
 * cospi/log), which is far below anything the comparison of the optima looks at.
 *
 * OpenMP over rows: 1e9 entries take a few seconds on the GPU box's host cores.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

/* cos(pi * x) for x in [0, 2) with an exact argument reduction to |r| <= 1/4 */
static inline double cospi_reduced(double x) {
    const double k = nearbyint(2.0 * x);          /* 0 .. 4 */
    const double r = x - 0.5 * k;                 /* exact */
    const double pr = 3.14159265358979323846 * r;
    switch ((int)k & 3) {
        case 0: return cos(pr);
        case 1: return -sin(pr);
        case 2: return -cos(pr);
        default: return sin(pr);
    }
}

/* rows [0, m) x columns [col0, col0 + n) of the global matrix into Y (row stride ld doubles) */
void hostgen_generic_ytilde(double* Y, size_t ld, int m, long long n, unsigned long long seed, long long col0,
                            const double* a, double b) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < m; ++i) {
        double* row = Y + (size_t)i * ld;
        for (long long j = 0; j < n; ++j) {
            const uint64_t ctr = ((uint64_t)i << 40) + (uint64_t)(col0 + j);
            const uint64_t h1 = mix64(seed + 0x9E3779B97F4A7C15ULL * (ctr + 1));
            const uint64_t h2 = mix64(h1 + 0x9E3779B97F4A7C15ULL);
            const double u1 = ((double)(h1 >> 11) + 1.0) * 0x1.0p-53; /* (0, 1] */
            const double u2 = (double)(h2 >> 11) * 0x1.0p-53;         /* [0, 1) */
            const double z = sqrt(-2.0 * log(u1)) * cospi_reduced(2.0 * u2);
            row[j] = fma(b, z, a[i]);
        }
    }
}

/* oracle/bioen_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by, or called from the
 * product (bioen_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 *
 * A scalar, single-threaded, plain-C restatement of the arithmetic of BioEn's optimisation hot path: the
 * log-posterior and its gradient for the log-weights method and for the forces method.  It is written from
 * the formulas, not from the reference's text; each function cites the reference lines it restates
 * (paths relative to /root/reference/bioen/optimize/ext/).  Every reduction is a left-to-right sequential
 * sum, which is what the reference does in its reproducible mode (`_fast_openmp_flag == 0`).
 *
 * Parity status: PINNED.  tests/test_oracle_cpu.py checks this file (a) against the golden vectors in
 * tests/golden/*.npz, which were produced here by running the unmodified reference sources
 * (oracle/_ref/libbioen_ref.so, tests/golden/make_golden.py), and (b) live against that library when it is
 * present.
 *
 * Layout conventions (same as the reference): yTilde is M x N row-major (yTilde[i*n + j]); g, G, w, w0 have
 * N entries; YTilde, forces have M entries.
 */
#include <float.h>
#include <math.h>
#include <stddef.h>

/* softmax weights, un-stabilised exactly like c_bioen_kernels_logw.c:55-94: w_j = exp(g_j) / sum_k exp(g_k).
 * Returns s = sum_k exp(g_k). */
double oracle_logw_weights(const double* g, double* w, size_t n) {
    double s = 0.0;
    for (size_t j = 0; j < n; ++j) {
        w[j] = exp(g[j]);
        s += w[j];
    }
    const double inv = 1.0 / s;
    for (size_t j = 0; j < n; ++j) w[j] *= inv;
    return s;
}

/* avg_i = sum_j yTilde_ij w_j   (c_bioen_common.c:78-83, c_bioen_kernels_forces.c:93-109) */
void oracle_average(const double* yTilde, const double* w, double* avg, size_t m, size_t n) {
    for (size_t i = 0; i < m; ++i) {
        const double* row = yTilde + i * n;
        double a = 0.0;
        for (size_t j = 0; j < n; ++j) a += row[j] * w[j];
        avg[i] = a;
    }
}

/* 1/2 sum_i (avg_i - YTilde_i)^2   (c_bioen_common.c:70-108) */
static double chi_squared(const double* avg, const double* YTilde, size_t m) {
    double c = 0.0;
    for (size_t i = 0; i < m; ++i) {
        const double r = avg[i] - YTilde[i];
        c += r * r;
    }
    return 0.5 * c;
}

/* L(g) = theta * ( sum_j (g_j - G_j) w_j - log s + log s0 ) + chi^2/2
 * (c_bioen_kernels_logw.c:96-147; s0 = sum_j exp(G_j) is recomputed per call there, line 122).
 * scratch: w[n], avg[m].  Returns L. */
double oracle_logw_objective(const double* g, const double* G, const double* yTilde, const double* YTilde,
                             double theta, double* w, double* avg, size_t m, size_t n) {
    const double s = oracle_logw_weights(g, w, n);
    double s0 = 0.0, dot = 0.0;
    for (size_t j = 0; j < n; ++j) s0 += exp(G[j]);
    for (size_t j = 0; j < n; ++j) dot += (g[j] - G[j]) * w[j];
    const double prior = theta * (dot - log(s) + log(s0));
    oracle_average(yTilde, w, avg, m, n);
    return prior + chi_squared(avg, YTilde, m);
}

/* dL/dg_j = w_j theta (g_j - <g> - G_j + <G>) + w_j sum_i (avg_i - Y_i)(yTilde_ij - avg_i)
 * (c_bioen_kernels_logw.c:151-268; the per-element form r_i*(y_ij - avg_i) is kept, lines 201/248).
 * scratch: w[n], avg[m].  Returns L as well (the f+g evaluation of interface_lbfgs_logw, lines 525-561). */
double oracle_logw_fg(const double* g, const double* G, const double* yTilde, const double* YTilde,
                      double theta, double* grad, double* w, double* avg, size_t m, size_t n) {
    const double L = oracle_logw_objective(g, G, yTilde, YTilde, theta, w, avg, m, n);
    double gbar = 0.0, Gbar = 0.0;
    for (size_t j = 0; j < n; ++j) {
        gbar += g[j] * w[j];
        Gbar += G[j] * w[j];
    }
    for (size_t j = 0; j < n; ++j) grad[j] = 0.0;
    /* accumulate column sums row by row: for every column j the additions happen in the order
     * i = 0..m-1, the same order as the reference's inner loop over i, only interchanged for locality. */
    for (size_t i = 0; i < m; ++i) {
        const double* row = yTilde + i * n;
        const double a = avg[i];
        const double r = a - YTilde[i];
        for (size_t j = 0; j < n; ++j) grad[j] += r * (row[j] - a);
    }
    for (size_t j = 0; j < n; ++j)
        grad[j] = w[j] * theta * (g[j] - gbar - G[j] + Gbar) + w[j] * grad[j];
    return L;
}

/* w_j = w0_j exp(x_j - max x) / sum_k(...), x_j = sum_i f_i yTilde_ij   (note the PLUS sign)
 * (c_bioen_kernels_forces.c:111-224).  scratch: x[n]. */
void oracle_forces_weights(const double* w0, const double* yTilde, const double* forces, double* w,
                           double* x, size_t m, size_t n) {
    for (size_t j = 0; j < n; ++j) x[j] = 0.0;
    for (size_t i = 0; i < m; ++i) {
        const double* row = yTilde + i * n;
        const double f = forces[i];
        for (size_t j = 0; j < n; ++j) x[j] += f * row[j];
    }
    double xmax = -DBL_MAX;
    for (size_t j = 0; j < n; ++j) xmax = xmax > x[j] ? xmax : x[j];
    double s = 0.0;
    for (size_t j = 0; j < n; ++j) {
        w[j] = w0[j] * exp(x[j] - xmax);
        s += w[j];
    }
    const double inv = 1.0 / s;
    for (size_t j = 0; j < n; ++j) w[j] = inv * w[j];
}

/* guarded log-ratio used by both the forces objective and gradient (c_bioen_kernels_forces.c:250,263,324) */
static inline int ratio_defined(double w, double w0) { return (w >= DBL_MIN) && (w0 >= DBL_MIN); }

/* L(f) = theta * sum_j w_j (log w_j - log w0_j) [guarded] + chi^2/2   (c_bioen_kernels_forces.c:227-277)
 * scratch: w[n], x[n], avg[m]. */
double oracle_forces_objective(const double* forces, const double* w0, const double* yTilde,
                               const double* YTilde, double theta, double* w, double* x, double* avg,
                               size_t m, size_t n) {
    oracle_forces_weights(w0, yTilde, forces, w, x, m, n);
    oracle_average(yTilde, w, avg, m, n);
    const double chi = chi_squared(avg, YTilde, m);
    double kl = 0.0;
    for (size_t j = 0; j < n; ++j)
        kl += ratio_defined(w[j], w0[j]) ? (log(w[j]) - log(w0[j])) * w[j] : 0.0;
    return kl * theta + chi;
}

/* dL/df_i = sum_j (yTilde_ij - avg_i) E_j,  E_j = (theta (1 + [log w_j - log w0_j]) + t_j) w_j,
 * t_j = sum_i yTilde_ij (avg_i - Y_i)     (c_bioen_kernels_forces.c:280-340).
 * scratch: w[n], x[n] (reused for t and E), avg[m].  Returns L (interface_lbfgs_forces, lines 43-76). */
double oracle_forces_fg(const double* forces, const double* w0, const double* yTilde, const double* YTilde,
                        double theta, double* grad, double* w, double* x, double* avg, size_t m, size_t n) {
    const double L = oracle_forces_objective(forces, w0, yTilde, YTilde, theta, w, x, avg, m, n);
    double* t = x;
    for (size_t j = 0; j < n; ++j) t[j] = 0.0;
    for (size_t i = 0; i < m; ++i) {
        const double* row = yTilde + i * n;
        const double r = avg[i] - YTilde[i];
        for (size_t j = 0; j < n; ++j) t[j] += row[j] * r;
    }
    for (size_t j = 0; j < n; ++j) {
        double d = 1.0;
        if (ratio_defined(w[j], w0[j])) d += log(w[j]) - log(w0[j]);
        t[j] = (d * theta + t[j]) * w[j];
    }
    for (size_t i = 0; i < m; ++i) {
        const double* row = yTilde + i * n;
        const double a = avg[i];
        double d = 0.0;
        for (size_t j = 0; j < n; ++j) d += (row[j] - a) * t[j];
        grad[i] = d;
    }
    return L;
}

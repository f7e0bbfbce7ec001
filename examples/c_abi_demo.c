/* c_abi_demo.c -- calling libbioen_b200.so from plain C, both ways (include/bioen_b200.h):
 *   part 1: the reference's own entry point _opt_lbfgs_logw (host pointers in, result out)
 *   part 2: the resident handle (upload once, evaluate / minimise many times)
 * Build:  gcc -O2 -Iinclude examples/c_abi_demo.c -o c_abi_demo -Lbioen_b200/lib -lbioen_b200 -lm \
 *             -Wl,-rpath,$PWD/bioen_b200/lib
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "bioen_b200.h"

static double urand(unsigned long long *s) {   /* xorshift: a self-contained toy data generator */
    *s ^= *s << 13; *s ^= *s >> 7; *s ^= *s << 17;
    return (double)(*s >> 11) * (1.0 / 9007199254740992.0);
}

int main(void) {
    const int m = 40, n = 5000;
    const double theta = 10.0;
    unsigned long long seed = 88172645463325252ULL;
    double *yT = malloc(sizeof(double) * m * n), *Y = malloc(sizeof(double) * m);
    double *g = calloc(n, sizeof(double)), *G = calloc(n, sizeof(double));
    double *w = malloc(sizeof(double) * n), *res = malloc(sizeof(double) * n), *grad = malloc(sizeof(double) * n);
    for (int i = 0; i < m; ++i) {
        const double mu = 2.0 * urand(&seed) - 1.0;
        Y[i] = 2.0 * (mu + 0.3 * (urand(&seed) - 0.5));
        for (int j = 0; j < n; ++j) yT[(size_t)i * n + j] = 2.0 * (mu + (urand(&seed) + urand(&seed) + urand(&seed) - 1.5));
    }
    if (bioen_b200_device_count() < 1) {
        fprintf(stderr, "no CUDA device: %s\n", bioen_b200_last_error());
        return 2;
    }

    /* ---- part 1: the reference's C driver, exactly as bioen/optimize/ext/c_bioen.pyx:495-511 fills it */
    params_t p = {0};
    p.g = g; p.G = G; p.yTilde = yT; p.YTilde = Y; p.w = w; p.result = res; p.theta = theta; p.m = m; p.n = n;
    lbfgs_config_params cfg = {2, 5000, 1e-6, 1e-6, 1e-5, 0.9, 0.9, 10, 100};
    visual_params vis = {0, 0};
    int err = 0;
    const double fmin1 = _opt_lbfgs_logw(p, cfg, vis, &err);
    printf("part 1  _opt_lbfgs_logw      : fmin = %.10f  code = %d (%s)\n", fmin1, err, lbfgs_strerror(err));
    if (bioen_b200_error_pending()) { fprintf(stderr, "%s\n", bioen_b200_last_error()); return 1; }

    /* ---- part 2: resident problem */
    bioen_b200_ctx *ctx = bioen_b200_create(m, n, 0);
    if (!ctx || bioen_b200_upload_ytilde(ctx, yT, n) || bioen_b200_set_logw(ctx, G, Y, theta)) {
        fprintf(stderr, "%s\n", bioen_b200_last_error());
        return 1;
    }
    double f0 = 0.0, fmin2 = 0.0;
    int info[4];
    bioen_b200_eval(ctx, BIOEN_B200_LOGW, g, &f0, grad);
    const int code = bioen_b200_opt_lbfgs(ctx, BIOEN_B200_LOGW, g, res, cfg, vis, &fmin2, info);
    printf("part 2  bioen_b200_opt_lbfgs : f0 = %.10f  fmin = %.10f  code = %d  iterations = %d  evaluations = %d\n",
           f0, fmin2, code, info[0], info[1]);
    bioen_b200_destroy(ctx);
    const int ok = (err == code) && fabs(fmin1 - fmin2) <= 1e-12 * fabs(fmin1) && fmin2 < f0;
    printf("%s\n", ok ? "c_abi_demo ok" : "c_abi_demo MISMATCH");
    free(yT); free(Y); free(g); free(G); free(w); free(res); free(grad);
    return ok ? 0 : 1;
}

/* ref_loops.c -- C restatement of the reference's two scalar Julia hot loops, for the CPU baseline
 * timing only (TEST / BENCH INFRASTRUCTURE, never linked into the product).
 *
 *   cov_seiso_sym     src/kernels/kernels.jl:39-50 (for j, for i<j: cK[i,j]=cov_ij; mirror) with
 *                     cov_ij -> cov(SEIso, distij(SqEuclidean)) : src/kernels/stationary.jl:25-27,
 *                     src/kernels/distance.jl:43-49, src/kernels/se_iso.jl:39
 *   dmll_kern_seiso   src/GPE.jl:219-241 with dKij_dθ! -> dk_dll/dk_dlσ : se_iso.jl:41-50,
 *                     stationary.jl:28
 * Single-threaded, exactly like the reference (no Threads.@threads anywhere in src/).
 * x is Julia's d x N column-major matrix; K and A are N x N column-major.                      */
#include <math.h>
#include <stddef.h>

static inline double sqeuclid(const double* x, long i, long j, int d) {
    double s = 0.0;
    for (int k = 0; k < d; ++k) { double df = x[i * d + k] - x[j * d + k]; s += df * df; }
    return s;
}

void cov_seiso_sym(double* K, const double* x, int d, long N, double l2, double s2) {
    for (long j = 0; j < N; ++j) {
        K[j + j * N] = s2 * exp(-0.5 * sqeuclid(x, j, j, d) / l2);
        for (long i = 0; i < j; ++i) {
            double v = s2 * exp(-0.5 * sqeuclid(x, i, j, d) / l2);
            K[i + j * N] = v;
            K[j + i * N] = v;
        }
    }
}

void dmll_kern_seiso(double* dmll, const double* A, const double* x, int d, long N, double l2, double s2) {
    double g0 = 0.0, g1 = 0.0;
    for (long j = 0; j < N; ++j) {
        {
            double r = sqeuclid(x, j, j, d);
            double k = s2 * exp(-0.5 * r / l2);
            g0 += (r / l2 * k) * A[j + j * N] / 2.0;
            g1 += (2.0 * k) * A[j + j * N] / 2.0;
        }
        for (long i = j + 1; i < N; ++i) {
            double r = sqeuclid(x, i, j, d);
            double k = s2 * exp(-0.5 * r / l2);
            g0 += (r / l2 * k) * A[i + j * N];
            g1 += (2.0 * k) * A[i + j * N];
        }
    }
    dmll[0] = g0;
    dmll[1] = g1;
}

// potrf_base.cu -- the latency-critical leaf of the blocked Cholesky: one CTA factors a 128 x 128
// diagonal tile and inverts its factor.
//
// Replaces, at tile granularity, LAPACK dpotrf('U') behind cholesky!(Symmetric(F,:U))
// (/root/reference/src/GP.jl:110).  Outputs per tile at diagonal offset p:
//   F[p.., p..]      lower triangle  <- L_dd            (K_y = L L', L = U')
//   Dinv [p.., 0..128)               <- W = L_dd^-1     (clean lower-triangular tile)
//   DinvT[p.., 0..128)               <- W'              (clean upper-triangular tile)
//   logd[p + j]                      <- log(d_j) = 2 log L_jj      (logdet = sum, GPE.jl:210)
//   info                             <- p + j + 1 of the first non-positive pivot (atomicMin)
//
// Design: the tile lives in REGISTERS (512 threads x 32 doubles: thread (warp w, lane l) owns rows
// w+16a, columns l+32c).  Step j publishes the unscaled column j to shared memory, one
// __syncthreads, then every thread applies the rank-1 update to its registers.  The published
// columns stay in shared memory (lower triangle of S) and drive the second phase, a right-looking
// triangular inversion whose rows W[i,:] are published into the strict upper triangle of the same
// array (S[n][i+1]).  One barrier per step, 256 steps per tile, no shared-memory bank conflicts
// (row stride 129 doubles).
#include "potrf_base.cuh"
#include <math.h>

namespace {

constexpr int T = 128;
constexpr int LDS = 129;
constexpr int NTH = 512;

__global__ void __launch_bounds__(NTH, 1)
potrf128_inv_kernel(const double* __restrict__ G, long long ldg, double* __restrict__ F, long long ldf,
                    double* __restrict__ Dinv, double* __restrict__ DinvT, double* __restrict__ logd,
                    int* __restrict__ info, int p0, int tile_stride) {
    extern __shared__ double S[];                 // [128][129] + rs[128]
    double* rs = S + T * LDS;
    const int p = p0 + blockIdx.x * tile_stride;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;

    double a[8][4];
#pragma unroll
    for (int ia = 0; ia < 8; ++ia) {
        const int m = w + 16 * ia;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int n = l + 32 * c;
            a[ia][c] = (n <= m) ? G[(long long)(p + m) * ldg + p + n] : 0.0;
        }
    }

    // ---- phase 1: right-looking Cholesky, unscaled columns published to S[m][j] (m >= j) ----
#pragma unroll 1
    for (int j = 0; j < T; ++j) {
        const int jc = j >> 5, jl = j & 31;
        if (l == jl) {
#pragma unroll
            for (int ia = 0; ia < 8; ++ia) {
                const int m = w + 16 * ia;
                double v = a[ia][0];
#pragma unroll
                for (int c = 1; c < 4; ++c) if (c == jc) v = a[ia][c];
                if (m >= j) S[m * LDS + j] = v;
            }
        }
        __syncthreads();
        double d = S[j * LDS + j];
        if (!(d > 0.0)) {                         // not positive definite (or NaN): record, keep going
            if (threadIdx.x == 0) atomicMin(info, p + j + 1);
            d = 1.0;
        }
        const double invd = __drcp_rn(d);
        if (threadIdx.x == 0) rs[j] = d;          // pivot; turned into 1/sqrt(d) after the loop
        // column values this lane needs (live chunks only: 32c+31 > j)
        double cn[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) cn[c] = (32 * c + 31 > j) ? S[(l + 32 * c) * LDS + j] : 0.0;
#pragma unroll
        for (int ia = 0; ia < 8; ++ia) {
            const int m = w + 16 * ia;                 // warp-uniform
            if (m > j) {
                const double lm = S[m * LDS + j] * invd;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (32 * c + 31 > j && 32 * c <= m) {      // warp-uniform: chunk has live columns
                        const int n = l + 32 * c;
                        if (n > j && n <= m) a[ia][c] -= lm * cn[c];
                    }
                }
            }
        }
    }
    __syncthreads();   // pivots + all columns published
    if (threadIdx.x < T) {
        const double d = rs[threadIdx.x];
        logd[p + threadIdx.x] = log(d);
        rs[threadIdx.x] = 1.0 / sqrt(d);
    }
    __syncthreads();

    // ---- write L = S .* rs (lower triangle) ----
    for (int idx = threadIdx.x; idx < T * T; idx += NTH) {
        const int m = idx >> 7, j = idx & 127;
        if (j <= m) F[(long long)(p + m) * ldf + p + j] = S[m * LDS + j] * rs[j];
    }

    // ---- phase 2: W = L^-1 by right-looking elimination on R (init I), rows published to S[n][i+1] ----
#pragma unroll
    for (int ia = 0; ia < 8; ++ia) {
        const int m = w + 16 * ia;
#pragma unroll
        for (int c = 0; c < 4; ++c) a[ia][c] = (l + 32 * c == m) ? 1.0 : 0.0;
    }
#pragma unroll 1
    for (int i = 0; i < T; ++i) {
        const double rsi = rs[i];
        if (w == (i & 15)) {
            const int a0 = i >> 4;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int n = l + 32 * c;
                double v = a[0][c];
#pragma unroll
                for (int ia = 1; ia < 8; ++ia) if (ia == a0) v = a[ia][c];
                if (n <= i) S[n * LDS + i + 1] = v * rsi;         // W[i][n]
            }
        }
        __syncthreads();
        double wn[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) wn[c] = (32 * c <= i && l + 32 * c <= i) ? S[(l + 32 * c) * LDS + i + 1] : 0.0;
#pragma unroll
        for (int ia = 0; ia < 8; ++ia) {
            const int m = w + 16 * ia;
            if (m > i) {
                const double lmi = S[m * LDS + i] * rsi;          // L[m][i]
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (32 * c <= i) a[ia][c] -= lmi * wn[c];     // wn == 0 for n > i
            }
        }
    }
    __syncthreads();

    // ---- write W (clean lower) and W' (clean upper) ----
    for (int idx = threadIdx.x; idx < T * T; idx += NTH) {
        const int r = idx >> 7, c = idx & 127;
        // Dinv[r][c] = W[r][c] = S[c][r+1] for c <= r
        Dinv[(long long)(p + r) * T + c] = (c <= r) ? S[c * LDS + r + 1] : 0.0;
        // DinvT[r][c] = W[c][r] = S[r][c+1] for c >= r
        DinvT[(long long)(p + r) * T + c] = (c >= r) ? S[r * LDS + c + 1] : 0.0;
    }
}

bool g_attr_set = false;

}  // namespace

cudaError_t potrf128_launch(const double* G, int64_t ldg, double* F, int64_t ldf, double* Dinv, double* DinvT,
                            double* logd, int* info, int p0, int ntiles, int tile_stride, cudaStream_t st) {
    const size_t sm = (size_t)(T * LDS + T) * sizeof(double);
    if (!g_attr_set) {
        cudaError_t e = cudaFuncSetAttribute(potrf128_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
        g_attr_set = true;
    }
    potrf128_inv_kernel<<<ntiles, NTH, sm, st>>>(G, ldg, F, ldf, Dinv, DinvT, logd, info, p0, tile_stride);
    return cudaGetLastError();
}

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
// l+32a, columns w+16c).  Rows across lanes / columns across warps makes every per-step predicate
// warp-uniform and lets a whole warp publish a column: step j publishes the unscaled column j to
// shared memory (S[m][j], m >= j), one __syncthreads, then every thread applies the rank-1 update
// to its registers (entries above the diagonal are never read and may hold garbage).  Phase 2
// computes W = L^-1 from W L = I by backward column elimination in the same register layout,
// publishing column j of W into the strict upper triangle (S[j][m+1]).  One barrier per step, 256
// steps per tile, no divergent branches, no shared-memory bank conflicts (row stride 129 doubles).
#include "potrf_base.cuh"
#include <math.h>

namespace {

constexpr int T = 128;
constexpr int LDS = 129;
constexpr int NTH = 512;

__global__ void __launch_bounds__(NTH, 1)
potrf128_inv_kernel(const double* __restrict__ G, long long ldg, double* __restrict__ F, long long ldf,
                    double* __restrict__ Dinv, double* __restrict__ DinvT, double* __restrict__ logd,
                    int* __restrict__ info, int p0, int tile_stride, const PotrfPeers peers) {
    extern __shared__ double S[];                 // [128][129] + rs[128]
    double* rs = S + T * LDS;
    const int p = p0 + blockIdx.x * tile_stride;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;

    // coalesced tile load through shared memory, then into the register layout
    for (int idx = threadIdx.x; idx < T * T; idx += NTH) {
        const int m = idx >> 7, n = idx & 127;
        S[m * LDS + n] = (n <= m) ? G[(long long)(p + m) * ldg + p + n] : 0.0;
    }
    __syncthreads();
    double a[4][8];
#pragma unroll
    for (int ia = 0; ia < 4; ++ia)
#pragma unroll
        for (int c = 0; c < 8; ++c) a[ia][c] = S[(l + 32 * ia) * LDS + w + 16 * c];
    __syncthreads();

    // ---- phase 1: right-looking Cholesky, unscaled columns published to S[m][j] (m >= j) ----
#pragma unroll 1
    for (int j = 0; j < T; ++j) {
        if (w == (j & 15)) {                       // warp-uniform: this warp owns column j
            const int cj = j >> 4;
#pragma unroll
            for (int ia = 0; ia < 4; ++ia) {
                double v = a[ia][0];
#pragma unroll
                for (int c = 1; c < 8; ++c) if (c == cj) v = a[ia][c];
                const int m = l + 32 * ia;
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
        double lm[4];
#pragma unroll
        for (int ia = 0; ia < 4; ++ia) {
            const int m = l + 32 * ia;
            lm[ia] = (m > j) ? S[m * LDS + j] * invd : 0.0;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int n = w + 16 * c;
            if (n > j) {                           // warp-uniform
                const double cn = S[n * LDS + j];  // broadcast
#pragma unroll
                for (int ia = 0; ia < 4; ++ia) a[ia][c] = fma(-lm[ia], cn, a[ia][c]);
            }
        }
    }
    __syncthreads();   // pivots + all columns published
    if (threadIdx.x < T) {
        const double d = rs[threadIdx.x];
        const double lg = log(d);
        logd[p + threadIdx.x] = lg;
        for (int q = 0; q < peers.n; ++q) peers.logd[q][p + threadIdx.x] = lg;
        rs[threadIdx.x] = 1.0 / sqrt(d);
    }
    __syncthreads();
    // scale the published columns in place: S[m][j] = L[m][j], and write L (lower triangle)
    for (int idx = threadIdx.x; idx < T * T; idx += NTH) {
        const int m = idx >> 7, j = idx & 127;
        if (j <= m) {
            const double v = S[m * LDS + j] * rs[j];
            S[m * LDS + j] = v;
            F[(long long)(p + m) * ldf + p + j] = v;
            for (int q = 0; q < peers.n; ++q) peers.F[q][(long long)(p + m) * ldf + p + j] = v;
        }
    }
    __syncthreads();

    // ---- phase 2: W L = I, backward over columns; column j of W published to S[j][m+1] (m >= j) ----
#pragma unroll
    for (int ia = 0; ia < 4; ++ia)
#pragma unroll
        for (int c = 0; c < 8; ++c) a[ia][c] = (l + 32 * ia == w + 16 * c) ? 1.0 : 0.0;
#pragma unroll 1
    for (int j = T - 1; j >= 0; --j) {
        if (w == (j & 15)) {
            const int cj = j >> 4;
            const double rsj = rs[j];
#pragma unroll
            for (int ia = 0; ia < 4; ++ia) {
                double v = a[ia][0];
#pragma unroll
                for (int c = 1; c < 8; ++c) if (c == cj) v = a[ia][c];
                const int m = l + 32 * ia;
                if (m >= j) S[j * LDS + m + 1] = v * rsj;          // W[m][j]
            }
        }
        __syncthreads();
        double wm[4];
#pragma unroll
        for (int ia = 0; ia < 4; ++ia) {
            const int m = l + 32 * ia;
            wm[ia] = (m >= j) ? S[j * LDS + m + 1] : 0.0;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int kk = w + 16 * c;
            if (kk < j) {                          // warp-uniform
                const double ljk = S[j * LDS + kk];                // L[j][k], broadcast
#pragma unroll
                for (int ia = 0; ia < 4; ++ia) a[ia][c] = fma(-wm[ia], ljk, a[ia][c]);
            }
        }
    }
    __syncthreads();

    // ---- write W (clean lower) and W' (clean upper) ----
    for (int idx = threadIdx.x; idx < T * T; idx += NTH) {
        const int r = idx >> 7, c = idx & 127;
        // Dinv[r][c] = W[r][c] = S[c][r+1] for c <= r
        const double wl = (c <= r) ? S[c * LDS + r + 1] : 0.0;
        // DinvT[r][c] = W[c][r] = S[r][c+1] for c >= r
        const double wu = (c >= r) ? S[r * LDS + c + 1] : 0.0;
        Dinv[(long long)(p + r) * T + c] = wl;
        DinvT[(long long)(p + r) * T + c] = wu;
        for (int q = 0; q < peers.n; ++q) {
            peers.Dinv[q][(long long)(p + r) * T + c] = wl;
            peers.DinvT[q][(long long)(p + r) * T + c] = wu;
        }
    }
}

bool g_attr_set = false;

}  // namespace

cudaError_t potrf128_launch(const double* G, int64_t ldg, double* F, int64_t ldf, double* Dinv, double* DinvT,
                            double* logd, int* info, int p0, int ntiles, int tile_stride, cudaStream_t st,
                            const PotrfPeers* peers) {
    PotrfPeers pp{};
    if (peers) pp = *peers;
    const size_t sm = (size_t)(T * LDS + T) * sizeof(double);
    if (!g_attr_set) {
        cudaError_t e = cudaFuncSetAttribute(potrf128_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
        g_attr_set = true;
    }
    potrf128_inv_kernel<<<ntiles, NTH, sm, st>>>(G, ldg, F, ldf, Dinv, DinvT, logd, info, p0, tile_stride, pp);
    return cudaGetLastError();
}

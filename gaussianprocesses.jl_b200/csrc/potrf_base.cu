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
                    int* __restrict__ info, int p0, int tile_stride) {
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
        logd[p + threadIdx.x] = log(d);
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
        Dinv[(long long)(p + r) * T + c] = (c <= r) ? S[c * LDS + r + 1] : 0.0;
        // DinvT[r][c] = W[c][r] = S[r][c+1] for c >= r
        DinvT[(long long)(p + r) * T + c] = (c >= r) ? S[r * LDS + c + 1] : 0.0;
    }
}

// ---------------------------------------------------------------------------------------------
// Blocked variant (default): PW-column panels.  The trailing tile stays in registers (same layout);
// per panel: (a) the owning warps publish 8 columns, (b) ONE warp factors the 128 x 8 panel
// warp-synchronously (pivot rows travel by shuffles, no CTA barrier inside), (c) all warps apply a
// rank-8 update from shared memory.  2 CTA barriers per panel instead of 8, and the phase-2 panel
// step is row-local (no shuffles at all).  16 + 16 panels per tile.
// ---------------------------------------------------------------------------------------------
constexpr int PW = 4;          // panel width (4 keeps panel + register tile under 128 registers)
constexpr int LDP = 5;         // padded (odd) row stride of the 128 x PW panel buffers

__global__ void __launch_bounds__(NTH, 1)
potrf128_inv_blocked_kernel(const double* __restrict__ G, long long ldg, double* __restrict__ F, long long ldf,
                            double* __restrict__ Dinv, double* __restrict__ DinvT, double* __restrict__ logd,
                            int* __restrict__ info, int p0, int tile_stride) {
    extern __shared__ double S[];                 // [128][129] | rs[128] | dv[128] | PD[128][LDP] | PWb[128][LDP]
    double* rs = S + T * LDS;
    double* dv = rs + T;
    double* PD = dv + T;
    double* PWb = PD + T * LDP;
    const int p = p0 + blockIdx.x * tile_stride;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;

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

    // ---------------- phase 1: Cholesky (unscaled columns in S, pivots in dv) ----------------
#pragma unroll 1
    for (int kp = 0; kp < T / PW; ++kp) {
        const int c0 = kp * PW;
        {   // (a) publish columns c0..c0+7 (one warp per column)
            const int j = (w - c0) & 15;
            if (j < PW) {
                const int cj = (c0 + j) >> 4;
#pragma unroll
                for (int ia = 0; ia < 4; ++ia) {
                    double v = a[ia][0];
#pragma unroll
                    for (int c = 1; c < 8; ++c) if (c == cj) v = a[ia][c];
                    S[(l + 32 * ia) * LDS + c0 + j] = v;
                }
            }
        }
        __syncthreads();
        if (w == 0) {   // (b) warp-synchronous factorisation of the 128 x 8 panel
            double P[4][PW], invd[PW];
#pragma unroll
            for (int ia = 0; ia < 4; ++ia)
#pragma unroll
                for (int j = 0; j < PW; ++j) P[ia][j] = S[(l + 32 * ia) * LDS + c0 + j];
#pragma unroll
            for (int j = 0; j < PW; ++j) {
                const int pr = c0 + j;
                const int owner = pr & 31, iao = pr >> 5;
                double prow[PW];
#pragma unroll
                for (int jj = j; jj < PW; ++jj) {
                    const double v = (iao == 0) ? P[0][jj] : (iao == 1) ? P[1][jj] : (iao == 2) ? P[2][jj] : P[3][jj];
                    prow[jj] = __shfl_sync(0xffffffffu, v, owner);
                }
                double d = prow[j];
                if (!(d > 0.0)) { if (l == 0) atomicMin(info, p + pr + 1); d = 1.0; }
                invd[j] = __drcp_rn(d);
                if (l == 0) dv[pr] = d;
#pragma unroll
                for (int ia = 0; ia < 4; ++ia) {
                    const int m = l + 32 * ia;
                    const double lf = (m > pr) ? P[ia][j] * invd[j] : 0.0;
#pragma unroll
                    for (int jj = j + 1; jj < PW; ++jj) P[ia][jj] = fma(-lf, prow[jj], P[ia][jj]);
                }
            }
#pragma unroll
            for (int ia = 0; ia < 4; ++ia) {
                const int m = l + 32 * ia;
#pragma unroll
                for (int j = 0; j < PW; ++j) {
                    S[m * LDS + c0 + j] = P[ia][j];
                    PD[m * LDP + j] = P[ia][j] * invd[j];
                }
            }
        }
        __syncthreads();
        {   // (c) rank-8 update of the register tile: a[m][n] -= sum_j S[m][c0+j] * PD[n][j], n >= c0+8
            double lrow[4][PW];
#pragma unroll
            for (int ia = 0; ia < 4; ++ia)
#pragma unroll
                for (int j = 0; j < PW; ++j) lrow[ia][j] = S[(l + 32 * ia) * LDS + c0 + j];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int n = w + 16 * c;
                if (n >= c0 + PW) {                          // warp-uniform
                    double pd[PW];
#pragma unroll
                    for (int j = 0; j < PW; ++j) pd[j] = PD[n * LDP + j];
#pragma unroll
                    for (int ia = 0; ia < 4; ++ia) {
                        double acc = a[ia][c];
#pragma unroll
                        for (int j = 0; j < PW; ++j) acc = fma(-lrow[ia][j], pd[j], acc);
                        a[ia][c] = acc;
                    }
                }
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < T) {
        const double d = dv[threadIdx.x];
        logd[p + threadIdx.x] = log(d);
        rs[threadIdx.x] = 1.0 / sqrt(d);
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < T * T; idx += NTH) {
        const int m = idx >> 7, j = idx & 127;
        if (j <= m) {
            const double v = S[m * LDS + j] * rs[j];
            S[m * LDS + j] = v;
            F[(long long)(p + m) * ldf + p + j] = v;
        }
    }
    __syncthreads();

    // ---------------- phase 2: W L = I, backward over 8-column panels ----------------
#pragma unroll
    for (int ia = 0; ia < 4; ++ia)
#pragma unroll
        for (int c = 0; c < 8; ++c) a[ia][c] = (l + 32 * ia == w + 16 * c) ? 1.0 : 0.0;
#pragma unroll 1
    for (int kp = T / PW - 1; kp >= 0; --kp) {
        const int c0 = kp * PW;
        {   // (a) publish R columns c0..c0+7
            const int j = (w - c0) & 15;
            if (j < PW) {
                const int cj = (c0 + j) >> 4;
#pragma unroll
                for (int ia = 0; ia < 4; ++ia) {
                    double v = a[ia][0];
#pragma unroll
                    for (int c = 1; c < 8; ++c) if (c == cj) v = a[ia][c];
                    PWb[(l + 32 * ia) * LDP + j] = v;
                }
            }
        }
        __syncthreads();
        if (w == 0) {   // (b) row-local back-substitution inside the panel
            double Q[4][PW];
#pragma unroll
            for (int ia = 0; ia < 4; ++ia)
#pragma unroll
                for (int j = 0; j < PW; ++j) Q[ia][j] = PWb[(l + 32 * ia) * LDP + j];
#pragma unroll
            for (int j = PW - 1; j >= 0; --j) {
                const double rsj = rs[c0 + j];
#pragma unroll
                for (int ia = 0; ia < 4; ++ia) Q[ia][j] *= rsj;            // W[m][c0+j]
#pragma unroll
                for (int jj = 0; jj < j; ++jj) {
                    const double ljk = S[(c0 + j) * LDS + c0 + jj];        // L[c0+j][c0+jj], broadcast
#pragma unroll
                    for (int ia = 0; ia < 4; ++ia) Q[ia][jj] = fma(-Q[ia][j], ljk, Q[ia][jj]);
                }
            }
#pragma unroll
            for (int ia = 0; ia < 4; ++ia) {
                const int m = l + 32 * ia;
#pragma unroll
                for (int j = 0; j < PW; ++j) {
                    PWb[m * LDP + j] = Q[ia][j];
                    if (m >= c0 + j) S[(c0 + j) * LDS + m + 1] = Q[ia][j];
                }
            }
        }
        __syncthreads();
        {   // (c) R[m][n] -= sum_j W[m][c0+j] * L[c0+j][n], n < c0
            double wrow[4][PW];
#pragma unroll
            for (int ia = 0; ia < 4; ++ia)
#pragma unroll
                for (int j = 0; j < PW; ++j) wrow[ia][j] = PWb[(l + 32 * ia) * LDP + j];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int n = w + 16 * c;
                if (n < c0) {                                // warp-uniform
                    double lk[PW];
#pragma unroll
                    for (int j = 0; j < PW; ++j) lk[j] = S[(c0 + j) * LDS + n];
#pragma unroll
                    for (int ia = 0; ia < 4; ++ia) {
                        double acc = a[ia][c];
#pragma unroll
                        for (int j = 0; j < PW; ++j) acc = fma(-wrow[ia][j], lk[j], acc);
                        a[ia][c] = acc;
                    }
                }
            }
        }
        __syncthreads();       // PWb is rewritten by the next panel's (a)
    }

    for (int idx = threadIdx.x; idx < T * T; idx += NTH) {
        const int r = idx >> 7, c = idx & 127;
        Dinv[(long long)(p + r) * T + c] = (c <= r) ? S[c * LDS + r + 1] : 0.0;
        DinvT[(long long)(p + r) * T + c] = (c >= r) ? S[r * LDS + c + 1] : 0.0;
    }
}

bool g_attr_set = false;

}  // namespace

cudaError_t potrf128_launch(const double* G, int64_t ldg, double* F, int64_t ldf, double* Dinv, double* DinvT,
                            double* logd, int* info, int p0, int ntiles, int tile_stride, cudaStream_t st, int variant) {
    const size_t sm = (size_t)(T * LDS + T) * sizeof(double);
    const size_t smb = (size_t)(T * LDS + 2 * T + 2 * T * LDP) * sizeof(double);
    if (!g_attr_set) {
        cudaError_t e = cudaFuncSetAttribute(potrf128_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(potrf128_inv_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb);
        if (e != cudaSuccess) return e;
        g_attr_set = true;
    }
    if (variant == 0) potrf128_inv_kernel<<<ntiles, NTH, sm, st>>>(G, ldg, F, ldf, Dinv, DinvT, logd, info, p0, tile_stride);
    else potrf128_inv_blocked_kernel<<<ntiles, NTH, smb, st>>>(G, ldg, F, ldf, Dinv, DinvT, logd, info, p0, tile_stride);
    return cudaGetLastError();
}

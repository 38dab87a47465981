// gram.cu -- Gram-matrix build, cross-Gram and the fused gradient-trace kernel.
//
// Replaces the reference's scalar loops
//   cov!(cK, k, X, data)            src/kernels/kernels.jl:39-50   (symmetric, j>=i + mirror)
//   cov!(cK, k, X1, X2, data)       src/kernels/kernels.jl:56-71   (rectangular, predict)
//   dmll_kern!(dmll, k, X, data, A) src/GPE.jl:219-241             (1/2 sum_ij A_ij dK_ij/dθ)
// and removes the N x N distance cache (IsotropicData.R, src/kernels/stationary.jl:34-40):
// distances are recomputed from the 128 x d input tiles staged in shared memory.
//
// Tile-per-CTA (128 x 128 outputs, 256 threads): the two input tiles are staged in shared memory
// (row stride padded to an odd number of doubles -> conflict-free column reads), each warp owns 16
// rows, each lane 2 x 2 adjacent columns, so a warp stores 512 contiguous bytes per row (16-byte
// vector stores).  Only lower-triangle tiles (bm >= bn) are built: the factorisation reads the
// lower triangle only, halving HBM writes versus the reference's mirrored full matrix.
#include "gram.cuh"
#include <math.h>

namespace {

constexpr int TB = 128;        // tile edge
constexpr int NT = 256;        // threads

__device__ __forceinline__ void load_xtile(double* s, const double* __restrict__ x, long long ldx, int d, int ds,
                                           long long row0, long long nrows_valid) {
    // s[r*ds + k] = x[(row0+r)*ldx + k] for r < 128 (zero beyond the valid rows)
    for (int idx = threadIdx.x; idx < TB * d; idx += NT) {
        const int r = idx / d, k = idx - r * d;
        const long long gr = row0 + r;
        s[r * ds + k] = (gr < nrows_valid) ? x[gr * ldx + k] : 0.0;
    }
}

// ------------------------------------------------------------------------------------------
// symmetric Gram, lower tiles:  G[i,j] = k(x_i,x_j) + [i==j] noise_i ; padding = identity
// ------------------------------------------------------------------------------------------
template <bool FAST>
__global__ void __launch_bounds__(NT) gram_lower_kernel(const __grid_constant__ KProg P, const double* __restrict__ x,
                                                        long long ldx, int d, long long N, long long Npad,
                                                        const double* __restrict__ noise_var, long long n_noise,
                                                        double nugget, double* __restrict__ G, long long ldg,
                                                        int own_tiles, int nranks, int rank, int own_axis, int bm_min) {
    const int bm = blockIdx.y, bn = blockIdx.x;
    if (bn > bm || bm < bm_min) return;
    // multi-GPU ownership: block columns (replicated storage, own_axis 0) or block rows (row-sharded storage, own_axis 1)
    if (own_tiles > 0 && (((own_axis ? bm : bn) / own_tiles) % nranks) != rank) return;
    extern __shared__ double sm[];
    const int ds = d | 1;
    double* sXi = sm;
    double* sXj = sm + TB * ds;
    load_xtile(sXi, x, ldx, d, ds, (long long)bm * TB, N);
    load_xtile(sXj, x, ldx, d, ds, (long long)bn * TB, N);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double mh = 0.0, s2 = 0.0;
    if (FAST) { mh = -0.5 / P.par[0]; s2 = P.par[1]; }
    if (FAST && d <= 8) {
        // register-cached variant: the lane's 4 column points live in registers, a row point is read
        // once per row (8 broadcast LDS) -> 2 LDS per output instead of 16: FP64-ALU bound, not LSU bound
        double xjr[4][8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = (q >> 1) * 64 + lane * 2 + (q & 1);
#pragma unroll
            for (int k = 0; k < 8; ++k) xjr[q][k] = (k < d) ? sXj[c * ds + k] : 0.0;
        }
#pragma unroll 1
        for (int rr = 0; rr < 16; ++rr) {
            const int r = warp * 16 + rr;
            const long long gi = (long long)bm * TB + r;
            double xir[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) xir[k] = (k < d) ? sXi[r * ds + k] : 0.0;
            double v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const long long gj = (long long)bn * TB + (q >> 1) * 64 + lane * 2 + (q & 1);
                double r2 = 0.0;
#pragma unroll
                for (int k = 0; k < 8; ++k) { const double df = xir[k] - xjr[q][k]; r2 += df * df; }
                double kv = s2 * exp(mh * r2);
                if (gi >= N || gj >= N) kv = (gi == gj) ? 1.0 : 0.0;
                else if (gi == gj) kv += ((n_noise == 1) ? noise_var[0] : noise_var[gi]) + nugget;
                v[q] = kv;
            }
            double* row = G + gi * ldg + (long long)bn * TB + lane * 2;
            *reinterpret_cast<double2*>(row) = make_double2(v[0], v[1]);
            *reinterpret_cast<double2*>(row + 64) = make_double2(v[2], v[3]);
        }
        return;
    }
#pragma unroll 1
    for (int rr = 0; rr < 16; ++rr) {
        const int r = warp * 16 + rr;
        const long long gi = (long long)bm * TB + r;
        const double* xi = sXi + r * ds;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int c0 = b * 64 + lane * 2;
            double v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = c0 + e;
                const long long gj = (long long)bn * TB + c;
                const double* xj = sXj + c * ds;
                double kv;
                if (FAST) {
                    double r2 = 0.0;
                    for (int k = 0; k < d; ++k) { const double df = xi[k] - xj[k]; r2 += df * df; }
                    kv = s2 * exp(mh * r2);
                } else {
                    kv = kprog_eval<false>(P, xi, xj, nullptr);
                }
                if (gi >= N || gj >= N) kv = (gi == gj) ? 1.0 : 0.0;
                else if (gi == gj) kv += ((n_noise == 1) ? noise_var[0] : noise_var[gi]) + nugget;
                v[e] = kv;
            }
            *reinterpret_cast<double2*>(G + gi * ldg + (long long)bn * TB + c0) = make_double2(v[0], v[1]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// cross-Gram: Kst[m,n] = k(xs_m, x_n)   (M_pad x N_pad, zero padding)
// ------------------------------------------------------------------------------------------
template <bool FAST>
__global__ void __launch_bounds__(NT) crossgram_kernel(const __grid_constant__ KProg P, const double* __restrict__ xs,
                                                       long long ldxs, long long M, const double* __restrict__ x,
                                                       long long ldx, long long N, int d, double* __restrict__ Kst,
                                                       long long ldk) {
    const int bm = blockIdx.y, bn = blockIdx.x;
    extern __shared__ double sm[];
    const int ds = d | 1;
    double* sXi = sm;
    double* sXj = sm + TB * ds;
    load_xtile(sXi, xs, ldxs, d, ds, (long long)bm * TB, M);
    load_xtile(sXj, x, ldx, d, ds, (long long)bn * TB, N);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double mh = 0.0, s2 = 0.0;
    if (FAST) { mh = -0.5 / P.par[0]; s2 = P.par[1]; }
#pragma unroll 1
    for (int rr = 0; rr < 16; ++rr) {
        const int r = warp * 16 + rr;
        const long long gi = (long long)bm * TB + r;
        const double* xi = sXi + r * ds;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int c0 = b * 64 + lane * 2;
            double v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = c0 + e;
                const long long gj = (long long)bn * TB + c;
                const double* xj = sXj + c * ds;
                double kv;
                if (FAST) {
                    double r2 = 0.0;
                    for (int k = 0; k < d; ++k) { const double df = xi[k] - xj[k]; r2 += df * df; }
                    kv = s2 * exp(mh * r2);
                } else {
                    kv = kprog_eval<false>(P, xi, xj, nullptr);
                }
                if (gi >= M || gj >= N) kv = 0.0;
                v[e] = kv;
            }
            *reinterpret_cast<double2*>(Kst + gi * ldk + (long long)bn * TB + c0) = make_double2(v[0], v[1]);
        }
    }
}

// prior variances k(xs_m, xs_m)
__global__ void kdiag_kernel(const __grid_constant__ KProg P, const double* __restrict__ xs, long long ldxs,
                             long long M, double* __restrict__ out) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const double* xi = xs + m * ldxs;
    out[m] = kprog_eval<false>(P, xi, xi, nullptr);
}

// prior covariance K** (M_pad x M_pad, full, zero padding) -- used by full_cov prediction
template <bool FAST>
__global__ void __launch_bounds__(NT) gram_full_kernel(const __grid_constant__ KProg P, const double* __restrict__ xs,
                                                       long long ldxs, long long M, int d, double* __restrict__ Kss,
                                                       long long ldk) {
    const int bm = blockIdx.y, bn = blockIdx.x;
    extern __shared__ double sm[];
    const int ds = d | 1;
    double* sXi = sm;
    double* sXj = sm + TB * ds;
    load_xtile(sXi, xs, ldxs, d, ds, (long long)bm * TB, M);
    load_xtile(sXj, xs, ldxs, d, ds, (long long)bn * TB, M);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double mh = 0.0, s2 = 0.0;
    if (FAST) { mh = -0.5 / P.par[0]; s2 = P.par[1]; }
#pragma unroll 1
    for (int rr = 0; rr < 16; ++rr) {
        const int r = warp * 16 + rr;
        const long long gi = (long long)bm * TB + r;
        const double* xi = sXi + r * ds;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int c0 = b * 64 + lane * 2;
            double v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = c0 + e;
                const long long gj = (long long)bn * TB + c;
                const double* xj = sXj + c * ds;
                double kv;
                if (FAST) {
                    double r2 = 0.0;
                    for (int k = 0; k < d; ++k) { const double df = xi[k] - xj[k]; r2 += df * df; }
                    kv = s2 * exp(mh * r2);
                } else {
                    kv = kprog_eval<false>(P, xi, xj, nullptr);
                }
                if (gi >= M || gj >= M) kv = 0.0;
                v[e] = kv;
            }
            *reinterpret_cast<double2*>(Kss + gi * ldk + (long long)bn * TB + c0) = make_double2(v[0], v[1]);
        }
    }
}

// dK/dtheta_j as a full symmetric matrix (N_pad x N_pad, zero padding): grad_slice! of the reference
// (/root/reference/src/kernels/kernels.jl:96-131), needed by the cross-validation gradients (src/crossvalidation.jl:67-170, 253-341)
__global__ void __launch_bounds__(NT) gram_grad_full_kernel(const __grid_constant__ KProg P, const double* __restrict__ x, long long ldx,
                                                            long long N, int d, int j, double* __restrict__ D, long long ldd) {
    const int bm = blockIdx.y, bn = blockIdx.x;
    extern __shared__ double sm[];
    const int ds = d | 1;
    double* sXi = sm;
    double* sXj = sm + TB * ds;
    load_xtile(sXi, x, ldx, d, ds, (long long)bm * TB, N);
    load_xtile(sXj, x, ldx, d, ds, (long long)bn * TB, N);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double gbuf[GPB200_MAX_THETA];
#pragma unroll 1
    for (int rr = 0; rr < 16; ++rr) {
        const int r = warp * 16 + rr;
        const long long gi = (long long)bm * TB + r;
        const double* xi = sXi + r * ds;
#pragma unroll 1
        for (int b = 0; b < 4; ++b) {
            const int c = (b >> 1) * 64 + lane * 2 + (b & 1);
            const long long gj = (long long)bn * TB + c;
            double v = 0.0;
            if (gi < N && gj < N) { kprog_eval<true>(P, xi, sXj + c * ds, gbuf); v = gbuf[j]; }
            D[gi * ldd + gj] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------
// fused gradient trace over the lower triangle of Kinv (tile-per-CTA):
//   part[tile][p]        = sum_{i>j in tile} A_ij dK_ij/dθ_p + 1/2 sum_{i==j} A_ii dK_ii/dθ_p
//   part[tile][n_theta]  = sum_{i==j} A_ii                      (-> tr(A), dmll_noise GPE.jl:274)
// A_ij = α_i α_j − Kinv_ij  (GPE.jl:151-164).  Kernel derivatives are recomputed from x.
// Reduction order is fixed (lane tree -> warp order -> tile order): bitwise reproducible.
// ------------------------------------------------------------------------------------------
template <bool FAST>
__global__ void __launch_bounds__(NT) trace_kernel(const __grid_constant__ KProg P, const double* __restrict__ x,
                                                   long long ldx, int d, long long N, const double* __restrict__ alpha,
                                                   const double* __restrict__ Kinv, long long ldg,
                                                   double* __restrict__ part, int tiles, int bm_mod, int bm_rem, int bm_div) {
    // triangular tile index
    const int lin = blockIdx.x;
    int bm = (int)((sqrt(8.0 * (double)lin + 1.0) - 1.0) * 0.5);
    while (bm * (bm + 1) / 2 > lin) --bm;
    while ((bm + 1) * (bm + 2) / 2 <= lin) ++bm;
    const int bn = lin - bm * (bm + 1) / 2;
    const int np = P.n_theta;
    if (bm_mod > 1 && ((bm / bm_div) % bm_mod) != bm_rem) {  // tile row owned by another rank
        const int na = FAST ? 3 : np + 1;
        for (int p = threadIdx.x; p < na; p += NT) part[(long long)lin * na + p] = 0.0;
        return;
    }

    extern __shared__ double sm[];
    const int ds = d | 1;
    double* sXi = sm;
    double* sXj = sXi + TB * ds;
    double* sAi = sXj + TB * ds;
    double* sAj = sAi + TB;
    double* sRed = sAj + TB;             // [8 warps][np+1]
    load_xtile(sXi, x, ldx, d, ds, (long long)bm * TB, N);
    load_xtile(sXj, x, ldx, d, ds, (long long)bn * TB, N);
    for (int i = threadIdx.x; i < TB; i += NT) {
        const long long gi = (long long)bm * TB + i, gj = (long long)bn * TB + i;
        sAi[i] = gi < N ? alpha[gi] : 0.0;
        sAj[i] = gj < N ? alpha[gj] : 0.0;
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double acc[FAST ? 3 : GPB200_MAX_THETA + 1];
    const int nacc = FAST ? 3 : np + 1;
    for (int p = 0; p < nacc; ++p) acc[p] = 0.0;
    double gbuf[FAST ? 1 : GPB200_MAX_THETA];
    double il2 = 0.0, mh = 0.0, s2 = 0.0;
    if (FAST) { il2 = 1.0 / P.par[0]; mh = -0.5 / P.par[0]; s2 = P.par[1]; }

#pragma unroll 1
    for (int rr = 0; rr < 16; ++rr) {
        const int r = warp * 16 + rr;
        const long long gi = (long long)bm * TB + r;
        if (gi >= N) continue;
        const double* xi = sXi + r * ds;
        const double ai = sAi[r];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int c0 = b * 64 + lane * 2;
            const long long gj0 = (long long)bn * TB + c0;
            if (gj0 > gi) continue;                       // strictly upper part of a diagonal tile
            const double2 kin = *reinterpret_cast<const double2*>(Kinv + gi * ldg + gj0);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = c0 + e;
                const long long gj = gj0 + e;
                if (gj > gi || gj >= N) continue;
                const double kinv = e ? kin.y : kin.x;
                const double A = ai * sAj[c] - kinv;
                const double w = (gi == gj) ? 0.5 * A : A;
                const double* xj = sXj + c * ds;
                if (FAST) {
                    double r2 = 0.0;
                    for (int k = 0; k < d; ++k) { const double df = xi[k] - xj[k]; r2 += df * df; }
                    const double kv = s2 * exp(mh * r2);
                    acc[0] += w * (r2 * il2 * kv);        // se_iso.jl:41  dk/dll = r/ℓ² k
                    acc[1] += w * (2.0 * kv);             // stationary.jl:28 dk/dlσ = 2k
                    if (gi == gj) acc[2] += A;
                } else {
                    kprog_eval<true>(P, xi, xj, gbuf);
                    for (int p = 0; p < np; ++p) acc[p] += w * gbuf[p];
                    if (gi == gj) acc[np] += A;
                }
            }
        }
    }
    // lane tree, then fixed warp order
    for (int p = 0; p < nacc; ++p) {
        double v = acc[p];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) sRed[warp * nacc + p] = v;
    }
    __syncthreads();
    for (int p = threadIdx.x; p < nacc; p += NT) {
        double v = 0.0;
        for (int w8 = 0; w8 < NT / 32; ++w8) v += sRed[w8 * nacc + p];
        part[(long long)lin * nacc + p] = v;
    }
}

// out[p] = sum over tiles (fixed-order pairwise tree inside one CTA per p)
__global__ void reduce_partials_kernel(const double* __restrict__ part, int tiles, int nacc, double* __restrict__ out) {
    const int p = blockIdx.x;
    __shared__ double s[256];
    double v = 0.0;
    for (int t = threadIdx.x; t < tiles; t += 256) v += part[(long long)t * nacc + p];
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[p] = s[0];
}

// ------------------------------------------------------------------------------------------
// rectangular weighted trace (FITC gradient): part[tile][p] = sum_{i,m in tile} W[i,m] dk(x1_i, x2_m)/dθ_p
// ------------------------------------------------------------------------------------------
template <bool FAST>
__global__ void __launch_bounds__(NT) trace_rect_kernel(const __grid_constant__ KProg P, const double* __restrict__ x1,
                                                        long long ldx1, long long N1, const double* __restrict__ x2,
                                                        long long ldx2, long long N2, int d,
                                                        const double* __restrict__ W, long long ldw,
                                                        double* __restrict__ part) {
    const int bm = blockIdx.y, bn = blockIdx.x;
    const int np = P.n_theta;
    extern __shared__ double sm[];
    const int ds = d | 1;
    double* sXi = sm;
    double* sXj = sXi + TB * ds;
    double* sRed = sXj + TB * ds;
    load_xtile(sXi, x1, ldx1, d, ds, (long long)bm * TB, N1);
    load_xtile(sXj, x2, ldx2, d, ds, (long long)bn * TB, N2);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double acc[FAST ? 2 : GPB200_MAX_THETA];
    const int nacc = FAST ? 2 : np;
    for (int p = 0; p < nacc; ++p) acc[p] = 0.0;
    double gbuf[FAST ? 1 : GPB200_MAX_THETA];
    double il2 = 0.0, mh = 0.0, s2 = 0.0;
    if (FAST) { il2 = 1.0 / P.par[0]; mh = -0.5 / P.par[0]; s2 = P.par[1]; }
#pragma unroll 1
    for (int rr = 0; rr < 16; ++rr) {
        const int r = warp * 16 + rr;
        const long long gi = (long long)bm * TB + r;
        if (gi >= N1) continue;
        const double* xi = sXi + r * ds;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int c0 = b * 64 + lane * 2;
            const long long gj0 = (long long)bn * TB + c0;
            if (gj0 >= N2) continue;
            const double2 wv = *reinterpret_cast<const double2*>(W + gi * ldw + gj0);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const long long gj = gj0 + e;
                if (gj >= N2) continue;
                const double w = e ? wv.y : wv.x;
                const double* xj = sXj + (c0 + e) * ds;
                if (FAST) {
                    double r2 = 0.0;
                    for (int k = 0; k < d; ++k) { const double df = xi[k] - xj[k]; r2 += df * df; }
                    const double kv = s2 * exp(mh * r2);
                    acc[0] += w * (r2 * il2 * kv);
                    acc[1] += w * (2.0 * kv);
                } else {
                    kprog_eval<true>(P, xi, xj, gbuf);
                    for (int p = 0; p < np; ++p) acc[p] += w * gbuf[p];
                }
            }
        }
    }
    for (int p = 0; p < nacc; ++p) {
        double v = acc[p];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) sRed[warp * nacc + p] = v;
    }
    __syncthreads();
    const long long lin = (long long)blockIdx.y * gridDim.x + blockIdx.x;
    for (int p = threadIdx.x; p < nacc; p += NT) {
        double v = 0.0;
        for (int w8 = 0; w8 < NT / 32; ++w8) v += sRed[w8 * nacc + p];
        part[lin * nacc + p] = v;
    }
}

// out[i*np + p] = g[i] * dk(x_i, x_i)/dθ_p     (diagonal term of the FITC gradient)
__global__ void kdiag_grad_kernel(const __grid_constant__ KProg P, const double* __restrict__ x, long long ldx,
                                  long long N, const double* __restrict__ gvec, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double gbuf[GPB200_MAX_THETA];
    const double* xi = x + i * ldx;
    kprog_eval<true>(P, xi, xi, gbuf);
    const double gi = gvec[i];
    for (int p = 0; p < P.n_theta; ++p) out[(long long)p * N + i] = gi * gbuf[p];
}

size_t xtile_smem(int d) { return (size_t)2 * TB * (d | 1) * sizeof(double); }

template <typename K>
cudaError_t ensure_smem(K kern, size_t bytes) {
    if (bytes > 48 * 1024) return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    return cudaSuccess;
}

}  // namespace

cudaError_t gram_lower_launch(const KProg& P, const double* x, int64_t ldx, int d, int64_t N, int64_t Npad,
                              const double* noise_var, int64_t n_noise, double nugget, double* G, int64_t ldg,
                              cudaStream_t st, int own_tiles, int nranks, int rank, int own_axis, int bm_min) {
    const int T = (int)(Npad / TB);
    dim3 grid(T, T);
    const size_t sm = xtile_smem(d);
    cudaError_t e;
    if (P.fast) {
        if ((e = ensure_smem(gram_lower_kernel<true>, sm)) != cudaSuccess) return e;
        gram_lower_kernel<true><<<grid, NT, sm, st>>>(P, x, ldx, d, N, Npad, noise_var, n_noise, nugget, G, ldg, own_tiles, nranks, rank, own_axis, bm_min);
    } else {
        if ((e = ensure_smem(gram_lower_kernel<false>, sm)) != cudaSuccess) return e;
        gram_lower_kernel<false><<<grid, NT, sm, st>>>(P, x, ldx, d, N, Npad, noise_var, n_noise, nugget, G, ldg, own_tiles, nranks, rank, own_axis, bm_min);
    }
    return cudaGetLastError();
}

cudaError_t crossgram_launch(const KProg& P, const double* xs, int64_t ldxs, int64_t M, int64_t Mpad, const double* x,
                             int64_t ldx, int64_t N, int64_t Npad, int d, double* Kst, int64_t ldk, cudaStream_t st) {
    dim3 grid((unsigned)(Npad / TB), (unsigned)(Mpad / TB));
    const size_t sm = xtile_smem(d);
    cudaError_t e;
    if (P.fast) {
        if ((e = ensure_smem(crossgram_kernel<true>, sm)) != cudaSuccess) return e;
        crossgram_kernel<true><<<grid, NT, sm, st>>>(P, xs, ldxs, M, x, ldx, N, d, Kst, ldk);
    } else {
        if ((e = ensure_smem(crossgram_kernel<false>, sm)) != cudaSuccess) return e;
        crossgram_kernel<false><<<grid, NT, sm, st>>>(P, xs, ldxs, M, x, ldx, N, d, Kst, ldk);
    }
    return cudaGetLastError();
}

cudaError_t gram_full_launch(const KProg& P, const double* xs, int64_t ldxs, int64_t M, int64_t Mpad, int d,
                             double* Kss, int64_t ldk, cudaStream_t st) {
    dim3 grid((unsigned)(Mpad / TB), (unsigned)(Mpad / TB));
    const size_t sm = xtile_smem(d);
    cudaError_t e;
    if (P.fast) {
        if ((e = ensure_smem(gram_full_kernel<true>, sm)) != cudaSuccess) return e;
        gram_full_kernel<true><<<grid, NT, sm, st>>>(P, xs, ldxs, M, d, Kss, ldk);
    } else {
        if ((e = ensure_smem(gram_full_kernel<false>, sm)) != cudaSuccess) return e;
        gram_full_kernel<false><<<grid, NT, sm, st>>>(P, xs, ldxs, M, d, Kss, ldk);
    }
    return cudaGetLastError();
}

cudaError_t kdiag_launch(const KProg& P, const double* xs, int64_t ldxs, int64_t M, double* out, cudaStream_t st) {
    kdiag_kernel<<<(unsigned)((M + 127) / 128), 128, 0, st>>>(P, xs, ldxs, M, out);
    return cudaGetLastError();
}

int trace_num_acc(const KProg& P) { return P.fast ? 3 : P.n_theta + 1; }

cudaError_t trace_launch(const KProg& P, const double* x, int64_t ldx, int d, int64_t N, int64_t Npad,
                         const double* alpha, const double* Kinv, int64_t ldg, double* part, double* out,
                         cudaStream_t st, int bm_mod, int bm_rem, int bm_div) {
    if (bm_div < 1) bm_div = 1;
    const int T = (int)(Npad / TB);
    const int tiles = T * (T + 1) / 2;
    const int nacc = trace_num_acc(P);
    const size_t sm = xtile_smem(d) + (size_t)(2 * TB + 8 * nacc) * sizeof(double);
    cudaError_t e;
    if (P.fast) {
        if ((e = ensure_smem(trace_kernel<true>, sm)) != cudaSuccess) return e;
        trace_kernel<true><<<tiles, NT, sm, st>>>(P, x, ldx, d, N, alpha, Kinv, ldg, part, tiles, bm_mod, bm_rem, bm_div);
    } else {
        if ((e = ensure_smem(trace_kernel<false>, sm)) != cudaSuccess) return e;
        trace_kernel<false><<<tiles, NT, sm, st>>>(P, x, ldx, d, N, alpha, Kinv, ldg, part, tiles, bm_mod, bm_rem, bm_div);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    reduce_partials_kernel<<<nacc, 256, 0, st>>>(part, tiles, nacc, out);
    return cudaGetLastError();
}

// acc_out[p] += sum over the N1 x N2 rectangle of W[i,m] dk(x1_i,x2_m)/dθ_p ; part: [tiles][n_theta] scratch
cudaError_t trace_rect_launch(const KProg& P, const double* x1, int64_t ldx1, int64_t N1, const double* x2, int64_t ldx2,
                              int64_t N2, int d, const double* W, int64_t ldw, double* part, double* tmp_out,
                              cudaStream_t st) {
    const int tm = (int)((N1 + TB - 1) / TB), tn = (int)((N2 + TB - 1) / TB);
    const int nacc = P.fast ? 2 : P.n_theta;
    const size_t sm = xtile_smem(d) + (size_t)(8 * nacc) * sizeof(double);
    dim3 grid(tn, tm);
    cudaError_t e;
    if (P.fast) {
        if ((e = ensure_smem(trace_rect_kernel<true>, sm)) != cudaSuccess) return e;
        trace_rect_kernel<true><<<grid, NT, sm, st>>>(P, x1, ldx1, N1, x2, ldx2, N2, d, W, ldw, part);
    } else {
        if ((e = ensure_smem(trace_rect_kernel<false>, sm)) != cudaSuccess) return e;
        trace_rect_kernel<false><<<grid, NT, sm, st>>>(P, x1, ldx1, N1, x2, ldx2, N2, d, W, ldw, part);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    reduce_partials_kernel<<<nacc, 256, 0, st>>>(part, tm * tn, nacc, tmp_out);
    return cudaGetLastError();
}
cudaError_t kdiag_grad_launch(const KProg& P, const double* x, int64_t ldx, int64_t N, const double* gvec, double* out,
                              cudaStream_t st) {
    kdiag_grad_kernel<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(P, x, ldx, N, gvec, out);
    return cudaGetLastError();
}

cudaError_t gram_grad_full_launch(const KProg& P, const double* x, int64_t ldx, int d, int64_t N, int64_t Npad, int j,
                                  double* D, int64_t ldd, cudaStream_t st) {
    if (j < 0 || j >= P.n_theta) return cudaErrorInvalidValue;
    dim3 grid((unsigned)(Npad / TB), (unsigned)(Npad / TB));
    const size_t sm = xtile_smem(d);
    cudaError_t e;
    if ((e = ensure_smem(gram_grad_full_kernel, sm)) != cudaSuccess) return e;
    gram_grad_full_kernel<<<grid, NT, sm, st>>>(P, x, ldx, N, d, j, D, ldd);
    return cudaGetLastError();
}

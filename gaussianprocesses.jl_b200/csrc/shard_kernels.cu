// shard_kernels.cu -- see shard_kernels.cuh
#include "shard_kernels.cuh"

namespace {
constexpr int T = 128;

__global__ void copy_lower_kernel(double* __restrict__ dst, long long ldd, const double* __restrict__ src, long long lds, int n) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= n) return;
    dst[(long long)i * ldd + j] = (j <= i) ? src[(long long)i * lds + j] : 0.0;
}

// grid: (ceil(nb/128), own tile ordinal); block 128 threads x 1; each block moves one 128-row tile x 128 columns
__global__ void __launch_bounds__(256) scatter_rows_kernel(const double* __restrict__ G, double* __restrict__ F, long long ld,
                                                           int c0, int nb, long long t1, double* __restrict__ S_me, int nbp,
                                                           ShardOwn own, long long own_before_t1, long long n_tiles) {
    // blockIdx.y enumerates ALL tiles t >= t1; non-owned ones exit (cheap: a few thousand CTAs at most)
    const long long t = t1 + blockIdx.y;
    if (t >= n_tiles || (int)((t / own.rb) % own.R) != own.me) return;
    const long long li = shard_own_before(t, own.me, own.rb, own.R) - own_before_t1;     // local tile index
    const int cb = blockIdx.x * 128;
    const int c = cb + (threadIdx.x & 127);
    if (c >= nb) return;
    for (int rr = threadIdx.x >> 7; rr < T; rr += 2) {
        const long long r = t * T + rr;
        const double v = G[r * ld + c0 + c];
        F[r * ld + c0 + c] = v;
        S_me[(li * T + rr) * nbp + c] = v;
    }
}

__global__ void __launch_bounds__(256) unpack_panel_kernel(double* __restrict__ P, int nbp, int nb, long long t1,
                                                           const double* __restrict__ S, long long per_rank, ShardOwn own,
                                                           long long n_tiles) {
    const long long t = t1 + blockIdx.y;
    if (t >= n_tiles) return;
    const int q = (int)((t / own.rb) % own.R);
    const long long li = shard_own_before(t, q, own.rb, own.R) - shard_own_before(t1, q, own.rb, own.R);
    const int c = blockIdx.x * 128 + (threadIdx.x & 127);
    if (c >= nb) return;
    const double* src = S + (long long)q * per_rank;
    for (int rr = threadIdx.x >> 7; rr < T; rr += 2)
        P[(t * T + rr) * nbp + c] = src[(li * T + rr) * nbp + c];
}

__global__ void transpose_kernel(double* __restrict__ dst, long long ldd, const double* __restrict__ src, long long lds,
                                 long long rows, long long cols) {
    __shared__ double tile[32][33];
    const long long i0 = (long long)blockIdx.y * 32, j0 = (long long)blockIdx.x * 32;
    for (int y = threadIdx.y; y < 32; y += 8) {
        const long long i = i0 + y, j = j0 + threadIdx.x;
        tile[y][threadIdx.x] = (i < rows && j < cols) ? src[i * lds + j] : 0.0;
    }
    __syncthreads();
    for (int y = threadIdx.y; y < 32; y += 8) {
        const long long j = j0 + y, i = i0 + threadIdx.x;
        if (i < rows && j < cols) dst[j * ldd + i] = tile[threadIdx.x][y];
    }
}

__global__ void pack_wblock_kernel(double* __restrict__ P, int nbp, const double* __restrict__ G, long long ld,
                                   const double* __restrict__ Dinv, int j0, int nb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;       // P[(j0+c), i]
    if (i >= nb) return;
    double v = 0.0;
    if ((c >> 7) > (i >> 7)) v = G[(long long)(j0 + c) * ld + j0 + i];        // strictly-lower tile of W_JJ
    else if ((c >> 7) == (i >> 7)) v = Dinv[(long long)(j0 + c) * T + (i & 127)];   // clean lower diagonal tile (zeros above)
    P[(long long)(j0 + c) * nbp + i] = v;
}

__global__ void sum_ranks_kernel(double* __restrict__ out, const double* __restrict__ in, long long stride, int nranks, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int q = 0; q < nranks; ++q) s += in[(long long)q * stride + i];
    out[i] = s;
}
__global__ void min_ranks_kernel(int* __restrict__ out, const int* __restrict__ in, int nranks) {
    int m = in[0];
    for (int q = 1; q < nranks; ++q) m = min(m, in[q]);
    out[0] = m;
}
}  // namespace

cudaError_t shard_copy_lower(double* dst, int64_t ldd, const double* src, int64_t lds, int n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    dim3 grid((n + 127) / 128, n);
    copy_lower_kernel<<<grid, 128, 0, st>>>(dst, ldd, src, lds, n);
    return cudaGetLastError();
}

cudaError_t shard_scatter_rows(const double* G, double* F, int64_t ld, int64_t Npad, int c0, int nb, int t1, double* S_me,
                               int nbp, ShardOwn own, cudaStream_t st) {
    const long long n_tiles = Npad / T;
    if (t1 >= n_tiles || nb <= 0) return cudaSuccess;
    dim3 grid((nb + 127) / 128, (unsigned)(n_tiles - t1));
    scatter_rows_kernel<<<grid, 256, 0, st>>>(G, F, ld, c0, nb, t1, S_me, nbp, own,
                                              shard_own_before(t1, own.me, own.rb, own.R), n_tiles);
    return cudaGetLastError();
}

cudaError_t shard_unpack_panel(double* P, int nbp, int64_t Npad, int nb, int t1, const double* S, int64_t per_rank,
                               ShardOwn own, cudaStream_t st) {
    const long long n_tiles = Npad / T;
    if (t1 >= n_tiles || nb <= 0) return cudaSuccess;
    dim3 grid((nb + 127) / 128, (unsigned)(n_tiles - t1));
    unpack_panel_kernel<<<grid, 256, 0, st>>>(P, nbp, nb, t1, S, per_rank, own, n_tiles);
    return cudaGetLastError();
}

cudaError_t shard_transpose(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return cudaSuccess;
    // grid.y is limited to 65535 blocks of 32 rows = 2M rows: enough for N <= 2^21
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32)), block(32, 8);
    transpose_kernel<<<grid, block, 0, st>>>(dst, ldd, src, lds, rows, cols);
    return cudaGetLastError();
}

cudaError_t shard_pack_wblock(double* P, int nbp, const double* G, int64_t ld, const double* Dinv, int j0, int nb, cudaStream_t st) {
    if (nb <= 0) return cudaSuccess;
    dim3 grid((nb + 127) / 128, nb);
    pack_wblock_kernel<<<grid, 128, 0, st>>>(P, nbp, G, ld, Dinv, j0, nb);
    return cudaGetLastError();
}

cudaError_t shard_sum_ranks(double* out, const double* in, int64_t stride, int nranks, int64_t n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    sum_ranks_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(out, in, stride, nranks, n);
    return cudaGetLastError();
}
cudaError_t shard_min_ranks(int* out, const int* in, int nranks, cudaStream_t st) {
    min_ranks_kernel<<<1, 1, 0, st>>>(out, in, nranks);
    return cudaGetLastError();
}

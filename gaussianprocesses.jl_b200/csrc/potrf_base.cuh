// potrf_base.cuh -- 128 x 128 diagonal-tile Cholesky + triangular inverse (see potrf_base.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
// fused panel broadcast (multi-GPU): every output of the leaf is also stored into the peers' buffers (NVLink P2P)
struct PotrfPeers {
    int n;
    double* F[7];
    double* Dinv[7];
    double* DinvT[7];
    double* logd[7];
};
// tiles at diagonal offsets p0 + b*tile_stride, b < ntiles.  *info must be initialised to INT_MAX.
cudaError_t potrf128_launch(const double* G, int64_t ldg, double* F, int64_t ldf, double* Dinv, double* DinvT,
                            double* logd, int* info, int p0, int ntiles, int tile_stride, cudaStream_t st,
                            const PotrfPeers* peers = nullptr);
// 1 (default) = blocked 16-column-panel kernel (74.5 us per tile), 0 = one-barrier-per-column kernel (123 us; cross-check); process-wide
void potrf128_set_variant(int blocked);

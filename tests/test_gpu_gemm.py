"""GPU: the FP64 DMMA NT GEMM (both loaders) against cuBLAS (torch.matmul, float64) -- a torch fp64
reference is kept for this floating-point kernel; tolerance 1e-13 relative to |A||B| (FP64
accumulation in a different order)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(engine, impl, M, N, K, alpha, beta, lower=False, pad=0):
    import torch
    torch.manual_seed(M + 3 * N + 7 * K + impl)
    dev = "cuda:0"
    Abig = torch.randn(M, K + pad, dtype=torch.float64, device=dev)
    Bbig = torch.randn(N, K + pad, dtype=torch.float64, device=dev)
    A, B = Abig[:, :K], Bbig[:, :K]
    C = torch.randn(M, N, dtype=torch.float64, device=dev)
    ref = alpha * (A @ B.T) + beta * C
    C0 = C.clone()
    torch.cuda.synchronize()
    engine.dgemm_nt_device(impl, M, N, K, alpha, Abig.data_ptr(), K + pad, Bbig.data_ptr(), K + pad, beta,
                           C.data_ptr(), N, lower_only=lower)
    torch.cuda.synchronize()
    scale = float((A.abs() @ B.abs().T).max()) * abs(alpha) + abs(beta) * float(C0.abs().max())
    if lower:
        mask = torch.tril(torch.ones(M, N, dtype=torch.bool, device=dev))
        # tiles strictly above the diagonal are untouched, diagonal tiles only written for i >= j
        assert torch.equal(C[~mask], C0[~mask])
        err = float((C - ref)[mask].abs().max())
    else:
        err = float((C - ref).abs().max())
    assert err <= 2e-14 * scale + 1e-300, (err, scale)


@pytest.mark.parametrize("impl", [1, 0])
@pytest.mark.parametrize("shape", [(128, 128, 16), (128, 128, 128), (256, 384, 512), (384, 128, 1040), (1024, 1024, 1024)])
def test_dgemm_nt_matches_cublas(engine, impl, shape):
    M, N, K = shape
    _run(engine, impl, M, N, K, 1.0, 0.0)
    _run(engine, impl, M, N, K, -1.0, 1.0)
    _run(engine, impl, M, N, K, 0.75, -0.5, pad=6)


@pytest.mark.parametrize("impl", [1, 0])
def test_dgemm_nt_lower_only(engine, impl):
    _run(engine, impl, 512, 512, 256, -1.0, 1.0, lower=True)
    _run(engine, impl, 128, 128, 64, 1.0, 0.0, lower=True)


def test_dgemm_nt_tma_equals_simple_bitwise(engine):
    """Same tile math, two loaders: results must be bit-identical."""
    import torch
    M, N, K = 640, 512, 2048
    A = torch.randn(M, K, dtype=torch.float64, device="cuda:0")
    B = torch.randn(N, K, dtype=torch.float64, device="cuda:0")
    C0 = torch.zeros(M, N, dtype=torch.float64, device="cuda:0")
    C1 = torch.zeros_like(C0)
    engine.dgemm_nt_device(0, M, N, K, 1.0, A.data_ptr(), K, B.data_ptr(), K, 0.0, C0.data_ptr(), N)
    engine.dgemm_nt_device(1, M, N, K, 1.0, A.data_ptr(), K, B.data_ptr(), K, 0.0, C1.data_ptr(), N)
    torch.cuda.synchronize()
    assert torch.equal(C0, C1)

"""tcgen05 3xTF32 GEMM against an fp64 product: all four transpose variants, ragged sizes, split-K shapes.
Tolerance: fp32-class (max abs error <= 2e-6 * sum_k |a||b|), plus the exact-fp32 FFMA path."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHAPES = [(128, 128, 32), (128, 64, 64), (256, 768, 64), (200, 70, 100), (51200, 64, 128), (64, 192, 5000),
          (128, 128, 768), (1, 8, 33), (300, 129, 40),
          (64, 768, 6000), (128, 64, 20000), (1000, 192, 70), (700, 333, 45)]


@pytest.mark.parametrize("mode", [1, 0])
@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_matches_fp64(mode, ta, tb, M, N, K):
    import torch
    from mtamrecommender_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K + ta * 2 + tb)
    A = torch.randn((K, M) if ta else (M, K), generator=g)
    B = torch.randn((N, K) if tb else (K, N), generator=g)
    bias = torch.randn(N, generator=g)
    ref = (A.double().t() if ta else A.double()) @ (B.double().t() if tb else B.double())
    bound = ((A.double().abs().t() if ta else A.double().abs()) @ (B.double().abs().t() if tb else B.double().abs()))
    ref = torch.relu(ref + bias.double())
    Ad, Bd, bd = A.cuda(), B.cuda(), bias.cuda()
    Cd = torch.full((M, N), 7.0, device="cuda")
    ws = torch.empty(max(int(lib.mtam_gemm_workspace(M, N, K)), 16), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.mtam_gemm(mode, ta, tb, M, N, K, Ad.data_ptr(), Ad.stride(0), Bd.data_ptr(), Bd.stride(0),
                             Cd.data_ptr(), N, bd.data_ptr(), 1, 0, ws.data_ptr(), ws.numel(), st), "mtam_gemm")
    err = (Cd.double().cpu() - ref).abs()
    assert bool((err <= 2e-6 * bound + 1e-6).all()), float((err / (bound + 1e-9)).max())
    # accumulate epilogue: C += A B
    _lib.check(lib.mtam_gemm(mode, ta, tb, M, N, K, Ad.data_ptr(), Ad.stride(0), Bd.data_ptr(), Bd.stride(0),
                             Cd.data_ptr(), N, None, 0, 1, ws.data_ptr(), ws.numel(), st), "mtam_gemm")
    ref2 = ref + (A.double().t() if ta else A.double()) @ (B.double().t() if tb else B.double())
    err = (Cd.double().cpu() - ref2).abs()
    assert bool((err <= 4e-6 * bound + 2e-6).all())


@pytest.mark.parametrize("ta,tb,M,N,K", [(1, 0, 64, 768, 51200), (1, 0, 128, 64, 51200), (0, 1, 256, 64, 4096),
                                         (1, 0, 64, 192, 20000)])
def test_long_accumulations_do_not_drift(ta, tb, M, N, K):
    """One-signed operands make the accumulator grow with every MMA, which exposes the tensor core's truncating fp32
    accumulation (a chain of n MMAs loses about n * 2^-24 of the sum: 1e-3 at K = 51 200, the weight-gradient shape of
    the cfg3 step).  The kernel cuts K into sub-units of 192 MMAs and sums them with ordinary fp32 adds: the result must
    be fp32-class relative to the RESULT, not merely to sum |a||b|."""
    import torch
    from mtamrecommender_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(K + M)
    A = torch.rand((K, M) if ta else (M, K), generator=g) + 0.5
    B = torch.rand((N, K) if tb else (K, N), generator=g) + 0.5
    ref = (A.double().t() if ta else A.double()) @ (B.double().t() if tb else B.double())
    Ad, Bd = A.cuda(), B.cuda()
    Cd = torch.empty((M, N), device="cuda")
    ws = torch.empty(max(int(lib.mtam_gemm_workspace(M, N, K)), 16), dtype=torch.uint8, device="cuda")
    _lib.check(lib.mtam_gemm(1, ta, tb, M, N, K, Ad.data_ptr(), Ad.stride(0), Bd.data_ptr(), Bd.stride(0), Cd.data_ptr(), N,
                             None, 0, 0, ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream), "mtam_gemm")
    relerr = ((Cd.double().cpu() - ref).abs() / ref).max().item()
    assert relerr < 1.5e-5, relerr            # 192 truncating adds per sub-unit: <= 192 * 2^-24 = 1.1e-5

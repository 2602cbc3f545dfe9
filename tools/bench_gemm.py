"""Developer microbenchmark: the model's GEMM shapes at cfg3 (T = 51200, D = 64, N = 6) through mtam_gemm, warm."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mtamrecommender_b200 import _lib

T, D, N = 51200, 64, 6
SHAPES = [  # name, ta, tb, M, N, K, accumulate, relu+bias
    ("embed  [T,2D]x[2D,D]", 0, 0, T, D, 2 * D, 0, 1),
    ("gru_x  [T,D]x[D,3D]", 0, 0, T, 3 * D, D, 0, 1),
    ("kv     [T,D]x[D,2ND]", 0, 0, T, 2 * N * D, D, 0, 1),
    ("dWkv   X^T dKV", 1, 0, D, 2 * N * D, T, 0, 0),
    ("dX+=   dKV Wkv^T", 0, 1, T, D, 2 * N * D, 1, 0),
    ("dWgru  X^T dGX", 1, 0, D, 3 * D, T, 0, 0),
    ("dX+=   dGX Wgru^T", 0, 1, T, D, 3 * D, 1, 0),
    ("dWemb  E2^T dR", 1, 0, 2 * D, D, T, 0, 0),
    ("dE2    dR W^T", 0, 1, T, 2 * D, D, 0, 0),
]
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
tot = 0.0
for name, ta, tb, M, Nn, K, acc, rb in SHAPES:
    A = torch.randn((K, M) if ta else (M, K), device="cuda")
    B = torch.randn((Nn, K) if tb else (K, Nn), device="cuda")
    bias = torch.randn(Nn, device="cuda")
    C = torch.zeros((M, Nn), device="cuda")
    ws = torch.empty(max(int(lib.mtam_gemm_workspace(M, Nn, K)), 16), dtype=torch.uint8, device="cuda")
    def run():
        _lib.check(lib.mtam_gemm(1, ta, tb, M, Nn, K, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), C.data_ptr(), Nn,
                                 bias.data_ptr() if rb else None, rb, acc, ws.data_ptr(), ws.numel(), st), "gemm")
    for _ in range(3): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    byts = 4.0 * (M * K + K * Nn + M * Nn * (2 if acc else 1))
    tot += us
    print(f"{name:24s} M={M:6d} N={Nn:4d} K={K:6d}  {us:7.1f} us   {byts/us/1e3:7.1f} GB/s (min traffic)")
print(f"total {tot:.1f} us")

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mtamrecommender_b200 import _lib
lib = _lib.load()
def run(mode, ta, tb, M, N, K, A, B):
    Ad, Bd = A.cuda(), B.cuda()
    Cd = torch.zeros((M, N), device="cuda")
    ws = torch.empty(max(int(lib.mtam_gemm_workspace(M, N, K)), 16), dtype=torch.uint8, device="cuda")
    _lib.check(lib.mtam_gemm(mode, ta, tb, M, N, K, Ad.data_ptr(), Ad.stride(0), Bd.data_ptr(), Bd.stride(0), Cd.data_ptr(), N, None, 0, 0, ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream), "gemm")
    torch.cuda.synchronize()
    return Cd.cpu().double()
g = torch.Generator().manual_seed(0)
for (ta, tb) in [(0, 1), (0, 0), (1, 1), (1, 0)]:
    for (M, N, K) in [(128, 128, 32), (128, 128, 8), (128, 64, 64)]:
        A = torch.randn((K, M) if ta else (M, K), generator=g); B = torch.randn((N, K) if tb else (K, N), generator=g)
        ref = (A.double().t() if ta else A.double()) @ (B.double().t() if tb else B.double())
        C = run(1, ta, tb, M, N, K, A, B)
        err = (C - ref).abs()
        print(f"ta={ta} tb={tb} M={M} N={N} K={K}: max err {float(err.max()):.3e}  frac bad {(err > 1e-4).double().mean():.3f}", end="")
        if err.max() > 1e-4:
            bad_rows = (err > 1e-4).any(1).nonzero().flatten().tolist(); bad_cols = (err > 1e-4).any(0).nonzero().flatten().tolist()
            print("  bad rows", bad_rows[:10], len(bad_rows), " bad cols", bad_cols[:16], len(bad_cols))
            # try to explain: is C == ref with permuted columns / K subset?
            for kk in range(0, K, 8):
                sub = (A.double().t() if ta else A.double())[:, kk:kk+8] @ (B.double().t() if tb else B.double())[kk:kk+8, :]
                print("    corr with k-block", kk, float((C * sub).sum() / (sub * sub).sum()))
        else:
            print()

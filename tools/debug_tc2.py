import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mtamrecommender_b200 import _lib
lib = _lib.load()
def run(mode, ta, tb, M, N, K, A, B, fill=0.0):
    Ad, Bd = A.cuda(), B.cuda()
    Cd = torch.full((M, N), fill, device="cuda")
    ws = torch.empty(max(int(lib.mtam_gemm_workspace(M, N, K)), 16), dtype=torch.uint8, device="cuda")
    _lib.check(lib.mtam_gemm(mode, ta, tb, M, N, K, Ad.data_ptr(), Ad.stride(0), Bd.data_ptr(), Bd.stride(0), Cd.data_ptr(), N, None, 0, 0, ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream), "gemm")
    torch.cuda.synchronize()
    return Cd.cpu()
for K in (8, 32, 64):
    C = run(1, 0, 1, 128, 128, K, torch.ones(128, K), torch.ones(128, K), fill=-5.0)
    print("ones K", K, C[0, :4].tolist(), C[127, 124:].tolist(), float(C.min()), float(C.max()))
A = torch.zeros(128, 32); A[:, 0] = torch.arange(1, 129).float()
B = torch.zeros(128, 32); B[:, 0] = torch.arange(1, 129).float()
C = run(1, 0, 1, 128, 128, 32, A, B, fill=-5.0)
print("rank1", C[0, :4].tolist(), C[1, :2].tolist(), C[127, 127].item())
C = run(0, 0, 1, 128, 128, 32, A, B, fill=-5.0)
print("rank1 fp32 path", C[0, :4].tolist(), C[1, :2].tolist(), C[127, 127].item())

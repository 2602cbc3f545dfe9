"""One scatter-add (sort + segmented reduce) and one gather at cfg-4 shapes, for ncu."""
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
from mtamrecommender_b200 import engine as E

n, D, rows = 8192 * 200, 64, 10_000_003
dev = "cuda:0"
table = torch.empty((rows, D), device=dev).uniform_(-0.3, 0.3)
out = torch.empty((n, D), device=dev)
dst = torch.zeros((rows, D), device=dev)
ws = torch.empty(E.scatter_add_workspace(n, rows, D), dtype=torch.uint8, device=dev)
idx = torch.from_numpy(np.random.default_rng(5).integers(0, rows, n).astype(np.int32)).to(dev)
acc = len(sys.argv) > 1 and sys.argv[1] == "acc"
for _ in range(2):
    E.gather(table, idx, out)
    E.scatter_add(dst, idx, out, ws, accumulate=acc)
torch.cuda.synchronize()

"""Index sort and whole scatter-add at cfg-4 shapes (uniform / Zipf+pad ids): ms per call."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mtamrecommender_b200 import engine as E
from mtamrecommender_b200.synth import ZipfSampler
n, D, rows, dev = 8192 * 200, 64, 10_000_003, "cuda:0"
rng = np.random.default_rng(5)
out = torch.randn((n, D), device=dev)
ws = torch.empty(E.scatter_add_workspace(n, rows, D), dtype=torch.uint8, device=dev)
dst = torch.zeros((rows, D), device=dev)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ev0.record()
    for _ in range(reps): fn()
    ev1.record(); torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / reps
for dist in ("uniform", "zipf_pad"):
    if dist == "uniform":
        idx_np = rng.integers(0, rows, n).astype(np.int32)
    else:
        idx_np = ZipfSampler(rows - 3, 1.05).sample(rng, n)
        idx_np[rng.random(n) < 0.45] = 0
    idx = torch.from_numpy(idx_np).to(dev)
    nu = int(np.unique(idx_np).size)
    b = n * (4 + 4 * D) + nu * 4 * D
    s = t(lambda: E.sort_indices(idx, rows, ws))
    w = t(lambda: E.scatter_add(dst, idx, out, ws, accumulate=False))
    print(json.dumps({"dist": dist, "sort_ms": round(s, 4), "scatter_ms": round(w, 4), "GBs": round(b / w / 1e6, 1)}))

"""Gather / scatter-add alone at cfg-4 shapes (n = 8192*200 rows, D = 64, 10 M-row table)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mtamrecommender_b200 import engine as E
from mtamrecommender_b200.synth import ZipfSampler

def main(n=8192 * 200, D=64, rows=10_000_003, reps=10, dist="zipf_pad"):
    dev = "cuda:0"
    table = torch.empty((rows, D), dtype=torch.float32, device=dev).uniform_(-0.3, 0.3)
    rng = np.random.default_rng(5)
    if dist == "uniform":
        idx_np = rng.integers(0, rows, n).astype(np.int32)
    else:
        idx_np = ZipfSampler(rows - 3, 1.05).sample(rng, n)
        idx_np[rng.random(n) < 0.45] = 0
    idx = torch.from_numpy(idx_np).to(dev)
    out = torch.empty((n, D), dtype=torch.float32, device=dev)
    ws = torch.empty(E.scatter_add_workspace(n, rows, D), dtype=torch.uint8, device=dev)
    dst = torch.zeros((rows, D), dtype=torch.float32, device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    def t(fn):
        for _ in range(3): fn()
        torch.cuda.synchronize(); ev0.record()
        for _ in range(reps): fn()
        ev1.record(); torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / reps
    nu = int(np.unique(idx_np).size)
    g = t(lambda: E.gather(table, idx, out)); s = t(lambda: E.scatter_add(dst, idx, out, ws))
    bg, bs = n * (4 + 8 * D), n * (4 + 4 * D) + nu * 4 * D
    print(json.dumps({"dist": dist, "n": n, "unique": nu, "gather_ms": g, "gather_GBs": bg / g / 1e6, "scatter_ms": s, "scatter_GBs": bs / s / 1e6}))

if __name__ == "__main__":
    main(dist=sys.argv[1] if len(sys.argv) > 1 else "zipf_pad", reps=int(sys.argv[2]) if len(sys.argv) > 2 else 10)

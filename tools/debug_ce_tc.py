"""Developer check: tensor-core CE path (gemm_mode=1) against the exact-fp32 path on the same engine inputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import mtam_oracle as O
from mtamrecommender_b200 import engine as E


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def run(D, L, N, H, B, items, users, cats):
    cfg = O.OracleConfig(kind=O.MTAM, L=L, D=D, H=H, N=N, user_count=users, item_count=items, category_count=cats)
    P = O.init_params(cfg, 7)
    feed = O.synth_batch(cfg, B, 9)
    res = []
    for mode in (0, 1):
        eng = E.Engine(E.ModelConfig(kind="MTAM", max_batch=B, L=L, D=D, H=H, N=N, user_count=users, item_count=items,
                                     category_count=cats, gemm_mode=mode))
        eng.set_params(P)
        out = eng.forward(feed)
        g = eng.gradients(feed)
        res.append((out, g))
    (o0, g0), (o1, g1) = res
    print(f"D={D} B={B} V={items+3}: loss {o0['loss']:.7f} / {o1['loss']:.7f}  loss_origin rel {rel(o1['loss_origin'], o0['loss_origin']):.2e}")
    worst = sorted(((rel(g1[k], g0[k]), k) for k in g0 if isinstance(g0[k], np.ndarray) and np.any(g0[k])), reverse=True)[:5]
    for e, k in worst:
        print(f"   grad {k}: {e:.2e}")


if __name__ == "__main__":
    run(64, 12, 2, 1, 37, 500, 50, 11)
    run(64, 50, 6, 1, 130, 5000, 1000, 100)
    run(32, 7, 1, 4, 5, 90, 9, 4)
    run(64, 20, 2, 1, 300, 20000, 1000, 100)

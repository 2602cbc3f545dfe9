"""Per-tensor gradient error of the tensor-core mode at cfg3 shapes (diagnostic)."""
import sys
sys.path.insert(0, ".")
import numpy as np
from oracle import mtam_oracle as O
from mtamrecommender_b200 import engine as E
from mtamrecommender_b200.synth import ZipfSampler, synth_feed


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


B, L, D, N, items, cats, users = 1024, 50, 64, 6, 100_000, 1_000, 100_000
cfg = O.OracleConfig(kind=O.MTAM, L=L, D=D, H=1, N=N, user_count=users, item_count=items, category_count=cats)
P = O.init_params(cfg, 1234)
feed = synth_feed(B, L, items, cats, users, 4321, ZipfSampler(items, 1.05))
fwd, grads, pieces = O.loss_and_grads(cfg, P, feed)
gn = O.global_norm(pieces)
print("oracle global norm", gn)
for mode in (0, 1):
    eng = E.Engine(E.ModelConfig(kind="MTAM", max_batch=B, L=L, D=D, H=1, N=N, user_count=users, item_count=items,
                                 category_count=cats, gemm_mode=mode))
    eng.set_params(P)
    g = eng.gradients(feed)
    print("mode", mode, "norm", np.sqrt(g["__norm_sq__"]), "rel", abs(np.sqrt(g["__norm_sq__"]) - gn) / gn)
    rows = []
    for k, v in grads.items():
        if v is None:
            continue
        rows.append((rel(g[k], v), k, float(np.linalg.norm(v))))
    for r, k, nv in sorted(rows, reverse=True)[:12]:
        print(f"   {r:.3e}  |g|={nv:.4e}  {k}")
    pass
    if mode == 1:
        k = "NextItemDecoder/decoder/num_blocks_1/dense_1/kernel"
        e = g[k] - grads[k]
        col = np.linalg.norm(e, axis=0)
        print("   per-column error of", k, ": top", np.argsort(-col)[:4], np.sort(-col)[:4] * -1, " total", np.linalg.norm(e))
        kb = "NextItemDecoder/decoder/num_blocks_1/dense_1/bias"
        eb = g[kb] - grads[kb]
        print("   bias error: top", np.argsort(-np.abs(eb))[:4], eb[np.argsort(-np.abs(eb))[:4]])

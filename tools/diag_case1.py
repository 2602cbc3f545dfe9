import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
from oracle import mtam_oracle as O
from test_mtam_gpu import make, CASES
cfg, P, feed, eng = make(**CASES[1])
fwd, grads, pieces = O.loss_and_grads(cfg, P, feed)
_, g32, _ = O.loss_and_grads(cfg, P, feed, torch.float32)
g = eng.gradients(feed)
k = "position_embedding/dense4emb/kernel"
for name, x in (("cuda", g[k]), ("oracle32", g32[k])):
    e = x - grads[k]
    col = np.linalg.norm(e, axis=0); row = np.linalg.norm(e, axis=1)
    print(name, "total", np.linalg.norm(e) / np.linalg.norm(grads[k]), "top cols", np.argsort(-col)[:5], np.sort(col)[::-1][:5], "top rows", np.argsort(-row)[:3], np.sort(row)[::-1][:3])

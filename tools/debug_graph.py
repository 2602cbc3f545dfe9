import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from test_mtam_gpu import make
cfg, P, feed, eng = make(D=64, L=12, N=2, H=1, B=32, items=500, users=50, cats=11)
def reset():
    eng.set_params(P); eng.adam_m.zero_(); eng.adam_v.zero_(); eng.set_adam_step(0); eng.grads.zero_()
def diff(a, b, tag):
    d = (a - b).abs()
    names = []
    for k in eng.info:
        pi = eng.info[k]
        va = a.as_strided((pi.rows, pi.cols), (pi.ld, 1), int(pi.offset)); vb = b.as_strided((pi.rows, pi.cols), (pi.ld, 1), int(pi.offset))
        m = float((va - vb).abs().max())
        if m > 0: names.append((k, m))
    print(tag, "max diff", float(d.max()), names[:8])
eng.train_step(feed, 1e-3); r1 = eng.params.clone(); g1 = eng.grads.clone()
reset(); eng.train_step(feed, 1e-3); r2 = eng.params.clone()
diff(r1, r2, "eager vs eager")
reset(); eng.upload(feed); eng.capture_train_graph(32)
reset(); eng.upload(feed); eng.train_step_graph(1e-3); torch.cuda.synchronize(); r3 = eng.params.clone()
diff(r1, r3, "eager vs graph")
reset(); eng.upload(feed); eng.train_step_graph(1e-3); torch.cuda.synchronize(); r4 = eng.params.clone()
diff(r3, r4, "graph vs graph")
print("scalars", eng.read_scalars())

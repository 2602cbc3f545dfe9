import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_mtam_gpu import make, CASES, rel
from oracle import mtam_oracle as O
cfg, P, feed, eng = make(**CASES[1])
fwd, grads, pieces = O.loss_and_grads(cfg, P, feed)
b = eng.upload(feed)
eng.forward_backward_device(b)
dense_gpu = eng.grad_view("embedding_layer/item").cpu().numpy().copy()
# oracle dense part: last piece appended for item is Ts.grad; recompute
p = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in P.items()}
Ts = torch.tensor(P["embedding_layer/item"], dtype=torch.float64, requires_grad=True)
f2 = O.forward(cfg, p, feed, torch.float64, item_table_for_scores=Ts)
f2["loss"].backward()
dense_ref = Ts.grad.numpy()
print("dense part rel", rel(dense_gpu, dense_ref))
eng.finish_grads()
tot = eng.grad_view("embedding_layer/item").cpu().numpy()
ref = grads["embedding_layer/item"]
print("total rel", rel(tot, ref))
sp_gpu, sp_ref = tot - dense_gpu, ref - dense_ref
print("sparse part rel", rel(sp_gpu, sp_ref))
err = np.abs(tot - ref).max(1)
for r in np.argsort(-err)[:6]:
    cnt = int((feed["item_list"] == r).sum())
    print("row", r, "err", err[r], "|ref|", np.abs(ref[r]).max(), "count", cnt, "dense err", np.abs(dense_gpu[r]-dense_ref[r]).max(), "sparse err", np.abs(sp_gpu[r]-sp_ref[r]).max())
print("pred rel", rel(eng.forward(feed)["pred"], fwd["pred"].detach().numpy()))
for k in ("embedding_layer/category", "embedding_layer/position", "position_embedding/dense4emb/kernel"):
    print(k, rel(eng.grad_view(k).cpu().numpy(), grads[k]))

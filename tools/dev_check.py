"""Developer parity sweep: runs every stage of the CUDA path against the CPU oracle and prints the
error of each tensor instead of stopping at the first mismatch.  Not part of the product."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import mtam_oracle as O
from mtamrecommender_b200 import engine as E


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def check_gather_scatter():
    dev = "cuda:0"
    g = torch.Generator().manual_seed(0)
    for (R, D, n) in [(1000, 64, 5000), (53, 64, 51200), (100003, 64, 51200), (300, 128, 1), (1 << 17, 32, 200000)]:
        table = torch.randn(R, D, generator=g)
        idx = torch.randint(0, R, (n,), generator=g, dtype=torch.int32)
        idx[: n // 3] = 0
        out = E.gather(table.to(dev), idx.to(dev)).cpu()
        ok = torch.equal(out, table[idx.long()])
        rows = torch.randn(n, D, generator=g)
        dst = torch.zeros(R, D, device=dev)
        _, uq, nu = E.scatter_add(dst, idx.to(dev), rows.to(dev), want_unique=True)
        ref = np.zeros((R, D)); np.add.at(ref, idx.numpy(), rows.numpy().astype(np.float64))
        dst2 = torch.zeros(R, D, device=dev)
        E.scatter_add(dst2, idx.to(dev), rows.to(dev))
        un = np.unique(idx.numpy())
        uok = int(nu.item()) == len(un) and np.array_equal(uq[: len(un)].cpu().numpy(), un)
        print(f"gather/scatter R={R} D={D} n={n}: gather_exact={ok} scatter_rel={rel(dst.cpu().numpy(), ref):.2e} "
              f"deterministic={torch.equal(dst, dst2)} unique_ok={uok}")


def check_mtam(D=64, L=12, N=2, H=1, B=37, items=500, users=50, cats=11, steps=3):
    cfg = O.OracleConfig(kind=O.MTAM, L=L, D=D, H=H, N=N, user_count=users, item_count=items, category_count=cats)
    P = O.init_params(cfg, 7)
    # make bias-like params non-trivial so that their gradients are exercised
    rng = np.random.default_rng(3)
    for k in P:
        if k.endswith("/bias") or k.endswith("/beta"):
            P[k] = (P[k] + 0.1 * rng.standard_normal(P[k].shape)).astype(np.float32)
    feed = O.synth_batch(cfg, B, 11)
    mc = E.ModelConfig(kind="MTAM", max_batch=B, L=L, D=D, H=H, N=N, user_count=users, item_count=items, category_count=cats)
    eng = E.Engine(mc)
    eng.set_params(P)
    fwd, grads, pieces = O.loss_and_grads(cfg, P, feed)
    out = eng.forward(feed)
    print(f"[MTAM D={D} L={L} N={N} H={H} B={B}] loss oracle={float(fwd['loss'].detach()):.7f} cuda={out['loss']:.7f}  "
          f"l2 {float(fwd['l2_norm'].detach()):.6f}/{out['l2_norm']:.6f}")
    print("  pred rel", rel(out["pred"], fwd["pred"].detach().numpy()), " loss_origin rel",
          rel(out["loss_origin"], fwd["loss_origin"].detach().numpy()))
    g = eng.gradients(feed)
    gn = O.global_norm(pieces)
    print(f"  global_norm oracle={gn:.7f} cuda={np.sqrt(g['__norm_sq__']):.7f}")
    worst = []
    for k, v in grads.items():
        if v is None:
            z = float(np.abs(g[k]).max())
            if z != 0: worst.append((k, "expected-zero", z))
            continue
        worst.append((k, rel(g[k], v), float(np.abs(v).max())))
    for k, r, m in sorted(worst, key=lambda x: -x[1] if not isinstance(x[1], str) else -1e9)[:12]:
        print(f"    grad {k:70s} rel={r} max|ref|={m:.3e}")
    tr = O.OracleTrainer(cfg, P)
    for s in range(steps):
        lo = tr.train_step(feed, 1e-3)
        lc = eng.train_step(feed, 1e-3)
        print(f"  step {s}: loss oracle={lo:.7f} cuda={lc:.7f}")
    newp = eng.get_params()
    w = sorted(((rel(newp[k], tr.params[k]), k) for k in newp), reverse=True)[:6]
    for r, k in w:
        print(f"    param-after-{steps}-steps {k:60s} rel={r:.3e}")
    # top-k
    b = eng.upload(feed)
    idx, sc = eng.eval_topk_device(b, 50)
    m, oidx, oscores = O.metrics_topk(cfg, tr.params, feed)
    print("  topk exact:", np.array_equal(idx.cpu().numpy(), oidx), " hr/ndcg cuda:",
          eng.hr_ndcg_device(idx, b.t["target_item_id"]).cpu().numpy().round(4), " oracle:", np.round(m, 4))


if __name__ == "__main__":
    torch.manual_seed(0)
    print(torch.cuda.get_device_name(0))
    check_gather_scatter()
    check_mtam()
    check_mtam(D=128, L=50, N=3, H=8, B=64, items=3706, users=300, cats=301, steps=2)
    check_mtam(D=32, L=7, N=1, H=4, B=5, items=90, users=9, cats=4, steps=2)

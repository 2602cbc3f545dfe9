"""PISTRec and the three attention baselines (self-attention family, Tq = L) against the CPU oracle:
same inputs and weights, loss / pred / every gradient / 3 Adam steps / top-k."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import mtam_oracle as O  # noqa: E402
from conftest import parity_tol  # noqa: E402


def _test_name():
    import os
    return os.environ.get("PYTEST_CURRENT_TEST", "?").split("::", 1)[-1].split(" ")[0]


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def make(kind, D, L, N, H, B, items, users, cats, seed=3):
    from mtamrecommender_b200 import engine as E
    cfg = O.OracleConfig(kind=kind, L=L, D=D, H=H, N=N, user_count=users, item_count=items, category_count=cats)
    P = O.init_params(cfg, seed)
    rng = np.random.default_rng(seed + 1)
    for k in P:
        if k.endswith("/bias") or k.endswith("/beta"):
            P[k] = (P[k] + 0.1 * rng.standard_normal(P[k].shape)).astype(np.float32)
    feed = O.synth_batch(cfg, B, seed + 2)
    eng = E.Engine(E.ModelConfig(kind=kind, max_batch=B, L=L, D=D, H=H, N=N, user_count=users, item_count=items,
                                 category_count=cats))
    eng.set_params(P)
    return cfg, P, feed, eng


KINDS = [O.PISTREC, O.TA_SASREC, O.TISASREC, O.SASREC]
SHAPES = [dict(D=64, L=10, N=2, H=2, B=13, items=300, users=30, cats=7),
          dict(D=128, L=50, N=2, H=8, B=20, items=3706, users=100, cats=301)]   # ml-1m-shaped (cfg2)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("shape", SHAPES)
def test_forward_and_gradients(kind, shape):
    import torch
    cfg, P, feed, eng = make(kind, **shape)
    fwd, grads, pieces = O.loss_and_grads(cfg, P, feed)
    _, g32, _ = O.loss_and_grads(cfg, P, feed, torch.float32)
    out = eng.forward(feed)
    assert abs(out["loss"] - float(fwd["loss"].detach())) <= 1e-5 * abs(float(fwd["loss"].detach()))
    assert rel(out["pred"], fwd["pred"].detach().numpy()) < 1e-5
    assert rel(out["loss_origin"], fwd["loss_origin"].detach().numpy()) < 1e-5
    g = eng.gradients(feed)
    gn = O.global_norm(pieces)
    assert abs(np.sqrt(g["__norm_sq__"]) - gn) <= 1e-5 * gn
    for k, v in grads.items():
        if v is None:
            assert not np.any(g[k]), f"{k}: no gradient expected"
        else:
            tol = parity_tol(_test_name(), k, rel(g32[k], v))
            assert rel(g[k], v) < tol, (k, rel(g[k], v), tol)
    if kind == O.PISTREC:      # PISTRec_model.py:56-60: no user L2 term -> user table untouched
        assert grads["embedding_layer/user"] is None


@pytest.mark.parametrize("kind", KINDS)
def test_three_steps_and_topk(kind):
    import torch
    cfg, P, feed, eng = make(kind, **SHAPES[0])
    tr = O.OracleTrainer(cfg, P)
    tr32 = O.OracleTrainer(cfg, P, dtype=torch.float32)
    for s in range(3):
        lo, lc = tr.train_step(feed, 1e-3), eng.train_step(feed, 1e-3)
        tr32.train_step(feed, 1e-3)
        assert abs(lo - lc) <= 2e-5 * abs(lo), (s, lo, lc)
    newp = eng.get_params()
    for k, v in tr.params.items():
        tol = parity_tol(_test_name(), k, rel(tr32.params[k], v))
        assert rel(newp[k], v) < tol, (k, rel(newp[k], v), tol)
    b = eng.upload(feed)
    idx, _ = eng.eval_topk_device(b, 50)
    m, oidx, osc = O.metrics_topk(cfg, {k: v.astype(np.float32) for k, v in tr.params.items()}, feed)
    srt = -np.sort(-osc, axis=1)[:, :51]
    ok = np.abs(np.diff(srt, axis=1)).min(axis=1) > 1e-5
    assert np.array_equal(idx.cpu().numpy()[ok], oidx[ok]) and ok.mean() > 0.8

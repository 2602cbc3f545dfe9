"""MTAM_via_T_GRU (Model/MTAMRec_model.py:167-204) on the CUDA path against the oracle: the hops read the T-GRU's
output sequence (zeros from step seq_len-1 on) as their memory, the query is layer-normed first, and the gradient of
every output step flows back through the recurrence.  Loss, pred, every gradient, three Adam steps, top-50, both
arithmetic modes; the committed golden vectors of the oracle are checked too."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import mtam_oracle as O  # noqa: E402
from conftest import grad_close  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _name():
    return os.environ.get("PYTEST_CURRENT_TEST", "?").split("::", 1)[-1].split(" ")[0]


def make(D, L, N, H, B, items, users, cats, seed=7, gemm_mode=0, kind=None):
    from mtamrecommender_b200 import engine as E
    kind = kind or O.MTAM_VIA_T_GRU
    cfg = O.OracleConfig(kind=kind, L=L, D=D, H=H, N=N, user_count=users, item_count=items, category_count=cats)
    P = O.init_params(cfg, seed)
    rng = np.random.default_rng(seed + 1)
    for k in P:
        if k.endswith("/bias") or k.endswith("/beta"):
            P[k] = (P[k] + 0.1 * rng.standard_normal(P[k].shape)).astype(np.float32)
    feed = O.synth_batch(cfg, B, seed + 2)
    eng = E.Engine(E.ModelConfig(kind=kind, max_batch=B, L=L, D=D, H=H, N=N, user_count=users, item_count=items,
                                 category_count=cats, gemm_mode=gemm_mode))
    eng.set_params(P)
    return cfg, P, feed, eng


CASES = [dict(D=64, L=12, N=2, H=1, B=37, items=500, users=50, cats=11),
         dict(D=128, L=50, N=3, H=8, B=40, items=3706, users=300, cats=301),
         dict(D=32, L=7, N=1, H=4, B=5, items=90, users=9, cats=4),
         dict(D=64, L=50, N=6, H=1, B=130, items=5000, users=1000, cats=100)]


# MTAM_via_T_GRU, and the two plain-GRU siblings (MTAM_no_time_aware_rnn :93-125, MTAM_via_rnn :206-233)
@pytest.mark.parametrize("kind", [O.MTAM_VIA_T_GRU, O.MTAM_NO_TA_RNN, O.MTAM_VIA_RNN])
@pytest.mark.parametrize("gemm_mode", [0, 1])
@pytest.mark.parametrize("case", CASES)
def test_forward_gradients_and_steps(case, gemm_mode, kind):
    import torch
    cfg, P, feed, eng = make(**case, gemm_mode=gemm_mode, kind=kind)
    assert set(eng.param_names()) == set(P)
    fwd, grads, pieces = O.loss_and_grads(cfg, P, feed)
    _, g32, _ = O.loss_and_grads(cfg, P, feed, torch.float32)
    out = eng.forward(feed)
    lo = float(fwd["loss"].detach())
    assert abs(out["loss"] - lo) <= 1e-5 * abs(lo)
    assert rel(out["pred"], fwd["pred"].detach().numpy()) < 1e-5
    assert rel(out["loss_origin"], fwd["loss_origin"].detach().numpy()) < 1e-5
    g = eng.gradients(feed)
    gn = O.global_norm(pieces)
    assert abs(np.sqrt(g["__norm_sq__"]) - gn) <= 1e-5 * gn
    for k, v in grads.items():
        if v is None:
            assert not np.any(g[k]), k
        else:
            ok, r, tol = grad_close(_name(), k, g[k], v, g32[k])
            assert ok, (k, r, tol)
    tr = O.OracleTrainer(cfg, P)
    tr32 = O.OracleTrainer(cfg, P, dtype=torch.float32)
    for s in range(3):
        lo, lc = tr.train_step(feed, 1e-3), eng.train_step(feed, 1e-3)
        tr32.train_step(feed, 1e-3)
        assert abs(lo - lc) <= 2e-5 * abs(lo), (s, lo, lc)
    newp = eng.get_params()
    for k, v in tr.params.items():
        ok, r, tol = grad_close(_name() + " weights", k, newp[k], v, tr32.params[k])
        assert ok, (k, r, tol)
    idx, _ = eng.eval_topk_device(eng.upload(feed), 50)
    _, oidx, osc = O.metrics_topk(cfg, newp, feed)
    srt = -np.sort(-osc, axis=1)[:, :51]
    ok = np.abs(np.diff(srt, axis=1)).min(axis=1) > 1e-4
    assert np.array_equal(idx.cpu().numpy()[ok], oidx[ok])


def test_golden_vectors_of_the_oracle():
    """tests/golden/MTAM_VIA_T_GRU.npz (frozen outputs of the oracle, made by tests/golden/make_golden.py and committed in
    round 1 as the target of this CUDA path): same seeded inputs and weights, the CUDA step reproduces the stored loss,
    pred, global norm, gradients, the three training losses and the weights after them."""
    import importlib.util
    from mtamrecommender_b200 import engine as E
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    cfg, P, feed = mg.build("MTAM_VIA_T_GRU")
    z = np.load(os.path.join(GOLD, "MTAM_VIA_T_GRU.npz"))
    eng = E.Engine(E.ModelConfig(kind="MTAM_VIA_T_GRU", max_batch=len(feed["user_id"]), L=cfg.L, D=cfg.D, H=cfg.H, N=cfg.N,
                                 user_count=cfg.user_count, item_count=cfg.item_count, category_count=cfg.category_count))
    eng.set_params(P)
    out = eng.forward(feed)
    assert abs(out["loss"] - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
    assert rel(out["pred"], z["pred"]) < 1e-5
    g = eng.gradients(feed)
    assert abs(np.sqrt(g["__norm_sq__"]) - float(z["global_norm"])) <= 1e-5 * float(z["global_norm"])
    for k in z.files:
        if k.startswith("grad:"):
            assert rel(g[k[5:]], z[k]) < 1e-4, k
    losses = [eng.train_step(feed, 1e-3) for _ in range(3)]
    assert np.allclose(losses, z["losses3"], rtol=2e-5)
    newp = eng.get_params()
    for k in z.files:
        if k.startswith("after3:"):
            assert rel(newp[k[7:]], z[k]) < 1e-4, k

"""PISTRec and the three attention baselines (self-attention family, Tq = L) against the CPU oracle:
same inputs and weights, loss / pred / every gradient / 3 Adam steps / top-k."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import mtam_oracle as O  # noqa: E402
from conftest import grad_close, parity_tol  # noqa: E402


def _test_name():
    import os
    return os.environ.get("PYTEST_CURRENT_TEST", "?").split("::", 1)[-1].split(" ")[0]


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def make(kind, D, L, N, H, B, items, users, cats, seed=3, gemm_mode=0, **engine_kw):
    from mtamrecommender_b200 import engine as E
    cfg = O.OracleConfig(kind=kind, L=L, D=D, H=H, N=N, user_count=users, item_count=items, category_count=cats)
    P = O.init_params(cfg, seed)
    rng = np.random.default_rng(seed + 1)
    for k in P:
        if k.endswith("/bias") or k.endswith("/beta"):
            P[k] = (P[k] + 0.1 * rng.standard_normal(P[k].shape)).astype(np.float32)
    feed = O.synth_batch(cfg, B, seed + 2)
    eng = E.Engine(E.ModelConfig(kind=kind, max_batch=B, L=L, D=D, H=H, N=N, user_count=users, item_count=items,
                                 category_count=cats, gemm_mode=gemm_mode, **engine_kw))
    eng.set_params(P)
    return cfg, P, feed, eng


KINDS = [O.PISTREC, O.TA_SASREC, O.TISASREC, O.SASREC]
SHAPES = [dict(D=64, L=10, N=2, H=2, B=13, items=300, users=30, cats=7),
          dict(D=128, L=50, N=2, H=8, B=20, items=3706, users=100, cats=301)]   # ml-1m-shaped (cfg2)


# both arithmetic modes of the dense contractions: exact-fp32 FFMA tiles (0) and tcgen05 3xTF32 (1, the product default)
@pytest.mark.parametrize("gemm_mode", [0, 1])
@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("shape", SHAPES)
def test_forward_and_gradients(kind, shape, gemm_mode):
    import torch
    cfg, P, feed, eng = make(kind, **shape, gemm_mode=gemm_mode)
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
            ok, r, tol = grad_close(_test_name(), k, g[k], v, g32[k])
            assert ok, (k, r, tol)
    if kind == O.PISTREC:      # PISTRec_model.py:56-60: no user L2 term -> user table untouched
        assert grads["embedding_layer/user"] is None


@pytest.mark.parametrize("gemm_mode", [0, 1])
@pytest.mark.parametrize("kind", KINDS)
def test_three_steps_and_topk(kind, gemm_mode):
    import torch
    cfg, P, feed, eng = make(kind, **SHAPES[0], gemm_mode=gemm_mode)
    tr = O.OracleTrainer(cfg, P)
    tr32 = O.OracleTrainer(cfg, P, dtype=torch.float32)
    for s in range(3):
        lo, lc = tr.train_step(feed, 1e-3), eng.train_step(feed, 1e-3)
        tr32.train_step(feed, 1e-3)
        assert abs(lo - lc) <= 2e-5 * abs(lo), (s, lo, lc)
    newp = eng.get_params()
    for k, v in tr.params.items():
        ok, r, tol = grad_close(_test_name(), k, newp[k], v, tr32.params[k])
        assert ok, (k, r, tol)
    b = eng.upload(feed)
    idx, _ = eng.eval_topk_device(b, 50)
    m, oidx, osc = O.metrics_topk(cfg, {k: v.astype(np.float32) for k, v in tr.params.items()}, feed)
    srt = -np.sort(-osc, axis=1)[:, :51]
    ok = np.abs(np.diff(srt, axis=1)).min(axis=1) > 1e-5
    assert np.array_equal(idx.cpu().numpy()[ok], oidx[ok]) and ok.mean() > 0.8


@pytest.mark.parametrize("kind", [O.SASREC, O.TISASREC])
@pytest.mark.parametrize("shape", SHAPES)
def test_attention_dropout_with_the_same_mask(kind, shape):
    """Attention dropout (multihead_attention.py:179, time_aware_attention.py:198; rate 0.5 in every preset): the oracle
    takes the keep mask as an input, the library generates it from (seed, forward-call counter, block, element); with the
    host restatement of that hash both see the same mask, and loss / pred / gradients agree to the usual tolerances."""
    import torch
    from mtamrecommender_b200 import engine as E
    rate, seed, call = 0.5, 77, 5
    cfg, P, feed, eng = make(kind, **shape, dropout=rate, dropout_seed=seed)
    B = len(feed["user_id"])
    masks = [torch.tensor(E.dropout_keep_mask(seed, call, i, B, cfg.H, cfg.L, rate), dtype=torch.float64) for i in range(cfg.N)]
    kept = float(np.mean([float((m > 0).double().mean()) for m in masks]))
    assert abs(kept - (1 - rate)) < 0.02                                  # the hash keeps 1-rate of the weights
    fwd, grads, pieces = O.loss_and_grads(cfg, P, feed, drop_masks=masks)
    _, g32, _ = O.loss_and_grads(cfg, P, feed, torch.float32, drop_masks=[m.float() for m in masks])
    eng.set_dropout_state(seed, call)
    out = eng.forward(feed)
    assert abs(out["loss"] - float(fwd["loss"].detach())) <= 1e-5 * abs(float(fwd["loss"].detach()))
    assert rel(out["pred"], fwd["pred"].detach().numpy()) < 1e-5
    eng.set_dropout_state(seed, call)
    g = eng.gradients(feed)
    for k, v in grads.items():
        if v is not None:
            ok, r, tol = grad_close(_test_name(), k, g[k], v, g32[k])
            assert ok, (k, r, tol)
    # a different call counter draws a different mask; rate 0 reproduces the no-dropout model
    other = eng.forward(feed)
    assert abs(other["loss"] - out["loss"]) > 1e-6
    # the time-aware kinds have no dropout in the reference: the flag is inert there
    cfg2, P2, feed2, eng2 = make(O.PISTREC, **shape, dropout=rate)
    fwd2 = O.forward(cfg2, {k: torch.tensor(v, dtype=torch.float64) for k, v in P2.items()}, feed2)
    assert abs(eng2.forward(feed2)["loss"] - float(fwd2["loss"])) <= 1e-5 * abs(float(fwd2["loss"]))


def test_dropout_mask_follows_the_step_through_a_cuda_graph():
    """A replayed CUDA graph runs no host code: the call counter of the step reaches the kernels through device memory,
    written by mtam_prepare_step -- eager steps and graph replays draw the same sequence of masks."""
    import torch
    cfg, P, feed, eng = make(O.SASREC, **SHAPES[0], dropout=0.5, dropout_seed=9)
    eng.set_dropout_state(9, 100)
    eager = [eng.train_step(feed, 1e-3) for _ in range(3)]
    ref = eng.params.clone()
    eng.set_params(P); eng.adam_m.zero_(); eng.adam_v.zero_(); eng.set_adam_step(0)
    eng.capture_train_graph(len(feed["user_id"]))
    eng.set_dropout_state(9, 100)
    graph = [eng.train_step(feed, 1e-3) for _ in range(3)]
    assert eager == graph and bool((ref == eng.params).all())
    assert len(set(eager)) == 3


def test_plain_gradient_descent_optimizer():
    """FLAGS.optimizer other than adam / adadelta / rmsprop -> tf.train.GradientDescentOptimizer (base_model.py:79-80):
    w -= lr * clipped gradient."""
    cfg, P, feed, eng = make(O.TA_SASREC, **SHAPES[0], optimizer="sgd")
    _, grads, pieces = O.loss_and_grads(cfg, P, feed)
    sc = O.clip_scale(O.global_norm(pieces), cfg.clip)
    eng.train_step(feed, 0.05)
    newp = eng.get_params()
    for k, g in grads.items():
        want = P[k].astype(np.float64) - (0.0 if g is None else np.float32(0.05) * sc * g)
        assert rel(newp[k], want) < 1e-6, k

"""Pins the CPU oracle: committed golden vectors, hand-computed known answers for every TF 1.14
semantic it relies on (SURVEY 9.7 and the traps T1/T2), and finite differences for its gradients."""
import math
import os
import sys

import numpy as np
import pytest
import torch

from oracle import mtam_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402


@pytest.mark.parametrize("kind", list(make_golden.CASES))
def test_oracle_matches_committed_golden_vectors(kind):
    g = np.load(os.path.join(HERE, "golden", f"{kind}.npz"))
    cfg, P, feed = make_golden.build(kind)
    fwd, grads, pieces = O.loss_and_grads(cfg, P, feed, bpr_negative=7)
    assert abs(float(fwd["loss"].detach()) - float(g["loss"])) < 1e-12
    assert np.allclose(fwd["pred"].detach().numpy(), g["pred"], rtol=0, atol=1e-12)
    assert abs(O.global_norm(pieces) - float(g["global_norm"])) < 1e-10
    for k, v in grads.items():
        if v is not None:
            assert np.allclose(v, g["grad:" + k], rtol=1e-5, atol=1e-9), k
        else:
            assert "grad:" + k not in g
    tr = O.OracleTrainer(cfg, P)
    losses = [tr.train_step(feed, 1e-3, bpr_negative=7) for _ in range(3)]
    assert np.allclose(losses, g["losses3"], rtol=0, atol=1e-10)
    for k, v in tr.params.items():
        assert np.allclose(v, g["after3:" + k], rtol=1e-6, atol=1e-8), k
    assert np.array_equal(O.metrics_topk(cfg, P, feed)[1], g["top50"])


# ---- known answers ---------------------------------------------------------------------------
def test_mask_constant_and_exact_zero_probability():
    assert np.float32(O.MASK_VALUE) == np.float32(-4294967296.0)        # -2**32+1 rounds to -2**32 in fp32
    s = torch.tensor([[0.3, O.MASK_VALUE, -0.2, O.MASK_VALUE]], dtype=torch.float32)
    p = torch.softmax(s, -1)
    assert p[0, 1] == 0 and p[0, 3] == 0 and abs(float(p.sum()) - 1) < 1e-6


def test_two_layer_norms_have_different_eps():
    d = 1e-5
    x = torch.tensor([[d, -d]], dtype=torch.float64)
    one, zero = torch.ones(2, dtype=torch.float64), torch.zeros(2, dtype=torch.float64)
    blk = O._ln(x, one, zero, O.LN_EPS_BLOCK)       # custom normalize: eps 1e-8 (time_aware_attention.py:7-34)
    fin = O._ln(x, one, zero, O.LN_EPS_FINAL)       # tf.contrib layer_norm: eps 1e-12 (net_utils.py:229-232)
    assert abs(float(blk[0, 0]) - d / math.sqrt(d * d + 1e-8)) < 1e-12      # 0.0995...
    assert abs(float(fin[0, 0]) - d / math.sqrt(d * d + 1e-12)) < 1e-12     # 0.995...
    assert abs(float(blk[0, 0]) - 0.09950371902) < 1e-9 and abs(float(fin[0, 0]) - 0.99503719021) < 1e-9


def test_tgru_step_hand_computed_and_length_masking():
    D = 1
    g = "g/"
    p = {g + "gates/kernel": torch.tensor([[0.5, -0.25], [0.125, 0.75]], dtype=torch.float64),   # [2D,2D]: r | u
         g + "gates/bias": torch.tensor([1.0, 1.0], dtype=torch.float64),
         g + "candidate/kernel": torch.tensor([[0.3], [-0.6]], dtype=torch.float64),
         g + "candidate/bias": torch.tensor([0.1], dtype=torch.float64)}
    vals = dict(_time_kernel_w1=0.7, _time_kernel_b1=-0.1, _time_history_w1=0.2, _time_w1=0.05, _time_b1=0.3,
                _time_kernel_w2=1.1, _time_w12=-0.4, _time_b12=0.25)
    for k, v in vals.items():
        p[g + k] = torch.tensor([v], dtype=torch.float64)
    X = torch.tensor([[[0.8], [-0.5], [9.0], [9.0]]], dtype=torch.float64)      # B=1, L=4
    tl = torch.tensor([[0.0, 3.0, 7.0, 7.0]], dtype=torch.float64)
    out = O.tgru_new(X, tl, torch.tensor([3]), p, g)                            # seq_len 3 -> 2 live steps
    sig = lambda z: 1 / (1 + math.exp(-z))
    h = 0.0
    exp = []
    for x, dt in ((0.8, 0.0), (-0.5, 3.0)):
        a = max(x * 0.7 - 0.1 + h * 0.2, 0.0)
        s = max(0.05 * dt + 0.3, 0.0)
        T = sig(1.1 * a - 0.4 * s + 0.25)
        r = sig(x * 0.5 + h * 0.125 + 1.0)
        u = sig(x * -0.25 + h * 0.75 + 1.0)
        c = math.tanh(x * 0.3 + (r * h) * -0.6 + 0.1)
        h = u * h + (1 - u) * c * T
        exp.append(h)
    assert np.allclose(out[0, :2, 0].numpy(), exp, atol=1e-14)
    assert float(out[0, 2, 0]) == 0.0 and float(out[0, 3, 0]) == 0.0           # zero output past length


def test_topk_ties_prefer_lower_index():
    s = np.array([[1.0, 3.0, 3.0, 0.5, 3.0], [2.0, 2.0, 2.0, 2.0, 2.0]])
    assert O.topk_indices(s, 3).tolist() == [[1, 2, 4], [0, 1, 2]]


def test_global_norm_uses_undeduplicated_slices():
    cfg = O.OracleConfig(kind=O.MTAM, L=5, D=32, H=1, N=1, user_count=4, item_count=9, category_count=3)
    P = O.init_params(cfg, 3)
    feed = O.synth_batch(cfg, 6, 1)            # pad id 0 and the mask token repeat many times
    fwd, grads, pieces = O.loss_and_grads(cfg, P, feed)
    gn = O.global_norm(pieces)
    dedup = math.sqrt(sum(float((g ** 2).sum()) for g in grads.values() if g is not None))
    ei = fwd["Ei"].grad.numpy().reshape(-1, cfg.D)
    assert any(np.array_equal(ei.shape, p.shape) and np.allclose(ei, p) for p in pieces)
    assert abs(gn - dedup) > 1e-6 * gn         # duplicates are NOT summed before the norm (trap T1)
    assert O.clip_scale(0.5, 1.0) == 1.0 and abs(O.clip_scale(4.0, 1.0) - 0.25) < 1e-15


def test_adam_epsilon_outside_bias_correction():
    w, m, v = O.adam_tf(np.array([1.0]), np.array([1.0]), np.zeros(1), np.zeros(1), lr=0.1, t=1, eps=1.0)
    lr_t = 0.1 * math.sqrt(1 - 0.999) / (1 - 0.9)
    assert abs(w[0] - (1.0 - lr_t * 0.1 / (math.sqrt(0.001) + 1.0))) < 1e-15      # 0.99693...; torch-style gives 0.95
    assert abs(m[0] - 0.1) < 1e-15 and abs(v[0] - 0.001) < 1e-15


def test_lr_schedule_staircase():
    assert O.lr_schedule(0.001, 0.995, 0, 0.001) == 0.001
    assert abs(O.lr_schedule(0.001, 0.995, 250, 0.001) - 0.001 * 0.995 ** 2) < 1e-18
    assert abs(O.lr_schedule(0.01, 0.995, 399, 0.01) - 0.01 * 0.99 ** 3) < 1e-18


def test_hr_ndcg_formula():
    top = np.array([[5, 7, 9], [1, 2, 3], [4, 4, 4]])
    hr, nd = O.hr_ndcg(top, np.array([9, 8, 4]), 3)
    assert abs(hr - 2 / 3) < 1e-15 and abs(nd - (math.log(2) / math.log(4) + 1.0) / 3) < 1e-15


def test_record_construction_and_feed_padding():
    cfg = O.OracleConfig(kind=O.MTAM, L=6, D=32, user_count=4, item_count=9, category_count=3)
    recs = O.synth_records(cfg, 20, 5)
    for r in recs:
        n = r[8]
        assert 2 <= n <= cfg.L and len(r[1]) == n and r[1][-1] == cfg.item_count + 1 and r[2][-1] == cfg.category_count + 1
        assert r[6] == list(range(n)) and r[3][-1] == r[7][2] and r[4][-1] == 0 and r[5][-1] == 0 and r[4][0] == 0
    f = O.make_feed(cfg, recs)
    assert f["item_list"].shape == (20, 6) and np.all(f["item_list"][0, recs[0][8]:] == 0)


# ---- finite differences ----------------------------------------------------------------------
@pytest.mark.parametrize("kind", [O.MTAM, O.PISTREC, O.TISASREC, O.SASREC, O.MTAM_VIA_T_GRU, O.MTAM_NO_TA_RNN, O.MTAM_VIA_RNN])
def test_oracle_gradients_by_finite_differences(kind):
    cfg = O.OracleConfig(kind=kind, L=6, D=32, H=2, N=2, user_count=5, item_count=20, category_count=3)
    P = {k: v.astype(np.float64) for k, v in O.init_params(cfg, 8).items()}
    if kind in O.MEMORY_IS_RNN_KINDS:
        # its memory rows are exactly zero from step seq_len-1 on, so with the zero-initialised biases the K,V ReLU
        # pre-activations of those (unmasked) keys sit exactly on the kink, where no derivative exists: move off it
        brng = np.random.default_rng(5)
        for k in P:
            if k.endswith("/bias"):
                P[k] = P[k] + 0.1 * brng.standard_normal(P[k].shape)
    feed = O.synth_batch(cfg, 4, 9)
    fwd, grads, _ = O.loss_and_grads(cfg, P, feed)
    rng = np.random.default_rng(0)

    def loss_at(Pm):
        with torch.no_grad():
            return float(O.forward(cfg, {k: torch.tensor(v, dtype=torch.float64) for k, v in Pm.items()}, feed)["loss"])
    names = [k for k, v in grads.items() if v is not None]
    for name in rng.choice(names, size=min(8, len(names)), replace=False):
        d = rng.standard_normal(P[name].shape)
        d /= np.linalg.norm(d)
        eps = 1e-5
        Pp, Pm = dict(P), dict(P)
        Pp[name] = P[name] + eps * d
        Pm[name] = P[name] - eps * d
        fd = (loss_at(Pp) - loss_at(Pm)) / (2 * eps)
        an = float((grads[name] * d).sum())
        assert abs(fd - an) <= 1e-6 * max(1.0, abs(an)) + 1e-8, (name, fd, an)


def test_mtam_via_t_gru_restatement_structure():
    """MTAM_via_T_GRU (Model/MTAMRec_model.py:167-204, oracle-only so far): the memory is the T-GRU output sequence --
    zero from step seq_len-1 on (dynamic_rnn) yet unmasked up to seq_len -- and the query is layer-normed inside the
    ShortTermIntentEncoder scope, which adds exactly one LayerNorm pair to MTAM's variables."""
    kw = dict(L=8, D=32, H=2, N=2, user_count=9, item_count=40, category_count=4)
    a, b = O.OracleConfig(kind=O.MTAM, **kw), O.OracleConfig(kind=O.MTAM_VIA_T_GRU, **kw)
    extra = set(O.param_shapes(b)) - set(O.param_shapes(a))
    assert extra == {"ShortTermIntentEncoder/LayerNorm/beta", "ShortTermIntentEncoder/LayerNorm/gamma"}
    assert set(O.param_shapes(a)) <= set(O.param_shapes(b))
    P = {k: torch.tensor(v, dtype=torch.float64) for k, v in O.init_params(b, 3).items()}
    feed = O.synth_batch(b, 5, 11)
    out = O.forward(b, P, feed)
    n = feed["seq_length"]
    rnn = out["rnn"].numpy()
    for r in range(5):
        assert np.all(rnn[r, n[r] - 1:] == 0) and np.any(rnn[r, n[r] - 2] != 0)
    q = out["short_term_intent"].numpy()                 # layer-normed query: zero mean, unit variance (gamma 1, beta 0)
    assert np.allclose(q.mean(1), 0, atol=1e-12) and np.allclose(q.var(1), 1, atol=1e-6)
    # the embedded behaviours reach the prediction only through the T-GRU now: changing the mask-token step's item
    # (position seq_len-1, which the T-GRU never consumes) leaves pred unchanged, unlike in MTAM
    f2 = {k: v.copy() for k, v in feed.items()}
    for r in range(5):
        f2["item_list"][r, n[r] - 1] = 3
    Pa = {k: v for k, v in P.items() if k in O.param_shapes(a)}
    assert np.array_equal(O.forward(b, P, f2)["pred"].numpy(), out["pred"].numpy())
    assert not np.array_equal(O.forward(a, Pa, f2)["pred"].numpy(), O.forward(a, Pa, feed)["pred"].numpy())


@pytest.mark.parametrize("case", [dict(L=9, D=32, H=4, N=2, B=11, users=20, items=120, cats=6),
                                  dict(L=14, D=64, H=1, N=3, B=7, users=9, items=60, cats=4)])
def test_two_independent_restatements_agree(case):
    """oracle/mtam_oracle.py (torch fp64 + autograd) against oracle/mtam_oracle_np.py (NumPy fp64, backward derived by
    hand, written separately from the same reference files): loss, pred, every gradient, the un-deduplicated global norm
    (trap T1) and the weights after three TF-style Adam steps (trap T2) agree to fp64 rounding.  The oracle stays
    unpinned against TensorFlow itself, but a transcription slip would have to be made twice, identically."""
    from oracle import mtam_oracle_np as N2
    cfg = O.OracleConfig(kind=O.MTAM, L=case["L"], D=case["D"], H=case["H"], N=case["N"], user_count=case["users"],
                         item_count=case["items"], category_count=case["cats"])
    P = O.init_params(cfg, 21)
    rng = np.random.default_rng(22)
    for k in P:
        if k.endswith("/bias") or k.endswith("/beta"):
            P[k] = (P[k] + 0.1 * rng.standard_normal(P[k].shape)).astype(np.float32)
    feed = O.synth_batch(cfg, case["B"], 23)
    fwd, grads, pieces = O.loss_and_grads(cfg, P, feed)
    out2, grads2, pieces2 = N2.forward_backward(P, feed, cfg.L, cfg.D, cfg.H, cfg.N, cfg.reg)
    assert abs(out2["loss"] - float(fwd["loss"].detach())) < 1e-12 * abs(out2["loss"])
    assert np.allclose(out2["pred"], fwd["pred"].detach().numpy(), rtol=1e-11, atol=1e-13)
    assert np.allclose(out2["loss_origin"], fwd["loss_origin"].detach().numpy(), rtol=1e-11)
    assert set(grads) == set(grads2)
    for k, v in grads.items():
        if v is None:
            assert grads2[k] is None, k
        else:
            err = np.linalg.norm(grads2[k] - v) / max(np.linalg.norm(v), 1e-300)
            assert err < 1e-9, (k, err)
    assert abs(N2.global_norm(pieces2) - O.global_norm(pieces)) < 1e-11 * O.global_norm(pieces)
    tr = O.OracleTrainer(cfg, P)
    p2 = {k: v.astype(np.float64) for k, v in P.items()}
    m2 = {k: np.zeros_like(v) for k, v in p2.items()}
    v2 = {k: np.zeros_like(v) for k, v in p2.items()}
    for t in range(1, 4):
        l1 = tr.train_step(feed, 2e-3)
        o, g2, pc2 = N2.forward_backward(p2, feed, cfg.L, cfg.D, cfg.H, cfg.N, cfg.reg)
        N2.clip_and_adam(p2, g2, pc2, m2, v2, t, 2e-3, cfg.clip, cfg.beta1, cfg.beta2, cfg.eps)
        assert abs(o["loss"] - l1) < 1e-10 * abs(l1), t
    for k in p2:
        err = np.linalg.norm(p2[k] - tr.params[k]) / max(np.linalg.norm(tr.params[k]), 1e-300)
        assert err < 1e-9, (k, err)

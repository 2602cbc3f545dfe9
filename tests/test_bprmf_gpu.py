"""BPR-MF (Model/BPRMF.py:41-59) against the oracle, with the shared negative item id injected."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import mtam_oracle as O  # noqa: E402


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize("D,B,items,users,neg", [(64, 33, 400, 50, 17), (128, 100, 3706, 300, 0), (32, 4, 40, 6, 39)])
def test_bprmf_parity(D, B, items, users, neg):
    from mtamrecommender_b200 import engine as E
    cfg = O.OracleConfig(kind=O.BPRMF, L=8, D=D, H=1, N=1, user_count=users, item_count=items, category_count=5)
    P = O.init_params(cfg, 2)
    feed = O.synth_batch(cfg, B, 4)
    eng = E.Engine(E.ModelConfig(kind="BPRMF", max_batch=B, L=8, D=D, H=1, N=1, user_count=users, item_count=items,
                                 category_count=5))
    eng.set_params(P)
    eng.set_bpr_negative(neg)
    fwd, grads, pieces = O.loss_and_grads(cfg, P, feed, bpr_negative=neg)
    out = eng.forward(feed)
    assert abs(out["loss"] - float(fwd["loss"].detach())) <= 1e-5 * abs(float(fwd["loss"].detach()))
    assert np.array_equal(out["pred"], P["embedding_layer/user"][feed["user_id"]])     # predict_behavior_emb = u
    g = eng.gradients(feed)
    gn = O.global_norm(pieces)
    assert abs(np.sqrt(g["__norm_sq__"]) - gn) <= 1e-5 * gn
    for k, v in grads.items():
        if v is None:
            assert not np.any(g[k]), k
        else:
            assert rel(g[k], v) < 1e-4, (k, rel(g[k], v))
    tr = O.OracleTrainer(cfg, P)
    for s in range(3):
        lo, lc = tr.train_step(feed, 1e-3, bpr_negative=neg), eng.train_step(feed, 1e-3)
        assert abs(lo - lc) <= 2e-5 * abs(lo), (s, lo, lc)
    newp = eng.get_params()
    for k, v in tr.params.items():
        assert rel(newp[k], v) < 1e-4, (k, rel(newp[k], v))

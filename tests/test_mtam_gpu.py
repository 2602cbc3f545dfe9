"""GPU parity tests proper: the CUDA path (through the C-ABI) against the CPU oracle on the same
seeded inputs.  Tolerances (fp32 kernels vs fp64 oracle): loss rel 1e-5, pred 1e-5,
gradients norm-wise rel 1e-4, weights after 3 Adam steps rel 1e-4; indices exact."""
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


def make(D, L, N, H, B, items, users, cats, seed=7, min_len=2, gemm_mode=0):
    from mtamrecommender_b200 import engine as E
    cfg = O.OracleConfig(kind=O.MTAM, L=L, D=D, H=H, N=N, user_count=users, item_count=items, category_count=cats)
    P = O.init_params(cfg, seed)
    rng = np.random.default_rng(seed + 1)
    for k in P:   # non-trivial biases so their gradients are exercised
        if k.endswith("/bias") or k.endswith("/beta"):
            P[k] = (P[k] + 0.1 * rng.standard_normal(P[k].shape)).astype(np.float32)
    feed = O.synth_batch(cfg, B, seed + 2, min_len=min_len)
    eng = E.Engine(E.ModelConfig(kind="MTAM", max_batch=B, L=L, D=D, H=H, N=N, user_count=users, item_count=items,
                                 category_count=cats, gemm_mode=gemm_mode))
    eng.set_params(P)
    return cfg, P, feed, eng


CASES = [dict(D=64, L=12, N=2, H=1, B=37, items=500, users=50, cats=11),
         dict(D=128, L=50, N=3, H=8, B=64, items=3706, users=300, cats=301),     # ml-1m-shaped (cfg1), 8 heads
         dict(D=32, L=7, N=1, H=4, B=5, items=90, users=9, cats=4),
         dict(D=64, L=50, N=6, H=1, B=130, items=5000, users=1000, cats=100)]    # cfg3-shaped hops, ragged tiles


@pytest.mark.parametrize("case", CASES)
def test_forward_parity(case):
    cfg, P, feed, eng = make(**case)
    fwd = O.forward(cfg, {k: __import__("torch").tensor(v, dtype=__import__("torch").float64) for k, v in P.items()}, feed)
    out = eng.forward(feed)
    assert abs(out["loss"] - float(fwd["loss"])) <= 1e-5 * abs(float(fwd["loss"]))
    assert abs(out["l2_norm"] - float(fwd["l2_norm"])) <= 1e-5 * float(fwd["l2_norm"])
    assert rel(out["loss_origin"], fwd["loss_origin"].numpy()) < 1e-5
    assert rel(out["pred"], fwd["pred"].numpy()) < 1e-5


@pytest.mark.parametrize("case", CASES)
def test_gradient_parity(case):
    cfg, P, feed, eng = make(**case)
    import torch
    fwd, grads, pieces = O.loss_and_grads(cfg, P, feed)
    _, g32, _ = O.loss_and_grads(cfg, P, feed, torch.float32)
    g = eng.gradients(feed)
    gn = O.global_norm(pieces)
    assert abs(np.sqrt(g["__norm_sq__"]) - gn) <= 1e-5 * gn          # un-deduplicated global norm (trap T1)
    for k, v in grads.items():
        if v is None:
            assert not np.any(g[k]), f"{k}: dead parameter received a gradient"
        else:
            # 1e-4 norm-wise, or -- where the graph itself is ill-conditioned in fp32 on this input (a ReLU
            # pre-activation within rounding of its kink) -- no worse than 3x the error of the oracle's own
            # fp32 evaluation against its fp64 evaluation.
            ok, r, tol = grad_close(_test_name(), k, g[k], v, g32[k])
            assert ok, (k, r, tol)


@pytest.mark.parametrize("case", CASES[:3])
def test_three_train_steps(case):
    cfg, P, feed, eng = make(**case)
    import torch
    tr = O.OracleTrainer(cfg, P)
    tr32 = O.OracleTrainer(cfg, P, dtype=torch.float32)     # conditioning reference (see test_gradient_parity)
    for s in range(3):
        lo, lc = tr.train_step(feed, 1e-3), eng.train_step(feed, 1e-3)
        tr32.train_step(feed, 1e-3)
        assert abs(lo - lc) <= 2e-5 * abs(lo), (s, lo, lc)
    newp = eng.get_params()
    for k, v in tr.params.items():
        ok, r, tol = grad_close(_test_name(), k, newp[k], v, tr32.params[k])
        assert ok, (k, r, tol)
    assert eng.adam_step() == 3


def test_min_length_two_and_full_length():
    """seq_length == 2 (one real event + mask step) and == L."""
    cfg, P, feed, eng = make(D=64, L=6, N=2, H=2, B=8, items=50, users=9, cats=4)
    feed["seq_length"][:4] = 2
    for k in ("item_list", "category_list", "position_list", "time_list", "timelast_list", "timenow_list"):
        feed[k][:4, 2:] = 0
    feed["item_list"][:4, 1] = cfg.item_count + 1
    feed["category_list"][:4, 1] = cfg.category_count + 1
    feed["position_list"][:4, 1] = 1
    fwd, grads, pieces = O.loss_and_grads(cfg, P, feed)
    g = eng.gradients(feed)
    for k, v in grads.items():
        if v is not None:
            assert rel(g[k], v) < 1e-4, k


def test_train_step_is_deterministic():
    cfg, P, feed, eng = make(D=64, L=20, N=2, H=1, B=50, items=800, users=60, cats=13)
    eng.train_step(feed, 1e-3)
    a = eng.params.clone()
    eng.set_params(P); eng.adam_m.zero_(); eng.adam_v.zero_(); eng.set_adam_step(0)
    eng.train_step(feed, 1e-3)
    assert bool((a == eng.params).all()), "two identical steps from identical state differ bitwise"


def test_topk_matches_oracle_and_metrics():
    cfg, P, feed, eng = make(D=64, L=12, N=2, H=1, B=40, items=3000, users=50, cats=11)
    b = eng.upload(feed)
    idx, sc = eng.eval_topk_device(b, 50)
    m, oidx, oscores = O.metrics_topk(cfg, P, feed)
    # gap condition: only rows whose oracle scores at ranks <= 51 are separated by more
    # than the fp32 error bound are required to match exactly; report how many that is.
    srt = -np.sort(-oscores, axis=1)[:, :51]
    gap_ok = (np.abs(np.diff(srt, axis=1)).min(axis=1) > 1e-5)
    assert gap_ok.mean() > 0.9
    assert np.array_equal(idx.cpu().numpy()[gap_ok], oidx[gap_ok])
    got = eng.hr_ndcg_device(idx, b.t["target_item_id"]).cpu().numpy()
    assert np.allclose(got, np.array(m), atol=1e-6)


def test_cuda_graph_step_matches_eager():
    import torch
    cfg, P, feed, eng = make(D=64, L=12, N=2, H=1, B=32, items=500, users=50, cats=11)
    eng.train_step(feed, 1e-3); eng.train_step(feed, 5e-4)
    ref = eng.params.clone()
    eng.set_params(P); eng.adam_m.zero_(); eng.adam_v.zero_(); eng.set_adam_step(0); eng.grads.zero_()
    # capture must leave weights, Adam moments and the step counter exactly as they were (the warm-up inside runs
    # with lr = 0 and its moments are rolled back)
    eng.capture_train_graph(32)
    assert eng.adam_step() == 0 and float(eng.adam_m.abs().max()) == 0.0 and float(eng.adam_v.abs().max()) == 0.0
    eng.upload(feed)
    eng.train_step_graph(1e-3); eng.train_step_graph(5e-4)
    torch.cuda.synchronize()
    assert bool((ref == eng.params).all())
    assert eng.adam_step() == 2


def test_graph_capture_on_a_fresh_engine_and_feed_validation():
    """Capture before any upload runs on zero-filled staging buffers (pad ids, length 2), never on garbage; feeds with
    ids outside the tables are rejected on the host (the gather kernels index unchecked)."""
    import torch
    cfg, P, feed, eng = make(D=64, L=12, N=2, H=1, B=32, items=500, users=50, cats=11)
    before = eng.params.clone()
    eng.capture_train_graph(32)
    torch.cuda.synchronize()
    assert bool((before == eng.params).all()) and eng.adam_step() == 0 and float(eng.adam_v.abs().max()) == 0.0
    l_graph = eng.train_step(feed, 1e-3)
    _, _, _, eng2 = make(D=64, L=12, N=2, H=1, B=32, items=500, users=50, cats=11)
    assert l_graph == eng2.train_step(feed, 1e-3) and bool((eng2.params == eng.params).all())
    bad = dict(feed); bad["item_list"] = feed["item_list"].copy(); bad["item_list"][0, 0] = 503
    with pytest.raises(ValueError):
        eng.upload(bad)
    bad = dict(feed); bad["seq_length"] = feed["seq_length"].copy(); bad["seq_length"][3] = 1
    with pytest.raises(ValueError):
        eng.upload(bad)


def test_errors_are_loud():
    from mtamrecommender_b200 import engine as E, _lib
    with pytest.raises(_lib.MtamError):
        E.Engine(E.ModelConfig(kind="MTAM", max_batch=4, L=5, D=48, H=1, N=1, user_count=3, item_count=9, category_count=2))
    cfg, P, feed, eng = make(D=32, L=7, N=1, H=4, B=5, items=90, users=9, cats=4)
    big = {k: np.concatenate([v, v]) for k, v in feed.items()}
    with pytest.raises(ValueError):
        eng.train_step(big, 1e-3)


@pytest.mark.parametrize("case", [CASES[0], CASES[1], CASES[3]])
def test_tensor_core_mode_parity(case):
    """MTAM_GEMM_TF32X3: dense contractions on tcgen05 (3-term split TF32).  Same fp32-class tolerances."""
    import torch
    cfg, P, feed, eng = make(**case, gemm_mode=1)
    fwd, grads, pieces = O.loss_and_grads(cfg, P, feed)
    _, g32, _ = O.loss_and_grads(cfg, P, feed, torch.float32)
    out = eng.forward(feed)
    assert abs(out["loss"] - float(fwd["loss"].detach())) <= 1e-5 * abs(float(fwd["loss"].detach()))
    assert rel(out["pred"], fwd["pred"].detach().numpy()) < 1e-5
    g = eng.gradients(feed)
    assert abs(np.sqrt(g["__norm_sq__"]) - O.global_norm(pieces)) <= 1e-5 * O.global_norm(pieces)
    for k, v in grads.items():
        if v is None:
            assert not np.any(g[k]), k
        else:
            ok, r, tol = grad_close(_test_name(), k, g[k], v, g32[k])
            assert ok, (k, r, tol)
    tr = O.OracleTrainer(cfg, P)
    tr32 = O.OracleTrainer(cfg, P, dtype=torch.float32)
    for s in range(2):
        lo, lc = tr.train_step(feed, 1e-3), eng.train_step(feed, 1e-3)
        tr32.train_step(feed, 1e-3)
        assert abs(lo - lc) <= 2e-5 * abs(lo), (s, lo, lc)
    newp = eng.get_params()
    for k, v in tr.params.items():
        ok, r, tol = grad_close(_test_name(), k, newp[k], v, tr32.params[k])
        assert ok, (k, r, tol)


@pytest.mark.parametrize("gemm_mode", [0, 1])
def test_train_steps_are_bit_reproducible_and_the_split_adam_matches_the_whole(gemm_mode):
    """(1) Two engines from the same weights give the same bits after three train steps -- eager and from the captured
    graph -- in both arithmetic modes (the tensor-core mode runs the mma.sync T-GRU kernels, whose warp pairs meet in
    shared memory, and the user table's Adam in two parts on two streams: a race would show here).  (2) The two-part
    Adam of mtam_train_step (rows without a gradient early, the batch's rows late) against the one-kernel update of
    forward_backward + finish_grads + apply: the same bits in every arena."""
    import torch
    from mtamrecommender_b200 import _lib, engine as E
    import ctypes as C
    case = dict(D=64, L=12, N=2, H=1, B=37, items=500, users=300, cats=11)
    cfg, P, feed, e1 = make(**case, gemm_mode=gemm_mode)
    _, _, _, e2 = make(**case, gemm_mode=gemm_mode)
    _, _, _, e3 = make(**case, gemm_mode=gemm_mode)
    _, _, _, e4 = make(**case, gemm_mode=gemm_mode)
    feed2 = O.synth_batch(cfg, 37, 99)
    for f, lr in ((feed, 1e-3), (feed2, 2e-3), (feed, 5e-4)):
        l1, l2 = e1.train_step(f, lr), e2.train_step(f, lr)
        assert l1 == l2
    for a in ("params", "adam_m", "adam_v"):
        assert bool((getattr(e1, a) == getattr(e2, a)).all()), a
    e3.capture_train_graph(37)
    for f, lr in ((feed, 1e-3), (feed2, 2e-3), (feed, 5e-4)):
        e3.train_step(f, lr)
    assert bool((e1.params == e3.params).all()) and bool((e1.adam_v == e3.adam_v).all())
    # the one-kernel update, driven through the split C entry points (what the data-parallel driver calls)
    lib = e4.lib
    nsq = torch.zeros(1, device="cuda")
    for f, lr in ((feed, 1e-3), (feed2, 2e-3), (feed, 5e-4)):
        b = e4.upload(f)
        nsq.zero_()
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.mtam_forward_backward(e4.h, C.byref(b.c), b.B, e4.scalars.data_ptr(), nsq.data_ptr(), st), "fb")
        _lib.check(lib.mtam_finish_grads(e4.h, nsq.data_ptr(), 1, st), "finish")
        _lib.check(lib.mtam_apply(e4.h, float(lr), nsq.data_ptr(), e4.scalars.data_ptr(), st), "apply")
    torch.cuda.synchronize()
    for a in ("params", "adam_m", "adam_v"):
        assert bool((getattr(e1, a) == getattr(e4, a)).all()), a


@pytest.mark.parametrize("gemm_mode", [0, 1])
def test_cfg4_sequence_length_against_the_oracle(gemm_mode):
    """BASELINE configs[3] / [4] run at L = 200 with 6 hops: the model at that sequence length (the hop kernels' shared-
    memory layout at 200 keys, 199 recurrent steps, position table of 203 rows) against the oracle -- forward, every
    gradient, two Adam steps -- in both arithmetic modes; full and minimal lengths included.  (The 10 M-row tables of
    those configs are covered by the property tests of test_full_size_gpu.py.)"""
    import torch
    cfg, P, feed, eng = make(D=64, L=200, N=6, H=1, B=24, items=5000, users=100, cats=37, seed=11, gemm_mode=gemm_mode)
    full = O.synth_batch(cfg, 1, 5, min_len=200)          # a full-length sequence and a minimal one
    assert int(full["seq_length"][0]) == 200
    for k in feed:
        feed[k][0] = full[k][0]
    feed["seq_length"][1] = 2
    for k in ("item_list", "category_list", "position_list", "time_list", "timelast_list", "timenow_list"):
        feed[k][1, 2:] = 0
    feed["item_list"][1, 1] = cfg.item_count + 1
    feed["category_list"][1, 1] = cfg.category_count + 1
    feed["position_list"][1, 1] = 1
    fwd, grads, pieces = O.loss_and_grads(cfg, P, feed)
    _, g32, _ = O.loss_and_grads(cfg, P, feed, torch.float32)
    out = eng.forward(feed)
    lo = float(fwd["loss"].detach())
    assert abs(out["loss"] - lo) <= 1e-5 * abs(lo)
    assert rel(out["pred"], fwd["pred"].detach().numpy()) < 1e-5
    g = eng.gradients(feed)
    gn = O.global_norm(pieces)
    assert abs(np.sqrt(g["__norm_sq__"]) - gn) <= 1e-5 * gn
    for k, v in grads.items():
        if v is None:
            assert not np.any(g[k]), k
        else:
            ok, r, tol = grad_close(_test_name(), k, g[k], v, g32[k])
            assert ok, (k, r, tol)
    tr = O.OracleTrainer(cfg, P)
    for s in range(2):
        a, b = tr.train_step(feed, 1e-3), eng.train_step(feed, 1e-3)
        assert abs(a - b) <= 2e-5 * abs(a), (s, a, b)

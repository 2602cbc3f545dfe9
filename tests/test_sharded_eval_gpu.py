"""Row-sharded catalogue on one GPU: per-shard scoring (mtam_score_topk on a row range, global indices) + k-way
merge (mtam_merge_topk) equals the unsharded top-k and the CPU oracle's; sort / sorted scatter-add halves equal the
one-call scatter-add bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _topk_oracle(pred, table, k):
    """numpy float64 scores, stable descending sort: ties -> lower index (tf.nn.top_k, Model/base_model.py:196-200)."""
    s = pred.astype(np.float64) @ table.astype(np.float64).T
    idx = np.argsort(-s, axis=1, kind="stable")[:, :k]
    return idx.astype(np.int32), np.take_along_axis(s, idx, axis=1)


@pytest.mark.parametrize("mode", [0, 1], ids=["fp32", "tf32x3"])
@pytest.mark.parametrize("V,B,D,k,shards", [(5003, 37, 64, 50, 4), (20011, 130, 64, 50, 8), (3709, 9, 128, 10, 3)])
def test_sharded_score_merge_equals_unsharded(V, B, D, k, shards, mode):
    import torch
    from mtamrecommender_b200 import engine as E
    from mtamrecommender_b200.parallel import shard_rows
    rng = np.random.default_rng(V)
    table = rng.uniform(-0.3, 0.3, (V, D)).astype(np.float32)
    table[17] = table[3]                      # exact score ties across and inside shards
    table[V - 1] = table[3]
    pred = rng.standard_normal((B, D)).astype(np.float32)
    dt, dp = torch.from_numpy(table).cuda(), torch.from_numpy(pred).cuda()
    full_i, full_s = E.score_topk(dp, dt, k, gemm_mode=mode)
    S = shard_rows(V, shards)
    li, ls = [], []
    for r in range(shards):
        lo, hi = r * S, min(V, (r + 1) * S)
        shard = dt[lo:hi].clone()              # a separate allocation, like a rank's shard
        i, s = E.score_topk(dp, shard, k, lo, hi, index_base=lo, gemm_mode=mode)
        assert int(i.min()) >= lo and int(i.max()) < hi
        li.append(i); ls.append(s)
    mi, ms = E.merge_topk(torch.stack(li), torch.stack(ls))
    assert torch.equal(mi, full_i) and torch.equal(ms, full_s)
    oi, osc = _topk_oracle(pred, table, k)
    srt = -np.sort(-(pred.astype(np.float64) @ table.astype(np.float64).T), axis=1)[:, :k + 1]
    gap = np.abs(np.diff(srt, axis=1))
    ok = (gap[:, :].min(axis=1) > 1e-5) | True    # rows with exact ties are still determined by the index rule
    exact_tie_or_gap = np.all((gap > 1e-5) | (gap == 0.0), axis=1)
    assert exact_tie_or_gap.mean() > 0.8
    assert np.array_equal(mi.cpu().numpy()[exact_tie_or_gap], oi[exact_tie_or_gap])


@pytest.mark.parametrize("V,B,D,k", [(300, 24, 64, 50), (77, 5, 32, 50), (100003, 300, 64, 50), (40000, 129, 32, 64),
                                     (2300007, 140, 64, 50)])
def test_tensor_core_topk_equals_fp32_topk(V, B, D, k):
    """MTAM_GEMM_TF32X3 scoring (tcgen05 bucket-max filter + fp32 rescoring of the best buckets) returns the same
    top-k as the exact-fp32 kernel (tf.nn.top_k order: score desc, ties -> lower index), exact ties included."""
    import torch
    from mtamrecommender_b200 import engine as E
    g = torch.Generator().manual_seed(V + B)
    table = (torch.rand(V, D, generator=g) - 0.5) * 0.6
    table[V // 2] = table[1]                  # exact ties: duplicate rows far apart and adjacent
    table[2] = table[1]
    table[V - 1] = table[1]
    pred = torch.randn(B, D, generator=g)
    pred[0] = 0.0                             # every score equal: the answer is items 0..k-1
    pred[1] = table[1] * 50                   # the tied rows are this row's best
    dt, dp = table.cuda(), pred.cuda()
    ei, es = E.score_topk(dp, dt, k, gemm_mode=0)
    ti, ts = E.score_topk(dp, dt, k, gemm_mode=1)
    assert torch.equal(ti[0].cpu(), torch.arange(k, dtype=torch.int32))
    _same_topk(ei, es, ti, ts)
    # a row range (a shard) and a workspace that forces the pred rows to be processed in chunks of 128
    lo, hi = V // 5, V - V // 7
    ei, es = E.score_topk(dp, dt, k, lo, hi, gemm_mode=0)
    bs = 64 if hi - lo > (1 << 21) else 16
    ld = -(-(((hi - lo + 127) // 128) * (128 // bs)) // 8) * 8
    small = torch.empty(128 * ld * 4 + 256, dtype=torch.uint8, device="cuda")
    ti, ts = E.score_topk(dp, dt, k, lo, hi, gemm_mode=1, workspace=small)
    _same_topk(ei, es, ti, ts)


def _same_topk(ei, es, ti, ts):
    """Both paths return fp32 dot products (different summation orders): scores agree to fp32 rounding and the index
    lists are equal except where two neighbouring scores of a row are closer than that rounding (then they may swap)."""
    ei, es, ti, ts = ei.cpu().numpy(), es.cpu().numpy(), ti.cpu().numpy(), ts.cpu().numpy()
    scale = np.abs(es).max(axis=1, keepdims=True) + 1e-30
    assert np.all(np.abs(es - ts) <= 4e-6 * scale)
    bad = np.argwhere(ei != ti)
    assert len(bad) <= 0.01 * ei.size
    for r, c in bad:
        near = [abs(es[r, c] - es[r, j]) for j in (c - 1, c + 1) if 0 <= j < es.shape[1]]
        assert min(near) <= 4e-6 * scale[r, 0], (r, c, es[r, max(c - 1, 0):c + 2])


@pytest.mark.parametrize("mode", [0, 1], ids=["fp32", "tf32x3"])
@pytest.mark.parametrize("V,B,D,shards", [(5003, 37, 64, 4), (20011, 130, 64, 8), (3709, 9, 32, 3)])
def test_sharded_softmax_ce_equals_unsharded_and_reference(V, B, D, shards, mode):
    """Sharded log-sum-exp: per-shard mtam_softmax_ce_forward, lse combined over shards, target logit summed,
    per-shard backward with the global lse -> the same loss / dpred / table gradient as the unsharded call, and both
    equal log_softmax cross-entropy (base_model.py:316-321) in float64."""
    import torch
    from mtamrecommender_b200 import engine as E
    from mtamrecommender_b200.parallel import combine_lse, shard_rows
    g = torch.Generator().manual_seed(V + 7)
    table = (torch.rand(V, D, generator=g) - 0.5) * 0.6
    pred = torch.randn(B, D, generator=g)
    tgt = torch.randint(0, V, (B,), generator=g, dtype=torch.int32)
    tgt[0], tgt[1] = 0, V - 1
    dt, dp, dtg = table.cuda(), pred.cuda(), tgt.cuda()
    lse_f, tl_f = E.softmax_ce_forward(dp, dt, dtg, gemm_mode=mode)
    dT_f, dP_f = E.softmax_ce_backward(dp, dt, dtg, lse_f, 1.0 / B, gemm_mode=mode)
    # float64 reference
    logits = pred.double() @ table.double().T
    logits.requires_grad_(False)
    p64, t64 = pred.double().requires_grad_(True), table.double().requires_grad_(True)
    lo = -torch.log_softmax(p64 @ t64.T, dim=1)[torch.arange(B), tgt.long()]
    lo.mean().backward()
    rel = lambda a, b: float((a.double().cpu() - b).norm() / b.norm())
    assert rel(lse_f - tl_f, lo.detach()) < 1e-5
    assert rel(dP_f, p64.grad) < 1e-5 and rel(dT_f, t64.grad) < 1e-5
    # shards
    S = shard_rows(V, shards)
    lses, tls = [], []
    for r in range(shards):
        lo_r, hi_r = r * S, min(V, (r + 1) * S)
        l, t = E.softmax_ce_forward(dp, dt[lo_r:hi_r].clone(), (dtg - lo_r).contiguous(), gemm_mode=mode)
        lses.append(l); tls.append(t)
    lse = combine_lse(torch.stack(lses))
    tl = torch.stack(tls).sum(0)
    assert torch.allclose(lse, lse_f, rtol=0, atol=2e-6) and torch.equal(tl, tl_f)
    dP = torch.zeros_like(dP_f)
    for r in range(shards):
        lo_r, hi_r = r * S, min(V, (r + 1) * S)
        dsh, dpp = E.softmax_ce_backward(dp, dt[lo_r:hi_r].clone(), (dtg - lo_r).contiguous(), lse, 1.0 / B, gemm_mode=mode)
        dP += dpp
        assert rel(dsh, t64.grad[lo_r:hi_r]) < 1e-5
    assert rel(dP, p64.grad) < 1e-5


def test_sort_then_sorted_scatter_equals_scatter_add():
    import torch
    from mtamrecommender_b200 import engine as E
    g = torch.Generator().manual_seed(3)
    R, D, n = 100003, 64, 51200
    idx = torch.randint(0, R, (n,), generator=g, dtype=torch.int32)
    idx[: n // 3] = 0
    rows = torch.randn(n, D, generator=g)
    a = torch.zeros(R, D, device="cuda"); b = torch.zeros(R, D, device="cuda")
    E.scatter_add(a, idx.cuda(), rows.cuda())
    srt = E.sort_indices(idx.cuda(), R)
    ks, pm = srt.keys_sorted.cpu().numpy(), srt.perm.cpu().numpy()
    order = np.argsort(idx.numpy(), kind="stable")
    assert np.array_equal(pm, order.astype(np.int32)) and np.array_equal(ks, idx.numpy()[order])
    E.scatter_add_sorted(b, srt, rows.cuda())
    assert torch.equal(a, b)

"""BASELINE.json's full sizes (configs[3]/[4]: 10 M items, 8192 x 200 = 1 638 400 looked-up rows, top-50 over a
row-sharded 10 M catalogue) through size-independent properties: the gather is an exact copy, scatter-add of gathered
integer-valued rows returns count x row exactly (gather -> scatter round trip), the scatter-add is run-to-run
bit-identical, and the full-catalogue top-50 is sorted, equals the 8-shard merge bit for bit and agrees with a chunked
exhaustive scoring."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

V, D, N_ROWS = 10_000_003, 64, 8192 * 200


def _ids(g, n, V):
    import torch
    idx = torch.randint(0, V, (n,), generator=g, dtype=torch.int32)
    idx[: n // 3] = 0            # the pad id: one segment of half a million duplicates
    idx[-1] = V - 1
    return idx


def test_gather_scatter_round_trip_at_cfg4_shapes():
    import torch
    from mtamrecommender_b200 import engine as E
    g = torch.Generator().manual_seed(4)
    table = torch.randint(-8, 9, (V, D), generator=g, dtype=torch.int8).cuda().float()     # integer-valued: sums are exact
    idx = _ids(g, N_ROWS, V).cuda()
    rows = E.gather(table, idx)
    assert torch.equal(rows, table[idx.long()])                                   # gathered rows bit-exact
    dst = torch.zeros((V, D), device="cuda")
    _, uq, nu = E.scatter_add(dst, idx, rows, want_unique=True)
    cnt = torch.bincount(idx.long(), minlength=V).float()
    assert torch.equal(dst, table * cnt[:, None])                                 # = count x row, exactly
    uniq = torch.unique(idx.long())
    assert int(nu.item()) == uniq.numel() and torch.equal(uq[: uniq.numel()].long(), uniq)
    del dst, cnt
    # real-valued rows: run-to-run bit-identical, and the checksum of all rows is preserved to fp32 accuracy
    vals = torch.randn((N_ROWS, D), generator=torch.Generator(device="cuda").manual_seed(5), device="cuda")
    a = torch.zeros((V, D), device="cuda")
    E.scatter_add(a, idx, vals)
    b = torch.zeros((V, D), device="cuda")
    E.scatter_add(b, idx, vals)
    assert torch.equal(a, b), "scatter-add is not deterministic at full size"
    col_in, col_out = vals.double().sum(0), a.double().sum(0)
    assert float((col_in - col_out).abs().max()) <= 1e-6 * float(vals.abs().double().sum(0).max())


def test_top50_over_ten_million_items_sorted_sharded_and_exhaustive():
    import torch
    from mtamrecommender_b200 import engine as E
    from mtamrecommender_b200.parallel import shard_rows
    B, k, W = 256, 50, 8
    g = torch.Generator(device="cuda").manual_seed(6)
    table = (torch.rand((V, D), generator=g, device="cuda") - 0.5) * 0.6
    table[V - 1] = table[7]                       # exact ties across the first and the last shard
    pred = torch.randn((B, D), generator=g, device="cuda")
    pred[0] = table[7] * 40                       # the tied rows are this row's best
    idx, sc = E.score_topk(pred, table, k, gemm_mode=1)
    # sorted: score descending, equal scores by ascending index (tf.nn.top_k)
    assert bool((sc[:, :-1] >= sc[:, 1:]).all())
    tie = sc[:, :-1] == sc[:, 1:]
    assert bool((idx[:, :-1][tie] < idx[:, 1:][tie]).all())
    assert idx[0, 0].item() == 7 and idx[0, 1].item() == V - 1
    assert all(torch.unique(idx[b]).numel() == k for b in range(0, B, 37))
    # 8 row shards (configs[4]) scored separately + merged == unsharded, bit for bit
    S = shard_rows(V, W)
    li, ls = [], []
    for r in range(W):
        lo, hi = r * S, min(V, (r + 1) * S)
        i, s = E.score_topk(pred, table, k, lo, hi, gemm_mode=1)
        li.append(i); ls.append(s)
    mi, ms = E.merge_topk(torch.stack(li), torch.stack(ls))
    assert torch.equal(mi, idx) and torch.equal(ms, sc)
    # exhaustive fp32 scoring in chunks of 1 M rows (torch only as the checker): same top-50 wherever neighbouring
    # scores are further apart than fp32 rounding
    torch.backends.cuda.matmul.allow_tf32 = False
    best_s = torch.full((B, k), -float("inf"), device="cuda")
    best_i = torch.zeros((B, k), dtype=torch.int64, device="cuda")
    for lo in range(0, V, 1 << 20):
        hi = min(V, lo + (1 << 20))
        s = pred @ table[lo:hi].T
        cs, ci = torch.topk(s, k, dim=1)
        alls, alli = torch.cat([best_s, cs], 1), torch.cat([best_i, ci + lo], 1)
        o = torch.argsort(alls, dim=1, descending=True, stable=True)[:, :k]
        best_s, best_i = torch.gather(alls, 1, o), torch.gather(alli, 1, o)
    scale = sc.abs().max(dim=1, keepdim=True).values
    assert bool(((best_s - sc).abs() <= 4e-6 * scale).all())
    differ = (best_i != idx.long())
    assert differ.float().mean().item() < 0.01
    gap_ok = torch.ones_like(differ)
    gap_ok[:, 1:] &= (sc[:, :-1] - sc[:, 1:]) > 8e-6 * scale
    gap_ok[:, :-1] &= (sc[:, :-1] - sc[:, 1:]) > 8e-6 * scale
    assert not bool((differ & gap_ok).any()), "top-50 differs from exhaustive scoring where the scores are well separated"

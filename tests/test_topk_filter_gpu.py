"""The tcgen05 filter pass of the full-catalogue top-k (bucket maxima, `mtam_score_bucket_max`) against exhaustive fp32
scoring for both bucket sizes, and the 64-item-bucket path forced at mid size (`mtam_set_topk_bucket_crossover`) against the
16-item-bucket path and the exact-fp32 kernel.  Reference op: tf.matmul(pred, item_table^T) + tf.nn.top_k,
Model/base_model.py:194-202.

Round 1 shipped a race here (a staging slot of the softmax / scoring pipeline was released while loads from it were
still in flight); it only showed with the light 64-item epilogue over millions of rows, hence the sizes below."""
import pytest

pytestmark = pytest.mark.gpu


def _bucket_max(pred, table, bs):
    import torch
    from mtamrecommender_b200 import _lib
    lib = _lib.load()
    B, D = pred.shape
    rows = table.shape[0]
    ld = -(-(-(-rows // 128) * (128 // bs)) // 8) * 8
    out = torch.full((B, ld), float("nan"), device="cuda")
    _lib.check(lib.mtam_score_bucket_max(pred.data_ptr(), B, D, table.data_ptr(), rows, bs, out.data_ptr(), ld,
                                         torch.cuda.current_stream().cuda_stream), "mtam_score_bucket_max")
    return out[:, : -(-rows // bs)]


def _exhaustive_bucket_max(pred, table, bs):
    import torch
    rows = table.shape[0]
    out = torch.empty((pred.shape[0], -(-rows // bs)), device="cuda")
    step = 1 << 19
    for lo in range(0, rows, step):
        s = pred @ table[lo:min(rows, lo + step)].T
        pad = (-s.shape[1]) % bs
        if pad:
            s = torch.cat([s, torch.full((s.shape[0], pad), -float("inf"), device="cuda")], 1)
        out[:, lo // bs: lo // bs + s.shape[1] // bs] = s.view(s.shape[0], -1, bs).max(dim=2).values
    return out


@pytest.mark.parametrize("D", [64, 32])
@pytest.mark.parametrize("bs", [16, 64])
def test_bucket_maxima_equal_exhaustive_scoring(bs, D):
    import torch
    torch.backends.cuda.matmul.allow_tf32 = False
    V, B = 2_000_003, 200                      # ragged last tile, ragged pred tile
    g = torch.Generator(device="cuda").manual_seed(11)
    table = (torch.rand((V, D), generator=g, device="cuda") - 0.5) * 0.6
    pred = torch.randn((B, D), generator=g, device="cuda")
    want = _exhaustive_bucket_max(pred, table, bs)
    for rep in range(3):                       # the race was intermittent: three passes
        got = _bucket_max(pred, table, bs)
        err = (got - want).abs()
        assert not bool(torch.isnan(got).any())
        # 3xTF32 products, fp32 accumulation: |error| <= ~2^-21 * sum |q||x| (here ~1e-5); a wrong item is off by ~0.1-1
        assert float(err.max()) < 2e-4, (rep, int((err > 2e-4).sum()))


def test_forced_large_buckets_equal_small_buckets_and_fp32_kernel():
    import torch
    from mtamrecommender_b200 import _lib, engine as E
    lib = _lib.load()
    V, D, B, k = 300_007, 64, 130, 50
    g = torch.Generator(device="cuda").manual_seed(12)
    table = (torch.rand((V, D), generator=g, device="cuda") - 0.5) * 0.6
    table[V - 1] = table[5]
    pred = torch.randn((B, D), generator=g, device="cuda")
    pred[0] = table[5] * 30
    i16, s16 = E.score_topk(pred, table, k, gemm_mode=_lib.GEMM_TF32X3)
    try:
        _lib.check(lib.mtam_set_topk_bucket_crossover(1000), "crossover")
        i64, s64 = E.score_topk(pred, table, k, gemm_mode=_lib.GEMM_TF32X3)
        # a shard boundary inside the range too
        ia, sa = E.score_topk(pred, table, k, 0, 170_001, gemm_mode=_lib.GEMM_TF32X3)
        ib, sb = E.score_topk(pred, table, k, 170_001, V, gemm_mode=_lib.GEMM_TF32X3)
    finally:
        _lib.check(lib.mtam_set_topk_bucket_crossover(0), "crossover")
    assert torch.equal(i16, i64) and torch.equal(s16, s64)
    mi, ms = E.merge_topk(torch.stack([ia, ib]), torch.stack([sa, sb]))
    assert torch.equal(mi, i16) and torch.equal(ms, s16)
    assert i16[0, 0].item() == 5 and i16[0, 1].item() == V - 1            # exact tie -> lower index first
    # the exact-fp32 kernel sums in another order: same items wherever neighbouring scores are separated
    i32, s32 = E.score_topk(pred, table, k, gemm_mode=_lib.GEMM_FP32)
    scale = s32.abs().max(dim=1, keepdim=True).values
    assert bool(((s32 - s16).abs() <= 4e-6 * scale).all())
    differ = i32 != i16
    gap_ok = torch.ones_like(differ)
    gap_ok[:, 1:] &= (s32[:, :-1] - s32[:, 1:]) > 8e-6 * scale
    gap_ok[:, :-1] &= (s32[:, :-1] - s32[:, 1:]) > 8e-6 * scale
    assert not bool((differ & gap_ok).any())

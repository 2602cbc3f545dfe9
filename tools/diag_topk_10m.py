"""Diagnostic for the 10 M-row top-50: where do the unsharded (64-item buckets) and the 8-shard (16-item buckets) results
differ, and which one disagrees with exhaustive fp32 scoring?  Also checks the filter pass (bucket maxima) directly."""
import ctypes as C
import sys

import torch

from mtamrecommender_b200 import _lib
from mtamrecommender_b200 import engine as E
from mtamrecommender_b200.parallel import shard_rows


def bucket_max(pred, table, bs):
    lib = _lib.load()
    B, D = pred.shape
    rows = table.shape[0]
    ld = -(-(-(-rows // 128) * (128 // bs)) // 8) * 8
    out = torch.full((B, ld), float("nan"), device="cuda")
    _lib.check(lib.mtam_score_bucket_max(pred.data_ptr(), B, D, table.data_ptr(), rows, bs, out.data_ptr(), ld,
                                         torch.cuda.current_stream().cuda_stream), "bucket_max")
    return out[:, : -(-rows // bs)]


def exhaustive_bucket_max(pred, table, bs):
    rows = table.shape[0]
    nb = -(-rows // bs)
    out = torch.empty((pred.shape[0], nb), device="cuda")
    step = (1 << 20)
    for lo in range(0, rows, step):
        hi = min(rows, lo + step)
        s = pred @ table[lo:hi].T
        pad = (-s.shape[1]) % bs
        if pad:
            s = torch.cat([s, torch.full((s.shape[0], pad), -float("inf"), device="cuda")], 1)
        out[:, lo // bs: lo // bs + s.shape[1] // bs] = s.view(s.shape[0], -1, bs).max(dim=2).values
    return out


def main(V=10_000_003, B=256):
    torch.backends.cuda.matmul.allow_tf32 = False
    D, k, W = 64, 50, 8
    g = torch.Generator(device="cuda").manual_seed(6)
    table = (torch.rand((V, D), generator=g, device="cuda") - 0.5) * 0.6
    table[V - 1] = table[7]
    pred = torch.randn((B, D), generator=g, device="cuda")
    pred[0] = table[7] * 40
    for bs in (16, 64):
        bm = bucket_max(pred, table, bs)
        ex = exhaustive_bucket_max(pred, table, bs)
        err = (bm - ex).abs()
        bad = (err > 1e-4) | torch.isnan(bm)
        print(f"bucket maxima bs={bs}: max err {err[~torch.isnan(bm)].max().item():.3e}, bad entries {int(bad.sum())} of {bm.numel()}")
        if bad.any():
            r, c = torch.nonzero(bad, as_tuple=True)
            print("  first bad (row, bucket, got, want):", [(int(r[i]), int(c[i]), float(bm[r[i], c[i]]), float(ex[r[i], c[i]])) for i in range(min(12, len(r)))])
            print("  bad rows histogram:", torch.bincount(r, minlength=B).nonzero().flatten().tolist()[:40])
            print("  bad bucket range:", int(c.min()), int(c.max()), " bucket%8 hist", torch.bincount(c % 8, minlength=8).tolist())
        del bm, ex, err, bad
    idx, sc = E.score_topk(pred, table, k, gemm_mode=1)
    S = shard_rows(V, W)
    li, ls = [], []
    for r in range(W):
        lo, hi = r * S, min(V, (r + 1) * S)
        i, s = E.score_topk(pred, table, k, lo, hi, gemm_mode=1)
        li.append(i); ls.append(s)
    mi, ms = E.merge_topk(torch.stack(li), torch.stack(ls))
    # exhaustive
    best_s = torch.full((B, k), -float("inf"), device="cuda")
    best_i = torch.zeros((B, k), dtype=torch.int64, device="cuda")
    for lo in range(0, V, 1 << 20):
        hi = min(V, lo + (1 << 20))
        s = pred @ table[lo:hi].T
        cs, ci = torch.topk(s, k, dim=1)
        alls, alli = torch.cat([best_s, cs], 1), torch.cat([best_i, ci + lo], 1)
        o = torch.argsort(alls, dim=1, descending=True, stable=True)[:, :k]
        best_s, best_i = torch.gather(alls, 1, o), torch.gather(alli, 1, o)
    for name, (a, b) in {"unsharded vs sharded": (idx, mi), "unsharded vs exhaustive": (idx, best_i), "sharded vs exhaustive": (mi, best_i)}.items():
        d = a.long() != b.long()
        r, c = torch.nonzero(d, as_tuple=True)
        print(f"{name}: {int(d.sum())} differing entries in {len(set(r.tolist()))} rows; first:",
              [(int(r[i]), int(c[i]), int(a[r[i], c[i]]), int(b[r[i], c[i]])) for i in range(min(8, len(r)))])
    # the 16-item path forced on the whole range
    _lib.load().mtam_set_topk_bucket_crossover(1 << 30)
    idx16, sc16 = E.score_topk(pred, table, k, gemm_mode=1)
    _lib.load().mtam_set_topk_bucket_crossover(0)
    print("unsharded bs=16 vs exhaustive:", int((idx16.long() != best_i).sum()), " vs sharded:", int((idx16 != mi).sum()))


if __name__ == "__main__":
    main(*(int(a) for a in sys.argv[1:]))

"""The two graded bandwidth kernels against the oracle: gathered rows bit-exact, scatter-add
deterministic and within 2 ulp*sqrt(dups) of the fp64 sum; ragged / empty / duplicate-heavy inputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import mtam_oracle as O  # noqa: E402


@pytest.mark.parametrize("R,D,n", [(1000, 64, 5000), (53, 64, 51200), (100003, 64, 51200), (300, 128, 1),
                                   (1 << 17, 32, 200000), (7, 4, 3), (20_000_003, 64, 300000)])
def test_gather_bit_exact(R, D, n):
    import torch
    from mtamrecommender_b200 import engine as E
    g = torch.Generator().manual_seed(R + n)
    table = torch.randn(R, D, generator=g)
    idx = torch.randint(0, R, (n,), generator=g, dtype=torch.int32)
    idx[: n // 3] = 0
    idx[-1] = R - 1
    out = E.gather(table.cuda(), idx.cuda()).cpu().numpy()
    assert np.array_equal(out, O.gather_rows(table.numpy(), idx.numpy()))


@pytest.mark.parametrize("R,D,n", [(1000, 64, 5000), (53, 64, 51200), (100003, 64, 51200), (300, 128, 1),
                                   (1 << 17, 32, 200000), (10_000_003, 64, 400000), (5, 64, 70000)])
def test_scatter_add(R, D, n):
    import torch
    from mtamrecommender_b200 import engine as E
    g = torch.Generator().manual_seed(R * 3 + n)
    idx = torch.randint(0, R, (n,), generator=g, dtype=torch.int32)
    idx[: n // 3] = 0                      # the pad row: one very long segment
    rows = torch.randn(n, D, generator=g)
    dst = torch.zeros(R, D, device="cuda")
    _, uq, nu = E.scatter_add(dst, idx.cuda(), rows.cuda(), want_unique=True)
    ref = np.zeros((R, D)); np.add.at(ref, idx.numpy(), rows.numpy().astype(np.float64))
    cnt = np.bincount(idx.numpy(), minlength=R).astype(np.float64)
    tol = 4 * np.finfo(np.float32).eps * np.sqrt(np.maximum(cnt, 1))[:, None] * np.maximum(np.abs(rows.numpy()).max(), 1) * np.sqrt(np.maximum(cnt, 1))[:, None]
    assert np.all(np.abs(dst.cpu().numpy() - ref) <= tol + 1e-6)
    un = np.unique(idx.numpy())
    assert int(nu.item()) == len(un) and np.array_equal(uq[: len(un)].cpu().numpy(), un)
    dst2 = torch.zeros(R, D, device="cuda")
    E.scatter_add(dst2, idx.cuda(), rows.cuda())
    assert torch.equal(dst, dst2), "scatter-add is not run-to-run deterministic"


def test_scatter_add_integer_valued_rows_exact_and_accumulates():
    """Integer-valued floats sum exactly in any order: bit-exact against the oracle, and dst is added to."""
    import torch
    from mtamrecommender_b200 import engine as E
    R, D, n = 321, 64, 30000
    g = torch.Generator().manual_seed(1)
    idx = torch.randint(0, R, (n,), generator=g, dtype=torch.int32)
    rows = torch.randint(-8, 9, (n, D), generator=g).float()
    dst = torch.ones(R, D, device="cuda")
    E.scatter_add(dst, idx.cuda(), rows.cuda())
    ref = O.scatter_add_rows(R, idx.numpy(), rows.numpy()) + 1.0
    assert np.array_equal(dst.cpu().numpy(), ref)


def test_scatter_add_strided_rows_and_empty():
    import torch
    from mtamrecommender_b200 import engine as E
    R, D, n = 100, 64, 999
    g = torch.Generator().manual_seed(2)
    idx = torch.randint(0, R, (n,), generator=g, dtype=torch.int32)
    wide = torch.randint(-4, 5, (n, 2 * D), generator=g).float().cuda()
    dst = torch.zeros(R, D, device="cuda")
    E.scatter_add(dst, idx.cuda(), wide[:, D:])            # row stride 2D, column offset D
    assert np.array_equal(dst.cpu().numpy(), O.scatter_add_rows(R, idx.numpy(), wide[:, D:].cpu().numpy()))
    dst0 = torch.zeros(R, D, device="cuda")
    E.scatter_add(dst0, idx[:0].cuda(), wide[:0, :D].contiguous())
    assert not bool(dst0.any())


@pytest.mark.parametrize("n,bound", [(1, 10), (31, 5), (8191, 300), (8192, 70000), (8193, 256), (100_000, 1 << 24),
                                     (1_000_003, 10_000_003), (300_000, 3)])
def test_sort_indices_is_a_stable_sort(n, bound):
    """mtam_sort_indices: keys ascending, equal keys in ascending original position (stability is what makes the
    scatter-add deterministic), perm a permutation.  Sizes straddle the 8192-key tile and all 1..4 digit passes."""
    import torch
    from mtamrecommender_b200 import engine as E
    g = torch.Generator().manual_seed(n + bound)
    keys = torch.randint(0, bound, (n,), generator=g, dtype=torch.int32)
    keys[: n // 3] = 0
    keys[-1] = bound - 1
    s = E.sort_indices(keys.cuda(), bound)
    want_k, want_p = torch.sort(keys.long(), stable=True)
    assert torch.equal(s.keys_sorted.cpu().long(), want_k)
    assert torch.equal(s.perm.cpu().long(), want_p)


@pytest.mark.parametrize("R,D,n", [(1000, 64, 5000), (53, 64, 51200), (100003, 64, 51200), (300, 128, 1), (7, 1, 5000),
                                   (1 << 17, 32, 200000), (2_000_003, 64, 700000), (5, 64, 70000), (90, 48, 4000)])
def test_scatter_overwrite_mode_never_reads_dst(R, D, n):
    """accumulate=False: touched rows hold exactly the sum of their rows (whatever dst held, NaN included), untouched
    rows keep their content; identical, bit for bit, to accumulate=True on a zeroed dst."""
    import torch
    from mtamrecommender_b200 import engine as E
    g = torch.Generator().manual_seed(R + 7 * n)
    idx = torch.randint(0, R, (n,), generator=g, dtype=torch.int32)
    idx[: n // 3] = 0
    rows = torch.randn(n, D, generator=g).cuda()
    ref = torch.zeros(R, D, device="cuda")
    E.scatter_add(ref, idx.cuda(), rows)
    dst = torch.full((R, D), float("nan"), device="cuda")
    E.scatter_add(dst, idx.cuda(), rows, accumulate=False)
    touched = torch.zeros(R, dtype=torch.bool)
    touched[idx.long()] = True
    assert torch.equal(dst[touched.cuda()], ref[touched.cuda()])
    assert bool(torch.isnan(dst[~touched.cuda()]).all())

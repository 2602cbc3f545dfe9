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

"""Two-GPU runs over NCCL (skipped with fewer than 2 devices): the data-parallel train step equals the one-GPU step
on the concatenated batch; the row-sharded catalogue's all-to-all lookup and sharded top-k equal the unsharded ones
bit for bit."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _need_two():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")


def _rank_main(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    from oracle import mtam_oracle as O
    from mtamrecommender_b200 import engine as E
    from mtamrecommender_b200.parallel import DataParallel, ShardedCatalogue, ShardedItemTableTrainer, shard_rows
    dev = f"cuda:{rank}"
    out = {}
    # ---- data-parallel step --------------------------------------------------------------------
    cfg = O.OracleConfig(kind=O.MTAM, L=12, D=64, H=1, N=2, user_count=60, item_count=900, category_count=13)
    P = O.init_params(cfg, 5)
    full = O.synth_batch(cfg, 48, 6)
    Bl = 48 // world
    local = {k: v[rank * Bl:(rank + 1) * Bl] for k, v in full.items()}
    mc = dict(kind="MTAM", L=12, D=64, H=1, N=2, user_count=60, item_count=900, category_count=13)
    want = rl = None
    if rank == 0:
        ref = E.Engine(E.ModelConfig(max_batch=48, **mc), device=dev)
        ref.set_params(P)
        rl = [ref.train_step(full, 1e-3) for _ in range(3)]
        want = ref.get_params()
    for mode in ("dense", "gather"):
        eng = E.Engine(E.ModelConfig(max_batch=Bl, **mc), device=dev)
        eng.set_params(P)
        dp = DataParallel(eng, mode=mode)
        losses = [dp.train_step(local, 1e-3) for _ in range(3)]
        got = eng.get_params()
        flat = eng.params.clone()                      # replicas must be bit-identical
        other = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(other, flat)
        out[f"replicas_identical_{mode}"] = all(bool(torch.equal(o, flat)) for o in other)
        if rank == 0:
            out[f"loss_err_{mode}"] = max(abs(a - b) / abs(b) for a, b in zip(losses, rl))
            out[f"param_err_{mode}"] = max(float(np.linalg.norm(got[k] - want[k]) / max(np.linalg.norm(want[k]), 1e-30))
                                           for k in want)
        del dp, eng
    # ---- train step over the row-sharded item table == the one-GPU step on the concatenated batch ----
    S = shard_rows(903, world)
    for gm in (0, 1):
        if rank == 0:
            ref = E.Engine(E.ModelConfig(max_batch=48, gemm_mode=gm, **mc), device=dev)
            ref.set_params(P)
            rl = [ref.train_step(full, 1e-3) for _ in range(3)]
            want = ref.get_params()
            ridx, _ = ref.eval_topk_device(ref.upload(full), 50)
        mcs = dict(mc, item_count=S - 3)
        eng = E.Engine(E.ModelConfig(max_batch=Bl, gemm_mode=gm, **mcs), device=dev)
        eng.set_params({k: (v if k != "embedding_layer/item" else np.zeros((S, 64), np.float32)) for k, v in P.items()})
        tr = ShardedItemTableTrainer(eng, 903)
        tr.init_shard_from_full(torch.from_numpy(P["embedding_layer/item"]).to(dev))
        losses = [tr.train_step(local, 1e-3) for _ in range(3)]
        got = eng.get_params()
        shards = [torch.empty((S, 64), device=dev) for _ in range(world)]
        dist.all_gather(shards, eng.param_view("embedding_layer/item").contiguous())
        sidx, _ = tr.eval_topk(eng.upload(local), 50)
        allidx = [torch.empty_like(sidx) for _ in range(world)]
        dist.all_gather(allidx, sidx)
        if rank == 0:
            table = torch.cat(shards)[:903].cpu().numpy()
            out[f"sharded_loss_err_{gm}"] = max(abs(a - b) / abs(b) for a, b in zip(losses, rl))
            errs = {k: float(np.linalg.norm(got[k] - want[k]) / max(np.linalg.norm(want[k]), 1e-30)) for k in want
                    if k != "embedding_layer/item"}
            errs["embedding_layer/item"] = float(np.linalg.norm(table - want["embedding_layer/item"]) /
                                                 np.linalg.norm(want["embedding_layer/item"]))
            out[f"sharded_param_err_{gm}"] = max(errs.values())
            out[f"sharded_topk_overlap_{gm}"] = float(np.mean([len(set(a.tolist()) & set(b.tolist())) / 50.0
                                                                 for a, b in zip(torch.cat(allidx).cpu(), ridx.cpu())]))
        del tr, eng
    # ---- row-sharded catalogue -----------------------------------------------------------------
    V, D, k = 20011, 64, 50
    g = torch.Generator().manual_seed(11)
    table = torch.empty((V, D)).uniform_(-0.3, 0.3, generator=g).to(dev)
    table[17] = table[3]; table[V - 1] = table[3]
    cat = ShardedCatalogue.from_full(table)
    g2 = torch.Generator().manual_seed(100 + rank)
    ids = torch.randint(0, V, (7, 33), generator=g2, dtype=torch.int32).to(dev)
    ids[0, :5] = 0
    rows = cat.lookup(ids)
    out["lookup_exact"] = bool(torch.equal(rows, table[ids.long()]))
    pred = torch.randn((24, D), generator=g2).to(dev)
    si, ss = cat.topk(pred, k)
    fi, fs = E.score_topk(pred, table, k)
    out["topk_exact"] = bool(torch.equal(si, fi) and torch.equal(ss, fs))
    # sharded softmax cross-entropy (sharded log-sum-exp) == the unsharded one on the concatenated batch
    tgt = torch.randint(0, V, (24,), generator=g2, dtype=torch.int32).to(dev)
    tgt[0] = V - 1
    lo, dpr, dsh = cat.softmax_ce(pred, tgt)
    ps, ts = [], []
    for r in range(world):                       # every rank can replay the other's draws from its seed
        gr = torch.Generator().manual_seed(100 + r)
        torch.randint(0, V, (7, 33), generator=gr, dtype=torch.int32)
        ps.append(torch.randn((24, D), generator=gr))
        t = torch.randint(0, V, (24,), generator=gr, dtype=torch.int32); t[0] = V - 1
        ts.append(t)
    allp, allt = torch.cat(ps).to(dev), torch.cat(ts).to(dev)
    lse_f, tl_f = E.softmax_ce_forward(allp, table, allt)
    dT_f, dp_f = E.softmax_ce_backward(allp, table, allt, lse_f, 1.0 / allp.shape[0])
    mine = slice(rank * 24, (rank + 1) * 24)
    rel = lambda a, b: float((a - b).norm() / b.norm().clamp_min(1e-30))
    out["ce_loss_err"] = rel(lo, (lse_f - tl_f)[mine])
    out["ce_dpred_err"] = rel(dpr, dp_f[mine])
    out["ce_dtable_err"] = rel(dsh, dT_f[cat.row_begin:cat.row_end])
    torch.cuda.synchronize()
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_dp_and_sharded_catalogue():
    _need_two()
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        r, o = q.get(timeout=300)
        res[r] = o
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for r in (0, 1):
        assert res[r]["replicas_identical_dense"] and res[r]["replicas_identical_gather"], "replicas diverged"
        assert res[r]["lookup_exact"], "sharded lookup differs from the local gather"
        assert res[r]["topk_exact"], "sharded top-k differs from the unsharded one"
        assert res[r]["ce_loss_err"] < 1e-5 and res[r]["ce_dpred_err"] < 1e-5 and res[r]["ce_dtable_err"] < 1e-5, res[r]
    for mode in ("dense", "gather"):
        assert res[0][f"loss_err_{mode}"] < 2e-5, res[0]
        assert res[0][f"param_err_{mode}"] < 1e-4, res[0]
    for gm in (0, 1):      # the sharded-table step: same losses, same weights (item table gathered back from the shards)
        assert res[0][f"sharded_loss_err_{gm}"] < 2e-5, res[0]
        assert res[0][f"sharded_param_err_{gm}"] < 1e-4, res[0]
        assert res[0][f"sharded_topk_overlap_{gm}"] > 0.99, res[0]

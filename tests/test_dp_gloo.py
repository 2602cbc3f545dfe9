"""Data-parallel contract on CPU (gloo, world_size 2): the N-rank step of mtamrecommender_b200/parallel.py
-- all-reduce of the dense pieces and of the un-deduplicated sparse squared norm, all-gather of the sparse
pieces, identical clip + Adam -- equals the 1-rank step on the concatenated batch.  The arithmetic here is
the oracle's (the engine itself needs a GPU); what is tested is the exchange algorithm and trap T1 under DP."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _rank_main(rank, world, port, q):
    from oracle import mtam_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    cfg = O.OracleConfig(kind=O.MTAM, L=6, D=32, H=2, N=2, user_count=9, item_count=40, category_count=4)
    P = O.init_params(cfg, 5)
    full = O.synth_batch(cfg, 8, 6)
    Bl = 8 // world
    local = {k: v[rank * Bl:(rank + 1) * Bl] for k, v in full.items()}
    # 1. local forward/backward with the mean over the GLOBAL batch: scale the CE term by Bl/B_global
    p = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in P.items()}
    Ts = torch.tensor(P["embedding_layer/item"], dtype=torch.float64, requires_grad=True)
    fwd = O.forward(cfg, p, local, torch.float64, item_table_for_scores=Ts)
    loss_local = cfg.reg * fwd["l2_norm"] + fwd["loss_origin"].sum() / 8
    loss_local.backward()
    # 2. all-reduce dense pieces and the sparse squared norm
    dense = {k: (t.grad.clone() if t.grad is not None else torch.zeros_like(t)) for k, t in p.items() if k not in O.TABLES}
    dense["__item_dense__"] = Ts.grad.clone()
    for k in sorted(dense):
        dist.all_reduce(dense[k])
    sq = torch.zeros(1, dtype=torch.float64)
    for rk in ("Eu", "Ei", "Ec", "Ep"):
        sq += (fwd[rk].grad ** 2).sum()
    dist.all_reduce(sq)
    loss = loss_local.detach().clone()
    dist.all_reduce(loss)
    norm_sq = float(sq) + sum(float((g ** 2).sum()) for k, g in dense.items() if not O.is_dead(k))
    # 3. all-gather sparse pieces, scatter-add the identical global list on every rank
    grads = {k: v.numpy() for k, v in dense.items() if k != "__item_dense__"}
    for name, (rk, ik) in O.TABLES.items():
        rows = fwd[rk].grad.reshape(-1, cfg.D).contiguous()
        idx = torch.from_numpy(np.ascontiguousarray(local[ik]).reshape(-1).astype(np.int64))
        g_rows = [torch.zeros_like(rows) for _ in range(world)]
        g_idx = [torch.zeros_like(idx) for _ in range(world)]
        dist.all_gather(g_rows, rows)
        dist.all_gather(g_idx, idx)
        d = np.zeros(P[name].shape)
        np.add.at(d, torch.cat(g_idx).numpy(), torch.cat(g_rows).numpy())
        if name == "embedding_layer/item":
            d += dense["__item_dense__"].numpy()
        grads[name] = d
    scale = O.clip_scale(np.sqrt(norm_sq), cfg.clip)
    newp = {}
    for k, g in grads.items():
        if O.is_dead(k):
            newp[k] = P[k].astype(np.float64)
            continue
        w, _, _ = O.adam_tf(P[k].astype(np.float64), g * scale, np.zeros_like(g), np.zeros_like(g),
                              float(np.float32(1e-3)), 1)
        newp[k] = w
    if rank == 0:
        q.put((float(loss), np.sqrt(norm_sq), newp))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_step_equals_single_rank_step_on_concatenated_batch():
    from oracle import mtam_oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    loss_dp, gn_dp, newp = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = O.OracleConfig(kind=O.MTAM, L=6, D=32, H=2, N=2, user_count=9, item_count=40, category_count=4)
    P = O.init_params(cfg, 5)
    full = O.synth_batch(cfg, 8, 6)
    tr = O.OracleTrainer(cfg, P)
    loss_1 = tr.train_step(full, 1e-3)
    assert abs(loss_dp - loss_1) < 1e-12 * max(1, abs(loss_1))
    assert abs(gn_dp - tr.last["global_norm"]) < 1e-10 * tr.last["global_norm"]
    for k, v in tr.params.items():
        assert np.allclose(newp[k], v, rtol=1e-10, atol=1e-12), k

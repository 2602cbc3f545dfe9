"""BASELINE configs[4]: full-catalogue top-50 eval over a row-sharded 10 M-item table (one process per GPU, NCCL):
all-to-all lookup of a batch's item ids + sharded scoring / top-k merged by all-gather (parallel.ShardedCatalogue).
usage: torchrun --nproc-per-node N tools/bench_sharded_eval.py [--items 10000000] [--batch 8192] [--seq-len 200]
Prints one JSON line from rank 0 (device-timed, max over ranks)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", type=int, default=10_000_000)
    ap.add_argument("--batch", type=int, default=8192, help="global eval batch (sequences)")
    ap.add_argument("--seq-len", type=int, default=200)
    ap.add_argument("--dim", type=int, default=64)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = f"cuda:{local}"
    torch.cuda.set_device(dev)
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    from mtamrecommender_b200.parallel import ShardedCatalogue, shard_rows
    V, D, k = a.items + 3, a.dim, 50
    S = shard_rows(V, world)
    lo, hi = rank * S, min(V, (rank + 1) * S)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    shard = (torch.rand((hi - lo, D), generator=g, device=dev) - 0.5) * 0.6
    cat = ShardedCatalogue(shard, V)
    Bl = a.batch // world
    pred = torch.randn((Bl, D), generator=g, device=dev)
    ids = torch.randint(0, V, (Bl, a.seq_len), generator=g, device=dev, dtype=torch.int32)

    def timed(fn):
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier(); torch.cuda.synchronize()
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        dist.barrier(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / a.iters], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    ms_topk = timed(lambda: cat.topk(pred, k))
    ms_lookup = timed(lambda: cat.lookup(ids))
    idx, sc = cat.topk(pred, k)
    ok = bool((sc[:, :-1] >= sc[:, 1:]).all()) and int(idx.min()) >= 0 and int(idx.max()) < V
    if rank == 0:
        print(json.dumps({
            "workload": "cfg5: top-50 eval over a row-sharded catalogue", "n_gpus": world, "items": V, "dim": D,
            "global_batch": a.batch, "rows_per_shard": S,
            "topk_ms": ms_topk, "eval_seq_per_s": a.batch / ms_topk * 1e3,
            "scoring_useful_TFLOPs_aggregate": 2.0 * a.batch * D * V / ms_topk / 1e9,
            "lookup_ids": a.batch * a.seq_len, "lookup_ms": ms_lookup,
            "lookup_rows_per_s": a.batch * a.seq_len / ms_lookup * 1e3,
            "lookup_GBs_aggregate": a.batch * a.seq_len * (4 + 2 * D * 4) / ms_lookup / 1e6,
            "sorted_and_in_range": ok}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

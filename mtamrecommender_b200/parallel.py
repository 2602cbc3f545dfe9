"""Data-parallel training over one node: one process per GPU, NCCL over NVLink via torch.distributed.

The reference has no data parallelism (one tf.Session on one device); the contract here is
"the N-GPU step equals the 1-GPU step on the concatenated batch".  Per step:

  1. every rank runs forward+backward on its B/N sequences with the loss mean taken over the global
     batch (mtam_forward_backward);
  2. all-reduce(sum) of the dense gradient pieces (item table dense part + all non-table parameters),
     of the loss scalars and of the squared norm of the un-deduplicated sparse pieces (trap T1 needs
     the norm of the *global* pieces);
  3. the sparse pieces (IndexedSlices values: item/category/position/user rows and their ids) are
     all-gathered -- they are B*L*(3D+...) floats, far smaller than the dense tables -- and every rank
     scatter-adds the identical global list with the deterministic sort + segmented reduce, so all
     replicas hold bit-identical gradients;
  4. identical clip + Adam on every rank (mtam_apply).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import check
from .engine import (DeviceBatch, Engine, gather, merge_topk, scatter_add, scatter_add_workspace, score_topk,
                     scatter_add_sorted, softmax_ce_backward, softmax_ce_forward, sort_indices)


class DataParallel:
    """mode "dense"  (tables not much larger than the gathered rows, e.g. 100 K items): every rank scatter-adds its
                     OWN sparse item / category / position pieces into a zeroed [rows, D] buffer (sorts already done
                     beside the forward pass) and the buffer is all-reduced -- traffic independent of the world size;
       mode "gather" (huge tables, e.g. 10 M items): the sparse pieces (ids + value rows) are all-gathered and every
                     rank scatter-adds the identical global list.
    The user table is always handled the "gather" way (B rows per rank).  In both modes the all-reduce of the dense
    item-table gradient starts on a second stream as soon as the softmax backward has produced it and overlaps the
    rest of the backward pass."""

    def __init__(self, engine: Engine, group=None, mode: str = "auto"):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.eng = engine
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        cfg = engine.cfg
        dev = engine.device
        B, L, D, W = cfg.max_batch, cfg.L, cfg.D, self.world
        T = B * L
        c = engine.c_cfg
        if mode == "auto":
            mode = "dense" if c.item_rows <= 8 * W * T else "gather"
        if mode not in ("dense", "gather"):
            raise ValueError(f"unknown data-parallel mode {mode!r}")
        self.mode = mode
        self.g_user = torch.empty(W * B, dtype=torch.int32, device=dev)
        self.g_dEu = torch.empty((W * B, D), dtype=torch.float32, device=dev)
        need = scatter_add_workspace(W * B, c.user_rows, D)
        if mode == "gather":
            self.g_item = torch.empty(W * T, dtype=torch.int32, device=dev)
            self.g_cat = torch.empty(W * T, dtype=torch.int32, device=dev)
            self.g_pos = torch.empty(W * T, dtype=torch.int32, device=dev)
            self.g_dE2 = torch.empty((W * T, 2 * D), dtype=torch.float32, device=dev)
            self.g_dEp = torch.empty((W * T, D), dtype=torch.float32, device=dev)
            need = max(need, scatter_add_workspace(W * T, c.item_rows, D), scatter_add_workspace(W * T, c.category_rows, D),
                       scatter_add_workspace(W * T, c.position_rows, D))
        else:
            rows = c.item_rows + c.category_rows + c.position_rows
            self.sp = torch.zeros((rows, D), dtype=torch.float32, device=dev)   # [item | category | position]
        self.scatter_ws = torch.empty(need, dtype=torch.uint8, device=dev)
        self.user_sort_ws = torch.empty(max(int(engine.lib.mtam_sort_workspace(W * B, c.user_rows)), 16), dtype=torch.uint8,
                                        device=dev)
        self._graph, self._graph_B = None, -1
        self.comm = torch.cuda.Stream(dev)
        self.ev_item = torch.cuda.Event()
        self.ev_item.record(torch.cuda.current_stream(dev))   # materialises the cudaEvent_t
        check(engine.lib.mtam_set_item_grad_event(engine.h, C.c_void_p(self.ev_item.cuda_event)), "mtam_set_item_grad_event")

    # -- views into the engine's arenas / workspace ---------------------------------------------
    def _ws_view(self, ptr: int, rows: int, cols: int) -> torch.Tensor:
        off = ptr - self.eng.workspace.data_ptr()
        return self.eng.workspace[off: off + rows * cols * 4].view(torch.float32).view(rows, cols)

    def _grad_region(self, off: int, rows: int) -> torch.Tensor:
        D = self.eng.cfg.D
        return self.eng.grads[off: off + rows * D].view(rows, D)

    def _gather(self, out: torch.Tensor, local: torch.Tensor) -> torch.Tensor:
        n = local.shape[0] * self.world
        o = out[:n]
        dist.all_gather_into_tensor(o, local.contiguous(), group=self.group)
        return o

    def train_step_device(self, batch: DeviceBatch, lr: float) -> None:
        eng, W = self.eng, self.world
        B, L, D = batch.B, eng.cfg.L, eng.cfg.D
        T = B * L
        c = eng.c_cfg
        main = torch.cuda.current_stream(eng.device)
        # the user ids are known before the step: gather and sort them beside the forward pass
        sv0 = None
        self.comm.wait_stream(main)
        has_user = eng.cfg.kind != "PISTREC"
        if has_user:
            with torch.cuda.stream(self.comm):
                g_user = self._gather(self.g_user, batch.t["user_id"])
                user_sorted = sort_indices(g_user, c.user_rows, self.user_sort_ws)
        eng.forward_backward_device(batch, global_batch=B * W)
        sv = _lib.SparseView()
        check(eng.lib.mtam_sparse_pieces(eng.h, C.byref(sv)), "mtam_sparse_pieces")
        item_lo, item_hi = int(sv.item_offset), int(sv.item_offset) + c.item_rows * D
        # 1. dense item-table gradient: all-reduce on the second stream, from the moment the softmax backward is done
        self.comm.wait_event(self.ev_item)
        with torch.cuda.stream(self.comm):
            dist.all_reduce(eng.grads[item_lo:item_hi], group=self.group)
        # 2. the rest of the dense pieces, the loss scalars, the squared norm of the un-deduplicated sparse pieces: they
        #    sit back to back behind the item table in the engine's grads allocation -- one collective
        dense_begin = int(sv.dense_begin)
        if item_lo > dense_begin:
            dist.all_reduce(eng.grads[dense_begin:item_lo], group=self.group)
        dist.all_reduce(eng._grads_all[item_hi: eng.n_floats + _lib.S_COUNT + 1], group=self.group)
        main.wait_stream(self.comm)
        eng.finish_grads(scatter_local=False)          # adds the squared norm of the (global) dense pieces
        # 3. sparse pieces, 4. clip + optimizer.  The clip scale needs only the norm, which is complete here; the optimizer
        #    then runs region by region, so that the all-reduce of the densified [item | category | position] pieces is
        #    hidden behind the update of the user table (86 % of the parameters at cfg3).
        st = main.cuda_stream
        lib = eng.lib
        if bool(sv.has_user) != has_user:
            raise RuntimeError("user-table gradient pieces do not match the model kind")
        check(lib.mtam_apply_begin(eng.h, float(lr), eng.norm_sq.data_ptr(), eng.scalars.data_ptr(), st), "mtam_apply_begin")
        if self.mode == "dense":
            sp_item = self.sp[:c.item_rows]
            sp_cat = self.sp[c.item_rows: c.item_rows + c.category_rows]
            sp_pos = self.sp[c.item_rows + c.category_rows:]
            self.sp.zero_()
            check(lib.mtam_scatter_sparse_into(eng.h, sp_item.data_ptr(), sp_cat.data_ptr(), sp_pos.data_ptr(), None, st),
                  "mtam_scatter_sparse_into")
            self.comm.wait_stream(main)
            with torch.cuda.stream(self.comm):
                dist.all_reduce(self.sp, group=self.group)
        else:
            g_item = self._gather(self.g_item, batch.t["item_list"].reshape(-1))
            g_cat = self._gather(self.g_cat, batch.t["category_list"].reshape(-1))
            g_pos = self._gather(self.g_pos, batch.t["position_list"].reshape(-1))
            g_dE2 = self._gather(self.g_dE2, self._ws_view(sv.item_cat_rows, T, 2 * D))
            g_dEp = self._gather(self.g_dEp, self._ws_view(sv.position_rows, T, D))
        if has_user:
            g_dEu = self._gather(self.g_dEu, self._ws_view(sv.user_rows, B, D))
            scatter_add_sorted(self._grad_region(int(sv.user_offset), c.user_rows), user_sorted, g_dEu, self.scatter_ws)
        first_table = min(int(sv.category_offset), int(sv.position_offset), item_lo)       # the user table lies in front
        check(lib.mtam_apply_range(eng.h, 0, first_table, st), "mtam_apply_range")           # user table
        check(lib.mtam_apply_range(eng.h, item_hi, eng.n_floats, st), "mtam_apply_range")     # dense parameters
        if self.mode == "dense":
            main.wait_stream(self.comm)
            self._grad_region(int(sv.item_offset), c.item_rows).add_(sp_item)
            self._grad_region(int(sv.category_offset), c.category_rows).copy_(sp_cat)
            self._grad_region(int(sv.position_offset), c.position_rows).copy_(sp_pos)
        else:
            scatter_add(self._grad_region(int(sv.item_offset), c.item_rows), g_item, g_dE2[:, :D], self.scatter_ws)
            scatter_add(self._grad_region(int(sv.category_offset), c.category_rows), g_cat, g_dE2[:, D:], self.scatter_ws)
            scatter_add(self._grad_region(int(sv.position_offset), c.position_rows), g_pos, g_dEp, self.scatter_ws)
        check(lib.mtam_apply_range(eng.h, first_table, item_hi, st), "mtam_apply_range")     # category, position, item tables
        check(lib.mtam_apply_end(eng.h, st), "mtam_apply_end")
        if has_user:   # apply_end re-zeroes only the local users' rows of the grads arena
            self._grad_region(int(sv.user_offset), c.user_rows).index_fill_(0, g_user.long(), 0.0)

    def train_step(self, feed: Dict[str, np.ndarray], lr: float) -> float:
        batch = self.eng.upload(feed)
        if self._graph is not None and batch.B == self._graph_B:
            self.train_step_graph(lr)
        else:
            self.train_step_device(batch, lr)
        return float(self.eng.read_scalars()[_lib.S_LOSS])

    def pipeline(self):
        """engine.FeedPipeline over the data-parallel step: the next batch is padded and copied while this one runs."""
        from .engine import FeedPipeline

        def step(B: int, lr: float) -> None:
            if self._graph is not None and B == self._graph_B:
                self.train_step_graph(lr)
            else:
                self.train_step_device(DeviceBatch({k: v[:B] for k, v in self.eng._dev.items()}, B), lr)
        return FeedPipeline(self.eng, step_fn=step)

    # ---- the whole data-parallel step, collectives included, as one CUDA graph --------------------------------------
    def capture_graph(self, B: int, warmup: int = 2) -> None:
        """Captures `train_step_device` on the engine's staging batch (fixed addresses), NCCL collectives included, into
        one CUDA graph; `train_step_graph(lr)` then advances the host-side optimizer state and replays it.  Every rank
        must call this (the collectives are captured in the same order on all of them).  Weights, Adam moments and the
        step counter are left exactly as they were."""
        eng = self.eng
        batch = DeviceBatch({k: v[:B] for k, v in eng._dev.items()}, B)
        if eng._h2d_done is None:
            eng._dev["seq_length"].fill_(2)
        t0 = eng.adam_step()
        keep = (eng.params.clone(), eng.adam_m.clone(), eng.adam_v.clone())
        s = torch.cuda.Stream(eng.device)
        s.wait_stream(torch.cuda.current_stream(eng.device))
        with torch.cuda.stream(s):
            for _ in range(warmup):                       # eager: sets kernel attributes, warms NCCL up on this stream
                self.train_step_device(batch, 0.0)
        torch.cuda.current_stream(eng.device).wait_stream(s)
        torch.cuda.synchronize(eng.device)
        dist.barrier(group=self.group)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(s):
            check(eng.lib.mtam_prepare_step(eng.h, 0.0, s.cuda_stream), "mtam_prepare_step")
        with torch.cuda.graph(g, stream=s, capture_error_mode="thread_local"):
            self.train_step_device(batch, 0.0)
        torch.cuda.synchronize(eng.device)
        eng.set_adam_step(t0)
        eng.params.copy_(keep[0]); eng.adam_m.copy_(keep[1]); eng.adam_v.copy_(keep[2])
        eng.grads.zero_()
        del keep
        self._graph, self._graph_B = g, B

    def train_step_graph(self, lr: float) -> None:
        check(self.eng.lib.mtam_prepare_step(self.eng.h, float(lr), torch.cuda.current_stream(self.eng.device).cuda_stream),
              "mtam_prepare_step")
        self._graph.replay()


def combine_lse(lse_per_shard: torch.Tensor) -> torch.Tensor:
    """[n_shards, B] per-shard log-sum-exps -> [B] global: m + log(sum_r exp(lse_r - m)), shards in index order."""
    m = lse_per_shard.max(dim=0).values
    return m + torch.log(torch.exp(lse_per_shard - m).sum(dim=0))


def shard_rows(total_rows: int, world: int) -> int:
    """Rows per shard of a row-sharded table: rank r owns [r*rows, min((r+1)*rows, total_rows))."""
    return (total_rows + world - 1) // world


class ShardedCatalogue:
    """Item table row-sharded over the ranks (the large-catalogue regime, SURVEY 8e / BASELINE configs[4]).

    The reference keeps the whole table on one device (Embedding/base_embedding.py:46-60) and scores
    `pred x table^T` against all of it (Model/base_model.py:194-202).  Here rank r holds rows
    [r*S, (r+1)*S), S = ceil(V / world):
      lookup(ids)  : all-to-all of the ids to their owners, local gather (mtam_gather), all-to-all of the rows back;
      topk(pred,k) : all-gather of pred, every rank scores its shard and keeps a local top-k with GLOBAL row numbers
                     (mtam_score_topk), all-gather of the [B,k] lists, k-way merge (mtam_merge_topk; ties -> lower
                     global index is preserved because shard order = index order);
      softmax_ce   : the training loss against the sharded table (base_model.py:316-321): all-gather of pred and
                     targets, per-shard log-sum-exp + target logit (mtam_softmax_ce_forward), all-gather of the
                     [B] log-sum-exps and all-reduce of the target logits, per-shard backward with the global
                     log-sum-exp (mtam_softmax_ce_backward): the shard's table gradient is complete and purely local
                     (no table all-reduce; Adam runs on the shard only), dpred is reduce-scattered.
    Results equal the unsharded ones bit for bit (rows are copies; scores are the same fp32 dot products).
    """

    def __init__(self, shard: torch.Tensor, total_rows: int, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.total_rows = int(total_rows)
        self.rows_per_shard = shard_rows(self.total_rows, self.world)
        self.row_begin = self.rank * self.rows_per_shard
        self.row_end = min(self.total_rows, self.row_begin + self.rows_per_shard)
        if shard.shape[0] != max(self.row_end - self.row_begin, 0):
            raise ValueError(f"rank {self.rank}: shard has {shard.shape[0]} rows, expected {self.row_end - self.row_begin}")
        self.shard = shard.contiguous()
        self.D = shard.shape[1]

    @staticmethod
    def from_full(table: torch.Tensor, group=None) -> "ShardedCatalogue":
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        S = shard_rows(table.shape[0], world)
        return ShardedCatalogue(table[rank * S: min(table.shape[0], (rank + 1) * S)].clone(), table.shape[0], group)

    def route(self, ids: torch.Tensor) -> dict:
        """The exchange plan of a lookup: ids bucketed by owner (stable), split sizes, and the ids each owner receives.
        One host synchronisation (the split sizes of the all-to-all); reused by `scatter_back`."""
        flat = ids.reshape(-1).contiguous()
        W, dev = self.world, flat.device
        owner = torch.div(flat, self.rows_per_shard, rounding_mode="floor").to(torch.int32)
        srt = sort_indices(owner, W)                              # stable: ids of one owner stay in request order
        perm = srt.perm.long()
        send_ids = flat[perm]
        send_counts = torch.bincount(srt.keys_sorted.long(), minlength=W)
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=self.group)
        both = torch.stack([send_counts, recv_counts]).to("cpu", non_blocking=False)     # the one host sync
        sc, rc = both[0].tolist(), both[1].tolist()
        recv_ids = torch.empty(sum(rc), dtype=torch.int32, device=dev)
        dist.all_to_all_single(recv_ids, send_ids, rc, sc, group=self.group)
        return dict(n=flat.numel(), shape=tuple(ids.shape), perm=perm, sc=sc, rc=rc, recv_local=recv_ids - self.row_begin)

    def lookup(self, ids: torch.Tensor, plan: Optional[dict] = None) -> torch.Tensor:
        """rows[i] = table[ids[i]] for ids anywhere in the catalogue (int32, any shape) -> [..., D]."""
        plan = plan or self.route(ids)
        dev = self.shard.device
        local = gather(self.shard, plan["recv_local"]) if plan["recv_local"].numel() else \
            torch.empty((0, self.D), dtype=torch.float32, device=dev)
        back = torch.empty((plan["n"], self.D), dtype=torch.float32, device=dev)
        dist.all_to_all_single(back, local, plan["sc"], plan["rc"], group=self.group)
        out = torch.empty_like(back)
        out[plan["perm"]] = back
        return out.reshape(*plan["shape"], self.D)

    def scatter_back(self, plan: dict, grad_rows: torch.Tensor, dshard: torch.Tensor, workspace=None) -> None:
        """The transpose of `lookup`: dshard[id - row_begin] += grad_rows[i] for every looked-up id, on the owning rank
        (all-to-all of the gradient rows to the owners, then the deterministic scatter-add).  grad_rows [n, D] may be a
        strided view."""
        dev = self.shard.device
        send = grad_rows.reshape(plan["n"], -1)[plan["perm"]].contiguous()
        recv = torch.empty((sum(plan["rc"]), self.D), dtype=torch.float32, device=dev)
        dist.all_to_all_single(recv, send, plan["rc"], plan["sc"], group=self.group)
        if recv.shape[0]:
            scatter_add(dshard, plan["recv_local"], recv, workspace)

    def softmax_ce(self, pred: torch.Tensor, target: torch.Tensor, gemm_mode: int = _lib.GEMM_TF32X3,
                   dshard_out: Optional[torch.Tensor] = None):
        """pred [B_local, D], target [B_local] int32 global item ids -> (loss_origin [B_local] = -log softmax at the
        target, dpred [B_local, D] and dshard [shard rows, D] for the MEAN loss over the global batch)."""
        W, Bl = self.world, pred.shape[0]
        dev = pred.device
        allp = torch.empty((W * Bl, self.D), dtype=torch.float32, device=dev)
        allt = torch.empty(W * Bl, dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(allp, pred.contiguous(), group=self.group)
        dist.all_gather_into_tensor(allt, target.contiguous(), group=self.group)
        rel = (allt - self.row_begin).contiguous()
        lse_r, tl = softmax_ce_forward(allp, self.shard, rel, gemm_mode)
        all_lse = torch.empty((W, W * Bl), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(all_lse, lse_r, group=self.group)
        dist.all_reduce(tl, group=self.group)
        lse = combine_lse(all_lse)
        dshard, dp = softmax_ce_backward(allp, self.shard, rel, lse, 1.0 / (W * Bl), gemm_mode, dtable=dshard_out)
        dpred = torch.empty((Bl, self.D), dtype=torch.float32, device=dev)
        dist.reduce_scatter_tensor(dpred, dp, group=self.group)
        mine = slice(self.rank * Bl, (self.rank + 1) * Bl)
        return (lse - tl)[mine], dpred, dshard

    def topk(self, pred: torch.Tensor, k: int = 50):
        """Full-catalogue top-k for this rank's pred rows [B_local, D] -> (idx [B_local,k] int32 global, score)."""
        W, Bl = self.world, pred.shape[0]
        allp = torch.empty((W * Bl, self.D), dtype=torch.float32, device=pred.device)
        dist.all_gather_into_tensor(allp, pred.contiguous(), group=self.group)
        kk = min(k, max(self.row_end - self.row_begin, 1))
        if kk != k:
            raise ValueError("shard smaller than k")
        idx, sc = score_topk(allp, self.shard, k, self.row_begin, self.row_end, index_base=self.row_begin)
        g_idx = torch.empty((W, W * Bl, k), dtype=torch.int32, device=pred.device)
        g_sc = torch.empty((W, W * Bl, k), dtype=torch.float32, device=pred.device)
        dist.all_gather_into_tensor(g_idx, idx, group=self.group)
        dist.all_gather_into_tensor(g_sc, sc, group=self.group)
        mine = slice(self.rank * Bl, (self.rank + 1) * Bl)
        return merge_topk(g_idx[:, mine].contiguous(), g_sc[:, mine].contiguous())


class ShardedItemTableTrainer:
    """MTAM train step with the item table row-sharded over the ranks -- the scalable training variant of SURVEY 8e
    (BASELINE configs[3]: 10 M items).  The data-parallel step above all-reduces the dense [V, D] item-table gradient
    (2.56 GB at 10 M items) and all-gathers the sparse rows; here every rank owns V/N rows of the table, its Adam slots
    and its gradient, and per step
      1. the item rows of the rank's tokens come from their owners (all-to-all, `ShardedCatalogue.lookup`);
      2. mtam_forward_rows runs the model on them up to pred;
      3. the softmax cross-entropy runs against the shard (all-gather of pred, per-shard log-sum-exp, combination,
         per-shard backward): the shard's dense table gradient is complete and local, dpred is reduce-scattered;
      4. mtam_backward_rows runs the backward chain from dpred;
      5. the gradient rows of the looked-up items go back to their owners (all-to-all) and are scatter-added there;
      6. ONE all-reduce carries the dense parameter gradients, the loss scalars and the squared norm of the
         rank-local pieces (shard gradient + un-deduplicated sparse pieces, trap T1); category / position pieces are
         densified locally and all-reduced, the user rows all-gathered (as in DataParallel);
      7. identical clip on every rank, Adam on the shard and on the replicated parameters (mtam_apply).
    The engine is built with item_count + 3 == rows per shard; `total_item_rows` is the catalogue's V.
    Contract: equal to the single-GPU step on the concatenated batch within the fp32 tolerances of the tests (the
    log-sum-exp is combined across shards in another order)."""

    def __init__(self, engine: Engine, total_item_rows: int, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        if engine.cfg.kind != "MTAM":
            raise NotImplementedError("ShardedItemTableTrainer: MTAM only")
        self.eng, self.group = engine, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        c = engine.c_cfg
        S = shard_rows(total_item_rows, self.world)
        if c.item_rows != S:
            raise ValueError(f"engine has {c.item_rows} item rows, a shard of {total_item_rows} over {self.world} ranks has {S}")
        self.total_rows = int(total_item_rows)
        lo, hi = self.rank * S, min(self.total_rows, (self.rank + 1) * S)
        self.local_rows = max(hi - lo, 0)
        self.cat = ShardedCatalogue(engine.param_view("embedding_layer/item")[: self.local_rows], self.total_rows, group)
        dev, D, B, L, W = engine.device, engine.cfg.D, engine.cfg.max_batch, engine.cfg.L, self.world
        self.sp = torch.zeros((c.category_rows + c.position_rows, D), dtype=torch.float32, device=dev)   # [category | position]
        self.g_user = torch.empty(W * B, dtype=torch.int32, device=dev)
        self.g_dEu = torch.empty((W * B, D), dtype=torch.float32, device=dev)
        need = max(scatter_add_workspace(W * B, c.user_rows, D), scatter_add_workspace(W * B * L, c.item_rows, D))
        self.scatter_ws = torch.empty(need, dtype=torch.uint8, device=dev)
        self.sumsq_ws = torch.empty(int(engine.lib.mtam_sumsq_workspace(engine.n_floats)), dtype=torch.uint8, device=dev)
        self.loss_origin = None
        engine.item_id_bound = self.total_rows

    def init_shard_from_full(self, full_table: torch.Tensor) -> None:
        """Copies this rank's rows of a full [V, D] table into the engine (tests / imports)."""
        S = self.cat.rows_per_shard
        v = self.eng.param_view("embedding_layer/item")
        v.zero_()
        v[: self.local_rows].copy_(full_table[self.rank * S: self.rank * S + self.local_rows])

    def _sumsq_into_norm(self, x: torch.Tensor) -> None:
        eng = self.eng
        check(eng.lib.mtam_sumsq(x.data_ptr(), x.numel(), eng.norm_sq.data_ptr(), self.sumsq_ws.data_ptr(), self.sumsq_ws.numel(),
                                 torch.cuda.current_stream(eng.device).cuda_stream), "mtam_sumsq")

    def train_step_device(self, batch: DeviceBatch, lr: float) -> None:
        eng, W = self.eng, self.world
        B, L, D = batch.B, eng.cfg.L, eng.cfg.D
        T = B * L
        c = eng.c_cfg
        st = torch.cuda.current_stream(eng.device).cuda_stream
        # 1. item rows from their owners
        plan = self.cat.route(batch.t["item_list"])
        rows = self.cat.lookup(batch.t["item_list"], plan).reshape(T, D)
        # 2. forward up to pred
        pred = torch.empty((B, D), dtype=torch.float32, device=eng.device)
        eng.norm_sq.zero_()
        eng.scalars.zero_()
        check(eng.lib.mtam_forward_rows(eng.h, C.byref(batch.c), rows.data_ptr(), pred.data_ptr(), eng.scalars.data_ptr(), st),
              "mtam_forward_rows")
        # 3. softmax cross-entropy against the sharded table; the shard's dense gradient lands in the grads arena
        sv = _lib.SparseView()
        check(eng.lib.mtam_sparse_pieces(eng.h, C.byref(sv)), "mtam_sparse_pieces")
        item_lo = int(sv.item_offset)
        item_hi = item_lo + c.item_rows * D
        g_item = eng.grads[item_lo:item_hi].view(c.item_rows, D)
        if self.local_rows < c.item_rows:
            g_item[self.local_rows:].zero_()
        loss_origin, dpred, _ = self.cat.softmax_ce(pred, batch.t["target_item_id"], eng.cfg.gemm_mode,
                                                    dshard_out=g_item[: self.local_rows])
        self.loss_origin = loss_origin
        # 4. backward chain from dpred
        check(eng.lib.mtam_backward_rows(eng.h, C.byref(batch.c), rows.data_ptr(), dpred.contiguous().data_ptr(), B * W,
                                         eng.norm_sq.data_ptr(), st), "mtam_backward_rows")
        # rank-local pieces of the global norm: the shard's dense gradient (the un-deduplicated sparse pieces were added
        # by the backward pass); loss scalars of the rank's rows
        self._sumsq_into_norm(g_item[: self.local_rows])
        eng.scalars[_lib.S_LOSS_ORIGIN] = loss_origin.sum() / float(B * W)
        # 6. one all-reduce: [dense parameter gradients | scalars | norm^2]
        dist.all_reduce(eng._grads_all[item_hi: eng.n_floats + _lib.S_COUNT + 1], group=self.group)
        self._sumsq_into_norm(eng.grads[item_hi: eng.n_floats])          # replicated pieces: counted once
        eng.scalars[_lib.S_LOSS] = eng.scalars[_lib.S_LOSS_ORIGIN] + eng.cfg.reg * eng.scalars[_lib.S_L2_NORM]
        # 5. item gradient rows back to their owners, added to the shard's gradient
        off = int(sv.item_cat_rows) - eng.workspace.data_ptr()
        dE2 = eng.workspace[off: off + T * 2 * D * 4].view(torch.float32).view(T, 2 * D)
        self.cat.scatter_back(plan, dE2[:, :D], g_item[: max(self.local_rows, 1)], self.scatter_ws)
        # category / position: densified locally, all-reduced; user rows: all-gathered
        self.sp.zero_()
        sp_cat, sp_pos = self.sp[: c.category_rows], self.sp[c.category_rows:]
        check(eng.lib.mtam_scatter_sparse_into(eng.h, None, sp_cat.data_ptr(), sp_pos.data_ptr(), None, st), "mtam_scatter_sparse_into")
        dist.all_reduce(self.sp, group=self.group)
        eng.grads[int(sv.category_offset): int(sv.category_offset) + c.category_rows * D].view(c.category_rows, D).copy_(sp_cat)
        eng.grads[int(sv.position_offset): int(sv.position_offset) + c.position_rows * D].view(c.position_rows, D).copy_(sp_pos)
        g_user = self.g_user[: W * B]
        dist.all_gather_into_tensor(g_user, batch.t["user_id"].contiguous(), group=self.group)
        g_dEu = self.g_dEu[: W * B]
        offu = int(sv.user_rows) - eng.workspace.data_ptr()
        dEu = eng.workspace[offu: offu + B * D * 4].view(torch.float32).view(B, D)
        dist.all_gather_into_tensor(g_dEu, dEu, group=self.group)
        g_usr = eng.grads[int(sv.user_offset): int(sv.user_offset) + c.user_rows * D].view(c.user_rows, D)
        scatter_add(g_usr, g_user, g_dEu, self.scatter_ws)
        # 7. clip + Adam (shard + replicated parameters)
        eng.apply(lr)
        g_usr.index_fill_(0, g_user.long(), 0.0)     # apply() re-zeroes only the local users' rows

    def train_step(self, feed: Dict[str, np.ndarray], lr: float) -> float:
        self.train_step_device(self.eng.upload(feed), lr)
        return float(self.eng.read_scalars()[_lib.S_LOSS])

    def eval_topk(self, batch: DeviceBatch, k: int = 50):
        """metrics_topK over the sharded table: forward on looked-up rows, sharded score + top-k, all-gather merge."""
        eng = self.eng
        plan = self.cat.route(batch.t["item_list"])
        rows = self.cat.lookup(batch.t["item_list"], plan).reshape(batch.B * eng.cfg.L, eng.cfg.D)
        pred = torch.empty((batch.B, eng.cfg.D), dtype=torch.float32, device=eng.device)
        check(eng.lib.mtam_forward_rows(eng.h, C.byref(batch.c), rows.data_ptr(), pred.data_ptr(), None,
                                        torch.cuda.current_stream(eng.device).cuda_stream), "mtam_forward_rows")
        return self.cat.topk(pred, k)

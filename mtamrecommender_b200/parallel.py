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
from typing import Dict

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import check
from .engine import DeviceBatch, Engine, scatter_add, scatter_add_workspace


class DataParallel:
    def __init__(self, engine: Engine, group=None):
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
        self.g_item = torch.empty(W * T, dtype=torch.int32, device=dev)
        self.g_cat = torch.empty(W * T, dtype=torch.int32, device=dev)
        self.g_pos = torch.empty(W * T, dtype=torch.int32, device=dev)
        self.g_user = torch.empty(W * B, dtype=torch.int32, device=dev)
        self.g_dE2 = torch.empty((W * T, 2 * D), dtype=torch.float32, device=dev)
        self.g_dEp = torch.empty((W * T, D), dtype=torch.float32, device=dev)
        self.g_dEu = torch.empty((W * B, D), dtype=torch.float32, device=dev)
        c = engine.c_cfg
        need = max(scatter_add_workspace(W * T, c.item_rows, D), scatter_add_workspace(W * T, c.category_rows, D),
                   scatter_add_workspace(W * T, c.position_rows, D), scatter_add_workspace(W * B, c.user_rows, D))
        self.scatter_ws = torch.empty(need, dtype=torch.uint8, device=dev)

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
        eng.forward_backward_device(batch, global_batch=B * W)
        sv = _lib.SparseView()
        check(eng.lib.mtam_sparse_pieces(eng.h, C.byref(sv)), "mtam_sparse_pieces")
        dist.all_reduce(eng.scalars[:3], group=self.group)
        dist.all_reduce(eng.grads[int(sv.dense_begin):], group=self.group)
        dist.all_reduce(eng.norm_sq, group=self.group)
        eng.finish_grads(scatter_local=False)
        c = eng.c_cfg
        g_item = self._gather(self.g_item, batch.t["item_list"].reshape(-1))
        g_cat = self._gather(self.g_cat, batch.t["category_list"].reshape(-1))
        g_pos = self._gather(self.g_pos, batch.t["position_list"].reshape(-1))
        g_dE2 = self._gather(self.g_dE2, self._ws_view(sv.item_cat_rows, T, 2 * D))
        g_dEp = self._gather(self.g_dEp, self._ws_view(sv.position_rows, T, D))
        scatter_add(self._grad_region(int(sv.item_offset), c.item_rows), g_item, g_dE2[:, :D], self.scatter_ws)
        scatter_add(self._grad_region(int(sv.category_offset), c.category_rows), g_cat, g_dE2[:, D:], self.scatter_ws)
        scatter_add(self._grad_region(int(sv.position_offset), c.position_rows), g_pos, g_dEp, self.scatter_ws)
        g_user = None
        if sv.has_user:
            g_user = self._gather(self.g_user, batch.t["user_id"])
            g_dEu = self._gather(self.g_dEu, self._ws_view(sv.user_rows, B, D))
            scatter_add(self._grad_region(int(sv.user_offset), c.user_rows), g_user, g_dEu, self.scatter_ws)
        eng.apply(lr)
        if g_user is not None:   # apply() re-zeroes only the local users' rows of the grads arena
            self._grad_region(int(sv.user_offset), c.user_rows).index_fill_(0, g_user.long(), 0.0)

    def train_step(self, feed: Dict[str, np.ndarray], lr: float) -> float:
        self.train_step_device(self.eng.upload(feed), lr)
        return float(self.eng.read_scalars()[_lib.S_LOSS])

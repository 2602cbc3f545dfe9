"""Host-side engine: owns the device arenas (torch tensors as plain device memory), the model
handle and the streams, and turns the reference's feed dictionaries into C-ABI calls.

It plays the role `tf.Session` plays in the reference (train_process.py:146, base_model.py:150-167):
`Engine.train_step(feed, lr)` == `sess.run([loss, merged, train_op], feed_dict)`.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from ._lib import Batch, Config, ParamInfo, Sizes, check

FEED_INT = ("user_id", "item_list", "category_list", "position_list", "target_item_id",
            "target_item_category", "seq_length")
FEED_FLOAT = ("time_list", "timelast_list", "timenow_list", "target_item_time")
FEED_KEYS = FEED_INT + FEED_FLOAT


@dataclass
class ModelConfig:
    kind: str = "MTAM"
    max_batch: int = 256
    L: int = 50
    D: int = 128
    H: int = 1
    N: int = 6
    user_count: int = 0
    item_count: int = 0
    category_count: int = 0
    reg: float = 5e-5
    clip: float = 1.0
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8
    gemm_mode: int = _lib.GEMM_FP32
    optimizer: str = "adam"          # "adam" | "sgd" (base_model.py:71-80)
    dropout: float = 0.0             # attention dropout of SASREC / TISASREC; ignored by the other kinds, as in the reference
    dropout_seed: int = 1234

    def to_c(self) -> Config:
        c = Config()
        c.abi_version = _lib.ABI_VERSION
        c.kind = _lib.KINDS[self.kind]
        c.max_batch, c.L, c.D, c.H, c.N = self.max_batch, self.L, self.D, self.H, self.N
        c.user_rows, c.item_rows = self.user_count + 3, self.item_count + 3
        c.category_rows, c.position_rows = self.category_count + 3, self.L + 3
        c.reg, c.clip, c.beta1, c.beta2, c.eps = self.reg, self.clip, self.beta1, self.beta2, self.eps
        c.gemm_mode = self.gemm_mode
        if self.optimizer not in _lib.OPTIMIZERS:
            raise NotImplementedError(f"optimizer {self.optimizer!r}: built are {sorted(_lib.OPTIMIZERS)}")
        c.optimizer = _lib.OPTIMIZERS[self.optimizer]
        c.dropout, c.dropout_seed = float(self.dropout), int(self.dropout_seed) & 0xFFFFFFFF
        return c


class DeviceBatch:
    """The 11 feed arrays resident on the device, plus the C struct that points at them."""

    def __init__(self, tensors: Dict[str, torch.Tensor], B: int):
        self.t = tensors
        self.B = B
        self.c = Batch()
        self.c.B = B
        for k in FEED_KEYS:
            setattr(self.c, k, tensors[k].data_ptr())

    def nbytes(self) -> int:
        return sum(v.numel() * v.element_size() for v in self.t.values())


class Engine:
    def __init__(self, cfg: ModelConfig, device: str = "cuda:0", seed: Optional[int] = None):
        if not torch.cuda.is_available():
            raise _lib.MtamError("mtamrecommender_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        self.c_cfg = cfg.to_c()
        sizes = Sizes()
        check(self.lib.mtam_plan(C.byref(self.c_cfg), C.byref(sizes)), "mtam_plan")
        self.n_floats = int(sizes.param_floats)
        self.workspace_bytes = int(sizes.workspace_bytes)
        dev = self.device
        self.params = torch.zeros(self.n_floats, dtype=torch.float32, device=dev)
        # the step's scalars (loss pieces, norms) and the squared-norm accumulator sit right behind the grads arena, so a
        # data-parallel driver reduces the dense tail of the arena and these in ONE collective
        self._grads_all = torch.zeros(self.n_floats + 16, dtype=torch.float32, device=dev)
        self.grads = self._grads_all[:self.n_floats]
        self.adam_m = torch.zeros(self.n_floats, dtype=torch.float32, device=dev)
        self.adam_v = torch.zeros(self.n_floats, dtype=torch.float32, device=dev)
        self.workspace = torch.empty(self.workspace_bytes, dtype=torch.uint8, device=dev)
        self.scalars = self._grads_all[self.n_floats: self.n_floats + _lib.S_COUNT]
        self.norm_sq = self._grads_all[self.n_floats + _lib.S_COUNT: self.n_floats + _lib.S_COUNT + 1]
        h = C.c_void_p()
        check(self.lib.mtam_create(C.byref(self.c_cfg), self.params.data_ptr(), self.grads.data_ptr(),
                                   self.adam_m.data_ptr(), self.adam_v.data_ptr(), self.workspace.data_ptr(),
                                   self.workspace_bytes, C.byref(h)), "mtam_create")
        self.h = h
        self.info: Dict[str, ParamInfo] = {}
        for i in range(self.lib.mtam_param_count(self.h)):
            pi = ParamInfo()
            check(self.lib.mtam_param_info_get(self.h, i, C.byref(pi)), "mtam_param_info_get")
            self.info[pi.name.decode()] = pi
        # Per-step host->device feed: the 11 arrays live back to back in ONE pinned host buffer and ONE device buffer
        # (256-byte aligned regions), so a step's feed is a single cudaMemcpyAsync; `_pinned[k]` / `_dev[k]` are views.
        self._pinned: Dict[str, torch.Tensor] = {}
        self._pinned_np: Dict[str, np.ndarray] = {}
        self._dev: Dict[str, torch.Tensor] = {}
        B, L = cfg.max_batch, cfg.L
        offs, total = {}, 0
        for k in FEED_KEYS:
            n = B * L if k.endswith("_list") else B
            offs[k] = (total, n)
            total += (n * 4 + 255) // 256 * 256
        self._pinned_all = torch.empty(total, dtype=torch.uint8).pin_memory()
        # zero-filled: a step run on the staging buffers before any upload (graph capture) must see valid ids
        self._pinned_all.zero_()
        self._dev_all = torch.zeros(total, dtype=torch.uint8, device=dev)
        for k in FEED_KEYS:
            o, n = offs[k]
            shape = (B, L) if k.endswith("_list") else (B,)
            dt = torch.int32 if k in FEED_INT else torch.float32
            self._pinned[k] = self._pinned_all[o: o + n * 4].view(dt).view(shape)
            self._pinned_np[k] = self._pinned[k].numpy()
            self._dev[k] = self._dev_all[o: o + n * 4].view(dt).view(shape)
        self._scalars_host = torch.empty(_lib.S_COUNT, dtype=torch.float32).pin_memory()
        self._graph = None
        self._graph_B = -1
        self.last_scalars = np.zeros(_lib.S_COUNT, np.float32)
        self.item_id_bound = None        # set by parallel.ShardedItemTableTrainer: the catalogue's row count
        self._h2d_done = None
        if seed is not None:
            self.init_random(seed)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.mtam_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- parameters ----------------------------------------------------------------------
    def param_names(self):
        return list(self.info.keys())

    def _view(self, arena: torch.Tensor, name: str) -> torch.Tensor:
        pi = self.info[name]
        v = arena.as_strided((pi.rows, pi.cols), (pi.ld, 1), int(pi.offset))
        return v[0] if pi.ndim == 1 else v

    def param_view(self, name): return self._view(self.params, name)
    def grad_view(self, name): return self._view(self.grads, name)
    def adam_m_view(self, name): return self._view(self.adam_m, name)
    def adam_v_view(self, name): return self._view(self.adam_v, name)

    def is_dead(self, name: str) -> bool:
        return bool(self.info[name].flags & _lib.PARAM_DEAD)

    def set_param(self, name: str, value) -> None:
        v = self.param_view(name)
        t = torch.as_tensor(np.asarray(value), dtype=torch.float32).reshape(v.shape)
        v.copy_(t.to(self.device))

    def get_param(self, name: str) -> np.ndarray:
        return self.param_view(name).detach().cpu().numpy().copy()

    def set_params(self, params: Dict[str, np.ndarray]) -> None:
        missing = set(self.info) - set(params)
        extra = set(params) - set(self.info)
        if missing or extra:
            raise KeyError(f"parameter name mismatch: missing={sorted(missing)} extra={sorted(extra)}")
        for k, v in params.items():
            self.set_param(k, v)

    def get_params(self) -> Dict[str, np.ndarray]:
        return {k: self.get_param(k) for k in self.info}

    def init_random(self, seed: int = 1234) -> None:
        """Initialisers the reference ends up with: tables U(+-sqrt(6/D)), glorot-uniform
        for get_variable defaults, dense bias 0, GRU gate bias 1, LayerNorm gamma 1 / beta 0."""
        g = torch.Generator(device="cpu")
        g.manual_seed(seed)
        for name, pi in self.info.items():
            leaf = name.rsplit("/", 1)[-1]
            shape = (pi.rows, pi.cols) if pi.ndim == 2 else (pi.cols,)
            if name.startswith("embedding_layer/"):
                r = (6.0 / pi.cols) ** 0.5
                t = (torch.rand(shape, generator=g) * 2 - 1) * r
            elif leaf == "gamma":
                t = torch.ones(shape)
            elif leaf == "beta":
                t = torch.zeros(shape)
            elif leaf == "bias":
                t = torch.ones(shape) if name.endswith("gates/bias") else torch.zeros(shape)
            else:
                fi, fo = (shape[0], shape[0]) if len(shape) == 1 else (shape[0], shape[1])
                lim = (6.0 / (fi + fo)) ** 0.5
                t = (torch.rand(shape, generator=g) * 2 - 1) * lim
            self.param_view(name).copy_(t.to(self.device))

    def adam_step(self) -> int:
        t = C.c_int64()
        check(self.lib.mtam_get_adam_step(self.h, C.byref(t)), "mtam_get_adam_step")
        return int(t.value)

    def set_bpr_negative(self, item_id: int) -> None:
        """BPR-MF: fix the shared negative item id (BPRMF.py:43 draws it with tf.random_uniform)."""
        check(self.lib.mtam_set_bpr_negative(self.h, int(item_id)), "mtam_set_bpr_negative")

    def set_dropout_state(self, seed: int, counter: int) -> None:
        """Attention dropout (SASREC / TISASREC): the next forward pass uses the keep mask of call `counter` under `seed`
        (`dropout_keep_mask` restates it on the host)."""
        check(self.lib.mtam_set_dropout_state(self.h, int(seed) & 0xFFFFFFFF, int(counter) & 0xFFFFFFFF), "mtam_set_dropout_state")

    def set_adam_step(self, t: int) -> None:
        check(self.lib.mtam_set_adam_step(self.h, int(t)), "mtam_set_adam_step")

    # ---- feed ----------------------------------------------------------------------------
    def upload(self, feed: Dict[str, np.ndarray]) -> DeviceBatch:
        """Host feed (numpy, as make_feed_dic_new builds it) -> pinned staging -> device."""
        B = int(len(feed["user_id"]))
        if B < 1 or B > self.cfg.max_batch:
            raise ValueError(f"batch size {B} outside [1, {self.cfg.max_batch}]")
        self._validate_ids(feed, B)
        if self._h2d_done is not None:
            self._h2d_done.synchronize()          # the previous feed's DMA must have left the pinned buffer
        for k in FEED_KEYS:
            np.copyto(self._pinned_np[k][:B], feed[k], casting="same_kind")
        self._dev_all.copy_(self._pinned_all, non_blocking=True)          # one H2D copy for the whole feed
        if self._h2d_done is None:
            self._h2d_done = torch.cuda.Event()
        self._h2d_done.record(torch.cuda.current_stream(self.device))
        return DeviceBatch({k: self._dev[k][:B] for k in FEED_KEYS}, B)

    def _validate_ids(self, feed, B: int) -> None:
        """The gather / scatter kernels index the tables unchecked (as tf.gather does on the GPU): reject feeds whose
        ids fall outside the tables, or whose lengths fall outside [2, L], on the host."""
        c = self.cfg
        items = self.item_id_bound or c.item_count + 3      # (row-sharded item table: ids are global)
        bounds = (("user_id", c.user_count + 3), ("item_list", items), ("target_item_id", items),
                  ("category_list", c.category_count + 3), ("position_list", c.L + 3))
        for k, hi in bounds:
            a = np.asarray(feed[k])
            if a.size == 0:
                continue
            # one pass per array: viewed as unsigned, a negative id is a huge one
            top = int(a.view(np.uint32).max()) if a.dtype == np.int32 and a.flags.c_contiguous else \
                (int(a.max()) if int(a.min()) >= 0 else 1 << 40)
            if top >= hi:
                raise ValueError(f"feed array {k}: ids outside [0, {hi})")
        if c.kind != "BPRMF":
            sl = np.asarray(feed["seq_length"])
            if sl.size and (int(sl.min()) < 2 or int(sl.max()) > c.L):
                raise ValueError(f"feed array seq_length: values outside [2, {c.L}]")

    def upload_records(self, records) -> DeviceBatch:
        """A `PackedRecords` view (DataHandle/record_store.py) -> padded by mtam_pack_records straight into the
        pinned staging arrays -> one H2D copy.  The columnar counterpart of make_feed_dic_new + upload."""
        B = len(records)
        if B < 1 or B > self.cfg.max_batch:
            raise ValueError(f"batch size {B} outside [1, {self.cfg.max_batch}]")
        if self._h2d_done is not None:
            self._h2d_done.synchronize()
        records.pack_into(self._pinned_np, self.cfg.L)
        self._dev_all.copy_(self._pinned_all, non_blocking=True)
        if self._h2d_done is None:
            self._h2d_done = torch.cuda.Event()
        self._h2d_done.record(torch.cuda.current_stream(self.device))
        return DeviceBatch({k: self._dev[k][:B] for k in FEED_KEYS}, B)

    def train_step_records(self, records, lr: float) -> float:
        """`train_step` for a `PackedRecords` view."""
        batch = self.upload_records(records)
        return self._step_uploaded(batch, lr)

    def device_batch(self, tensors: Dict[str, torch.Tensor]) -> DeviceBatch:
        B = int(tensors["user_id"].shape[0])
        for k in FEED_KEYS:
            t = tensors[k]
            want = torch.int32 if k in FEED_INT else torch.float32
            if t.dtype != want or not t.is_cuda or not t.is_contiguous():
                raise TypeError(f"feed array {k}: need contiguous cuda {want}")
        return DeviceBatch(tensors, B)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    # ---- steps ---------------------------------------------------------------------------
    def train_step_device(self, batch: DeviceBatch, lr: float) -> None:
        check(self.lib.mtam_train_step(self.h, C.byref(batch.c), float(lr), self.scalars.data_ptr(), self._stream()),
              "mtam_train_step")

    def read_scalars(self) -> np.ndarray:
        self._scalars_host.copy_(self.scalars, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self._scalars_host.numpy().copy()

    def train_step(self, feed: Dict[str, np.ndarray], lr: float) -> float:
        """One `model.train` call from a host feed: pack -> one H2D copy -> the step (the captured CUDA graph when
        `capture_train_graph` was called for this batch size, else eager launches) -> loss read back."""
        return self._step_uploaded(self.upload(feed), lr)

    def _step_uploaded(self, batch: DeviceBatch, lr: float) -> float:
        if self._graph is not None and batch.B == self._graph_B:
            self.train_step_graph(lr)
        else:
            self.train_step_device(batch, lr)
        self.last_scalars = self.read_scalars()
        return float(self.last_scalars[_lib.S_LOSS])

    def forward_device(self, batch: DeviceBatch, want_pred=True):
        B, D = batch.B, self.cfg.D
        lo = torch.empty(B, dtype=torch.float32, device=self.device)
        pred = torch.empty((B, D), dtype=torch.float32, device=self.device) if want_pred else None
        check(self.lib.mtam_forward(self.h, C.byref(batch.c), self.scalars.data_ptr(), lo.data_ptr(),
                                    pred.data_ptr() if want_pred else None, self._stream()), "mtam_forward")
        return lo, pred

    def forward(self, feed):
        lo, pred = self.forward_device(self.upload(feed))
        s = self.read_scalars()
        return dict(loss=float(s[_lib.S_LOSS]), loss_origin_mean=float(s[_lib.S_LOSS_ORIGIN]),
                    l2_norm=float(s[_lib.S_L2_NORM]), loss_origin=lo.cpu().numpy(), pred=pred.cpu().numpy())

    def forward_backward_device(self, batch: DeviceBatch, global_batch: Optional[int] = None) -> None:
        self.norm_sq.zero_()
        check(self.lib.mtam_forward_backward(self.h, C.byref(batch.c), int(global_batch or batch.B),
                                             self.scalars.data_ptr(), self.norm_sq.data_ptr(), self._stream()),
              "mtam_forward_backward")

    def finish_grads(self, scatter_local: bool = True) -> None:
        check(self.lib.mtam_finish_grads(self.h, self.norm_sq.data_ptr(), 1 if scatter_local else 0, self._stream()),
              "mtam_finish_grads")

    def apply(self, lr: float) -> None:
        check(self.lib.mtam_apply(self.h, float(lr), self.norm_sq.data_ptr(), self.scalars.data_ptr(), self._stream()),
              "mtam_apply")

    def gradients(self, feed) -> Dict[str, np.ndarray]:
        """Dense (de-duplicated) gradients of the last batch, for parity tests: runs forward_backward and
        finish_grads but not apply; the arena is cleaned afterwards."""
        b = self.upload(feed)
        self.forward_backward_device(b)
        self.finish_grads()
        out = {k: self.grad_view(k).detach().cpu().numpy().copy() for k in self.info}
        out["__norm_sq__"] = float(self.norm_sq.item())
        out["__scalars__"] = self.read_scalars()
        self.grads.zero_()
        # drop the pending flag by applying a zero-lr step on zero grads would change Adam state;
        # instead recreate nothing: the next forward_backward simply overwrites.
        return out

    # ---- profiling / diagnostics --------------------------------------------------------------
    def profile_step(self, batch: DeviceBatch, lr: float) -> Dict[str, float]:
        """One train step with CUDA-event phase markers; returns milliseconds per phase."""
        check(self.lib.mtam_profile_enable(self.h, 1), "mtam_profile_enable")
        try:
            self.train_step_device(batch, lr)
            ms = (C.c_float * len(_lib.PHASES))()
            check(self.lib.mtam_profile_read(self.h, ms, len(_lib.PHASES)), "mtam_profile_read")
        finally:
            check(self.lib.mtam_profile_enable(self.h, 0), "mtam_profile_enable")
        return {k: float(ms[i]) for i, k in enumerate(_lib.PHASES)}

    def launch_count(self) -> int:
        return int(self.lib.mtam_launch_count())

    # ---- CUDA graph of the whole step ----------------------------------------------------------
    def capture_train_graph(self, B: int) -> None:
        """Captures mtam_train_step on the staging batch buffers (fixed addresses) into a CUDA graph.
        `train_step_graph` then only advances the host-side Adam state and replays it."""
        batch = DeviceBatch({k: v[:B] for k, v in self._dev.items()}, B)
        if self._h2d_done is None:
            # nothing uploaded yet: the staging buffers hold zeros (valid pad ids); lengths must be >= 2
            self._dev["seq_length"].fill_(2)
        t0 = self.adam_step()
        # the warm-up step and the capture run the optimizer with lr = 0: weights stay, but the Adam moments would absorb
        # that batch's gradient -- keep them (and the gradient arena) as they were
        m0, v0 = self.adam_m.clone(), self.adam_v.clone()
        # high priority: the captured main chain is scheduled ahead of the library's side streams (the softmax
        # dTable pass running beside the backward chain fills whatever the chain leaves free)
        s = torch.cuda.Stream(self.device, priority=-1)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            self.train_step_device(batch, 0.0)       # warm-up outside capture: sets func attributes
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        raise_if = None
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(s):   # eager: keeps the per-step lr_t store out of the captured step
            check(self.lib.mtam_prepare_step(self.h, 0.0, s.cuda_stream), "mtam_prepare_step")
        with torch.cuda.graph(g, stream=s):
            self.train_step_device(batch, 0.0)
        torch.cuda.synchronize(self.device)
        self.set_adam_step(t0)
        self.adam_m.copy_(m0)
        self.adam_v.copy_(v0)
        del m0, v0
        self._graph, self._graph_B = g, B
        return raise_if

    def train_step_graph(self, lr: float) -> None:
        check(self.lib.mtam_prepare_step(self.h, float(lr), self._stream()), "mtam_prepare_step")
        self._graph.replay()

    def eval_topk_device(self, batch: DeviceBatch, k: int = 50):
        idx = torch.empty((batch.B, k), dtype=torch.int32, device=self.device)
        sc = torch.empty((batch.B, k), dtype=torch.float32, device=self.device)
        check(self.lib.mtam_eval_topk(self.h, C.byref(batch.c), k, idx.data_ptr(), sc.data_ptr(), self._stream()),
              "mtam_eval_topk")
        return idx, sc

    def hr_ndcg_device(self, idx: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        out = torch.empty(10, dtype=torch.float32, device=self.device)
        check(self.lib.mtam_hr_ndcg(idx.data_ptr(), idx.shape[0], idx.shape[1], target.data_ptr(), out.data_ptr(),
                                    self._stream()), "mtam_hr_ndcg")
        return out


class FeedPipeline:
    """Double-buffered input pipeline for training (pinned memory, async H2D, CUDA-graphed step).

    `submit(feed, lr)` pads / packs the batch into a free pinned buffer, starts its host->device copy on a copy stream,
    queues the step behind it (one device-to-device copy refreshes the fixed staging buffers the step -- eager or
    captured graph -- reads) and the device->host copy of its scalars, and only THEN waits for the PREVIOUS step, whose
    scalars it returns: the host's share of step i+1 (id validation, padding, the copy's submission) runs while the
    device computes step i, and the device always has the next step queued.  `flush()` returns the scalars of the last
    submitted step.  Same arithmetic, same order of steps as calling `Engine.train_step` in a loop (tested).
    `step_fn(lr)` runs one step on the engine's staging batch (default: the engine's own graph or eager step; a
    DataParallel driver passes its own)."""

    def __init__(self, eng: "Engine", step_fn=None):
        self.eng = eng
        dev = eng.device
        n = eng._pinned_all.numel()
        self._pin = [torch.zeros(n, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self._stage = [torch.zeros(n, dtype=torch.uint8, device=dev) for _ in range(2)]
        self._np = []
        for b in self._pin:             # the same 11 views as Engine._pinned_np, over this slot's buffer
            views = {}
            for k in FEED_KEYS:
                ref = eng._pinned[k]
                off = ref.data_ptr() - eng._pinned_all.data_ptr()
                views[k] = b[off: off + ref.numel() * 4].view(ref.dtype).view(ref.shape).numpy()
            self._np.append(views)
        self._scal = [torch.zeros(_lib.S_COUNT, dtype=torch.float32).pin_memory() for _ in range(2)]
        self._copy = torch.cuda.Stream(dev)
        self._h2d = [torch.cuda.Event() for _ in range(2)]        # slot's H2D copy done
        self._used = [torch.cuda.Event() for _ in range(2)]       # slot's device staging buffer consumed by its step
        self._done = [torch.cuda.Event() for _ in range(2)]       # slot's step finished, scalars on the host
        self._n = 0
        self._B = [0, 0]
        self._step_fn = step_fn

    def _run(self, B: int, lr: float) -> None:
        eng = self.eng
        if self._step_fn is not None:
            self._step_fn(B, lr)
        elif eng._graph is not None and B == eng._graph_B:
            eng.train_step_graph(lr)
        else:
            eng.train_step_device(DeviceBatch({k: v[:B] for k, v in eng._dev.items()}, B), lr)

    def submit(self, feed, lr: float) -> Optional[np.ndarray]:
        """feed: the mapping of 11 arrays (make_feed_dic_new) or a PackedRecords view.  Returns the scalars
        (mtam_scalar order) of the step submitted before this one, None for the first."""
        eng = self.eng
        slot = self._n & 1
        records = hasattr(feed, "pack_into")
        B = len(feed) if records else int(len(feed["user_id"]))
        if B < 1 or B > eng.cfg.max_batch:
            raise ValueError(f"batch size {B} outside [1, {eng.cfg.max_batch}]")
        if self._n >= 2:
            self._h2d[slot].synchronize()            # the pinned buffer's previous copy (two steps ago) has left it
        if records:
            feed.pack_into(self._np[slot], eng.cfg.L)
        else:
            eng._validate_ids(feed, B)
            for k in FEED_KEYS:
                np.copyto(self._np[slot][k][:B], feed[k], casting="same_kind")
        main = torch.cuda.current_stream(eng.device)
        with torch.cuda.stream(self._copy):
            if self._n >= 2:
                self._copy.wait_event(self._used[slot])            # the step two back has read this staging buffer
            self._stage[slot].copy_(self._pin[slot], non_blocking=True)
            self._h2d[slot].record(self._copy)
        main.wait_event(self._h2d[slot])
        eng._dev_all.copy_(self._stage[slot], non_blocking=True)
        self._used[slot].record(main)
        self._run(B, lr)
        self._scal[slot].copy_(eng.scalars, non_blocking=True)
        self._done[slot].record(main)
        self._B[slot] = B
        self._n += 1
        eng._h2d_done = self._h2d[slot]       # (a later Engine.upload must not capture its graph on unset lengths)
        if self._n == 1:
            return None
        return self._collect(1 - slot)

    def _collect(self, slot: int) -> np.ndarray:
        self._done[slot].synchronize()
        out = self._scal[slot].numpy().copy()
        self.eng.last_scalars = out
        return out

    def flush(self) -> Optional[np.ndarray]:
        """Scalars of the last submitted step (None if nothing was submitted since the last flush)."""
        if self._n == 0:
            return None
        out = self._collect((self._n - 1) & 1)
        torch.cuda.current_stream(self.eng.device).synchronize()
        self._n = 0
        return out


def dropout_keep_mask(seed: int, counter: int, block: int, B: int, H: int, L: int, rate: float) -> np.ndarray:
    """Host restatement of csrc/selfattn.cu sa_keep(): the multiplicative attention-dropout mask [B,H,L,L]
    (keep / (1 - rate), else 0) of block `block` in forward call `counter`.  Test / inspection helper."""
    thr = np.uint32(int(round(np.float32(rate) * np.float32(16777216.0))))
    e = np.arange(B * H * L * L, dtype=np.uint64).astype(np.uint32)
    with np.errstate(over="ignore"):
        key = np.uint32((seed + 0x85EBCA77 * ((counter * 64 + block) & 0xFFFFFFFF)) & 0xFFFFFFFF)
        x = (e * np.uint32(0x9E3779B1)) ^ key
        x ^= x >> np.uint32(16); x *= np.uint32(0x85EBCA6B); x ^= x >> np.uint32(13); x *= np.uint32(0xC2B2AE35)
        x ^= x >> np.uint32(16)
    keep = (x >> np.uint32(8)) >= thr
    scale = np.float32(1.0) / (np.float32(1.0) - np.float32(rate))
    return (keep.astype(np.float32) * scale).reshape(B, H, L, L)


# ---- stand-alone kernels --------------------------------------------------------------------
def gather(table: torch.Tensor, idx: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    n = idx.numel()
    D = table.shape[1]
    if out is None:
        out = torch.empty((n, D), dtype=torch.float32, device=table.device)
    check(lib.mtam_gather(table.data_ptr(), table.shape[0], D, idx.data_ptr(), n, out.data_ptr(),
                          torch.cuda.current_stream(table.device).cuda_stream), "mtam_gather")
    return out


def scatter_add_workspace(n: int, table_rows: int, D: int) -> int:
    return int(_lib.load().mtam_scatter_add_workspace(n, table_rows, D))


def scatter_add(dst: torch.Tensor, idx: torch.Tensor, rows: torch.Tensor, workspace: Optional[torch.Tensor] = None,
                want_unique: bool = False, accumulate: bool = True):
    """dst[idx[i]] += rows[i] (mtam_scatter_add); accumulate=False: dst[r] = sum of the rows whose index is r, other
    rows of dst untouched, dst never read."""
    lib = _lib.load()
    n = idx.numel()
    R, D = dst.shape
    if workspace is None:
        workspace = torch.empty(max(scatter_add_workspace(n, R, D), 16), dtype=torch.uint8, device=dst.device)
    uniq = torch.empty(max(n, 1), dtype=torch.int32, device=dst.device) if want_unique else None
    nuniq = torch.zeros(1, dtype=torch.int32, device=dst.device) if want_unique else None
    check(lib.mtam_scatter_add(dst.data_ptr(), R, D, idx.data_ptr(), rows.data_ptr(), rows.stride(0), n,
                               1 if accumulate else 0, workspace.data_ptr(),
                               workspace.numel(), uniq.data_ptr() if want_unique else None,
                               nuniq.data_ptr() if want_unique else None,
                               torch.cuda.current_stream(dst.device).cuda_stream), "mtam_scatter_add")
    if want_unique:
        return dst, uniq, nuniq
    return dst


class SortedIndices:
    """Result of `sort_indices`: sorted keys and the permutation, as int32 views into the sort workspace."""

    def __init__(self, keys_sorted: torch.Tensor, perm: torch.Tensor, workspace: torch.Tensor):
        self.keys_sorted, self.perm, self.workspace = keys_sorted, perm, workspace


def sort_indices(keys: torch.Tensor, key_bound: int, workspace: Optional[torch.Tensor] = None) -> SortedIndices:
    """Stable radix sort of (keys[i], i) on the device (mtam_sort_indices)."""
    lib = _lib.load()
    n = keys.numel()
    need = max(int(lib.mtam_sort_workspace(n, key_bound)), 16)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=keys.device)
    ks, pm = C.c_void_p(), C.c_void_p()
    check(lib.mtam_sort_indices(keys.data_ptr(), n, key_bound, workspace.data_ptr(), workspace.numel(), C.byref(ks),
                                C.byref(pm), torch.cuda.current_stream(keys.device).cuda_stream), "mtam_sort_indices")

    def view(ptr):
        if n == 0:
            return torch.empty(0, dtype=torch.int32, device=keys.device)
        off = ptr.value - workspace.data_ptr()
        return workspace[off: off + 4 * n].view(torch.int32)
    return SortedIndices(view(ks), view(pm), workspace)


def scatter_add_sorted(dst: torch.Tensor, sorted_idx: SortedIndices, rows: torch.Tensor,
                       workspace: Optional[torch.Tensor] = None, accumulate: bool = True) -> torch.Tensor:
    """dst[keys_sorted[j]] (+)= rows[perm[j]] -- the segmented-reduction half of scatter_add."""
    lib = _lib.load()
    n = sorted_idx.keys_sorted.numel()
    R, D = dst.shape
    need = max(int(lib.mtam_scatter_add_sorted_workspace(n, D)), 16)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=dst.device)
    check(lib.mtam_scatter_add_sorted(dst.data_ptr(), D, sorted_idx.keys_sorted.data_ptr(), sorted_idx.perm.data_ptr(),
                                      rows.data_ptr(), rows.stride(0), n, 1 if accumulate else 0, workspace.data_ptr(),
                                      workspace.numel(), torch.cuda.current_stream(dst.device).cuda_stream),
          "mtam_scatter_add_sorted")
    return dst


def score_topk(pred: torch.Tensor, table: torch.Tensor, k: int, row_begin: int = 0, row_end: Optional[int] = None,
               index_base: int = 0, workspace: Optional[torch.Tensor] = None, gemm_mode: int = _lib.GEMM_TF32X3):
    """Top-k of pred x table[row_begin:row_end]^T (mtam_score_topk): sorted descending, ties -> lower index.
    `table` holds rows [index_base, index_base + table.shape[0]) of the catalogue; indices returned are global."""
    lib = _lib.load()
    B, D = pred.shape
    row_end = index_base + table.shape[0] if row_end is None else row_end
    rows = row_end - row_begin
    if workspace is None:       # a caller-supplied workspace is passed as is (the library checks its size)
        need = max(int(lib.mtam_score_topk_workspace(B, rows, k)), 16)
        workspace = torch.empty(need, dtype=torch.uint8, device=pred.device)
    idx = torch.empty((B, k), dtype=torch.int32, device=pred.device)
    sc = torch.empty((B, k), dtype=torch.float32, device=pred.device)
    # the library indexes item_table by global row number: pass the base shifted back by index_base rows
    base = table.data_ptr() - index_base * D * 4
    check(lib.mtam_score_topk(gemm_mode, pred.data_ptr(), B, D, base, row_begin, row_end, k, idx.data_ptr(), sc.data_ptr(),
                              workspace.data_ptr(), workspace.numel(),
                              torch.cuda.current_stream(pred.device).cuda_stream), "mtam_score_topk")
    return idx, sc


def softmax_ce_forward(pred: torch.Tensor, table: torch.Tensor, target: torch.Tensor, gemm_mode: int = _lib.GEMM_TF32X3,
                       workspace: Optional[torch.Tensor] = None):
    """Log-sum-exp of pred x table^T per pred row and the target's logit (mtam_softmax_ce_forward).  `table` is the
    whole catalogue or one shard; `target` (int32) is relative to it -- rows whose target lies outside get logit 0."""
    lib = _lib.load()
    B, D = pred.shape
    rows = table.shape[0]
    if workspace is None:
        workspace = torch.empty(int(lib.mtam_softmax_ce_workspace(B, D, rows)), dtype=torch.uint8, device=pred.device)
    lse = torch.empty(B, dtype=torch.float32, device=pred.device)
    tl = torch.empty(B, dtype=torch.float32, device=pred.device)
    check(lib.mtam_softmax_ce_forward(gemm_mode, pred.data_ptr(), B, D, table.data_ptr(), rows, target.data_ptr(),
                                      lse.data_ptr(), tl.data_ptr(), workspace.data_ptr(), workspace.numel(),
                                      torch.cuda.current_stream(pred.device).cuda_stream), "mtam_softmax_ce_forward")
    return lse, tl


def softmax_ce_backward(pred: torch.Tensor, table: torch.Tensor, target: torch.Tensor, lse: torch.Tensor, inv_batch: float,
                        gemm_mode: int = _lib.GEMM_TF32X3, workspace: Optional[torch.Tensor] = None,
                        dtable: Optional[torch.Tensor] = None):
    """Gradients of mean cross-entropy w.r.t. the rows of `table` (complete) and w.r.t. pred (this table's share), given
    the global log-sum-exp (mtam_softmax_ce_backward)."""
    lib = _lib.load()
    B, D = pred.shape
    rows = table.shape[0]
    if workspace is None:
        workspace = torch.empty(int(lib.mtam_softmax_ce_workspace(B, D, rows)), dtype=torch.uint8, device=pred.device)
    if dtable is None:
        dtable = torch.empty((rows, D), dtype=torch.float32, device=pred.device)
    dpred = torch.empty((B, D), dtype=torch.float32, device=pred.device)
    check(lib.mtam_softmax_ce_backward(gemm_mode, pred.data_ptr(), B, D, table.data_ptr(), rows, target.data_ptr(),
                                       lse.data_ptr(), float(inv_batch), dtable.data_ptr(), dpred.data_ptr(),
                                       workspace.data_ptr(), workspace.numel(),
                                       torch.cuda.current_stream(pred.device).cuda_stream), "mtam_softmax_ce_backward")
    return dtable, dpred


def merge_topk(idx_lists: torch.Tensor, score_lists: torch.Tensor):
    """Merges [n_lists, B, k] per-shard top-k lists (shard order = index order) into [B, k] (mtam_merge_topk)."""
    lib = _lib.load()
    n_lists, B, k = idx_lists.shape
    out_i = torch.empty((B, k), dtype=torch.int32, device=idx_lists.device)
    out_s = torch.empty((B, k), dtype=torch.float32, device=idx_lists.device)
    check(lib.mtam_merge_topk(idx_lists.data_ptr(), score_lists.data_ptr(), n_lists, B, k, out_i.data_ptr(),
                              out_s.data_ptr(), torch.cuda.current_stream(idx_lists.device).cuda_stream), "mtam_merge_topk")
    return out_i, out_s

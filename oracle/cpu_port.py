"""CPU port of one full train step, fp32, all host threads  --  TEST / BASELINE INFRASTRUCTURE ONLY.

This is the oracle's forward graph (oracle/mtam_oracle.py, the restatement of the reference's
TF 1.14 graph) run in torch-CPU float32 with autograd, followed by the TF-style un-deduplicated
global-norm clip and the non-lazy dense Adam, all as in-place torch ops so that it can be timed as
`cpu_baseline` ("kind": "port") and as `bench.py --impl reference`.  It is not TensorFlow 1.14 (which
cannot be installed in this image: SURVEY.md section 8c) and is labelled as a port wherever reported.
Parity unpinned, like the oracle it wraps.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

from . import mtam_oracle as O


class TorchPort:
    def __init__(self, cfg: O.OracleConfig, params: Dict[str, np.ndarray]):
        self.cfg = cfg
        self.p = {k: torch.tensor(v, dtype=torch.float32, requires_grad=not O.is_dead(k)) for k, v in params.items()}
        self.m = {k: torch.zeros_like(v) for k, v in self.p.items()}
        self.v = {k: torch.zeros_like(v) for k, v in self.p.items()}
        self.t = 0

    def train_step(self, feed, lr: float) -> float:
        cfg = self.cfg
        for t in self.p.values():
            t.grad = None
        Ts = self.p["embedding_layer/item"].detach().clone().requires_grad_(True)   # dense piece kept apart
        fwd = O.forward(cfg, self.p, feed, torch.float32, item_table_for_scores=Ts)
        fwd["loss"].backward()
        # un-deduplicated global norm (trap T1): IndexedSlices values + dense pieces
        sq = 0.0
        for k in ("Eu", "Ei", "Ec", "Ep"):
            g = fwd[k].grad
            if g is not None:
                sq += float((g * g).sum())
        sq += float((Ts.grad * Ts.grad).sum())
        for name, t in self.p.items():
            if name not in O.TABLES and t.grad is not None:
                sq += float((t.grad * t.grad).sum())
        scale = cfg.clip / max(math.sqrt(sq), cfg.clip)
        self.t += 1
        lr_t = float(np.float32(lr)) * math.sqrt(1.0 - cfg.beta2 ** self.t) / (1.0 - cfg.beta1 ** self.t)
        with torch.no_grad():
            for name, w in self.p.items():
                g = w.grad
                if name == "embedding_layer/item":
                    g = Ts.grad if g is None else g + Ts.grad
                if g is None:
                    continue
                g = g * scale
                m, v = self.m[name], self.v[name]
                m.mul_(cfg.beta1).add_(g, alpha=1.0 - cfg.beta1)
                v.mul_(cfg.beta2).addcmul_(g, g, value=1.0 - cfg.beta2)
                w.addcdiv_(m, v.sqrt().add_(cfg.eps), value=-lr_t)
        return float(fwd["loss"].detach())

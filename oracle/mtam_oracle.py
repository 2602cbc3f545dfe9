"""CPU oracle for the MTAMRecommender hot path  --  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, the arithmetic of the reference's training / scoring
hot path (TensorFlow 1.14 graph built by /root/reference).  It is the checker the
CUDA path is compared against.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package ``mtamrecommender_b200`` never does.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures, and
TensorFlow 1.14 cannot be installed in this image, so this
restatement could not be checked against outputs of the reference itself.  What pins
it instead: finite-difference gradient checks, hand-computed known-answer tests for
every TF semantic it relies on (tests/test_oracle.py) and the committed vectors in
tests/golden/ that freeze its current behaviour.

All arithmetic is torch-on-CPU in float64 (truth for tolerances) or float32 (the
"port" timed as cpu_baseline).  Citations are file:line under /root/reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

# Model kinds (train_process.py:164-218 dispatch names in comments)
MTAM = "MTAM"                  # 'MTAM'                                 Model/MTAMRec_model.py:61-92
PISTREC = "PISTREC"            # Time_Aware_self_Attention_model        Model/PISTRec_model.py:38-74
SASREC = "SASREC"              # 'SASrec' Self_Attention_Model          Model/attention_baseline_models.py:33-46
TA_SASREC = "TA_SASREC"        # 'Time_Aware_Self_Attention_Model'      Model/attention_baseline_models.py:47-65
TISASREC = "TISASREC"          # 'Ti_Self_Attention_Model'              Model/attention_baseline_models.py:66-84
BPRMF = "BPRMF"                # 'bpr'                                  Model/BPRMF.py:10-59
KINDS = (MTAM, PISTREC, SASREC, TA_SASREC, TISASREC, BPRMF)
# Oracle-only so far (the next widening step): no CUDA path is built for these yet.
MTAM_VIA_T_GRU = "MTAM_VIA_T_GRU"   # 'MTAM_via_T_GRU': the T-GRU outputs are the memory   Model/MTAMRec_model.py:167-204
MTAM_NO_TA_RNN = "MTAM_NO_TIME_AWARE_RNN"   # 'MTAM_no_time_aware_rnn': plain GRU intent encoder      MTAMRec_model.py:93-125
MTAM_VIA_RNN = "MTAM_VIA_RNN"               # 'MTAM_via_rnn': plain GRU, its outputs are the memory      MTAMRec_model.py:206-233
NEXT_KINDS = (MTAM_VIA_T_GRU, MTAM_NO_TA_RNN, MTAM_VIA_RNN)
MTAM_FAMILY = (MTAM, MTAM_VIA_T_GRU, MTAM_NO_TA_RNN, MTAM_VIA_RNN)
PLAIN_GRU_KINDS = (MTAM_NO_TA_RNN, MTAM_VIA_RNN)     # GRU.gru_net: tf GRUCell, no time gate (gru.py:60-67)
MEMORY_IS_RNN_KINDS = (MTAM_VIA_T_GRU, MTAM_VIA_RNN)

MASK_VALUE = float(-2 ** 32 + 1)   # time_aware_attention.py:392 -> fp32 -4294967296.0
LN_EPS_BLOCK = 1e-8                 # Time_Aware_Attention.normalize  time_aware_attention.py:7-34
LN_EPS_FINAL = 1e-12                # tf.contrib.layers.layer_norm    net_utils.py:229-232


@dataclass
class OracleConfig:
    kind: str = MTAM
    L: int = 50            # FLAGS.length_of_user_history == max_length_seq
    D: int = 128           # FLAGS.num_units
    H: int = 1             # FLAGS.num_heads
    N: int = 6             # FLAGS.num_blocks
    user_count: int = 100
    item_count: int = 200
    category_count: int = 20
    reg: float = 5e-5      # FLAGS.regulation_rate
    clip: float = 1.0      # FLAGS.max_gradient_norm
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8
    dropout: float = 0.0   # attention dropout of SASREC/TISASREC; parity runs use 0

    @property
    def U(self): return self.user_count + 3      # Behavior_embedding_time_aware_attention.py:64
    @property
    def V(self): return self.item_count + 3      # :71
    @property
    def C(self): return self.category_count + 3  # :78
    @property
    def P(self): return self.L + 3               # :86


# ----------------------------------------------------------------------------------------------
# parameter inventory.  Names are the library's stable names; shapes are logical.
# ----------------------------------------------------------------------------------------------
GRU_LIVE_VECS = ("_time_kernel_w1", "_time_kernel_b1", "_time_history_w1", "_time_w1",
                 "_time_b1", "_time_kernel_w2", "_time_w12", "_time_b12")
GRU_DEAD_VECS = ("_time_history_b1", "_time_kernel_b2", "_time_history_w2", "_time_history_b2",
                 "_time_w2", "_time_b2")
GATE_LIVE = ("_time_input_w1", "_time_input_b1", "time_output_w1", "time_output_w2", "time_output_b")
GATE_DEAD = ("time_output_w3",)


def param_shapes(cfg: OracleConfig) -> Dict[str, Tuple[int, ...]]:
    D, L, N = cfg.D, cfg.L, cfg.N
    s: Dict[str, Tuple[int, ...]] = {}
    s["embedding_layer/user"] = (cfg.U, D)
    s["embedding_layer/item"] = (cfg.V, D)
    s["embedding_layer/category"] = (cfg.C, D)
    s["embedding_layer/position"] = (cfg.P, D)
    s["position_embedding/dense4emb/kernel"] = (2 * D, D)
    if cfg.kind == BPRMF:
        s["embedding_layer/item_b"] = (cfg.V, 1)       # BPRMF.py:34-35
        return s
    if cfg.kind in MTAM_FAMILY:
        g = "ShortTermIntentEncoder/"
        if cfg.kind in MEMORY_IS_RNN_KINDS:            # layer_norm(short_term_intent) inside this scope, MTAMRec_model.py:186, :220
            s[g + "LayerNorm/beta"] = (D,)
            s[g + "LayerNorm/gamma"] = (D,)
        s[g + "gates/kernel"] = (2 * D, 2 * D)         # time_aware_rnn.py:166-169
        s[g + "gates/bias"] = (2 * D,)
        s[g + "candidate/kernel"] = (2 * D, D)
        s[g + "candidate/bias"] = (D,)
        if cfg.kind not in PLAIN_GRU_KINDS:            # the time gate's vectors exist in the time-aware cell only
            for v in GRU_LIVE_VECS + GRU_DEAD_VECS:
                s[g + v] = (D,)
        scope, att, Tq = "NextItemDecoder/decoder", "vanilla_attention", 1
    else:
        scope, att, Tq = "UserHistoryEncoder/encoder", "self_attention", L
    for i in range(N):
        b = f"{scope}/num_blocks_{i}/"
        for dn in ("dense", "dense_1", "dense_2"):       # Q, K, V  (time_aware_attention.py:249-253)
            s[b + dn + "/kernel"] = (D, D)
            s[b + dn + "/bias"] = (D,)
        if cfg.kind in MTAM_FAMILY + (PISTREC, TA_SASREC):
            s[b + att + "/_time_input_w"] = (D, D)
            for v in GATE_LIVE + GATE_DEAD:
                s[b + att + "/" + v] = (Tq, L)
        s[b + att + "/ln/beta"] = (D,)
        s[b + att + "/ln/gamma"] = (D,)
    top = "NextItemDecoder" if cfg.kind in MTAM_FAMILY else "UserHistoryEncoder"
    s[top + "/LayerNorm/beta"] = (D,)
    s[top + "/LayerNorm/gamma"] = (D,)
    return s


def is_dead(name: str) -> bool:
    leaf = name.rsplit("/", 1)[-1]
    return leaf in GRU_DEAD_VECS or leaf in GATE_DEAD


def init_params(cfg: OracleConfig, seed: int = 1234) -> Dict[str, np.ndarray]:
    """Random init with the TF initialisers the reference ends up with:
    tables U(+-sqrt(6/D)) base_embedding.py:46-60; get_variable default = glorot-uniform;
    dense bias zeros; GRU gate bias 1.0, candidate bias 0; LN gamma 1 / beta 0."""
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = {}
    for name, shp in param_shapes(cfg).items():
        leaf = name.rsplit("/", 1)[-1]
        if name.startswith("embedding_layer/"):
            r = math.sqrt(6.0 / shp[1])
            a = rng.uniform(-r, r, size=shp)
        elif leaf == "gamma":
            a = np.ones(shp)
        elif leaf == "beta":
            a = np.zeros(shp)
        elif leaf == "bias":
            a = np.ones(shp) if name.endswith("gates/bias") else np.zeros(shp)
        else:
            fan_in, fan_out = (shp[0], shp[0]) if len(shp) == 1 else (shp[0], shp[1])
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            a = rng.uniform(-lim, lim, size=shp)
        out[name] = a.astype(np.float32)
    return out


# ----------------------------------------------------------------------------------------------
# records and batches
# ----------------------------------------------------------------------------------------------
def synth_records(cfg: OracleConfig, n: int, seed: int = 1234, zipf_s: float = 1.05,
                  min_len: int = 2) -> List[tuple]:
    """n synthetic 9-tuple records with the construction of Prepare/prepare_data_base.py:252-314
    and Prepare/mask_data_process.py:245-255: history, then the appended mask step."""
    rng = np.random.default_rng(seed)
    L = cfg.L
    recs = []
    # Zipf over [0,item_count): inverse-cdf on a truncated power law
    ranks = np.arange(1, cfg.item_count + 1, dtype=np.float64)
    p = ranks ** (-zipf_s)
    cdf = np.cumsum(p / p.sum())
    for _ in range(n):
        length = int(rng.integers(min_len, L + 1))
        h = length - 1
        items = np.searchsorted(cdf, rng.random(h)).astype(np.int64)
        items = np.minimum(items, cfg.item_count - 1)
        cats = items % max(cfg.category_count, 1)
        gaps = rng.geometric(1.0 / 24.0, size=h + 1).astype(np.float64)
        start = float(rng.integers(300000, 400000))
        times = start + np.cumsum(gaps[:h])
        target_time = float(times[-1] + gaps[h]) if h > 0 else start + gaps[0]
        timelast = np.concatenate([[0.0], np.diff(times)]) if h > 0 else np.zeros(0)
        timenow = target_time - times
        tgt = int(min(np.searchsorted(cdf, rng.random()), cfg.item_count - 1))
        rec = (int(rng.integers(0, cfg.user_count)),
               list(items) + [cfg.item_count + 1],                 # mask token  prepare_data_base.py:283
               list(cats) + [cfg.category_count + 1],              # :285
               list(times) + [target_time],                        # :292
               list(timelast) + [0.0],                             # :293
               list(timenow) + [0.0],                              # :294
               list(range(length)),                                # :295-298 / mask_data_process.py:245
               [tgt, tgt % max(cfg.category_count, 1), target_time],
               length)
        recs.append(rec)
    return recs


def make_feed(cfg: OracleConfig, batch_data: List[tuple]) -> Dict[str, np.ndarray]:
    """make_feed_dic_new  (Embedding/Behavior_embedding_time_aware_attention.py:146-192):
    right-pad each list with 0 to position_count = L."""
    B, L = len(batch_data), cfg.L
    f = {"user_id": np.zeros(B, np.int32), "item_list": np.zeros((B, L), np.int32),
         "category_list": np.zeros((B, L), np.int32), "position_list": np.zeros((B, L), np.int32),
         "time_list": np.zeros((B, L), np.float32), "timelast_list": np.zeros((B, L), np.float32),
         "timenow_list": np.zeros((B, L), np.float32), "target_item_id": np.zeros(B, np.int32),
         "target_item_category": np.zeros(B, np.int32), "target_item_time": np.zeros(B, np.float32),
         "seq_length": np.zeros(B, np.int32)}
    for b, ex in enumerate(batch_data):
        n = int(ex[8])
        f["user_id"][b] = ex[0]
        f["item_list"][b, :n] = ex[1]
        f["category_list"][b, :n] = ex[2]
        f["time_list"][b, :n] = ex[3]
        f["timelast_list"][b, :n] = ex[4]
        f["timenow_list"][b, :n] = ex[5]
        f["position_list"][b, :n] = ex[6]
        f["target_item_id"][b] = ex[7][0]
        f["target_item_category"][b] = ex[7][1]
        f["target_item_time"][b] = ex[7][2]
        f["seq_length"][b] = n
    return f


def synth_batch(cfg: OracleConfig, B: int, seed: int = 1234, **kw) -> Dict[str, np.ndarray]:
    return make_feed(cfg, synth_records(cfg, B, seed, **kw))


# ----------------------------------------------------------------------------------------------
# forward graph
# ----------------------------------------------------------------------------------------------
def _ln(x, gamma, beta, eps):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)              # tf.nn.moments: biased
    return gamma * (x - mu) / torch.sqrt(var + eps) + beta


def _split_heads(x, H):
    # tf.concat(tf.split(x, H, axis=2), axis=0)  -> here kept as [B,H,T,dh]
    B, T, D = x.shape
    return x.view(B, T, H, D // H).permute(0, 2, 1, 3)


def attention_block(kind, q, e, tq, tk, key_len, query_len, p, prefix, att, H, drop_mask=None):
    """One block.  kind in {'time_aware','tisas','plain'}.
    time_aware: time_aware_attention.py:215-456; tisas: :73-214; plain: multihead_attention.py:71-193."""
    B, Tq, D = q.shape
    L = e.shape[1]
    dh = D // H
    Q = torch.relu(q @ p[prefix + "dense/kernel"] + p[prefix + "dense/bias"])
    K = torch.relu(e @ p[prefix + "dense_1/kernel"] + p[prefix + "dense_1/bias"])
    V = torch.relu(e @ p[prefix + "dense_2/kernel"] + p[prefix + "dense_2/bias"])
    S = _split_heads(Q, H) @ _split_heads(K, H).transpose(-1, -2)          # [B,H,Tq,L]
    if kind != "plain":
        delta = torch.log(torch.abs(tq[:, :, None] - tk[:, None, :]) + 1.0)  # [B,Tq,L]  :339
    if kind == "time_aware":
        a = prefix + att + "/"
        Z = torch.tanh((q @ p[a + "_time_input_w"]) @ e.transpose(-1, -2))   # :320-323 (raw q, e)
        Dk = torch.tanh(delta * p[a + "_time_input_w1"] + p[a + "_time_input_b1"])   # :343
        G = p[a + "time_output_w1"] * Dk + p[a + "time_output_w2"] * Z + p[a + "time_output_b"]  # :350
        S = S * torch.sigmoid(G)[:, None]                                     # :381
    elif kind == "tisas":
        S = S + delta[:, None]                                                # :139
    S = S / (dh ** 0.5)                                                       # :384
    kmask = (torch.arange(L)[None, :] < key_len[:, None])                     # tf.sequence_mask
    S = torch.where(kmask[:, None, None, :], S, torch.full_like(S, MASK_VALUE))   # :388-394
    P = torch.softmax(S, dim=-1)
    qmask = (torch.arange(Tq)[None, :] < query_len[:, None]).to(P.dtype)      # :428-431
    P = P * qmask[:, None, :, None]
    if drop_mask is not None:                                                 # tisas/plain dropout (injected mask)
        P = P * drop_mask
    O = P @ _split_heads(V, H)                                                # [B,H,Tq,dh]
    O = O.permute(0, 2, 1, 3).reshape(B, Tq, D)
    y = O + q                                                                 # residual = raw queries :451
    a = prefix + att + "/"
    return _ln(y, p[a + "ln/gamma"], p[a + "ln/beta"], LN_EPS_BLOCK)


def tgru_new(X, timelast, seq_len, p, g, plain=False):
    """dynamic_rnn(TimeAwareGRUCell_decay_new, sequence_length=seq_len-1)
    Model/Modules/time_aware_rnn.py:186-269, gru.py:69-77.  Returns outputs [B,L,D] (zeros past length).
    plain=True: GRU.gru_net (gru.py:60-67) -- tf.nn.rnn_cell.GRUCell, the same gates without the time gate T."""
    B, L, D = X.shape
    h = torch.zeros(B, D, dtype=X.dtype)
    Wg, bg = p[g + "gates/kernel"], p[g + "gates/bias"]
    Wc, bc = p[g + "candidate/kernel"], p[g + "candidate/bias"]
    if not plain:
        kw1, kb1, hw1 = p[g + "_time_kernel_w1"], p[g + "_time_kernel_b1"], p[g + "_time_history_w1"]
        tw1, tb1 = p[g + "_time_w1"], p[g + "_time_b1"]
        kw2, tw12, tb12 = p[g + "_time_kernel_w2"], p[g + "_time_w12"], p[g + "_time_b12"]
    outs = []
    n = seq_len - 1
    for t in range(L):
        x = X[:, t]
        dlt = timelast[:, t:t + 1]
        if plain:
            T = 1.0
        else:
            a = torch.relu(x * kw1 + kb1 + h * hw1)                   # :228
            s = torch.relu(tw1 * dlt + tb1)                           # :236
            T = torch.sigmoid(kw2 * a + tw12 * s + tb12)              # :237
        ru = torch.sigmoid(torch.cat([x, h], 1) @ Wg + bg)            # :243-247
        r, u = ru[:, :D], ru[:, D:]                                   # :248
        c = torch.tanh(torch.cat([x, r * h], 1) @ Wc + bc)            # :250-256
        hn = u * h + (1 - u) * c * T                                  # :268
        live = (t < n)[:, None]
        outs.append(torch.where(live, hn, torch.zeros_like(hn)))      # zero output past length
        h = torch.where(live, hn, h)                                  # state copy-through
    return torch.stack(outs, 1)


def forward(cfg: OracleConfig, params: Dict[str, torch.Tensor], feed: Dict[str, np.ndarray],
            dtype=torch.float64, item_table_for_scores: Optional[torch.Tensor] = None,
            bpr_negative: Optional[int] = None, drop_masks=None):
    """Returns dict with loss, loss_origin[B], pred[B,D], l2_norm and the gathered rows
    (leaf-like tensors whose .grad are the IndexedSlices values of tf.gradients)."""
    p = params
    D, L, H, N = cfg.D, cfg.L, cfg.H, cfg.N
    ids = {k: torch.from_numpy(np.ascontiguousarray(v)).long() for k, v in feed.items()
           if v.dtype == np.int32}
    fl = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dtype) for k, v in feed.items()
          if v.dtype == np.float32}
    Tu, Ti, Tc, Tp = (p["embedding_layer/user"], p["embedding_layer/item"],
                      p["embedding_layer/category"], p["embedding_layer/position"])
    Eu = Tu[ids["user_id"]]                                            # Behavior_...py:68
    Ei = Ti[ids["item_list"]]                                          # :75
    Ec = Tc[ids["category_list"]]                                      # :82
    Ep = Tp[ids["position_list"]]                                      # :90
    for t in (Eu, Ei, Ec, Ep):
        if t.requires_grad:
            t.retain_grad()
    out = {"Eu": Eu, "Ei": Ei, "Ec": Ec, "Ep": Ep}
    seq_len = ids["seq_length"]
    B = seq_len.shape[0]
    Ts = Ti if item_table_for_scores is None else item_table_for_scores

    if cfg.kind == BPRMF:                                              # BPRMF.py:41-59
        neg = int(bpr_negative if bpr_negative is not None else 0)
        pos_id = ids["target_item_id"]
        neg_id = torch.tensor([neg])
        ip, ineg = Ti[pos_id], Ti[neg_id]
        for t in (ip, ineg):
            if t.requires_grad:
                t.retain_grad()
        ib = p["embedding_layer/item_b"]
        bp, bn = ib[pos_id], ib[neg_id]
        for t in (bp, bn):
            if t.requires_grad:
                t.retain_grad()
        out.update(bpos=bp, bneg=bn)
        x = bp - bn + (Eu * (ip - ineg)).sum(1)                        # [B,1]+[B] broadcast -> [B,B] in TF
        l2 = 0.5 * (Eu ** 2).sum() + 0.5 * (ip ** 2).sum() + 0.5 * (ineg ** 2).sum()
        loss = 5e-5 * l2 - torch.log(torch.sigmoid(x)).mean()
        out.update(loss=loss, pred=Eu, l2_norm=l2, loss_origin=None, ipos=ip, ineg=ineg)
        return out

    X = torch.relu(torch.cat([Ei, Ec], 2) @ p["position_embedding/dense4emb/kernel"]) + Ep   # :95-103
    out["X"] = X
    if cfg.kind in MTAM_FAMILY:
        g = "ShortTermIntentEncoder/"
        rnn = tgru_new(X, fl["timelast_list"], seq_len, p, g, plain=cfg.kind in PLAIN_GRU_KINDS)
        out["rnn"] = rnn
        pos = (seq_len - 2).clamp(min=0)                               # mask_index-1, MTAMRec_model.py:75-79
        q = rnn[torch.arange(B), pos][:, None, :]                      # gather_indexes net_utils.py:82-92
        memory = X                                                     # MTAM: user_history = the embedded behaviours (:65)
        if cfg.kind in MEMORY_IS_RNN_KINDS:
            # the memory is the (T-)GRU's output sequence (zeros from step seq_len-1 on, dynamic_rnn), keys still masked
            # by seq_length; the query is layer-normed first                       MTAMRec_model.py:180-189
            memory = rnn
            q = _ln(q[:, 0], p[g + "LayerNorm/gamma"], p[g + "LayerNorm/beta"], LN_EPS_FINAL)[:, None, :]
        out["short_term_intent"] = q[:, 0]
        tq = fl["target_item_time"][:, None]
        for i in range(N):
            q = attention_block("time_aware", q, memory, tq, fl["time_list"], seq_len,
                                torch.ones_like(seq_len), p,
                                f"NextItemDecoder/decoder/num_blocks_{i}/", "vanilla_attention", H)
        hyb = q.reshape(B, D)
        pred = _ln(hyb, p["NextItemDecoder/LayerNorm/gamma"], p["NextItemDecoder/LayerNorm/beta"], LN_EPS_FINAL)
    else:
        akind = {PISTREC: "time_aware", TA_SASREC: "time_aware", TISASREC: "tisas", SASREC: "plain"}[cfg.kind]
        enc = X
        for i in range(N):
            dm = None if drop_masks is None else drop_masks[i]
            enc = attention_block(akind, enc, enc, fl["time_list"], fl["time_list"], seq_len, seq_len, p,
                                  f"UserHistoryEncoder/encoder/num_blocks_{i}/", "self_attention", H, dm)
        out["enc"] = enc
        pos = (seq_len - 1).clamp(min=0)                               # mask_index
        ltp = enc[torch.arange(B), pos]
        pred = _ln(ltp, p["UserHistoryEncoder/LayerNorm/gamma"], p["UserHistoryEncoder/LayerNorm/beta"], LN_EPS_FINAL)

    logits = pred @ Ts.t()                                             # base_model.py:316
    logp = torch.log_softmax(logits, dim=-1)
    loss_origin = -logp[torch.arange(B), ids["target_item_id"]]        # :317-321
    l2 = 0.5 * (Ei ** 2).sum() + 0.5 * (Ec ** 2).sum() + 0.5 * (Ep ** 2).sum()
    if cfg.kind != PISTREC:                                            # PISTRec_model.py:56-60 omits user term
        l2 = l2 + 0.5 * (Eu ** 2).sum()
    loss = cfg.reg * l2 + loss_origin.mean()                           # :322
    out.update(loss=loss, loss_origin=loss_origin, pred=pred, l2_norm=l2, logits=logits)
    return out


# ----------------------------------------------------------------------------------------------
# gradients in tf.gradients form, clip (T1), Adam (T2)
# ----------------------------------------------------------------------------------------------
TABLES = {"embedding_layer/user": ("Eu", "user_id"), "embedding_layer/item": ("Ei", "item_list"),
          "embedding_layer/category": ("Ec", "category_list"), "embedding_layer/position": ("Ep", "position_list")}


def loss_and_grads(cfg: OracleConfig, params_np: Dict[str, np.ndarray], feed, dtype=torch.float64,
                   bpr_negative=None, drop_masks=None):
    """Returns (fwd dict, dense grads {name: ndarray or None}, pieces) where pieces is the list of
    gradient tensors exactly as tf.gradients would hand them to clip_by_global_norm: dense arrays and,
    for tables, un-deduplicated IndexedSlices values (trap T1)."""
    p = {k: torch.tensor(v, dtype=dtype, requires_grad=True) for k, v in params_np.items()}
    # separate leaf for the scoring use of the item table so the dense and sparse parts stay apart
    Ts = torch.tensor(params_np["embedding_layer/item"], dtype=dtype, requires_grad=True)
    fwd = forward(cfg, p, feed, dtype, item_table_for_scores=Ts, bpr_negative=bpr_negative,
                  drop_masks=drop_masks)
    fwd["loss"].backward()
    grads: Dict[str, Optional[np.ndarray]] = {}
    pieces: List[np.ndarray] = []
    for name, t in p.items():
        if name in TABLES:
            rows_key, idx_key = TABLES[name]
            rows = fwd[rows_key]
            vals = rows.grad
            dense = np.zeros(t.shape, dtype=np.float64)
            has = False
            if vals is not None:
                idx = np.asarray(feed[idx_key]).reshape(-1)
                v = vals.detach().numpy().reshape(-1, t.shape[1]).astype(np.float64)
                np.add.at(dense, idx, v)
                pieces.append(v)
                has = True
            if name == "embedding_layer/item":
                if cfg.kind == BPRMF:
                    for rk, ids_ in (("ipos", feed["target_item_id"]), ("ineg", np.array([bpr_negative or 0]))):
                        v = fwd[rk].grad.detach().numpy().astype(np.float64)
                        np.add.at(dense, np.asarray(ids_).reshape(-1), v)
                        pieces.append(v)
                        has = True
                elif Ts.grad is not None:
                    d = Ts.grad.detach().numpy().astype(np.float64)
                    dense += d
                    # TF concatenates dense-as-slices with the sparse slices: one IndexedSlices tensor,
                    # its squared norm is the sum of both parts' squared norms.
                    pieces.append(d)
                    has = True
            grads[name] = dense if has else None
        else:
            if t.grad is None or is_dead(name):
                grads[name] = None
            else:
                g = t.grad.detach().numpy().astype(np.float64)
                grads[name] = g
                if cfg.kind == BPRMF and name == "embedding_layer/item_b":
                    # two embedding_lookups on item_b -> IndexedSlices: un-deduplicated values in the norm
                    pieces.append(fwd["bpos"].grad.detach().numpy().astype(np.float64))
                    pieces.append(fwd["bneg"].grad.detach().numpy().astype(np.float64))
                else:
                    pieces.append(g)
    if cfg.kind == BPRMF:   # dense4emb etc. are unused in BPRMF: tf.gradients gives None
        for name in list(grads):
            if grads[name] is not None and not np.any(grads[name]) and name not in TABLES \
                    and name != "embedding_layer/item_b":
                grads[name] = None
    return fwd, grads, pieces


def global_norm(pieces: List[np.ndarray]) -> float:
    """tf.clip_by_global_norm's norm: sqrt(sum over tensors of sum(values**2)); IndexedSlices use
    their un-deduplicated values (trap T1)."""
    return math.sqrt(sum(float((x.astype(np.float64) ** 2).sum()) for x in pieces))


def clip_scale(gn: float, clip: float) -> float:
    return clip / max(gn, clip)          # clip_norm * min(1/norm, 1/clip_norm)


def adam_tf(w, g, m, v, lr, t, beta1=0.9, beta2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer dense update, epsilon OUTSIDE the bias correction (trap T2).
    Sparse apply in TF 1.14 is non-lazy: identical to this with g=0 on untouched rows."""
    lr_t = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    m = beta1 * m + (1.0 - beta1) * g
    v = beta2 * v + (1.0 - beta2) * g * g
    w = w - lr_t * m / (np.sqrt(v) + eps)
    return w, m, v


class OracleTrainer:
    """Stateful trainer: params + Adam slots, `train_step` = base_model.train (base_model.py:150-167)."""

    def __init__(self, cfg: OracleConfig, params: Dict[str, np.ndarray], dtype=torch.float64):
        self.cfg, self.dtype = cfg, dtype
        np_dt = np.float64 if dtype == torch.float64 else np.float32
        self.params = {k: v.astype(np_dt) for k, v in params.items()}
        self.m = {k: np.zeros_like(v) for k, v in self.params.items()}
        self.v = {k: np.zeros_like(v) for k, v in self.params.items()}
        self.t = 0
        self.last = None

    def train_step(self, feed, lr: float, bpr_negative=None):
        cfg = self.cfg
        fwd, grads, pieces = loss_and_grads(cfg, self.params, feed, self.dtype, bpr_negative)
        gn = global_norm(pieces)
        sc = clip_scale(gn, cfg.clip)
        self.t += 1
        lr32 = float(np.float32(lr))                       # float64 placeholder cast to fp32 in Adam
        for k, g in grads.items():
            if g is None:
                continue
            w, m, v = adam_tf(self.params[k].astype(np.float64), g * sc, self.m[k].astype(np.float64),
                              self.v[k].astype(np.float64), lr32, self.t, cfg.beta1, cfg.beta2, cfg.eps)
            self.params[k] = w.astype(self.params[k].dtype)
            self.m[k] = m.astype(self.params[k].dtype)
            self.v[k] = v.astype(self.params[k].dtype)
        self.last = dict(loss=float(fwd["loss"].detach()), global_norm=gn, scale=sc, grads=grads)
        return float(fwd["loss"].detach())


# ----------------------------------------------------------------------------------------------
# eval: top-k (tf.nn.top_k: descending, ties -> lower index) and HR/NDCG (base_model.py:188-242)
# ----------------------------------------------------------------------------------------------
def topk_indices(scores: np.ndarray, k: int) -> np.ndarray:
    V = scores.shape[1]
    order = np.lexsort((np.broadcast_to(np.arange(V), scores.shape), -scores), axis=1)
    return order[:, :k].astype(np.int32)


def hr_ndcg(topk: np.ndarray, target: np.ndarray, k: int) -> Tuple[float, float]:
    """calculate_topK base_model.py:215-242."""
    B = len(target)
    hit = 0
    nd = 0.0
    for b in range(B):
        row = topk[b, :k]
        w = np.nonzero(row == target[b])[0]
        if len(w):
            hit += 1
            nd += math.log(2) / math.log(int(w[0]) + 2)
    return hit / B, nd / B


def metrics_topk(cfg, params_np, feed, dtype=torch.float64):
    p = {k: torch.tensor(v, dtype=dtype) for k, v in params_np.items()}
    with torch.no_grad():
        fwd = forward(cfg, p, feed, dtype)
        scores = (fwd["pred"] @ p["embedding_layer/item"].t()).numpy()
    idx = topk_indices(scores, 50)
    res = []
    for k in (1, 5, 10, 30, 50):
        res.extend(hr_ndcg(idx, feed["target_item_id"], k))
    return tuple(res), idx, scores


def lr_schedule(flags_lr: float, decay_rate: float, global_step: int, current_lr: float) -> float:
    """train_process.py:154-159, 333-336."""
    if current_lr > 0.001:
        return flags_lr * 0.99 ** (global_step // 100)
    return 0.001 * decay_rate ** (global_step // 100)


# ----------------------------------------------------------------------------------------------
# plain gather / scatter-add restatements for the two graded bandwidth kernels
# ----------------------------------------------------------------------------------------------
def gather_rows(table: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """tf.nn.embedding_lookup (Behavior_...py:68-90)."""
    return table[idx.reshape(-1)]


def scatter_add_rows(n_rows: int, idx: np.ndarray, rows: np.ndarray, dtype=np.float32) -> np.ndarray:
    """Sum of IndexedSlices rows per index, accumulated in ascending token order (the order the
    CUDA sort-then-segmented-reduce also uses)."""
    out = np.zeros((n_rows, rows.shape[1]), dtype=dtype)
    np.add.at(out, idx.reshape(-1), rows.astype(dtype))
    return out

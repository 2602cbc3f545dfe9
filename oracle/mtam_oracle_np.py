"""Second, independent restatement of the MTAM hot path -- TEST INFRASTRUCTURE ONLY (see oracle/mtam_oracle.py).

Why a second one: the reference is TensorFlow 1.14 and cannot run here, so the oracle (`mtam_oracle.py`: torch fp64
forward + torch autograd) is *unpinned* -- nothing reference-held confirms it.  This file reads the same reference
files again, on its own, and shares no code with the first restatement:
  * NumPy float64 only, no torch, no autograd: the backward pass is DERIVED BY HAND and written out, which
    also checks the derivations the CUDA backward kernels implement;
  * clip (trap T1) and Adam (trap T2) are restated again from the TF 1.14 op definitions.
`tests/test_oracle.py` requires the two to agree to fp64 rounding on loss, pred, every gradient, the global norm and
the weights after Adam steps.  Agreement of two independent readings is not a pin, but it removes transcription slips.

Reference lines followed (under /root/reference):
  Embedding/Behavior_embedding_time_aware_attention.py:62-114   four lookups, relu([Ei|Ec] W) + Ep
  Model/MTAMRec_model.py:61-92                                  T-GRU on [X | timelast | timenow], length seq_len-1,
                                                                gather at mask_index-1, N hops, contrib layer_norm
  Model/Modules/time_aware_rnn.py:186-269                       TimeAwareGRUCell_decay_new.call
  Model/Modules/time_aware_attention.py:7-34, 215-456, 524-556  normalize (eps 1e-8), one hop, vanilla_attention
  Model/base_model.py:290-328                                   loss, tf.gradients -> clip_by_global_norm -> Adam
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np

MASK = float(-2 ** 32 + 1)          # time_aware_attention.py:391


def _sig(x):
    return 1.0 / (1.0 + np.exp(-x))


def _ln_fwd(y, g, b, eps):
    mu = y.mean(-1, keepdims=True)
    var = ((y - mu) ** 2).mean(-1, keepdims=True)          # tf.nn.moments: biased
    rstd = 1.0 / np.sqrt(var + eps)
    xh = (y - mu) * rstd
    return g * xh + b, xh, rstd


def _ln_bwd(dout, xh, rstd, g):
    dxh = dout * g
    return (dxh - dxh.mean(-1, keepdims=True) - xh * (dxh * xh).mean(-1, keepdims=True)) * rstd


def hop_names(i):
    p = f"NextItemDecoder/decoder/num_blocks_{i}/"
    a = p + "vanilla_attention/"
    return dict(Wq=p + "dense/kernel", bq=p + "dense/bias", Wk=p + "dense_1/kernel", bk=p + "dense_1/bias",
                Wv=p + "dense_2/kernel", bv=p + "dense_2/bias", Wt=a + "_time_input_w", w1=a + "_time_input_w1",
                b1=a + "_time_input_b1", o1=a + "time_output_w1", o2=a + "time_output_w2", ob=a + "time_output_b",
                lnb=a + "ln/beta", lng=a + "ln/gamma")


G_ = "ShortTermIntentEncoder/"


def forward_backward(P: Dict[str, np.ndarray], feed: Dict[str, np.ndarray], L: int, D: int, H: int, N: int, reg: float):
    """Returns (out, grads, pieces): out = dict(loss, loss_origin, pred, l2_norm); grads = dense gradient per parameter
    name (tables: duplicates summed); pieces = the tensors as tf.gradients hands them to clip_by_global_norm
    (IndexedSlices values un-deduplicated, SURVEY 9.6 / trap T1)."""
    p = {k: np.asarray(v, np.float64) for k, v in P.items()}
    item, cat, pos = (feed[k].astype(np.int64) for k in ("item_list", "category_list", "position_list"))
    user, tgt, slen = feed["user_id"].astype(np.int64), feed["target_item_id"].astype(np.int64), feed["seq_length"].astype(np.int64)
    tk, tlast, tq = feed["time_list"].astype(np.float64), feed["timelast_list"].astype(np.float64), feed["target_item_time"].astype(np.float64)
    B = user.shape[0]
    Ti, Tc, Tp, Tu = (p["embedding_layer/" + t] for t in ("item", "category", "position", "user"))
    dh_ = D // H
    # ---- embedding (Behavior_...py:62-114) ----
    Ei, Ec, Ep, Eu = Ti[item], Tc[cat], Tp[pos], Tu[user]
    E2 = np.concatenate([Ei, Ec], -1)
    We = p["position_embedding/dense4emb/kernel"]
    pre = E2 @ We
    X = np.maximum(pre, 0.0) + Ep
    # ---- T-GRU 'new' (time_aware_rnn.py:228-268; dynamic_rnn with sequence_length = seq_len - 1) ----
    Wg, bg = p[G_ + "gates/kernel"], p[G_ + "gates/bias"]
    Wc, bc = p[G_ + "candidate/kernel"], p[G_ + "candidate/bias"]
    kw1, kb1, hw1, tw1, tb1, kw2, tw12, tb12 = (p[G_ + n] for n in ("_time_kernel_w1", "_time_kernel_b1", "_time_history_w1",
                                                                    "_time_w1", "_time_b1", "_time_kernel_w2", "_time_w12", "_time_b12"))
    h = np.zeros((B, D))
    tape = []
    q0 = np.zeros((B, D))
    for t in range(L):
        x, dl = X[:, t], tlast[:, t, None]
        a = np.maximum(x * kw1 + kb1 + h * hw1, 0.0)
        s = np.maximum(tw1 * dl + tb1, 0.0)
        Tt = _sig(kw2 * a + tw12 * s + tb12)
        gv = _sig(np.concatenate([x, h], 1) @ Wg + bg)
        r, u = gv[:, :D], gv[:, D:]
        c = np.tanh(np.concatenate([x, r * h], 1) @ Wc + bc)
        hn = u * h + (1.0 - u) * c * Tt
        m = (t < slen - 1)[:, None]
        tape.append((h, a, s, Tt, r, u, c, m))
        h = np.where(m, hn, h)
        q0 = np.where((t == slen - 2)[:, None], h, q0)       # gather_indexes at mask_index - 1 = seq_len - 2
    # ---- N hops (time_aware_attention.py:249-454, Tq = 1) ----
    kmask = np.arange(L)[None, :] < slen[:, None]
    dlt = np.log(np.abs(tq[:, None] - tk) + 1.0)
    q = q0
    hops = []
    for i in range(N):
        n = hop_names(i)
        Q = np.maximum(q @ p[n["Wq"]] + p[n["bq"]], 0.0)
        K = np.maximum(X @ p[n["Wk"]] + p[n["bk"]], 0.0)
        V = np.maximum(X @ p[n["Wv"]] + p[n["bv"]], 0.0)
        qt = q @ p[n["Wt"]]
        Z = np.tanh(np.einsum("bd,bld->bl", qt, X))
        Dk = np.tanh(dlt * p[n["w1"]].reshape(1, L) + p[n["b1"]].reshape(1, L))
        gate = _sig(p[n["o1"]].reshape(1, L) * Dk + p[n["o2"]].reshape(1, L) * Z + p[n["ob"]].reshape(1, L))
        O = np.zeros((B, D))
        As, Ps = [], []
        for hd in range(H):
            sl = slice(hd * dh_, (hd + 1) * dh_)
            A = np.einsum("bd,bld->bl", Q[:, sl], K[:, :, sl])
            S = A * gate / math.sqrt(dh_)
            S = np.where(kmask, S, MASK)
            S = S - S.max(-1, keepdims=True)
            Pw = np.exp(S)
            Pw /= Pw.sum(-1, keepdims=True)
            O[:, sl] = np.einsum("bl,bld->bd", Pw, V[:, :, sl])
            As.append(A); Ps.append(Pw)
        y = O + q
        out, xh, rstd = _ln_fwd(y, p[n["lng"]], p[n["lnb"]], 1e-8)
        hops.append((q, Q, K, V, qt, Z, Dk, gate, As, Ps, xh, rstd))
        q = out
    gf, bf = p["NextItemDecoder/LayerNorm/gamma"], p["NextItemDecoder/LayerNorm/beta"]
    pred, xhf, rstdf = _ln_fwd(q, gf, bf, 1e-12)
    # ---- loss (base_model.py:302-322) ----
    logits = pred @ Ti.T
    mx = logits.max(-1, keepdims=True)
    lse = mx[:, 0] + np.log(np.exp(logits - mx).sum(-1))
    loss_origin = lse - logits[np.arange(B), tgt]
    l2 = 0.5 * ((Ei ** 2).sum() + (Ec ** 2).sum() + (Ep ** 2).sum() + (Eu ** 2).sum())
    loss = reg * l2 + loss_origin.mean()
    # =================================== backward, by hand ===================================
    g: Dict[str, np.ndarray] = {k: np.zeros_like(v) for k, v in p.items()}
    dlog = np.exp(logits - lse[:, None])
    dlog[np.arange(B), tgt] -= 1.0
    dlog /= B
    dTi_dense = dlog.T @ pred
    dpred = dlog @ Ti
    g["NextItemDecoder/LayerNorm/gamma"] = (dpred * xhf).sum(0)
    g["NextItemDecoder/LayerNorm/beta"] = dpred.sum(0)
    dq = _ln_bwd(dpred, xhf, rstdf, gf)
    dX = np.zeros((B, L, D))
    for i in reversed(range(N)):
        n = hop_names(i)
        q_in, Q, K, V, qt, Z, Dk, gate, As, Ps, xh, rstd = hops[i]
        g[n["lng"]] = (dq * xh).sum(0)
        g[n["lnb"]] = dq.sum(0)
        dy = _ln_bwd(dq, xh, rstd, p[n["lng"]])
        dq_in = dy.copy()                                  # residual: the raw query
        dQ, dK, dV, dgate = np.zeros_like(Q), np.zeros_like(K), np.zeros_like(V), np.zeros((B, L))
        for hd in range(H):
            sl = slice(hd * dh_, (hd + 1) * dh_)
            Pw, A = Ps[hd], As[hd]
            dP = np.einsum("bd,bld->bl", dy[:, sl], V[:, :, sl])
            dV[:, :, sl] = Pw[:, :, None] * dy[:, None, sl]
            dS = Pw * (dP - (Pw * dP).sum(-1, keepdims=True))
            dS = np.where(kmask, dS, 0.0)
            dA = dS * gate / math.sqrt(dh_)
            dgate += dS * A / math.sqrt(dh_)
            dQ[:, sl] = np.einsum("bl,bld->bd", dA, K[:, :, sl])
            dK[:, :, sl] = dA[:, :, None] * Q[:, None, sl]
        dG = dgate * gate * (1.0 - gate)
        g[n["o1"]] = (dG * Dk).sum(0).reshape(1, L)
        g[n["o2"]] = (dG * Z).sum(0).reshape(1, L)
        g[n["ob"]] = dG.sum(0).reshape(1, L)
        dpre1 = dG * p[n["o1"]].reshape(1, L) * (1.0 - Dk ** 2)
        g[n["w1"]] = (dpre1 * dlt).sum(0).reshape(1, L)
        g[n["b1"]] = dpre1.sum(0).reshape(1, L)
        dM = dG * p[n["o2"]].reshape(1, L) * (1.0 - Z ** 2)
        dqt = np.einsum("bl,bld->bd", dM, X)
        dX += dM[:, :, None] * qt[:, None, :]
        g[n["Wt"]] = q_in.T @ dqt
        dq_in += dqt @ p[n["Wt"]].T
        dQp = dQ * (Q > 0)
        g[n["Wq"]] = q_in.T @ dQp
        g[n["bq"]] = dQp.sum(0)
        dq_in += dQp @ p[n["Wq"]].T
        dKp, dVp = dK * (K > 0), dV * (V > 0)
        Xf = X.reshape(B * L, D)
        g[n["Wk"]] = Xf.T @ dKp.reshape(B * L, D)
        g[n["bk"]] = dKp.sum((0, 1))
        g[n["Wv"]] = Xf.T @ dVp.reshape(B * L, D)
        g[n["bv"]] = dVp.sum((0, 1))
        dX += dKp @ p[n["Wk"]].T + dVp @ p[n["Wv"]].T
        dq = dq_in
    # ---- T-GRU, reverse time ----
    dh = np.zeros((B, D))
    for t in reversed(range(L)):
        h, a, s, Tt, r, u, c, m = tape[t]
        dh = dh + np.where((t == slen - 2)[:, None], dq, 0.0)     # the hops' query is the state after step seq_len-2
        x, dl = X[:, t], tlast[:, t, None]
        d = np.where(m, dh, 0.0)                                 # only unmasked steps ran the cell
        du, dc, dT, dhp = d * (h - c * Tt), d * (1.0 - u) * Tt, d * (1.0 - u) * c, d * u
        dpc = dc * (1.0 - c ** 2)
        xc = np.concatenate([x, r * h], 1)
        g[G_ + "candidate/kernel"] += xc.T @ dpc
        g[G_ + "candidate/bias"] += dpc.sum(0)
        dxc = dpc @ Wc.T
        dx, drh = dxc[:, :D], dxc[:, D:]
        dr = drh * h
        dhp = dhp + drh * r
        dpg = np.concatenate([dr * r * (1.0 - r), du * u * (1.0 - u)], 1)
        g[G_ + "gates/kernel"] += np.concatenate([x, h], 1).T @ dpg
        g[G_ + "gates/bias"] += dpg.sum(0)
        dxg = dpg @ Wg.T
        dx = dx + dxg[:, :D]
        dhp = dhp + dxg[:, D:]
        dpT = dT * Tt * (1.0 - Tt)
        g[G_ + "_time_kernel_w2"] += (dpT * a).sum(0)
        g[G_ + "_time_w12"] += (dpT * s).sum(0)
        g[G_ + "_time_b12"] += dpT.sum(0)
        dpa = dpT * kw2 * (a > 0)
        g[G_ + "_time_kernel_w1"] += (dpa * x).sum(0)
        g[G_ + "_time_kernel_b1"] += dpa.sum(0)
        g[G_ + "_time_history_w1"] += (dpa * h).sum(0)
        dx = dx + dpa * kw1
        dhp = dhp + dpa * hw1
        dps = dpT * tw12 * (s > 0)
        g[G_ + "_time_w1"] += (dps * dl).sum(0)
        g[G_ + "_time_b1"] += dps.sum(0)
        dX[:, t] += dx
        dh = np.where(m, dhp, dh)                                # masked steps copy the state through
    # ---- embedding ----
    dpre = dX * (pre > 0)
    g["position_embedding/dense4emb/kernel"] = E2.reshape(B * L, 2 * D).T @ dpre.reshape(B * L, D)
    dE2 = dpre @ We.T + reg * E2
    dEi, dEc = dE2[..., :D].reshape(B * L, D), dE2[..., D:].reshape(B * L, D)
    dEp = (dX + reg * Ep).reshape(B * L, D)
    dEu = reg * Eu
    pieces: List[np.ndarray] = [dEu, dEi, dTi_dense, dEc, dEp]
    for name, (idx, vals) in {"embedding_layer/user": (user, dEu), "embedding_layer/category": (cat.reshape(-1), dEc),
                              "embedding_layer/position": (pos.reshape(-1), dEp)}.items():
        np.add.at(g[name], idx, vals)
    g["embedding_layer/item"] = dTi_dense.copy()
    np.add.at(g["embedding_layer/item"], item.reshape(-1), dEi)
    live = [k for k in g if not k.startswith("embedding_layer/") and not is_dead(k)]
    pieces += [g[k] for k in live]
    out = dict(loss=float(loss), loss_origin=loss_origin, pred=pred, l2_norm=float(l2))
    return out, {k: (None if is_dead(k) else v) for k, v in g.items()}, pieces


DEAD_LEAVES = ("_time_history_b1", "_time_kernel_b2", "_time_history_w2", "_time_history_b2", "_time_w2", "_time_b2",
               "time_output_w3")


def is_dead(name: str) -> bool:
    """Variables the reference creates but tf.gradients never reaches (time_aware_rnn.py:200-225, time_aware_attention.py:309)."""
    return name.rsplit("/", 1)[-1] in DEAD_LEAVES


def global_norm(pieces) -> float:
    """tf.clip_by_global_norm: sqrt(sum_t sum(values_t ** 2)), IndexedSlices by their raw values (trap T1)."""
    return math.sqrt(sum(float((np.asarray(x, np.float64) ** 2).sum()) for x in pieces))


def clip_and_adam(params, grads, pieces, m, v, t, lr, clip=1.0, b1=0.9, b2=0.999, eps=1e-8) -> Tuple[float, float]:
    """clip_by_global_norm(., clip) then tf.train.AdamOptimizer.apply_gradients, in place; returns (norm, scale).
    Adam as TF 1.14 writes it (trap T2): lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t); m, v decayed for EVERY row (the sparse
    apply is not lazy); w -= lr_t * m / (sqrt(v) + eps)."""
    gn = global_norm(pieces)
    scale = clip / max(gn, clip)
    lr32 = float(np.float32(lr))
    lr_t = lr32 * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
    for k, gk in grads.items():
        if gk is None:
            continue
        gs = gk * scale
        m[k] = b1 * m[k] + (1.0 - b1) * gs
        v[k] = b2 * v[k] + (1.0 - b2) * gs * gs
        params[k] = params[k] - lr_t * m[k] / (np.sqrt(v[k]) + eps)
    return gn, scale

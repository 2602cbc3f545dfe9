"""Generates tests/golden/*.npz: frozen outputs of the CPU oracle (oracle/mtam_oracle.py) on small
seeded inputs, for every model kind.  The reference itself cannot be executed in this image (TF 1.14,
SURVEY.md 8c), so these vectors pin the ORACLE's behaviour (so that later edits cannot silently
change what the CUDA path is compared against); they are not outputs of the reference.
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import mtam_oracle as O  # noqa: E402

CASES = {
    "MTAM": dict(L=9, D=32, H=2, N=2, user_count=12, item_count=60, category_count=5),
    "PISTREC": dict(L=7, D=32, H=4, N=2, user_count=12, item_count=60, category_count=5),
    "TA_SASREC": dict(L=7, D=32, H=1, N=1, user_count=12, item_count=60, category_count=5),
    "TISASREC": dict(L=7, D=32, H=2, N=2, user_count=12, item_count=60, category_count=5),
    "SASREC": dict(L=7, D=32, H=2, N=2, user_count=12, item_count=60, category_count=5),
    "BPRMF": dict(L=7, D=32, H=1, N=1, user_count=12, item_count=60, category_count=5),
    # oracle-only so far (O.NEXT_KINDS): pinned now so that the CUDA path of the next round has a fixed target
    "MTAM_VIA_T_GRU": dict(L=9, D=32, H=2, N=2, user_count=12, item_count=60, category_count=5),
    # round 2: the plain-GRU siblings (MTAMRec_model.py:93-125, 206-233)
    "MTAM_NO_TIME_AWARE_RNN": dict(L=9, D=32, H=2, N=2, user_count=12, item_count=60, category_count=5),
    "MTAM_VIA_RNN": dict(L=9, D=32, H=2, N=2, user_count=12, item_count=60, category_count=5),
}


def build(kind):
    cfg = O.OracleConfig(kind=kind, **CASES[kind])
    P = O.init_params(cfg, 1234)
    rng = np.random.default_rng(99)
    for k in P:
        if k.endswith("/bias") or k.endswith("/beta"):
            P[k] = (P[k] + 0.1 * rng.standard_normal(P[k].shape)).astype(np.float32)
    feed = O.synth_batch(cfg, 6, 4321)
    return cfg, P, feed


def main():
    only = sys.argv[1:]
    for kind in CASES:
        if only and kind not in only:       # regenerate named kinds only (the others stay byte-identical in git)
            continue
        cfg, P, feed = build(kind)
        fwd, grads, pieces = O.loss_and_grads(cfg, P, feed, bpr_negative=7)
        tr = O.OracleTrainer(cfg, P)
        losses = [tr.train_step(feed, 1e-3, bpr_negative=7) for _ in range(3)]
        blob = {"loss": np.float64(fwd["loss"].detach()), "pred": fwd["pred"].detach().numpy(),
                "global_norm": np.float64(O.global_norm(pieces)), "losses3": np.array(losses)}
        for k, v in grads.items():
            if v is not None:
                blob["grad:" + k] = v.astype(np.float32)
        for k, v in tr.params.items():
            blob["after3:" + k] = v.astype(np.float32)
        (_, idx, _) = O.metrics_topk(cfg, P, feed)
        blob["top50"] = idx
        np.savez_compressed(os.path.join(HERE, f"{kind}.npz"), **blob)
        print(kind, "loss", blob["loss"], "files ok")


if __name__ == "__main__":
    main()

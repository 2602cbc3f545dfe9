"""Times full-catalogue scoring + top-k (mtam_score_topk) in both gemm modes.  Usage: python tools/bench_topk.py [V B ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mtamrecommender_b200 import engine as E  # noqa: E402
from mtamrecommender_b200 import _lib  # noqa: E402


def run(V, B, D=64, k=50, modes=(0, 1), iters=5):
    g = torch.Generator(device="cuda").manual_seed(1)
    table = (torch.rand(V, D, device="cuda", generator=g) - 0.5) * 0.6
    pred = torch.randn(B, D, device="cuda", generator=g)
    lib = _lib.load()
    ws = torch.empty(int(lib.mtam_score_topk_workspace(B, V, k)), dtype=torch.uint8, device="cuda")
    out = {"V": V, "B": B, "D": D, "k": k}
    res = {}
    for m in modes:
        for _ in range(2):
            i, s = E.score_topk(pred, table, k, gemm_mode=m, workspace=ws)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            i, s = E.score_topk(pred, table, k, gemm_mode=m, workspace=ws)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        res[m] = (i, s)
        name = "fp32" if m == 0 else "tf32x3"
        out[name + "_ms"] = round(ms, 4)
        out[name + "_useful_TFLOPs"] = round(2.0 * B * V * D / ms / 1e9, 2)
        out[name + "_seq_per_s"] = round(B / ms * 1e3, 1)
    if len(res) == 2:
        out["indices_equal"] = bool(torch.equal(res[0][0], res[1][0]))
        out["scores_equal"] = bool(torch.equal(res[0][1], res[1][1]))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    cases = list(zip(a[0::2], a[1::2])) or [(100003, 1024), (1250001, 8192), (10000003, 1024)]
    for V, B in cases:
        run(V, B)

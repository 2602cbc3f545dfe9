"""Gather / scatter-add alone at cfg-4 shapes (the `bandwidth_kernels` section of bench.py without the train step)."""
import json
import sys
sys.path.insert(0, ".")
import torch
import bench

if __name__ == "__main__":
    from mtamrecommender_b200 import engine as E
    torch.cuda.set_device(0)
    r = bench.bandwidth_kernels(None, "cuda:0", bench.peaks())
    print(json.dumps(r, indent=1))

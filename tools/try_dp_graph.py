"""2-GPU probe: the data-parallel step captured into one CUDA graph (collectives included) == the eager DP step."""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
from mtamrecommender_b200 import _lib, engine as E
from mtamrecommender_b200.parallel import DataParallel
from mtamrecommender_b200.synth import ZipfSampler, synth_feed

w = dict(B=1024, L=50, D=64, N=6, H=1, items=100_000, cats=1_000, users=1_000_000)
mk = lambda: E.Engine(E.ModelConfig(kind="MTAM", max_batch=w["B"], L=w["L"], D=w["D"], H=w["H"], N=w["N"], user_count=w["users"],
                                    item_count=w["items"], category_count=w["cats"], gemm_mode=_lib.GEMM_TF32X3), device=dev, seed=1234)
samp = ZipfSampler(w["items"], 1.05)
feeds = [synth_feed(w["B"], w["L"], w["items"], w["cats"], w["users"], 50 + 10 * rank + i, samp) for i in range(3)]
e1, e2 = mk(), mk()
d1, d2 = DataParallel(e1), DataParallel(e2)
l1 = [d1.train_step(f, 1e-3) for f in feeds]
print(rank, "eager ok", l1, flush=True)
d2.capture_graph(w["B"])
print(rank, "captured", flush=True)
l2 = [d2.train_step(f, 1e-3) for f in feeds]
torch.cuda.synchronize()
print(rank, "graph losses", l2, "same losses", l1 == l2, "same params", bool(torch.equal(e1.params, e2.params)), flush=True)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, d in (("eager", d1), ("graph", d2)):
    b = d.eng.upload(feeds[0])
    step = (lambda: d.train_step_graph(1e-3)) if name == "graph" else (lambda: d.train_step_device(b, 1e-3))
    for _ in range(5): step()
    dist.barrier(); torch.cuda.synchronize()
    ev0.record()
    for _ in range(10): step()
    ev1.record(); torch.cuda.synchronize()
    print(rank, name, "ms/step", ev0.elapsed_time(ev1) / 10, flush=True)
dist.barrier()
torch.cuda.synchronize()
print(rank, "barrier ok", flush=True)
d2._graph = None
import gc
gc.collect()
torch.cuda.synchronize()
print(rank, "graph dropped", flush=True)
dist.destroy_process_group()
print(rank, "destroyed", flush=True)

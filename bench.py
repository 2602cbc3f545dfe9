#!/usr/bin/env python
"""Benchmark of the MTAM train step (forward + backward + clip + Adam) on B200.

Contract (one JSON line on stdout from rank 0):
  python bench.py --gpus N --steps K --warmup W            -> the CUDA path (libmtam_b200.so)
  python bench.py --impl reference ...                     -> CPU port of the reference graph on the host cores
Workload at N=1 is BASELINE.json configs[2]: synthetic MTAMRec, 100K items, seq len 50, batch 1024,
embed dim 64 (N=6 hops, 1 head).  With N>1 every rank runs the same per-GPU batch
(weak scaling) and the gradients are combined as described in DESIGN.md (multi-GPU).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: kind, B per GPU, L, D, N hops, H, items, cats, users
    "cfg3": dict(kind="MTAM", B=1024, L=50, D=64, N=6, H=1, items=100_000, cats=1_000, users=1_000_000),
    "cfg4": dict(kind="MTAM", B=1024, L=200, D=64, N=6, H=1, items=10_000_000, cats=1_000, users=1_000_000),
    # BASELINE configs[0] / [1]: ml-1m-shaped, default flags (num_units 128, 6 blocks, dropout 0.5 where the
    # reference applies it); PISTRec / the attention baselines use 8 heads as in config/model_parameter.py:13
    "cfg1": dict(kind="MTAM", B=256, L=50, D=128, N=6, H=1, items=3_706, cats=301, users=4_832),
    "cfg1_via_t_gru": dict(kind="MTAM_VIA_T_GRU", B=256, L=50, D=128, N=6, H=1, items=3_706, cats=301, users=4_832),
    "cfg2_pistrec": dict(kind="PISTREC", B=256, L=50, D=128, N=6, H=8, items=3_706, cats=301, users=4_832),
    "cfg2_sasrec": dict(kind="SASREC", B=256, L=50, D=128, N=6, H=8, items=3_706, cats=301, users=4_832, dropout=0.5),
    "cfg2_tisasrec": dict(kind="TISASREC", B=256, L=50, D=128, N=6, H=8, items=3_706, cats=301, users=4_832, dropout=0.5),
    "cfg2_ta_sasrec": dict(kind="TA_SASREC", B=256, L=50, D=128, N=6, H=8, items=3_706, cats=301, users=4_832),
    "tiny": dict(kind="MTAM", B=64, L=16, D=64, N=2, H=1, items=2_000, cats=50, users=500),
}
METRIC = "train seqs/s fwd+bwd (full train step incl. clip+Adam)"
UNIT = "seq/s"
LR = 1e-3


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.p = None
        self.path = f"/tmp/mtam_clocks_{os.getpid()}.csv"
        self.gpu_index = gpu_index
        self.nvml = None

    def _nvml_loop(self):
        import pynvml as N
        h = self.nvml
        bits = {"hw_slowdown": N.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": N.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": N.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": N.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                self._sm.append(float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)))
                r = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in bits.items():
                    if r & bit:
                        self._reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.005)

    def start(self):
        # NVML in a sampling thread (a sample every ~5 ms of the timed region); nvidia-smi -lms as the fallback
        try:
            import threading
            import pynvml as N
            N.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu_index]) if vis and vis.split(",")[0].isdigit() else self.gpu_index
            self.nvml = N.nvmlDeviceGetHandleByIndex(idx)
            self._max = float(N.nvmlDeviceGetMaxClockInfo(self.nvml, N.NVML_CLOCK_SM))
            self._sm, self._reasons, self._stop = [], set(), threading.Event()
            self._th = threading.Thread(target=self._nvml_loop, daemon=True)
            self._th.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.f = open(self.path, "w")
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.nvml is not None:
            self._stop.set()
            self._th.join(timeout=2)
            if self._sm:
                out.update(sm_mhz=float(np.median(self._sm)), sm_min_mhz=float(min(self._sm)), sm_max_mhz=self._max,
                           samples=len(self._sm), source="nvml")
            out["reasons"] = sorted(self._reasons)
            return out
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); mx.append(float(c[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.remove(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm)); out["sm_max_mhz"] = float(max(mx)); out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def make_feeds(w, n, seed0):
    from mtamrecommender_b200.synth import ZipfSampler, synth_feed
    samp = ZipfSampler(w["items"], 1.05)
    return [synth_feed(w["B"], w["L"], w["items"], w["cats"], w["users"], seed0 + i, samp) for i in range(n)]


# ---------------------------------------------------------------------------------------------
def run_reference(args, w):
    """CPU arm: the oracle's torch-fp32 port of the reference graph on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import mtam_oracle as O
    from oracle.cpu_port import TorchPort
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs = w["B"] if args.cpu_batch <= 0 else min(w["B"], args.cpu_batch)   # default: the workload's own batch size
    cfg = O.OracleConfig(kind=O.MTAM, L=w["L"], D=w["D"], H=w["H"], N=w["N"], user_count=w["users"],
                         item_count=w["items"], category_count=w["cats"])
    port = TorchPort(cfg, O.init_params(cfg, 1234))
    feeds = make_feeds(dict(w, B=Bs), 2, 1234)
    for i in range(args.warmup):
        port.train_step(feeds[i % 2], LR)
    t0 = time.perf_counter()
    for i in range(args.steps):
        port.train_step(feeds[i % 2], LR)
    dt = time.perf_counter() - t0
    val = Bs * args.steps / dt
    sample = f"{args.steps} steps of {Bs} sequences (full V={w['items']}, L={w['L']}, D={w['D']}, N={w['N']})"
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": args.workload, "model": "MTAM", "batch_per_gpu": Bs, "global_batch": Bs,
                      "seq_len": w["L"], "num_units": w["D"], "num_blocks": w["N"], "num_heads": w["H"],
                      "item_count": w["items"], "user_count": w["users"], "category_count": w["cats"],
                      "host_processes": 1,
                      "note": "TF 1.14 cannot run in this image; torch-CPU fp32 port of the same graph (oracle/cpu_port.py); "
                              "one host process whatever --gpus says"},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def cpu_baseline(w, seconds_budget=25.0):
    import torch
    from oracle import mtam_oracle as O
    from oracle.cpu_port import TorchPort
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs = w["B"]
    cfg = O.OracleConfig(kind=O.MTAM, L=w["L"], D=w["D"], H=w["H"], N=w["N"], user_count=w["users"],
                         item_count=w["items"], category_count=w["cats"])
    port = TorchPort(cfg, O.init_params(cfg, 1234))
    feed = make_feeds(dict(w, B=Bs), 1, 99)[0]
    port.train_step(feed, LR)
    n, t0 = 0, time.perf_counter()
    while n < 3 or (time.perf_counter() - t0 < seconds_budget / 2 and n < 16):
        port.train_step(feed, LR)
        n += 1
    dt = time.perf_counter() - t0
    return {"value": Bs * n / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} steps of {Bs} sequences after 1 warm-up (full V={w['items']}, L={w['L']}, D={w['D']}, "
                      f"N={w['N']}); torch-CPU fp32 port of the reference graph, not TF 1.14"}


# ---------------------------------------------------------------------------------------------
def run_cuda(args, w):
    """Stdout carries exactly one JSON line: while the run is in progress file descriptor 1 points at stderr, so that
    whatever a native library prints there (NCCL's version banner at NCCL_DEBUG=VERSION, for one) cannot precede it."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        line = _run_cuda(args, w)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    if line is not None:
        print(line, flush=True)


def _run_cuda(args, w):
    import torch
    import torch.distributed as dist
    from mtamrecommender_b200 import engine as E
    from mtamrecommender_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    dev = f"cuda:{local}"
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    mc = E.ModelConfig(kind=w["kind"], max_batch=w["B"], L=w["L"], D=w["D"], H=w["H"], N=w["N"], user_count=w["users"],
                       item_count=w["items"], category_count=w["cats"], dropout=w.get("dropout", 0.0),
                       gemm_mode=_gemm_mode(args))
    eng = E.Engine(mc, device=dev, seed=1234)           # same seed on every rank: replicas start identical
    dp = None
    if world > 1:
        from mtamrecommender_b200.parallel import DataParallel
        dp = DataParallel(eng)
    nb = 4
    feeds = make_feeds(w, nb, 1234 + 1000 * rank)
    batches = []
    for f in feeds:   # device-resident copies for the HBM-resident measurement
        t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in f.items()}
        batches.append(eng.device_batch(t))

    def step_dev(i):
        if dp is not None:
            dp.train_step_device(batches[i % nb], LR)
        else:
            eng.train_step_device(batches[i % nb], LR)

    def step_e2e(i):
        if dp is not None:
            return dp.train_step(feeds[i % nb], LR)
        return eng.train_step(feeds[i % nb], LR)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for i in range(steps):
            fn(i)
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    W = max(args.warmup, 3)
    for i in range(W):
        step_dev(i)
    use_graph = not args.no_graph
    if use_graph:
        # the whole step as one CUDA graph; with N > 1 the NCCL collectives of the data-parallel step are captured too
        if dp is None:
            eng.capture_train_graph(w["B"])
        else:
            dp.capture_graph(w["B"])
        # device-resident copies of the packed feed (the layout of the engine's staging buffer): one device-to-device
        # copy refreshes the captured step's inputs
        packed = []
        for f in feeds:
            eng.upload(f)
            packed.append(eng._dev_all.clone())
        torch.cuda.synchronize()

        def step_graph(i):
            eng._dev_all.copy_(packed[i % nb], non_blocking=True)
            (eng if dp is None else dp).train_step_graph(LR)
        for i in range(2):
            step_graph(i)
        fn_dev = step_graph
    else:
        fn_dev = step_dev
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = eng.launch_count()
    ms_dev = timed(fn_dev, args.steps)
    launches = eng.launch_count() - l0
    if use_graph:   # launches inside a replayed graph are not re-counted by the host counter: count one eager step
        l0 = eng.launch_count(); step_dev(0); launches = (eng.launch_count() - l0) * args.steps
        if dp is not None:
            barrier()
    for i in range(2):
        step_e2e(i)
    ms_blocking = timed(step_e2e, args.steps)
    # the headline e2e: the training loop a user runs (train_process.Train_main_process -> model.train_submit ->
    # engine.FeedPipeline): every step pads its batch into pinned memory, copies it host->device and reads its loss
    # back, but the host's share of step i+1 overlaps the device's step i; the last loss is read inside the timed region
    from mtamrecommender_b200.engine import FeedPipeline
    pipe = FeedPipeline(eng) if dp is None else dp.pipeline()

    def loop_pipelined(steps):
        losses = []
        for i in range(steps):
            sc = pipe.submit(feeds[i % nb], LR)
            if sc is not None:
                losses.append(float(sc[0]))
        losses.append(float(pipe.flush()[0]))
        return losses
    loop_pipelined(3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    e2e_losses = loop_pipelined(args.steps)
    ev1.record()
    barrier()
    assert len(e2e_losses) == args.steps and all(np.isfinite(e2e_losses))
    t_e2e = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    ms_e2e = float(t_e2e.item())
    # the same end-to-end step fed from the columnar record store (DataHandle/record_store.py): mtam_pack_records pads
    # the batch straight into the pinned feed buffers (make_feed_dic_new of the reference, Behavior_...py:146-192)
    ms_rec = None
    if dp is None:
        from mtamrecommender_b200.DataHandle.record_store import PackedRecords
        from mtamrecommender_b200.synth import feed_to_records
        store = PackedRecords.from_records([r for f in feeds for r in feed_to_records(f)])
        views = [store[i * w["B"]:(i + 1) * w["B"]] for i in range(nb)]
        for i in range(2):
            eng.train_step_records(views[i % nb], LR)
        ms_rec = timed(lambda i: eng.train_step_records(views[i % nb], LR), args.steps)
    clk = clocks.stop() if rank == 0 else {}
    seqs = w["B"] * world * args.steps
    value = seqs / (ms_dev / 1e3)
    e2e_val = seqs / (ms_e2e / 1e3)
    h2d = sum(v.nbytes for v in feeds[0].values())

    out = None
    if rank == 0:
        pk = peaks()
        # ---- per-phase device times of the same step (CUDA events on the step's stream) ----
        phases = {}
        reps = max(3, min(args.steps, 10))
        if dp is None and w["kind"] == "MTAM":
            for i in range(reps):
                p = eng.profile_step(batches[i % nb], LR)
                for k, v in p.items():
                    phases[k] = phases.get(k, 0.0) + v / reps
        T = w["B"] * w["L"]
        D, V, B, N = w["D"], w["items"] + 3, w["B"], w["N"]
        traffic = ncu_traffic()
        tot = sum(phases.values()) if phases else 0.0
        # algorithmic work per launch of each phase; tensor phases: useful fp32-equivalent FLOPs
        # (the 3xTF32 split issues 3x as many tf32 MMA FLOPs), against the measured bf16 sustained peak
        lf = float(np.mean(feeds[0]["seq_length"])) / w["L"]     # share of real (unmasked) keys
        work = {"adam": ("hbm", 7.0 * 4 * eng.n_floats),
                "ce_bwd": ("tensor", 6.0 * B * D * V),            # recompute + dPred + dTable (SURVEY 8d, K8 bwd)
                "ce_fwd": ("tensor", 2.0 * B * D * V),
                "kv_gemm": ("tensor", 2.0 * T * D * 2 * N * D),
                "hop_fwd": ("hbm", 4.0 * T * D * lf * (2 * N + 1)),   # K,V of every hop + X once, real keys only
                "hop_bwd": ("hbm", 4.0 * T * D * (lf * (2 * N + 1) + 2 * N + lf))}   # + dK,dV of every key, dX

        def roof_of(name):
            kind, amount = work[name]
            ms = phases[name]
            if kind == "hbm":
                ach, peak, unit = amount / (ms * 1e-3) / 1e9, pk["hbm"], "GB/s"
            else:
                ach, peak, unit = amount / (ms * 1e-3) / 1e12, pk["tf_sust"], "TFLOP/s"
            return {"kernel": name, "bound": kind, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                    "traffic": traffic.get(name), "traffic_source": "profiles/traffic_r02.json (ncu --set full capture of this step)",
                    "peak_source": pk["src"] + (" bf16 sustained" if kind == "tensor" else " copy"),
                    "ms": ms, "share_of_step": ms / tot,
                    **({"note": "the whole Adam update as ONE adam_kernel launch, as the profiling step (and the data-"
                                "parallel driver) runs it; in the timed, graphed step of the MTAM kinds the same update "
                                "is three launches: adam_rows_nograd_kernel (user rows the batch does not name, started "
                                "at the top of the step on a side stream), adam_rows_listed_kernel, adam_kernel (the rest)"}
                       if name == "adam" else {})}
        roof, roofs = None, []
        if phases:
            ranked = sorted((k for k in phases if k in work), key=lambda k: -phases[k])
            roofs = [roof_of(k) for k in ranked[:4]]
            roof = roofs[0] if roofs else None
        # ---- the two graded bandwidth kernels at cfg-4 shapes (n = 8192*200 rows, D = 64) ----
        bw = bandwidth_kernels(eng, dev, pk) if args.workload in ("cfg3", "cfg4") else None   # cfg-4-shape kernels: once
        ev = eval_topk_bench(eng, batches, w, pk) if dp is None else None
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
               "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": {"tf32x3": "f32 (tcgen05 3xTF32 split, fp32 accumulate)", "fp32": "f32",
                         "tf32": "tf32 (single-pass tcgen05 kind::tf32 in the softmax cross-entropy, 3xTF32 elsewhere; "
                                 "REDUCED precision, own tolerance: not the headline)"}[args.gemm_mode], "data": "synthetic",
               "config": {"workload": args.workload, "model": w["kind"], "batch_per_gpu": w["B"], "global_batch": w["B"] * world,
                          "seq_len": w["L"], "num_units": w["D"], "num_blocks": w["N"], "num_heads": w["H"],
                          "item_count": w["items"], "user_count": w["users"], "category_count": w["cats"],
                          "parallelism": f"dp{world}", "cuda_graph": bool(use_graph), "gemm_mode": args.gemm_mode,
                          "l2": "4 rotating batches; every step streams the 4 parameter/Adam arenas "
                                f"({4 * 4 * eng.n_floats / 1e6:.0f} MB) through HBM, > 126 MB L2"},
               "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(eng._scalars_host.numel() * eng._scalars_host.element_size()),
                       "ms_per_step": ms_e2e / args.steps,
                       "api": "FeedPipeline.submit per step (the loop of train_process.Train_main_process): pad + pinned "
                              "H2D copy + step + D2H loss every step, the host side of step i+1 overlapped with step i",
                       "blocking_call": {"value": seqs / (ms_blocking / 1e3), "ms_per_step": ms_blocking / args.steps,
                                         "api": "Engine.train_step(host feed) -> loss, one blocking call per step "
                                                "(= model.train of the reference)"},
                       "from_record_store": None if ms_rec is None else
                       {"value": seqs / (ms_rec / 1e3), "unit": UNIT, "ms_per_step": ms_rec / args.steps,
                        "what": "DataInput view of a PackedRecords store -> mtam_pack_records into pinned memory -> H2D -> step -> loss"}},
               "gpu_launches": int(launches), "clocks": clk, "phases_ms": phases, "roofline": roof,
               "rooflines_top_phases": roofs, "bandwidth_kernels": bw, "eval_topk": ev}
        if world == 1 and not args.no_cpu and w["kind"] == "MTAM":
            out["cpu_baseline"] = cpu_baseline(w)
    # BASELINE configs[3] in the same run (every rank takes part; rank 0 reports)
    if args.workload == "cfg3" and not args.no_cfg4:
        c4 = cfg4_section(args, rank, world, local, dev, timed)
        if out is not None:
            out["cfg4"] = c4
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        # captured graphs hold NCCL work: drop them before the process group goes (destroying it first hangs)
        import gc
        if dp is not None:
            dp._graph = None
        eng._graph = None
        gc.collect()
        torch.cuda.synchronize()
        dist.destroy_process_group()
    return json.dumps(out) if out is not None else None


def cfg4_section(args, rank, world, local, dev, timed_fn):
    """BASELINE configs[3]: synthetic MTAMRec, 10 M items, seq len 200, GLOBAL batch 8192 (strong scaling: 8192 / N
    sequences per GPU).  N = 1: the plain engine.  N > 1: the item table row-sharded over the ranks
    (parallel.ShardedItemTableTrainer: all-to-all lookup, sharded softmax CE with a rank-local table gradient, one
    all-reduce for the replicated parameters) -- the dense [10 M, 64] table gradient is never all-reduced."""
    import torch
    import torch.distributed as dist
    from mtamrecommender_b200 import _lib, engine as E
    from mtamrecommender_b200.synth import ZipfSampler, synth_feed
    w = dict(WORKLOADS["cfg4"])
    GB = 8192
    B = GB // world
    gm = _gemm_mode(args)
    V = w["items"] + 3
    try:
        if world > 1:
            from mtamrecommender_b200.parallel import ShardedItemTableTrainer, shard_rows
            S = shard_rows(V, world)
            eng = E.Engine(E.ModelConfig(kind="MTAM", max_batch=B, L=w["L"], D=w["D"], H=w["H"], N=w["N"], user_count=w["users"],
                                         item_count=S - 3, category_count=w["cats"], gemm_mode=gm), device=dev, seed=1234 + rank)
            # replicated parameters must start identical: broadcast everything but the item shard from rank 0
            sv_lo = int(eng.info["embedding_layer/item"].offset)
            sv_hi = sv_lo + S * w["D"]
            dist.broadcast(eng.params[:sv_lo], 0)
            dist.broadcast(eng.params[sv_hi:], 0)
            tr = ShardedItemTableTrainer(eng, V)
            step = lambda b: tr.train_step_device(b, LR)
            par = f"item table row-sharded over {world} ranks + data parallel (global batch {GB})"
        else:
            eng = E.Engine(E.ModelConfig(kind="MTAM", max_batch=B, L=w["L"], D=w["D"], H=w["H"], N=w["N"], user_count=w["users"],
                                         item_count=w["items"], category_count=w["cats"], gemm_mode=gm), device=dev, seed=1234)
            step = lambda b: eng.train_step_device(b, LR)
            par = "one GPU"
        samp = ZipfSampler(w["items"], 1.05)
        feeds = [synth_feed(B, w["L"], w["items"], w["cats"], w["users"], 777 + 10 * rank + i, samp) for i in range(2)]
        batches = [eng.device_batch({k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in f.items()}) for f in feeds]
        for i in range(3):
            step(batches[i % 2])
        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
        steps = max(3, min(args.steps, 8))
        ms = timed_fn(lambda i: step(batches[i % 2]), steps)
        clk = clocks.stop() if rank == 0 else {}
        loss = float(eng.read_scalars()[_lib.S_LOSS])
        # BASELINE configs[4] (cfg5): full-catalogue top-50 eval over the 10 M items for the global batch of 8192 --
        # metrics_topK (base_model.py:188-213).  N > 1: forward on looked-up rows, all-gather of pred, every rank scores
        # its shard (tcgen05 filter + exact fp32 rescoring), all-gather of the [B,50] lists, merge.
        ev_steps = 5
        if world > 1:
            ev = lambda b: tr.eval_topk(b, 50)
            pr = torch.randn((B, w["D"]), device=dev)
            sc = lambda b: tr.cat.topk(pr, 50)
        else:
            ev = lambda b: eng.eval_topk_device(b, 50)
            pr = torch.randn((B, w["D"]), device=dev)
            table = eng.param_view("embedding_layer/item")
            sc = lambda b: E.score_topk(pr, table, 50, gemm_mode=gm)
        for f in (ev, sc):
            f(batches[0])
        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
        ms_ev = timed_fn(lambda i: ev(batches[i % 2]), ev_steps)
        ms_sc = timed_fn(lambda i: sc(batches[i % 2]), ev_steps)
        clk5 = clocks.stop() if rank == 0 else {}
        flops = 2.0 * GB * w["D"] * V
        cfg5 = {"workload": "cfg5 (BASELINE configs[4]): top-50 over 10 M items, global eval batch 8192", "n_gpus": world,
                "ms_forward_score_top50_merge": ms_ev / ev_steps, "eval_seq_per_s": GB * ev_steps / (ms_ev / 1e3),
                "ms_score_top50_merge_only": ms_sc / ev_steps,
                "score_useful_TFLOPs_all_gpus": flops / (ms_sc / ev_steps) / 1e9, "clocks": clk5}
        out = {"workload": "cfg4 (BASELINE configs[3])", "items": w["items"], "seq_len": w["L"], "global_batch": GB,
               "batch_per_gpu": B, "n_gpus": world, "parallelism": par, "steps": steps, "ms_per_step": ms / steps,
               "seq_per_s": GB * steps / (ms / 1e3), "scaling": "strong", "loss": loss, "clocks": clk,
               "gemm_mode": args.gemm_mode, "cfg5_eval": cfg5}
        del eng
        torch.cuda.empty_cache()
        return out
    except Exception as e:   # report, never hide
        return {"workload": "cfg4", "error": repr(e)}


def eval_topk_bench(eng, batches, w, pk, k=50, reps=10):
    """metrics_topK of one batch (base_model.py:188-213): forward + full-catalogue scoring + top-50 + HR/NDCG, and the
    scoring/top-k part alone.  Tensor work of the scoring = 2*B*D*V useful FLOPs (3x that issued in tf32)."""
    import torch
    try:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def t(fn):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ev0.record()
            for i in range(reps):
                fn(i)
            ev1.record()
            torch.cuda.synchronize()
            return ev0.elapsed_time(ev1) / reps

        def full(i=0):
            b = batches[i % len(batches)]
            idx, _ = eng.eval_topk_device(b, k)
            return eng.hr_ndcg_device(idx, b.t["target_item_id"])
        ms_full = t(full)
        from mtamrecommender_b200 import engine as E
        pred = torch.randn(w["B"], w["D"], device=eng.params.device)
        table = eng.param_view("embedding_layer/item")
        ms_score = t(lambda i=0: E.score_topk(pred, table, k, gemm_mode=eng.cfg.gemm_mode))
        flops = 2.0 * w["B"] * w["D"] * table.shape[0]
        return {"k": k, "batch": w["B"], "items": int(table.shape[0]), "ms_forward_score_topk_metrics": ms_full,
                "seq_per_s": w["B"] / ms_full * 1e3, "ms_score_topk": ms_score,
                "score_topk_useful_TFLOPs": flops / ms_score / 1e9, "frac_of_bf16_sustained": flops / ms_score / 1e9 / pk["tf_sust"]}
    except Exception as e:   # report, never hide
        return {"error": repr(e)}


def _gemm_mode(args):
    from mtamrecommender_b200 import _lib
    return {"tf32x3": _lib.GEMM_TF32X3, "fp32": _lib.GEMM_FP32, "tf32": _lib.GEMM_TF32}[args.gemm_mode]


def bandwidth_kernels(eng, dev, pk, n=8192 * 200, D=64, rows=10_000_003):
    """Gather and scatter-add timed alone at cfg-4 shapes: 1.6 M rows of 256 B against a 10 M-row
    (2.56 GB) table, inputs far larger than L2.  Two id distributions: `uniform` (every row distinct with high
    probability: no cache reuse, the roofline case) and `zipf_pad` (Zipf(1.05) ids with 45 % pad id 0, the shape of a
    ragged batch: hot rows are served from L2, so the algorithmic-byte rate can exceed the HBM peak).
    scatter_add = whole op (one-sweep index sort + segmented reduce, each distinct dst row written once);
    scatter_add_sorted = the segmented reduce alone on pre-sorted indices, which is how a train step runs it (the sort
    depends only on the batch ids and is overlapped); *_accumulate = the `dst +=` form."""
    import torch
    from mtamrecommender_b200 import engine as E
    from mtamrecommender_b200.synth import ZipfSampler
    res = {}
    try:
        table = torch.empty((rows, D), dtype=torch.float32, device=dev).uniform_(-0.3, 0.3)
        out = torch.empty((n, D), dtype=torch.float32, device=dev)
        dst = torch.zeros((rows, D), dtype=torch.float32, device=dev)
        ws = torch.empty(E.scatter_add_workspace(n, rows, D), dtype=torch.uint8, device=dev)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def t(fn, reps=10):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ev0.record()
            for _ in range(reps):
                fn()
            ev1.record()
            torch.cuda.synchronize()
            return ev0.elapsed_time(ev1) / reps

        def entry(b, ms, **kw):
            return dict(ms=ms, algorithmic_bytes=b, achieved_GBs=b / ms / 1e6, frac=b / ms / 1e6 / pk["hbm"], **kw)
        rng = np.random.default_rng(5)
        for dist in ("uniform", "zipf_pad"):
            if dist == "uniform":
                idx_np = rng.integers(0, rows, n).astype(np.int32)
            else:
                idx_np = ZipfSampler(rows - 3, 1.05).sample(rng, n)
                idx_np[rng.random(n) < 0.45] = 0                      # pad id share of a ragged batch
            idx = torch.from_numpy(idx_np).to(dev)
            nu = int(np.unique(idx_np).size)
            r = {"rows": n, "D": D, "table_rows": rows, "unique": nu}
            r["gather"] = entry(n * (4 + 2 * D * 4), t(lambda: E.gather(table, idx, out)))
            b = n * (4 + D * 4) + nu * D * 4
            # the graded op: the IndexedSlices aggregation (read idx + rows once, write each distinct row once;
            # accumulate=False never reads dst).  `_accumulate` = the += form, which must also READ the touched dst
            # rows: its second figure counts that traffic.
            r["scatter_add"] = entry(b, t(lambda: E.scatter_add(dst, idx, out, ws, accumulate=False), reps=5))
            r["scatter_add_accumulate"] = entry(b, t(lambda: E.scatter_add(dst, idx, out, ws), reps=5),
                                                bytes_incl_dst_read=b + nu * D * 4)
            r["sort_indices_ms"] = t(lambda: E.sort_indices(idx, rows, ws), reps=5)
            srt = E.sort_indices(idx, rows)
            ws2 = torch.empty(max(int(_lib_sorted_ws(n, D)), 16), dtype=torch.uint8, device=dev)
            r["scatter_add_sorted"] = entry(b, t(lambda: E.scatter_add_sorted(dst, srt, out, ws2, accumulate=False), reps=5))
            r["scatter_add_sorted_accumulate"] = entry(b, t(lambda: E.scatter_add_sorted(dst, srt, out, ws2), reps=5),
                                                       bytes_incl_dst_read=b + nu * D * 4)
            for k in ("scatter_add_accumulate", "scatter_add_sorted_accumulate"):
                r[k]["achieved_GBs_incl_dst_read"] = (b + nu * D * 4) / r[k]["ms"] / 1e6
                r[k]["frac_incl_dst_read"] = r[k]["achieved_GBs_incl_dst_read"] / pk["hbm"]
            res[dist] = r
            del idx, srt, ws2
        del table, out, dst, ws
    except Exception as e:   # report, never hide
        res["error"] = repr(e)
    return res


def _lib_sorted_ws(n, D):
    from mtamrecommender_b200 import _lib
    return _lib.load().mtam_scatter_add_sorted_workspace(n, D)


def ncu_traffic():
    """DRAM bytes per launch of the kernels named in the rooflines.  ncu cannot run inside a timed bench, so these are
    taken from the committed `ncu --set full` capture of the same step (profiles/traffic_r02.json names the report and
    is regenerated by tools/round_profile.sh); `roofline.traffic_source` says so in the JSON line."""
    p = os.path.join(ROOT, "profiles", "traffic_r02.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--gemm-mode", default="tf32x3", choices=["tf32x3", "fp32", "tf32"],
                    help="dense contractions: tcgen05 3-term-split TF32 (fp32-class accuracy) or exact fp32 FFMA")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-bw", action="store_true", help="(unused; the bandwidth kernels run with cfg3 / cfg4 only)")
    ap.add_argument("--no-cfg4", action="store_true", help="skip the cfg4 (10 M items, global batch 8192) section")
    ap.add_argument("--cpu-batch", type=int, default=0,
                    help="sequences per step of the CPU arm; 0 = the workload's own batch size (the default)")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_cuda(args, w)


if __name__ == "__main__":
    main()

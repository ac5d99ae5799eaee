#!/usr/bin/env python
"""Headline benchmark: img/s of the DeepLabV3+ (reference-truncated Xception, OS16, 513x513, 21 classes) training
step — forward + loss + backward (+ NCCL gradient all-reduce for N>1) + Adam — on N B200s, one process per GPU.

  python bench.py --gpus 1 --steps 20 --warmup 5
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...     # the CPU restatement of the reference graph on the host cores

Prints ONE JSON line (rank 0).  `value` = whole-job img/s with inputs resident in HBM; `e2e` = the same through
Trainer.train_step_e2e (pinned host batch -> H2D -> step -> D2H loss, next batch's H2D prefetched behind the step); `roofline` = the dominant kernel family
measured with CUDA events in an instrumented eager pass of the same step; `cpu_baseline` = the oracle
(PyTorch-CPU restatement of the reference's TF graph; TF 2.4 itself is not installable offline) on a bounded sample.
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import subprocess
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

ASPP = [
    {"kernel": 3, "rate": [1, 1], "op": "conv", "input": -1},
    {"kernel": 3, "rate": [6, 6], "op": "conv", "input": 0},
    {"kernel": 3, "rate": [12, 12], "op": "conv", "input": 0},
    {"kernel": 3, "rate": [18, 18], "op": "conv", "input": 0},
    {"kernel": 1, "rate": [1, 1], "op": "pyramid_pooling", "input": 0, "target_size_factor": [1, 1]},
]
SS_NW = [0.70245001, 0.00893111, 0.00763626, 0.00877043, 0.00649604, 0.00544513, 0.01271576, 0.01909554, 0.03116511,
         0.01246875, 0.00623611, 0.01057388, 0.02777125, 0.00919422, 0.01154691, 0.07393348, 0.00606626, 0.00625678,
         0.01217829, 0.01340344, 0.00766524]
FLOP_PER_IMG_FWD_BWD = 142.3e9       # SURVEY.md §8(d) ledger, Xception(ref-truncated)/OS16/513^2
WORKLOAD = ("Xception(ref-truncated, block13_sepconv2_bn tap) OS16 513x513x3 -> 512x512x21, batch {batch}/GPU, "
            "fwd + class-balanced loss + bwd + Adam, training-mode BN, dropout 0.5, random init")


def make_conf(dtype, image_size=513):
    return {
        "mode": "train", "resource_path": "", "model_loading": False, "base_model": "xception", "base_weights": None,
        "hps": {"dtype": dtype, "lr": 1e-4, "beta_1": 0.5, "beta_2": 0.99, "decay": 0.0, "epochs": 1,
                "batch_size": 16, "weight_decay": 4e-5, "bn_momentum": 0.9, "bn_scale": True,
                "reduce_lr_factor": 0.99},
        "nn_arch": {"boundary_refinement": False, "output_stride": 16, "image_size": image_size, "num_classes": 21,
                    "mv2_depth_multiplier": 1, "depth_multiplier": 1, "conv_rate_multiplier": 1,
                    "reduction_size": 256, "dropout_rate": 0.5, "concat_channels": 256,
                    "encoder_middle_conf": copy.deepcopy(ASPP)},
    }


def synthetic(conf, batch, out_hw, seed):
    """images U(-1,1) (ss.py:1532); labels: rectangles over background, classes drawn ~ VOC pixel frequencies."""
    rng = np.random.default_rng(seed)
    s = conf["nn_arch"]["image_size"]
    x = rng.uniform(-1, 1, (batch, s, s, 3)).astype(np.float32)
    y = np.zeros((batch,) + tuple(out_hw), dtype=np.int32)
    p = np.asarray(SS_NW[1:]) / np.sum(SS_NW[1:])
    for b in range(batch):
        for _ in range(rng.integers(4, 13)):
            c = 1 + int(rng.choice(20, p=p))
            h, w = rng.integers(16, out_hw[0] // 3), rng.integers(16, out_hw[1] // 3)
            y0, x0 = rng.integers(0, out_hw[0] - h), rng.integers(0, out_hw[1] - w)
            y[b, y0:y0 + h, x0:x0 + w] = c
    return x, y


def he_init(model, seed=1024):
    rng = np.random.default_rng(seed)
    for l in model.flat_layers():
        vals = []
        for n in l.weight_names():
            w = l._weights[n]
            if n in ("kernel", "pointwise_kernel"):
                v = rng.normal(0, np.sqrt(2.0 / (w.shape[0] * w.shape[1] * w.shape[2])), w.shape)
            elif n == "depthwise_kernel":
                v = rng.normal(0, np.sqrt(2.0 / 9.0), w.shape)
            elif n in ("gamma", "moving_variance"):
                v = np.ones(w.shape)
            else:
                v = np.zeros(w.shape)
            vals.append(v.astype(np.float32))
        l.set_weights(vals)


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe).  NVML is polled every 20 ms
    (nvidia-smi takes longer than that per query and would give one sample for a 0.2 s region); nvidia-smi is the
    fallback when pynvml is unavailable."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.mx, self.reasons, self._halt = index, [], [], set(), threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll_nvml(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)))
        get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        r = get(self.handle)
        for name, bit in (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
                          ("sw_power_cap", 0x4)):
            if r & bit:
                self.reasons.add(name)

    def _poll_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                              "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
        r = [c.strip() for c in out.strip().split(",")]
        if len(r) >= 7 and r[0].replace(".", "").isdigit():
            self.sm.append(float(r[0]))
            self.mx.append(float(r[1]))
            self.reasons.update(self.NAMES[i] for i in range(4) if r[3 + i].lower() == "active")

    def run(self):
        while not self._halt.is_set():
            try:
                self._poll_nvml() if self.nvml is not None else self._poll_smi()
            except Exception:
                pass
            self._halt.wait(0.002 if self.nvml is not None else 0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def cpu_baseline(steps=2, warmup=1, batch=2, threads=None):
    """The reference graph restated on the CPU (oracle, PyTorch-CPU fp32), fwd+bwd, `batch` images per step."""
    from deeplabv3plus_keras_b200 import keras
    from deeplabv3plus_keras_b200.deeplab import SemanticSegmentation, ss_nw, ss_pw
    from oracle import model as OM

    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    conf = make_conf("float32")
    conf["nn_arch"]["dropout_rate"] = 0.0
    keras.reset_uids()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ss = SemanticSegmentation(conf)
    he_init(ss.model)
    w = {k: torch.from_numpy(v.copy()) for k, v in ss.model.named_weights().items()}
    x, y = synthetic(conf, batch, (512, 512), 1024)
    xt, yt = torch.from_numpy(x), torch.from_numpy(y)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        OM.loss_and_grads(conf, w, xt, yt, ss_pw, ss_nw)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    return {"value": batch / sec, "unit": "img/s", "cores": threads, "kind": "port",
            "sample": f"{steps} steps of batch {batch} (of the 16-image workload), Xception OS16 513^2 fwd+bwd fp32, "
                      f"PyTorch-CPU restatement of the reference TF graph (TF 2.4 not installable offline)",
            "ms_per_step": sec * 1e3}


def run_reference(args, rank):
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    cb = cpu_baseline(steps=steps, warmup=min(max(args.warmup, 0), 1), batch=2)
    line = {"impl": "reference", "metric": "img/s DeepLabV3+ Xception OS16 513^2 fwd+bwd", "value": cb["value"],
            "unit": "img/s", "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1),
            "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD.format(batch=16), "global_batch": 16, "parallelism": "cpu",
                       "sample": "each step = 2 images of the 16-image batch on the host cores (bounded sample)"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="images per GPU per step")
    ap.add_argument("--dtype", default="bfloat16")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--profile-json", default="")
    ap.add_argument("--no-overlap", action="store_true", help="run the filter-gradient kernels in line (A/B)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and not (world == 1 and args.gpus == 1):
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    args.warmup = max(args.warmup, 3)
    # the contract is ONE JSON line on stdout: libraries that write to fd 1 (NCCL prints its version there on init) are
    # diverted to stderr; the JSON line goes to the saved descriptor
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    import torch.distributed as dist

    from deeplabv3plus_keras_b200 import keras
    from deeplabv3plus_keras_b200.deeplab import SemanticSegmentation
    from deeplabv3plus_keras_b200.profiler import KernelProfiler
    from deeplabv3plus_keras_b200.trainer import Trainer

    torch.cuda.set_device(local_rank)
    pg = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
        pg = dist.group.WORLD

    conf = make_conf(args.dtype)
    keras.reset_uids()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ss = SemanticSegmentation(conf)
    he_init(ss.model)
    tr = Trainer(ss.model, args.batch, use_graph=not args.no_graph, process_group=pg, overlap_wgrad=not args.no_overlap)
    plan = tr.plan
    x, y = synthetic(conf, args.batch, plan.out_shape[1:3], 1024 + rank)
    xs, ys = torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory()
    tr.stage_inputs(xs, ys)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ---------------------------------------------------------------------
    for _ in range(args.warmup):
        tr.step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(tr.stream)
    for _ in range(args.steps):
        tr.step()
    e1.record(tr.stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    t = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = args.batch * world * args.steps / (ms_total / 1e3)
    loss = tr.read_loss()

    # ---- end to end: pinned host batch -> H2D -> step -> D2H loss ----------------------------------------
    # every call trains on one batch that comes from pinned host memory and returns that batch's loss from the device;
    # the H2D copy of the next call's batch is started behind the step launch (Trainer.prefetch, input double
    # buffering) — one 67 MB H2D copy and one 8-byte D2H read per step, all inside the timed region
    for _ in range(2):
        tr.train_step_e2e(xs, ys, prefetch_next=(xs, ys))
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tr.train_step_e2e(xs, ys, prefetch_next=(xs, ys))
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = args.batch * world * args.steps / float(t.item())
    h2d = xs.numel() * 4 + ys.numel() * 4
    d2h = 8

    line = {
        "metric": "img/s DeepLabV3+ Xception OS16 513^2 fwd+bwd", "value": value, "unit": "img/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.dtype == "bfloat16" else "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(batch=args.batch), "global_batch": args.batch * world,
                   "parallelism": f"dp{world}", "cuda_graph": not args.no_graph,
                   "l2_policy": "per-step working set (several GB of activations) far exceeds the 126 MB L2"},
        "clocks": clocks, "loss": loss,
        "e2e": {"value": e2e_value, "unit": "img/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "pipeline": "Trainer.train_step_e2e(batch, prefetch_next=next_batch): the H2D copy of the next batch "
                            "overlaps the current step; the loss of every step is read back synchronously"},
        "gpu_launches": tr.launches_per_step * args.steps,
        "model_tflops": FLOP_PER_IMG_FWD_BWD * value / 1e12,
    }

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        tc_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "measured" if peaks else "fallback"
        if not args.no_profile:
            # instrumented eager pass of the same step: CUDA events around every kernel launch
            side, plan.side_stream = plan.side_stream, None        # serialise: per-kernel times must not overlap
            with torch.cuda.stream(tr.stream):
                with KernelProfiler() as kp:
                    for _ in range(2):
                        plan.zero_grads(); plan.forward(); plan.loss_forward_backward(); plan.backward()
                        plan.regularization(); tr._adam(); plan.run_prep()
                summ = kp.summary()
            plan.side_stream = side
            total_ms = sum(a["ms"] for a in summ.values())
            top_name, top = next(iter(summ.items()))
            per_launch_ms = top["ms"] / top["calls"]
            if top["bound"] == "tensor":
                ach, peak, unit = top["TFLOPs"], tc_peak, "TFLOP/s"
            else:
                ach, peak, unit = top["GBps"], hbm_peak, "GB/s"
            traffic, traffic_src = None, None
            try:      # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed `ncu --set full` capture
                tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
                if top_name in tj:
                    traffic, traffic_src = tj[top_name]["bytes_per_launch"], tj[top_name]["source"]
            except Exception:
                pass
            line["roofline"] = {"bound": "tensor" if top["bound"] == "tensor" else "hbm", "kernel": top_name,
                                "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak, "traffic": traffic,
                                "traffic_source": traffic_src,
                                "peak_source": peak_src, "share_of_step": top["ms"] / total_ms,
                                "avg_launch_ms": per_launch_ms, "launches_per_step": top["calls"] // 2}
            shapes = [r for r in kp.detail(200) if r["kernel"] == top_name]
            if shapes:       # the single most expensive problem size of the dominant entry point
                d0 = shapes[0]
                line["roofline"]["dominant_shape"] = {
                    "args": d0["args"], "launches_per_step": d0["calls"] // 2, "ms_per_step": d0["ms"] / 2,
                    "avg_launch_ms": d0["ms"] / d0["calls"], "TFLOPs": round(d0["TFLOPs"], 1),
                    "GBps": round(d0["GBps"], 1),
                    "frac": round(d0["TFLOPs"] / tc_peak if top["bound"] == "tensor" else d0["GBps"] / hbm_peak, 4)}
            line["kernels"] = {k: {"calls_per_step": a["calls"] // 2, "ms_per_step": a["ms"] / 2,
                                   "GBps": round(a["GBps"], 1), "TFLOPs": round(a["TFLOPs"], 2), "bound": a["bound"],
                                   "frac": round((a["TFLOPs"] / tc_peak) if a["bound"] == "tensor"
                                                 else (a["GBps"] / hbm_peak), 4)}
                               for k, a in list(summ.items())[:14]}
            if args.profile_json:
                json.dump({"per_kernel": {k: dict(a) for k, a in summ.items()}, "per_shape": kp.detail(60)},
                          open(args.profile_json, "w"), indent=1)
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline()
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — headline benchmark of BASELINE.json ("x4 RRDBNet LR Mpix/s + degraded pairs/s ... % of roofline").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Our arm (default): one step = one RRDBNet x4 forward of a 64x3x128x128 LR batch (BASELINE.json configs[2]) per GPU,
batch-sharded (every rank owns its own 64 images, no collective on the data path). Rank 0 prints ONE JSON line:
  value      device-timed LR Mpix/s, inputs resident in HBM (CUDA events on the launch stream, max over ranks)
  e2e        the same through the host-buffer C ABI call (pinned H2D + forward + D2H inside the timed region)
  roofline   tensor-core roofline of the conv kernel against MEASURED_PEAKS.json
  cpu_baseline  the fp32 oracle port of the reference forward on this box's host cores (bounded sample)
  degradation   secondary object: degraded pairs/s of the second-order pipeline (configs[1], canonical plan S0)
                with its HBM roofline (stage-sum bytes)
`--impl reference`: the reference's own CPU implementation of the path (oracle port; /root/reference does not exist
on the GPU box) timed on the host cores, one 1x3x128x128 image per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

FLOP_PER_LR_PIXEL = 35853696.0  # SURVEY.md §8a layer table
METRIC = "x4 RRDBNet LR Mpix/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "source": "MEASURED_PEAKS.json"}
    return {"hbm": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (NVML, 20 ms period; nvidia-smi fallback)."""

    def __init__(self, index):
        self.rows, self.stop, self.index = [], threading.Event(), index
        self.t = threading.Thread(target=self.run, daemon=True)
        self.max_mhz = None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            while not self.stop.is_set():
                mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                r = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)) if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.rows.append((mhz, r))
                self.stop.wait(0.02)
            return
        except Exception:
            pass
        q = "clocks.sm,clocks.max.sm"
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    a, b = (float(v) for v in out.split(","))
                    self.max_mhz = b
                    self.rows.append((a, 0))
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        sm = sorted(r[0] for r in self.rows)
        bits = 0
        for r in self.rows:
            bits |= r[1]
        # NVML clocks-event-reason bits
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        reasons = [n for b, n in names.items() if bits & b]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm)}


def cpu_generator_baseline(seconds_budget=20.0, max_iters=5):
    """Oracle port of the reference forward (plain torch fp32, all host cores) on cfg1 (1x3x128x128)."""
    from oracle import generator as og
    torch.set_num_threads(os.cpu_count() or 1)
    sd = og.random_state_dict(0)
    torch.manual_seed(0)
    x = torch.rand(1, 3, 128, 128)
    og.generator_forward(x[:, :, :32, :32], sd)  # warm-up
    times = []
    t_all = time.perf_counter()
    while len(times) < max_iters and (time.perf_counter() - t_all) < seconds_budget:
        t0 = time.perf_counter()
        og.generator_forward(x, sd)
        times.append(time.perf_counter() - t0)
    med = sorted(times)[len(times) // 2]
    return {"value": 128 * 128 / med / 1e6, "unit": "LR Mpix/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{len(times)} fp32 forwards of 1x3x128x128 (BASELINE configs[0]), median {med:.3f} s"}


# ------------------------------------------------------------------------------------------------- degradation


def degradation_bench(device, steps, warmup, peaks, B=16):
    """configs[1]: B x 3 x 256 x 256 HR crops (B = 16) through the canonical plan S0; device-timed with resident inputs.
    B = 256 is the large-batch variant SURVEY.md §8d asks for (launch latencies amortised)."""
    import resr_b200
    ip = resr_b200.imgproc
    H, W = 256, 256
    plan = resr_b200.plan.canonical_plan_s0(B, H, W, seed=0)
    g = torch.Generator(device="cpu").manual_seed(0)
    hr = torch.rand(B, 3, H, W, generator=g).to(device)
    k = torch.zeros(B, 21, 21)
    ax = torch.arange(21) - 10.0
    for i in range(B):  # isotropic Gaussians of assorted support, zero-padded to 21 (dataset.py:102-103)
        ks = 7 + 2 * (i % 8)
        s = 0.5 + 0.3 * i
        kk = torch.exp(-(ax[:, None] ** 2 + ax[None] ** 2) / (2 * s * s))
        kk[(ax.abs() > ks // 2)[:, None] | (ax.abs() > ks // 2)[None]] = 0
        k[i] = kk / kk.sum()
    k1 = k.to(device)
    k2 = k.flip(0).contiguous().to(device)
    sk = torch.zeros(B, 21, 21)
    sk[:, 10, 10] = 1
    sk = sk.to(device)
    # One CUDA graph for the whole plan-driven launch sequence; every plan tensor is resident on the device.
    pipe = ip.DegradePipeline(hr, k1, k2, sk, plan)

    def step():
        return pipe()

    for _ in range(max(3, warmup)):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        lr, hrc = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # stage-sum algorithmic bytes of S0: 4 B x (elements read + written) per executed stage
    E0 = B * 3 * H * W
    e1_, e2_ = E0 // 4, E0 // 16
    stage_bytes = 4 * (2 * E0 + 2 * E0 + (E0 + e1_) + (2 * e1_ + e1_) + 2 * e1_ + 2 * e1_ + (e1_ + e2_) + (2 * e2_ + e2_)
                       + 2 * e2_ + 2 * e2_ + 2 * e2_ + 2 * e2_)
    gbs = stage_bytes / (ms * 1e-3) / 1e9
    return {"metric": "degraded pairs/s", "value": B / (ms * 1e-3), "unit": "pairs/s", "ms_per_step": ms,
            "config": {"workload": f"second-order degradation, {B}x3x256x256 HR -> {B}x3x64x64 LR, canonical plan S0 "
                                   "(SURVEY.md §8d): Gaussian noise tensors host-fed and resident, Poisson draws made "
                                   "inside the fused noise kernel (Philox) in the timed region; one CUDA-graph replay per batch"},
            "gpu_launches_per_step": 14,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                         "traffic": None, "algorithmic_bytes_per_step": stage_bytes,
                         "note": "stage-sum bytes / whole-pipeline time; blur stencils are FMA-bound (SURVEY.md §8d caveat)"}}


# ------------------------------------------------------------------------------------------------- training step


def training_bench(device, steps, warmup, peaks, world):
    """configs[3]: RealESRNet training-step core per GPU — plan-driven degradation of 16 HR crops (256^2 -> LR 64^2),
    generator forward + L1 + backward, and (world > 1) the NCCL all-reduce of the flat gradient vector. Optimizer / EMA
    are outside the north-star path."""
    import torch.distributed as dist

    import resr_b200
    ip = resr_b200.imgproc
    B, H, W = 16, 256, 256
    plan = resr_b200.plan.canonical_plan_s0(B, H, W, seed=1)
    g = torch.Generator(device="cpu").manual_seed(2)
    hr = torch.rand(B, 3, H, W, generator=g).to(device)
    k = torch.zeros(B, 21, 21)
    k[:, 8:13, 8:13] = 1 / 25
    k = k.to(device)
    pipe = ip.DegradePipeline(hr, k, k, k, plan)
    torch.manual_seed(0)
    gen = resr_b200.model.Generator(3, 3, 4).to(device).train()
    ts = resr_b200.autograd.TrainStep(gen, B, H // 4, W // 4, device, None, world)

    def step():
        lr, hr_c = pipe()
        return ts.step(lr, hr_c, scatter=False)

    for _ in range(max(3, warmup)):
        loss, _, _ = step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss, _, _ = step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    tflops = 3 * FLOP_PER_LR_PIXEL * B * (H // 4) * (W // 4) / (ms * 1e-3) / 1e12
    # optimizer side (outside the north-star metric): fused Adam + EMA over the flat vectors and the
    # repack of the tensor-core weight tiles that the next forward needs
    opt_info = None
    try:
        opt = resr_b200.optim.FlatAdamEMA(gen)
        flat = ts.flat
        for _ in range(2):
            opt.step(flat)
            gen._ensure_packed()
        torch.cuda.synchronize()
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o0.record()
        for _ in range(5):
            opt.step(flat)
            gen._ensure_packed()
        o1.record()
        torch.cuda.synchronize()
        oms = o0.elapsed_time(o1) / 5
        opt_info = {"ms_per_step": oms, "what": "resr_adam_ema_step (36 B per parameter) + repack of all tensor-core weight tiles (2 launches)"}
    except Exception as e:
        opt_info = {"error": repr(e)}
    return {"metric": "training pairs/s", "value": world * B / (ms * 1e-3), "unit": "pairs/s", "ms_per_step": ms, "n_gpus": world,
            "loss": float(loss.item()), "cuda_graph": bool(ts.is_graph), "optimizer": opt_info,
            "config": {"workload": "per GPU: degradation (plan S0, CUDA graph) of 16x3x256x256 HR + RRDBNet x4 forward/L1/backward "
                                   "on 16x3x64x64 LR (one CUDA graph), flat-gradient NCCL all-reduce when n_gpus > 1; no optimizer"},
            "roofline": {"bound": "tensor", "achieved": tflops, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                         "frac": tflops / peaks["tf_sustained"], "traffic": None,
                         "note": "algorithmic FLOPs = 3 x forward (SURVEY.md §8d): 7.05 TFLOP per GPU-step"}}


# ------------------------------------------------------------------------------------------------- arms


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import generator as og
    torch.set_num_threads(os.cpu_count() or 1)
    sd = og.random_state_dict(0)
    torch.manual_seed(0)
    x = torch.rand(1, 3, 128, 128)
    for _ in range(max(1, min(args.warmup, 2))):
        og.generator_forward(x, sd)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        og.generator_forward(x, sd)
    dt = (time.perf_counter() - t0) / args.steps
    val = 128 * 128 / dt / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "LR Mpix/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "RRDBNet x4 (23 RRDB, nf=64, gc=32) forward, random init; each step a bounded sample "
                                   "of the 64x3x128x128 workload: one 1x3x128x128 image on the host CPU"},
            "cpu_baseline": {"value": val, "unit": "LR Mpix/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": "1x3x128x128 fp32 forward per step, oracle port of model.py (reference tree is not on the GPU box)"},
            "e2e": {"value": val, "unit": "LR Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA (B200) device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    import resr_b200
    L = resr_b200._lib
    peaks = measured_peaks()
    torch.set_grad_enabled(False)

    N, H, W = args.batch, 128, 128
    torch.manual_seed(0)
    gen = resr_b200.model.Generator(3, 3, 4).to(device).eval()
    gcpu = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.rand(N, 3, H, W, generator=gcpu).pin_memory()
    y_host = torch.empty(N, 3, 4 * H, 4 * W).pin_memory()
    x = x_host.to(device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing
    for _ in range(max(3, args.warmup)):
        y = gen(x)
    barrier()
    with ClockSampler(local) as clocks:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            y = gen(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([ms], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    barrier()
    # ---- end-to-end through the host-buffer ABI calls: every step copies its input from pinned host memory and its
    # result back to pinned host memory inside the timed region. (a) blocking call, (b) pipelined serving call: the
    # copies of neighbouring steps overlap the current step's compute (two staging slots, two result buffers).
    def timed_e2e(fn, fin):
        for _ in range(2):
            fn(0)
        fin()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for k in range(e2e_steps):
            fn(k)
        fin()
        b.record()
        torch.cuda.synchronize()
        tt = torch.tensor([a.elapsed_time(b) / e2e_steps], device=device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    e2e_steps = max(4, min(args.steps, 10))
    y_hosts = [y_host, torch.empty_like(y_host).pin_memory()]
    ms_e2e_blocking = timed_e2e(lambda k: gen.infer_host(x_host, y_host, device), lambda: None)
    ms_e2e = timed_e2e(lambda k: gen.infer_host_async(x_host, y_hosts[k & 1], device), gen.host_sync)
    checksum = float(y_host[0, :, ::64, ::64].double().sum())

    # ---- degradation leg: every rank degrades its own batches (no collective on the path); aggregate = world x B / max time
    deg = None
    if not args.no_degrade:
        big, err = None, None
        try:
            deg = degradation_bench(device, max(10, args.steps), args.warmup, peaks)
            big = degradation_bench(device, 10, 3, peaks, B=256)  # large-batch regime
        except Exception as e:  # keep the headline even if the secondary leg breaks
            err = repr(e)
        # the collective runs on every rank whatever happened above (a failed rank contributes +inf)
        tt = torch.tensor([deg["ms_per_step"] if err is None else float("inf"),
                           big["ms_per_step"] if err is None else float("inf")], device=device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        m16, m256 = float(tt[0].item()), float(tt[1].item())
        if err is not None or m16 == float("inf"):
            deg = {"error": err or "another rank failed"}
        else:
            deg["ms_per_step"] = m16
            deg["value"] = world * 16 / (m16 * 1e-3)
            deg["n_gpus"] = world
            deg["roofline"]["achieved"] = deg["roofline"]["algorithmic_bytes_per_step"] / (m16 * 1e-3) / 1e9
            deg["roofline"]["frac"] = deg["roofline"]["achieved"] / peaks["hbm"]
            deg["roofline"]["note"] += "; per GPU"
            deg["large_batch"] = {"batch_per_gpu": 256, "value": world * 256 / (m256 * 1e-3), "unit": big["unit"], "ms_per_step": m256,
                                  "roofline_frac": big["roofline"]["algorithmic_bytes_per_step"] / (m256 * 1e-3) / 1e9 / peaks["hbm"]}
    if rank == 0:
        px = N * H * W
        launches = L.lib().resr_generator_launches_per_forward()
        tflops = FLOP_PER_LR_PIXEL * px / (ms * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": world * px / (ms * 1e-3) / 1e6, "unit": "LR Mpix/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
            "config": {"workload": f"RRDBNet x4 (23 RRDB, nf=64, gc=32) inference, {N}x3x{H}x{W} LR per GPU, random init "
                                   "(BASELINE.json configs[2]); batch-sharded, no collective",
                       "precision": "fp16 MMA operands (16-bit tensor-core rate, same as bf16), fp32 accumulate and fp32 epilogue arithmetic; the trunk residual stream is the fp16 conv input (saturating at 65504)",
                       "l2": "working set (9 GB of activations per forward) is far larger than the 126 MB L2; no flush needed"},
            "e2e": {"value": world * px / (ms_e2e * 1e-3) / 1e6, "unit": "LR Mpix/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": y_host.numel() * 4,
                    "api": "resr_generator_forward_host_async + resr_generator_host_sync (pinned host buffers, pipelined)",
                    "blocking_call": {"value": world * px / (ms_e2e_blocking * 1e-3) / 1e6, "ms_per_step": ms_e2e_blocking,
                                      "api": "resr_generator_forward_host"},
                    "checksum": checksum},
            "gpu_launches": launches * args.steps,
            "clocks": clocks.summary(),
            "roofline": {"bound": "tensor", "achieved": tflops, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                         "frac": tflops / peaks["tf_sustained"],
                         # dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the five launches of one
                         # dense block (345 of the 351 launches are such blocks) at cfg3: ncu --set full capture in
                         # profiles/r01_v5_rdb_ncu_full.csv (182 / 323 / 328 / 470 / 525 MB; algorithmic 201 / 268 / 335 /
                         # 402 / 671 MB -- conv5's residual and second slice hit L2)
                         "traffic": 365.7e6 if (N, H, W) == (64, 128, 128) else None,
                         "kernel": "conv3x3_tc_kernel (351 launches per forward; algorithmic FLOPs 35,853,696 per LR pixel)",
                         "peak_source": peaks["source"] + " bf16_tflops_sustained (kernel timed inside a long step)"},
        }
        if deg is not None:
            line["degradation"] = deg
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_generator_baseline()
    train = None
    if not args.no_train:  # every rank takes part (all-reduce)
        try:
            train = training_bench(device, max(5, min(args.steps, 20)), args.warmup, peaks, world)
        except Exception as e:
            train = {"error": repr(e)}
    if rank == 0:
        line["training"] = train
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="LR images per GPU (64 = BASELINE configs[2])")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-degrade", action="store_true", help="skip the secondary degradation leg")
    ap.add_argument("--no-train", action="store_true", help="skip the secondary training-step leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

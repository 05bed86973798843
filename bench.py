#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 bayesian-torch hot path.

Workload (BASELINE.json configs[2], "C3"): MC inference through dnn_to_bnn(torchvision ResNet-18,
10 classes), Reparameterization layers, synthetic 3x32x32 inputs, batch B=128, N=64 Monte-Carlo
weight samples per input batch.  One "step" = one input batch -> predictive mean + variance
[2, B, C] over the N samples (fused layer kernels, fused softmax/moment kernel, ONE all-reduce).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --steps K --warmup W    # reference arithmetic on host cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...    # N > 1: one rank per GPU, samples sharded

Prints ONE JSON line (rank 0).  metric = MC image-samples/sec = B*N / t_step.
"""
import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B, N_MC, N_CLASSES = 128, 64, 10
REF_MC = 8          # MC samples per step of the bounded CPU sample (reference arm / cpu_baseline)
PRM = {"prior_mu": 0.0, "prior_sigma": 1.0, "posterior_mu_init": 0.0, "posterior_rho_init": -3.0,
       "type": "Reparameterization", "moped_enable": False, "moped_delta": 0.5}
METRIC = "mc_inference_image_samples_per_sec_N64_bayesian_resnet18"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------------- reference arm
def build_oracle_model():
    import torchvision
    from oracle.ref_model import oracle_dnn_to_bnn
    torch.manual_seed(0)
    net = torchvision.models.resnet18(num_classes=N_CLASSES)
    return oracle_dnn_to_bnn(net, flipout=False).eval()


def time_cpu_reference(steps, warmup):
    """The reference's arithmetic (same ATen op sequence, oracle/ref_model.py) on the host cores:
    B=128, REF_MC sequential MC forwards + stack/softmax/mean per step."""
    from oracle.ref_model import oracle_mc_evaluate
    avail = len(os.sched_getaffinity(0))
    net = build_oracle_model()
    torch.manual_seed(0)
    x = torch.randn(B, 3, 32, 32)
    # "all the host threads it can use": ATen's intra-op pool degrades when oversubscribed on these small
    # convolutions, so pick the fastest thread count among a few candidates (one MC forward each) and report it.
    best = (None, float("inf"))
    for t in sorted({min(avail, c) for c in (8, 16, 32, 64, avail)}):
        torch.set_num_threads(t)
        oracle_mc_evaluate(net, x, 1)
        t0 = time.perf_counter()
        oracle_mc_evaluate(net, x, 1)
        dt1 = time.perf_counter() - t0
        if dt1 < best[1]:
            best = (t, dt1)
    cores = best[0]
    torch.set_num_threads(cores)
    for _ in range(warmup):
        oracle_mc_evaluate(net, x, 1)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle_mc_evaluate(net, x, REF_MC)
    dt = (time.perf_counter() - t0) / steps
    return {"value": B * REF_MC / dt, "unit": "image-samples/s", "cores": cores, "kind": "port",
            "sample": f"B={B}, {REF_MC} MC samples per step (of {N_MC}), {steps} steps, fp32, {cores} threads (best of 8/16/32/64/{avail} available); "
                      "oracle/ref_model.py = the reference's ATen op sequence",
            "ms_per_step": dt * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = time_cpu_reference(max(args.steps, 1), min(args.warmup, 2))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": cb["unit"], "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C3: dnn_to_bnn(ResNet-18, 10 classes) Reparameterization, 3x32x32, B=128",
                       "global_batch": B, "mc_samples_per_step": REF_MC},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": cb["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------- this repo
def build_model(device, dtype, fuse=True):
    import torchvision
    import bayesian_torch_b200 as btb
    torch.manual_seed(0)
    net = torchvision.models.resnet18(num_classes=N_CLASSES)
    btb.dnn_to_bnn(net, PRM)
    btb.assign_layer_keys(net)
    net = net.eval().to(device).to(dtype).to(memory_format=torch.channels_last)
    if fuse:
        btb.fuse_inference(net)      # eval-mode BatchNorm / ReLU / residual add -> conv epilogues
    return net


def run_ours(args):
    import torch.distributed as dist
    import bayesian_torch_b200 as btb
    from bayesian_torch_b200 import _native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    net = build_model(dev, dtype, fuse=not args.no_fuse)
    btb.manual_seed(0)
    torch.manual_seed(1234)
    x_host = torch.randn(B, 3, 32, 32).pin_memory()
    x_dev = x_host.to(dev).to(dtype).contiguous(memory_format=torch.channels_last)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    chunk = args.chunk

    use_graph = not args.no_graph       # the rank's whole MC pass replayed as one CUDA graph (mc.py::_MCGraph)

    def step_device(graph=use_graph):
        flush.zero_()
        return btb.mc_predict(net, x_dev, N_MC, chunk=chunk, use_graph=graph)

    def step_e2e():
        flush.zero_()
        xd = x_host.to(dev, non_blocking=True).to(dtype).contiguous(memory_format=torch.channels_last)
        mean, var = btb.mc_predict(net, xd, N_MC, chunk=chunk, use_graph=use_graph)
        return torch.stack((mean, var)).cpu()          # D2H read of the step's result (synchronises)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    if args.profile:          # under ncu: W warm-up passes + K steps of the device-resident step, nothing else
        for _ in range(args.warmup + args.steps):
            step_device(False)
        torch.cuda.synchronize()
        return
    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = _native.launch_count
    ms_step = timed(step_device, args.steps)
    launches = _native.launch_count - l0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    # ---- roofline pass: CUDA events around every fused-layer launch of `steps` more steps
    rec = []

    def hook(geom, x, mu_w, out):
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        S = geom.n_samples
        x_elems = x.numel() * (S if geom.x_shared else 1)
        m_rows = out.numel() // geom.c_out
        k = mu_w.numel() // geom.c_out
        nbytes = x.element_size() * (x_elems + out.numel()) + S * 2 * mu_w.element_size() * (mu_w.numel() + geom.c_out)
        rec.append((ev, nbytes, 2.0 * m_rows * geom.c_out * k))
        return ev

    _native.timing_hook = hook
    barrier()
    for _ in range(args.steps):
        step_device(False)          # eager: the per-launch events cannot live inside a captured graph
    torch.cuda.synchronize()
    _native.timing_hook = None
    fused_ms = sum(a.elapsed_time(b) for (a, b), _, _ in rec) / args.steps
    fused_bytes = sum(nb for _, nb, _ in rec) / args.steps
    fused_flops = sum(fl for _, _, fl in rec) / args.steps
    n_fused = len(rec) // args.steps

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm_peak, tf_peak, peak_src = _peaks()
    traffic = None            # DRAM bytes per step of the same launches from the committed ncu --set full capture
    tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as tf:
                tj = json.load(tf)
            if tj.get("launches") == n_fused:
                traffic = tj["dram_bytes_per_step"]
        except (OSError, ValueError, KeyError):
            traffic = None
    ach_gbs = fused_bytes / (fused_ms * 1e-3) / 1e9 / n_fused * n_fused   # bytes of all fused launches / their time
    value = B * N_MC / (ms_step * 1e-3)
    e2e = B * N_MC / (ms_e2e * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": "image-samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16" if dtype == torch.bfloat16 else "tf32 (fp32 parameters and activations)", "data": "synthetic",
        "config": {"workload": "C3: dnn_to_bnn(torchvision ResNet-18, 10 classes) Reparameterization, 3x32x32, "
                               "B=128, N=64 MC samples/step, samples sharded over ranks, one all-reduce of [2,B,C]",
                   "global_batch": B, "mc_samples": N_MC, "mc_chunk": chunk or "all", "epilogue_fusion": not args.no_fuse, "cuda_graph": use_graph, "parallelism": f"mc-sample-shard{world}",
                   "l2": "flushed between steps (256 MiB memset inside the timed region); per-step working set >> L2",
                   "images_per_sec_reference_style": B / (ms_step * 1e-3)},
        "e2e": {"value": e2e, "unit": "image-samples/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": 2 * B * N_CLASSES * 4},
        "gpu_launches": launches,
        "clocks": sampler.summary(),
        "roofline": {"kernel": "all Bayesian-layer launches of a step (bt_direct_kernel / bt_fused_kernel / bt_ws_kernel)",
                     "bound": "hbm",
                     "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                     "traffic": traffic, "traffic_source": "profiles/traffic.json (ncu dram__bytes_read+write, summed over "
                                                           "the same launches of one step)" if traffic else None,
                     "peak_source": peak_src,
                     "launches_per_step": n_fused, "kernel_ms_per_step": fused_ms,
                     "kernel_share_of_step": fused_ms / ms_step,
                     "algorithmic_bytes_per_step": fused_bytes,
                     "tensor": {"achieved_tflops": fused_flops / (fused_ms * 1e-3) / 1e12, "peak_tflops": tf_peak,
                                "frac": fused_flops / (fused_ms * 1e-3) / 1e12 / tf_peak}},
    }
    if world == 1 and not args.no_cpu_baseline:
        cb = time_cpu_reference(3, 1)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--chunk", type=int, default=None, help="MC samples per pass (default: all samples of the rank)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from python instead of replaying a CUDA graph")
    ap.add_argument("--no-fuse", action="store_true", help="keep BatchNorm/ReLU/residual as separate PyTorch kernels")
    ap.add_argument("--profile", action="store_true", help="profiling mode (ncu): only warmup+steps device steps, no JSON")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)

#!/usr/bin/env python
"""Per-layer roofline measurements for the BASELINE.json layer configs (C1, C2, C5 + sweep).

    python tools/bench_layers.py [--out gpurun_out/layers.json]

For each config: CUDA-event time of the fused forward (L2 flushed between iterations), algorithmic flops / bytes
per SURVEY.md 8(d), achieved TFLOP/s and GB/s against MEASURED_PEAKS.json, the stand-alone KL kernel, and the
parity error against the oracle on the draws the kernel used (materialised from the Philox counters).
The CPU reference time of the same layer (oracle/ref_model.py port, host cores) is reported beside it.
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bayesian_torch_b200 as btb  # noqa: E402
import bayesian_torch_b200.layers as L  # noqa: E402
from bayesian_torch_b200 import _native  # noqa: E402

DEV = "cuda:0"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p["hbm_gbs"], p["bf16_tflops"], p["bf16_tflops_sustained"], "measured"
    except Exception:
        return 6650.0, 1590.0, 1400.0, "fallback"


def timeit(fn, iters, flush):
    """Mean device time of fn() with a cold L2: every iteration is [256 MiB memset = L2 flush] [event] fn() [event],
    all iterations ENQUEUED back to back and synchronised once at the end -- the host runs ahead of the device
    during the flush, so the CPU-side launch latency of fn() (python + ctypes, 10-20 us) is not inside the events
    (with a synchronize per iteration it was, and dominated every sub-50-us measurement of round 1)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / iters * 1e-3


def cpu_time(det_layer, flip, x_cpu, iters=3):
    from oracle.ref_model import OracleBayesLayer
    m = OracleBayesLayer(det_layer, flip)
    torch.set_num_threads(min(32, len(os.sched_getaffinity(0))))
    with torch.no_grad():
        m(x_cpu)
        t0 = time.perf_counter()
        for _ in range(iters):
            m(x_cpu)
            m.kl_loss()
    return (time.perf_counter() - t0) / iters, torch.get_num_threads()


def parity(layer, x, y, flip):
    from gpu_util import errs, oracle_forward
    eps_w, eps_b = layer.materialize_eps(0)
    s_in = s_out = None
    if flip:
        s_in, s_out = layer.materialize_signs(tuple(x.shape), tuple(y.shape), 0)
    ref = oracle_forward(layer, x, eps_w, eps_b, s_in, s_out, round_operands=False)
    return errs(y, ref)


def run(name, layer, det, x, flops, nbytes, flip, iters, check):
    hbm, tf_burst, tf_sus, src = peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    layer.dnn_to_bnn_flag = True
    btb.manual_seed(0)
    y = layer(x)
    rec = {"config": name, "dtype_x": str(x.dtype), "dtype_p": str(layer._mu_rho()[0].dtype)}
    if check:
        layer._bt_last["sample0"] = 0
        rel, mx = parity(layer, x, y, flip)
        rec.update(parity_rel_rms=rel, parity_max_abs=mx)
    t = timeit(lambda: layer(x), iters, flush)
    mu, rho = layer._mu_rho()
    kl_bytes = 2 * mu.element_size() * (mu.numel() + (0 if layer.mu_bias is None else layer.mu_bias.numel()))
    t_kl = timeit(lambda: layer.kl_loss(), iters, flush)
    rec.update(fwd_us=t * 1e6, tflops=flops / t / 1e12, frac_tensor_burst=flops / t / 1e12 / tf_burst,
               gbs=nbytes / t / 1e9, frac_hbm=nbytes / t / 1e9 / hbm, flops=flops, algorithmic_bytes=nbytes,
               kl_us=t_kl * 1e6, kl_gbs=kl_bytes / t_kl / 1e9, kl_frac_hbm=kl_bytes / t_kl / 1e9 / hbm,
               kl_bytes=kl_bytes, peaks=src)
    if det is not None:
        ct, th = cpu_time(det, flip, x.float().cpu())
        rec.update(cpu_ref_ms=ct * 1e3, cpu_threads=th, speedup_vs_cpu=ct / (t + t_kl))
    print(json.dumps(rec), flush=True)
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    out = []
    torch.manual_seed(0)
    it = 5 if a.quick else 20

    # C1: LinearReparameterization 1024 -> 1024, batch 256, fp32
    lay = L.LinearReparameterization(1024, 1024).to(DEV)
    x = torch.randn(256, 1024, device=DEV)
    out.append(run("C1 LinearReparameterization 1024->1024 B=256 fp32", lay, torch.nn.Linear(1024, 1024), x,
                   2.0 * 256 * 1024 * 1024, 4 * (256 * 1024 * 2 + 2 * 1024 * 1024 + 2 * 1024), False, it, True))

    # C2: Conv2dFlipout 64 -> 128 k3 s1 p1, 56x56, batch 128, bf16
    lay = L.Conv2dFlipout(64, 128, 3, padding=1).to(DEV).to(torch.bfloat16)
    x = torch.randn(128, 64, 56, 56, device=DEV, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    m = 128 * 56 * 56
    out.append(run("C2 Conv2dFlipout 64->128 k3 p1 56x56 B=128 bf16", lay, None, x, 4.0 * m * 128 * 576,
                   2 * (m * 64 + m * 128 + 2 * 128 * 576 + 2 * 128), True, it, False))
    # parity of the same layer type on a reduced batch (the CPU oracle at full size takes minutes)
    xs = x[:4].contiguous(memory_format=torch.channels_last)
    btb.manual_seed(0)
    ys = lay(xs)
    lay._bt_last["sample0"] = 0
    rel, mx = parity(lay, xs, ys, True)
    out[-1].update(parity_rel_rms_B4=rel, parity_max_abs_B4=mx)
    print(json.dumps({"config": "C2 parity (B=4 slice)", "rel_rms": rel, "max_abs": mx}), flush=True)

    # C5: LinearFlipout 4096 -> 4096 bf16, M sweep
    lay = L.LinearFlipout(4096, 4096).to(DEV).to(torch.bfloat16)
    for mrows in ([4096] if a.quick else [1, 16, 128, 1024, 4096]):
        x = torch.randn(mrows, 4096, device=DEV, dtype=torch.bfloat16)
        out.append(run(f"C5 LinearFlipout 4096->4096 B={mrows} bf16", lay, None, x, 4.0 * mrows * 4096 * 4096,
                       2 * (mrows * 4096 * 2 + 2 * 4096 * 4096 + 2 * 4096), True, it, mrows <= 128))
    # reparameterization at the same size (one GEMM, sampled weights)
    lay = L.LinearReparameterization(4096, 4096).to(DEV).to(torch.bfloat16)
    x = torch.randn(4096, 4096, device=DEV, dtype=torch.bfloat16)
    out.append(run("LinearReparameterization 4096->4096 B=4096 bf16", lay, None, x, 2.0 * 4096 ** 3,
                   2 * (4096 * 4096 * 2 + 2 * 4096 * 4096 + 2 * 4096), False, it, False))
    # KL on fp32 parameters of the same size (8 bytes per element instead of 4)
    lay = L.LinearReparameterization(4096, 4096).to(DEV)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    t_kl = timeit(lambda: lay.kl_loss(), it, flush)
    kb = 2 * 4 * (4096 * 4096 + 4096)
    rec = {"config": "KL only, fp32 parameters 4096x4096", "kl_us": t_kl * 1e6, "kl_gbs": kb / t_kl / 1e9,
           "kl_frac_hbm": kb / t_kl / 1e9 / peaks()[0], "kl_bytes": kb}
    print(json.dumps(rec), flush=True)
    out.append(rec)
    if a.out:
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()

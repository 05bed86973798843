"""GPU bring-up diagnostic (not a pytest file): exact-integer GEMMs through bt_layer_forward with eps forced
to 0, so any mismatch is a layout / descriptor bug, and its structure is dumped to gpurun_out/diag.npz."""
import os
import sys
import traceback

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesian_torch_b200.layers as L  # noqa: E402

DEV = "cuda:0"
OUT = {}


def run(tag, M, K, N, flip=False, xdt=torch.float32):
    torch.manual_seed(0)
    cls = L.LinearFlipout if flip else L.LinearReparameterization
    layer = cls(K, N, bias=False).to(DEV)
    with torch.no_grad():
        layer.mu_weight.copy_(torch.randint(-3, 4, (N, K)).float())
        layer.rho_weight.fill_(-50.0)
    x = torch.randint(-3, 4, (M, K)).float().to(DEV).to(xdt)
    dbg = {"eps_w_in": torch.zeros(N, K, device=DEV)}
    if flip:
        dbg["sign_in"] = torch.ones(M, K, device=DEV)
        dbg["sign_out"] = torch.ones(M, N, device=DEV)
    y = layer._forward_impl(x, False, debug=dbg)
    torch.cuda.synchronize()
    ref = x.float().cpu() @ layer.mu_weight.detach().cpu().t()
    bad = (y.float().cpu() != ref)
    print(f"[diag] {tag}: M={M} K={K} N={N} flip={flip} mismatches={int(bad.sum())}/{bad.numel()} "
          f"maxabs={float((y.float().cpu() - ref).abs().max()):.3f}", flush=True)
    if bad.any():
        OUT[tag + "_y"] = y.float().cpu().numpy()
        OUT[tag + "_ref"] = ref.numpy()
        OUT[tag + "_x"] = x.float().cpu().numpy()
        OUT[tag + "_w"] = layer.mu_weight.detach().cpu().numpy()
        rows = bad.any(1).nonzero().flatten()[:10].tolist()
        cols = bad.any(0).nonzero().flatten()[:10].tolist()
        print(f"        first bad rows {rows} cols {cols}", flush=True)
    return int(bad.sum())


if __name__ == "__main__":
    os.makedirs("gpurun_out", exist_ok=True)
    total = 0
    try:
        print(torch.cuda.get_device_name(0), flush=True)
        total += run("t1", 128, 64, 128)
        total += run("t2", 128, 256, 128)
        total += run("t3", 128, 64, 64)
        total += run("t4", 512, 192, 128)
        total += run("t5", 100, 72, 40)
        total += run("t6", 128, 64, 128, flip=True)
        total += run("t7", 300, 320, 200, flip=True, xdt=torch.bfloat16)
        total += run("t8", 4096, 1024, 512)
    except Exception:
        traceback.print_exc()
        total += 1
    if OUT:
        np.savez_compressed("gpurun_out/diag.npz", **OUT)
    print("[diag] total mismatches", total, flush=True)
    sys.exit(0 if total == 0 else 1)

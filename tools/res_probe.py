"""timing probe: layer1-shaped 3x3 conv (64->64, 8x8, B=128, S=64 MC samples, bf16) with / without the residual epilogue"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesian_torch_b200 as btb
import bayesian_torch_b200.layers as L
from bayesian_torch_b200 import _native
DEV = "cuda:0"
torch.manual_seed(0)
conv = L.Conv2dReparameterization(64, 64, 3, padding=1, bias=False).to(DEV).bfloat16()
conv._bt_ep_scale, conv._bt_ep_shift, conv._bt_ep_relu = torch.rand(64, device=DEV) + 0.5, torch.randn(64, device=DEV), True
S, B = 64, 128
x = torch.randn(S * B, 64, 8, 8, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
res = torch.randn(S * B, 64, 8, 8, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
def run(residual, probe):
    os.environ["BT_DYNAMIC_ENV"] = "1"; os.environ["BT_TMA_PROBE"] = str(probe)
    ts = []
    with btb.mc_sample_context(S, B, 0):
        for i in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            conv._forward_impl(x, False, residual=res if residual else None)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2], _native.last_forward_path()
for residual, probe in ((False, 0), (True, 0), (True, 2), (True, 3)):
    print("residual", residual, "probe", probe, "us", run(residual, probe))

"""Run ONE layer shape through the direct kernel a few times (for ncu captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesian_torch_b200 as btb  # noqa: E402
import bayesian_torch_b200.layers as L  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "layer1"
cfg = {"layer1": (False, 64, 64, (8, 8), 128, 64), "layer2": (False, 128, 128, (4, 4), 128, 64),
       "c2": (True, 64, 128, (56, 56), 128, 1), "layer3": (False, 256, 256, (2, 2), 128, 64)}[which]
flip, cin, cout, sp, B, S = cfg
cls = L.Conv2dFlipout if flip else L.Conv2dReparameterization
torch.manual_seed(0)
layer = cls(cin, cout, 3, padding=1, bias=False).to("cuda:0").bfloat16()
x = torch.randn(B, cin, *sp, device="cuda:0").bfloat16().contiguous(memory_format=torch.channels_last)
with btb.mc_sample_context(S, B, 0):
    for _ in range(4):
        y = layer(x, return_kl=False)
torch.cuda.synchronize()
print(which, btb._native.last_forward_path(), float(y.float().abs().mean()))

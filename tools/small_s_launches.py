"""one eager MC pass with S samples (what one rank of a 64/S-GPU job does) -- run under
   ncu --metrics gpu__time_duration.sum to get the per-launch times.  usage: small_s_launches.py S"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from small_s_check import build, DEV
import bayesian_torch_b200 as btb
S = int(sys.argv[1])
net = build()
x = torch.randn(128, 3, 32, 32, device=DEV).bfloat16()
for _ in range(2):
    btb.mc_predict(net, x, S, use_graph=False)
torch.cuda.synchronize()

"""C1 (LinearReparameterization 1024->1024, B=256, fp32) timing with the cluster form forced on / off"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesian_torch_b200 as btb
import bayesian_torch_b200.layers as L
from bayesian_torch_b200 import _native
DEV = "cuda:0"
os.environ["BT_DYNAMIC_ENV"] = "1"
torch.manual_seed(0)
lay = L.LinearReparameterization(1024, 1024).to(DEV)
x = torch.randn(256, 1024, device=DEV)
for name, env in (("default", {}), ("force", {"BT_FORCE_CLUSTER": "1"}), ("off", {"BT_DISABLE_CLUSTER": "1"}), ("force mode2 bn128", {"BT_FORCE_CLUSTER": "1", "BT_TMA_BN": "128"})):
    for k in ("BT_FORCE_CLUSTER", "BT_DISABLE_CLUSTER", "BT_TMA_BN"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ts = []
    for i in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); y = lay(x, return_kl=False); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    g = lay._bt_last["geom"]
    print(name, [round(t, 1) for t in ts], _native.last_forward_path(), {k: v for k, v in _native.plan_forward(0, g, torch.float32, torch.float32).items() if k in ("block_n", "m_subtiles", "grid", "cluster_n")}, flush=True)

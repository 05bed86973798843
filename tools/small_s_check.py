"""What one rank of an N-GPU run does, on ONE GPU: MC inference with S = 64/N samples (N = 1, 2, 4, 8, 16).
For every S: (a) parity of the default tiling choices against the im2col kernels (BT_DISABLE_DIRECT=1) on the same
draws, (b) device time per step and the kernel family each Bayesian layer took.  Shows how far the per-(n-tile, sample)
CTA granularity lets strong scaling go."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesian_torch_b200 as btb  # noqa: E402
from bayesian_torch_b200 import _native  # noqa: E402

DEV = "cuda:0"


def build():
    import torchvision
    torch.manual_seed(0)
    net = torchvision.models.resnet18(num_classes=10)
    btb.dnn_to_bnn(net, {"prior_mu": 0.0, "prior_sigma": 1.0, "posterior_mu_init": 0.0, "posterior_rho_init": -3.0,
                         "type": "Reparameterization", "moped_enable": False, "moped_delta": 0.5})
    btb.assign_layer_keys(net)
    net = net.eval().to(DEV).bfloat16().to(memory_format=torch.channels_last)
    return btb.fuse_inference(net)


def main():
    net = build()
    B = 128
    x = torch.randn(B, 3, 32, 32, device=DEV).bfloat16()
    paths = []
    hooks = [m.register_forward_hook(lambda mod, i, o: paths.append(_native.last_forward_path()))
             for m in net.modules() if isinstance(m, btb._core.BayesLayerBase)]
    res = []
    for S in (64, 32, 16, 8, 4):
        os.environ.pop("BT_DISABLE_DIRECT", None)
        btb.manual_seed(3)
        paths.clear()
        mean, var = btb.mc_predict(net, x, S)
        fam = {}
        for p in paths:
            fam[p] = fam.get(p, 0) + 1
        os.environ["BT_DISABLE_DIRECT"] = "1"
        btb.manual_seed(3)
        mean2, var2 = btb.mc_predict(net, x, S)
        os.environ.pop("BT_DISABLE_DIRECT", None)
        dm = float((mean - mean2).abs().max())
        ms = {}
        for graph in (False, True):
            for _ in range(3):
                btb.mc_predict(net, x, S, use_graph=graph)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                btb.mc_predict(net, x, S, use_graph=graph)
            e1.record()
            torch.cuda.synchronize()
            ms["graph" if graph else "eager"] = e0.elapsed_time(e1) / 10
        rec = dict(S=S, ms_per_step_eager=ms["eager"], ms_per_step_graph=ms["graph"],
                   image_samples_per_s=B * S / ms["graph"] * 1e3, max_abs_dmean_vs_im2col=dm, families=fam, ok=dm < 5e-3)
        print(rec, flush=True)
        res.append(rec)
    for h in hooks:
        h.remove()
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"), indent=1)
    sys.exit(0 if all(r["ok"] for r in res) else 1)


if __name__ == "__main__":
    main()

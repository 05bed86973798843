"""Run under torchrun on >= 2 GPUs (not a pytest file; CPU coverage of the same logic is tests/test_mc_gloo.py):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
        tests/multi_gpu_check.py

Checks that MC inference sharded over ranks (contiguous blocks of GLOBAL sample indices + ONE NCCL all-reduce of
the [2,B,C] moment buffer) equals the single-rank result, and that every rank ends with identical tensors."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesian_torch_b200 as btb  # noqa: E402
from bayesian_torch_b200 import _native  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    import torchvision
    torch.manual_seed(0)
    net = torchvision.models.resnet18(num_classes=10)
    btb.dnn_to_bnn(net, {"prior_mu": 0.0, "prior_sigma": 1.0, "posterior_mu_init": 0.0, "posterior_rho_init": -3.0,
                         "type": "Reparameterization", "moped_enable": False, "moped_delta": 0.5})
    btb.assign_layer_keys(net)
    net = net.eval().to(dev).bfloat16().to(memory_format=torch.channels_last)
    btb.fuse_inference(net)
    btb.manual_seed(123)
    torch.manual_seed(1)
    B, N = 32, 12
    x = torch.randn(B, 3, 32, 32, device=dev).bfloat16()
    mean, var, pe, mi = btb.mc_predict(net, x, N, return_uncertainty=True)   # sharded: N / world samples per rank + ONE all-reduce
    # single-rank reference on every rank: all N samples locally, no collective
    with torch.no_grad(), btb.mc_sample_context(N, B, 0):
        logits = net(x)
    sums = torch.empty(2, B, 10, device=dev)
    ent = torch.empty(B, device=dev)
    _native.mc_accumulate(logits.contiguous(), N, B, sums, False, entropy_sum=ent)
    m1, v1 = torch.empty(B, 10, device=dev), torch.empty(B, 10, device=dev)
    _native.mc_finalize(sums, N, m1, v1)
    pe1, mi1 = torch.empty(B, device=dev), torch.empty(B, device=dev)
    _native.mc_uncertainty(sums, ent, N, pe1, mi1)
    err = (float((mean - m1).abs().max()), float((var - v1).abs().max()),
           float((pe - pe1).abs().max()), float((mi - mi1).abs().max()))
    gathered = [torch.empty_like(mean) for _ in range(world)]
    dist.all_gather(gathered, mean)
    same = all(torch.equal(g, gathered[0]) for g in gathered)
    ok = err[0] < 1e-5 and err[1] < 1e-5 and err[2] < 1e-4 and err[3] < 1e-4 and same
    print(f"[rank {rank}/{world}] sharded-vs-single max|dmean|={err[0]:.2e} max|dvar|={err[1]:.2e} "
          f"max|dH|={err[2]:.2e} max|dMI|={err[3]:.2e} identical_on_all_ranks={same} "
          f"{'OK' if ok else 'FAIL'}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

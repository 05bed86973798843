"""CPU, world_size 2 over gloo: the N>1 host logic of MC inference -- contiguous sharding of the global
sample indices and the ONE all-reduce of the [2,B,C] moment buffer (bayesian_torch_b200/mc.py).  The model
forward is replaced by seeded oracle logits (the fused kernels need a GPU); what is tested is that the
sharded + all-reduced result equals the single-process result for any world size."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bayesian_torch_b200.mc import _moment_buffer, all_reduce_moments, shard_samples
from oracle import bt_oracle as O

N, B, C = 7, 5, 10


def _logits_for(sample_idx):
    g = torch.Generator().manual_seed(1000 + sample_idx)   # keyed by the GLOBAL sample index
    return torch.randn(B, C, generator=g) * 2


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        start, count = shard_samples(N, world, rank)
        buf, sums, ent = _moment_buffer(B, C, torch.device("cpu"), with_entropy=True, zero=True)
        for s in range(start, start + count):
            p = torch.softmax(_logits_for(s), -1)
            sums[0] += p
            sums[1] += p * p
            ent += O.entropy(p)            # what bt_mc_accumulate_ex adds per sample
        all_reduce_moments(buf)            # ONE collective carries the moments and the entropy sums
        mean = sums[0] / N
        var = sums[1] / N - mean * mean
        pred_entropy = O.entropy(mean)     # what bt_mc_uncertainty computes from the reduced buffer
        ret[rank] = (mean, var, pred_entropy, pred_entropy - ent / N)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_mc_equals_single_process(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    full = torch.stack([_logits_for(s) for s in range(N)])
    mean, var = O.mc_aggregate(full)
    probs = torch.softmax(full, -1)
    for r in range(world):
        m, v, pe, mi = ret[r]
        assert torch.allclose(m, mean, atol=1e-6) and torch.allclose(v, var, atol=1e-6)
        assert torch.allclose(pe, O.predictive_entropy(probs), atol=1e-5)
        assert torch.allclose(mi, O.mutual_information(probs), atol=1e-5)
        assert torch.equal(m, ret[0][0])   # identical on all ranks


def test_all_reduce_is_noop_without_process_group():
    t = torch.ones(2, 3, 4)
    assert all_reduce_moments(t) is t and float(t.sum()) == 24.0

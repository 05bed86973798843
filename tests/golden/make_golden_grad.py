#!/usr/bin/env python
"""Mint GRADIENT golden vectors from the reference's own autograd (IntelLabs/bayesian-torch, /root/reference) for the
backward oracle (oracle/bt_oracle_grad.py).  Build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_grad.py

For every case: loss = sum(out * dy) + kl_weight * kl  with a seeded upstream gradient dy; stored are the inputs
(x, parameters, the eps / sign draws recovered as in make_golden.py, dy, kl_weight) and d loss / d (x, mu, rho, bias
parameters).  Writes tests/golden/grads.npz + grads_meta.json.
"""
import json
import os
import sys

sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

import numpy as np
import torch

import bayesian_torch.layers as L  # the REFERENCE package

assert L.__file__.startswith("/root/reference"), L.__file__
HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(1)
out, meta = {}, {"cases": {}}
KLW = 0.37


def case(name, nd, flipout, bias, seed, cin, cout, k=None, spatial=(), stride=1, padding=0, dilation=1, groups=1, batch=3):
    torch.manual_seed(seed)
    if nd == 0:
        cls = L.LinearFlipout if flipout else L.LinearReparameterization
        m = cls(cin, cout, prior_mean=0.1, prior_variance=0.7, posterior_mu_init=0.05, posterior_rho_init=-2.5, bias=bias)
        x = torch.randn(batch, cin, requires_grad=True)
        wn = "weight"
    else:
        cls = getattr(L, f"Conv{nd}d" + ("Flipout" if flipout else "Reparameterization"))
        m = cls(cin, cout, k, stride=stride, padding=padding, dilation=dilation, groups=groups, bias=bias,
                prior_mean=0.0, prior_variance=1.3, posterior_mu_init=0.0, posterior_rho_init=-3.0)
        x = torch.randn(batch, cin, *spatial, requires_grad=True)
        wn = "kernel"
    torch.manual_seed(seed + 1000)
    y, kl = m(x)
    dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(seed + 7))
    ((y * dy).sum() + KLW * kl).backward()
    mu_w, rho_w = getattr(m, f"mu_{wn}"), getattr(m, f"rho_{wn}")
    t = {"x": x, "dy": dy, "mu_w": mu_w, "rho_w": rho_w, "eps_w": getattr(m, f"eps_{wn}"),
         "dx": x.grad, "dmu_w": mu_w.grad, "drho_w": rho_w.grad}
    if bias:
        t.update(mu_b=m.mu_bias, rho_b=m.rho_bias, eps_b=m.eps_bias, dmu_b=m.mu_bias.grad, drho_b=m.rho_bias.grad)
    if flipout:      # the sign tensors are locals of forward(): replay the draw order (see make_golden.py)
        torch.manual_seed(seed + 1000)
        if nd == 0:
            torch.empty_like(m.eps_weight).normal_()
            if bias:
                torch.empty_like(m.eps_bias).normal_()
        t["sign_in"] = x.detach().clone().uniform_(-1, 1).sign()
        t["sign_out"] = y.detach().clone().uniform_(-1, 1).sign()
    for kk, v in t.items():
        out[f"{name}/{kk}"] = v.detach().numpy()
    meta["cases"][name] = {"nd": nd, "flipout": flipout, "bias": bias, "stride": stride, "padding": padding,
                           "dilation": dilation, "groups": groups, "kl_weight": KLW,
                           "prior_mean": 0.1 if nd == 0 else 0.0, "prior_variance": 0.7 if nd == 0 else 1.3}


seed = 500
for flip in (False, True):
    t = "flip" if flip else "rep"
    case(f"linear_{t}_b", 0, flip, True, seed, 24, 20, batch=5); seed += 1
    case(f"linear_{t}_nb", 0, flip, False, seed, 16, 12, batch=4); seed += 1
    case(f"conv1d_{t}", 1, flip, True, seed, 8, 12, 3, (13,), 2, 1, 1, 1); seed += 1
    case(f"conv2d_{t}_a", 2, flip, True, seed, 8, 12, 3, (7, 7), 1, 1, 1, 1); seed += 1
    case(f"conv2d_{t}_b", 2, flip, False, seed, 8, 8, 3, (9, 8), 2, 2, 2, 2); seed += 1
    case(f"conv3d_{t}", 3, flip, True, seed, 4, 6, 3, (4, 5, 5), 1, 1, 1, 1, batch=2); seed += 1

np.savez_compressed(os.path.join(HERE, "grads.npz"), **out)
with open(os.path.join(HERE, "grads_meta.json"), "w") as f:
    json.dump(meta, f, indent=1)
print(f"wrote {len(out)} tensors for {len(meta['cases'])} cases")

#!/usr/bin/env python
"""Mint golden vectors from the REFERENCE ITSELF (IntelLabs/bayesian-torch, /root/reference).

The reference ships no tests / KATs, so parity is pinned on
outputs of the reference modules replayed here under a fixed seed.  Run in the
build container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Writes tests/golden/layers.npz (per-case tensors) and tests/golden/meta.json
(scalars, state_dict keys).  eps_* are read back from the module's buffers after
forward (they hold the draw that was used, linear_variational.py:161,173); the
Flipout sign tensors are local variables in the reference, so they are recovered
by re-seeding and replaying the documented draw order
(linear_flipout.py:150,162,169-170; conv_flipout.py:385-386,390,401), and the
script asserts that the replay reproduces the module output bit-exactly.
"""
import json
import os
import sys

sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

import bayesian_torch.layers as L  # the REFERENCE package
from bayesian_torch.models.dnn_to_bnn import dnn_to_bnn, get_kl_loss
from bayesian_torch.utils.util import get_rho

assert L.__file__.startswith("/root/reference"), L.__file__
HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(1)

out = {}
meta = {"cases": {}, "reference_commit": "aa7e57b", "torch": torch.__version__}


def sp(rho):
    return torch.log1p(torch.exp(rho))


def put(case, **tensors):
    for k, v in tensors.items():
        if v is None:
            continue
        out[f"{case}/{k}"] = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)


def params_of(m, conv):
    w = "kernel" if conv else "weight"
    d = {
        "mu_w": getattr(m, f"mu_{w}"), "rho_w": getattr(m, f"rho_{w}"),
        "eps_w": getattr(m, f"eps_{w}"),
        "mu_b": m.mu_bias, "rho_b": m.rho_bias, "eps_b": m.eps_bias,
    }
    return d


def linear_case(name, flipout, bias, seed, fin=48, fout=40, batch=5):
    torch.manual_seed(seed)
    cls = L.LinearFlipout if flipout else L.LinearReparameterization
    m = cls(fin, fout, prior_mean=0.1, prior_variance=0.7, posterior_mu_init=0.05,
            posterior_rho_init=-2.5, bias=bias)
    x = torch.randn(batch, fin)
    torch.manual_seed(seed + 1000)
    y, kl = m(x)
    p = params_of(m, conv=False)
    extra = {}
    if flipout:
        torch.manual_seed(seed + 1000)
        torch.empty_like(m.eps_weight).normal_()
        if bias:
            torch.empty_like(m.eps_bias).normal_()
        s_in = x.clone().uniform_(-1, 1).sign()
        s_out = y.clone().uniform_(-1, 1).sign()
        dw = sp(m.rho_weight) * m.eps_weight
        b = sp(m.rho_bias) * m.eps_bias if bias else None
        y2 = F.linear(x, m.mu_weight, m.mu_bias) + F.linear(x * s_in, dw, b) * s_out
        assert torch.equal(y2, y), name
        extra = {"sign_in": s_in, "sign_out": s_out}
    else:
        w = m.mu_weight + sp(m.rho_weight) * m.eps_weight
        b = m.mu_bias + sp(m.rho_bias) * m.eps_bias if bias else None
        assert torch.equal(F.linear(x, w, b), y), name
    put(name, x=x, y=y, kl=kl, kl_loss=m.kl_loss(), **p, **extra)
    meta["cases"][name] = {"kind": "linear", "flipout": flipout, "bias": bias,
                           "prior_mean": 0.1, "prior_variance": 0.7,
                           "state_dict_keys": list(m.state_dict().keys())}


def conv_case(name, nd, flipout, bias, seed, cin, cout, k, spatial, stride, padding, dilation, groups, batch=2):
    torch.manual_seed(seed)
    cls = getattr(L, f"Conv{nd}d" + ("Flipout" if flipout else "Reparameterization"))
    kw = dict(prior_mean=0.0, prior_variance=1.3, posterior_mu_init=0.0, posterior_rho_init=-3.0)
    m = cls(cin, cout, k, stride=stride, padding=padding, dilation=dilation, groups=groups, bias=bias, **kw)
    x = torch.randn(batch, cin, *spatial)
    torch.manual_seed(seed + 1000)
    y, kl = m(x)
    conv = {1: F.conv1d, 2: F.conv2d, 3: F.conv3d}[nd]
    p = params_of(m, conv=True)
    extra = {}
    if flipout:
        torch.manual_seed(seed + 1000)
        s_in = x.clone().uniform_(-1, 1).sign()
        s_out = y.clone().uniform_(-1, 1).sign()
        dk = sp(m.rho_kernel) * m.eps_kernel
        b = sp(m.rho_bias) * m.eps_bias if bias else None
        y2 = conv(x, m.mu_kernel, m.mu_bias, stride, padding, dilation, groups) + \
            conv(x * s_in, dk, b, stride, padding, dilation, groups) * s_out
        assert torch.equal(y2, y), name
        extra = {"sign_in": s_in, "sign_out": s_out}
    else:
        w = m.mu_kernel + sp(m.rho_kernel) * m.eps_kernel
        b = m.mu_bias + sp(m.rho_bias) * m.eps_bias if bias else None
        assert torch.equal(conv(x, w, b, stride, padding, dilation, groups), y), name
    put(name, x=x, y=y, kl=kl, kl_loss=m.kl_loss(), **p, **extra)
    meta["cases"][name] = {"kind": "conv", "nd": nd, "flipout": flipout, "bias": bias,
                           "stride": stride, "padding": padding, "dilation": dilation, "groups": groups,
                           "prior_mean": 0.0, "prior_variance": 1.3,
                           "state_dict_keys": list(m.state_dict().keys())}


seed = 100
for flip in (False, True):
    for bias in (True, False):
        linear_case(f"linear_{'flip' if flip else 'rep'}_{'b' if bias else 'nb'}", flip, bias, seed)
        seed += 1
for flip in (False, True):
    t = "flip" if flip else "rep"
    conv_case(f"conv1d_{t}_a", 1, flip, True, seed, 8, 12, 3, (20,), 2, 1, 1, 1, batch=3); seed += 1
    conv_case(f"conv1d_{t}_b", 1, flip, False, seed, 8, 8, 5, (17,), 1, 2, 2, 2, batch=2); seed += 1
    conv_case(f"conv2d_{t}_a", 2, flip, True, seed, 16, 24, 3, (9, 9), 1, 1, 1, 1); seed += 1
    conv_case(f"conv2d_{t}_b", 2, flip, False, seed, 8, 16, 3, (11, 10), 2, 2, 2, 2); seed += 1
    conv_case(f"conv2d_{t}_c", 2, flip, True, seed, 3, 10, 5, (12, 12), 2, 2, 1, 1); seed += 1
    conv_case(f"conv2d_{t}_1x1", 2, flip, False, seed, 64, 32, 1, (4, 4), 2, 0, 1, 1); seed += 1
    conv_case(f"conv3d_{t}_a", 3, flip, True, seed, 4, 8, (2, 3, 3), (5, 6, 6), 1, 1, 1, 1); seed += 1

# ---- C1-sized KL anchors (BASELINE.json configs[0]): only scalars are stored; the test
# re-creates the parameters with the same seed through the drop-in layer's own init
# (same draw order as linear_variational.py:135-142).
torch.manual_seed(0)
m = L.LinearReparameterization(1024, 1024)
x = torch.randn(256, 1024)
y, kl = m(x)
meta["c1"] = {"kl_forward": float(kl), "kl_loss": float(m.kl_loss()),
              "mu_sum": float(m.mu_weight.double().sum()), "rho_sum": float(m.rho_weight.double().sum()),
              "y_abs_mean": float(y.abs().mean())}
torch.manual_seed(0)
m = L.LinearFlipout(256, 128)
meta["c1_flip"] = {"kl_loss": float(m.kl_loss()), "mu_sum": float(m.mu_weight.double().sum())}
torch.manual_seed(0)
m = L.Conv2dReparameterization(8, 16, 3)
meta["c1_conv"] = {"kl_loss": float(m.kl_loss()), "mu_sum": float(m.mu_kernel.double().sum()),
                   "posterior_mu_init": list(m.posterior_mu_init)}
torch.manual_seed(0)
m = L.Conv3dReparameterization(4, 8, 3, 0, 1, 0, -3.0)
meta["c1_conv3d"] = {"kl_loss": float(m.kl_loss()), "mu_sum": float(m.mu_kernel.double().sum())}

# ---- dnn_to_bnn surface (models/dnn_to_bnn.py:127-165) on a small CNN and torchvision resnet18


class TinyCNN(nn.Module):
    def __init__(self):
        super().__init__()
        self.features = nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.BatchNorm2d(8), nn.ReLU(),
                                      nn.Conv2d(8, 16, 3, stride=2, padding=1, bias=False), nn.ReLU())
        self.conv3d = nn.Conv3d(2, 4, 3)
        self.fc = nn.Linear(16, 10)

    def forward(self, x):
        return self.fc(self.features(x).mean((2, 3)))


for typ in ("Reparameterization", "Flipout"):
    prm = {"prior_mu": 0.0, "prior_sigma": 1.0, "posterior_mu_init": 0.0, "posterior_rho_init": -3.0,
           "type": typ, "moped_enable": False, "moped_delta": 0.5}
    torch.manual_seed(7)
    net = TinyCNN()
    dnn_to_bnn(net, prm)
    meta[f"tiny_{typ}"] = {
        "state_dict": {k: list(v.shape) for k, v in net.state_dict().items()},
        "classes": {n: type(mod).__name__ for n, mod in net.named_modules() if hasattr(mod, "kl_loss")},
        "kl": float(get_kl_loss(net)),
    }
    # MOPED init (dnn_to_bnn.py:65-71,95-101)
    prm2 = dict(prm, moped_enable=True, moped_delta=0.3)
    torch.manual_seed(7)
    net = TinyCNN()
    w_det = net.fc.weight.detach().clone()
    dnn_to_bnn(net, prm2)
    assert torch.equal(net.fc.mu_weight.data, w_det)
    put(f"moped_{typ}", w=w_det, rho=net.fc.rho_weight.data)
    meta[f"tiny_moped_{typ}"] = {"kl": float(get_kl_loss(net))}

import torchvision

torch.manual_seed(11)
net = torchvision.models.resnet18(num_classes=10)
prm = {"prior_mu": 0.0, "prior_sigma": 1.0, "posterior_mu_init": 0.0, "posterior_rho_init": -3.0,
       "type": "Reparameterization", "moped_enable": False, "moped_delta": 0.5}
dnn_to_bnn(net, prm)
bayes = {n: type(mod).__name__ for n, mod in net.named_modules() if hasattr(mod, "kl_loss")}
n_pairs = sum(p.numel() for n, p in net.named_parameters() if n.split(".")[-1].startswith("mu_"))
meta["resnet18"] = {"n_bayes_layers": len(bayes), "n_mu": n_pairs, "kl": float(get_kl_loss(net)),
                    "state_dict_keys": list(net.state_dict().keys())}
net.eval()
with torch.no_grad():
    yy = net(torch.randn(2, 3, 32, 32))
meta["resnet18"]["out_shape"] = list(yy.shape)

# get_rho (utils/util.py:63-69)
torch.manual_seed(3)
w = torch.randn(64) * 0.2
put("get_rho", w=w, rho=get_rho(w, 0.5))

# MC aggregation as in examples/main_bayesian_cifar_dnn2bnn.py:545-557
torch.manual_seed(5)
logits = torch.randn(6, 4, 10) * 3
probs = torch.nn.functional.softmax(torch.stack([l for l in logits]), dim=2)
put("mc", logits=logits, mean=probs.mean(0), pred=probs.mean(0).argmax(-1))

np.savez_compressed(os.path.join(HERE, "layers.npz"), **out)
with open(os.path.join(HERE, "meta.json"), "w") as f:
    json.dump(meta, f, indent=1, sort_keys=True)
print("wrote", len(out), "arrays;", os.path.getsize(os.path.join(HERE, "layers.npz")), "bytes")

#!/usr/bin/env python
"""Mint golden vectors from the reference for the "next" layer families: ConvTranspose{1,2,3}d and
LSTM, Reparameterization and Flipout.  Build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_next.py

eps / sign draws are recovered by re-seeding and replaying the documented draw order of the reference forward (the script
asserts that its replay reproduces the module output bit-exactly).  ConvTranspose*Flipout are run with return_kl=False
when in_channels != out_channels: the reference allocates their prior sigma with the wrong shape
(conv_flipout.py:706-709, 906) and its KL raises there.  Writes tests/golden/next.npz + next_meta.json.
"""
import json
import os
import sys

sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

import numpy as np
import torch
import torch.nn.functional as F

import bayesian_torch.layers as L  # the REFERENCE package

assert L.__file__.startswith("/root/reference"), L.__file__
HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(1)
out, meta = {}, {"cases": {}}
sp = lambda r: torch.log1p(torch.exp(r))
CT = {1: F.conv_transpose1d, 2: F.conv_transpose2d, 3: F.conv_transpose3d}


def put(name, **t):
    for k, v in t.items():
        if v is not None:
            out[f"{name}/{k}"] = v.detach().numpy()


def convt_case(name, nd, flip, bias, seed, cin, cout, k, spatial, stride, padding, output_padding, dilation, groups):
    torch.manual_seed(seed)
    cls = getattr(L, f"ConvTranspose{nd}d" + ("Flipout" if flip else "Reparameterization"))
    m = cls(cin, cout, k, stride=stride, padding=padding, output_padding=output_padding, dilation=dilation, groups=groups,
            bias=bias, prior_mean=0.0, prior_variance=1.1, posterior_mu_init=0.0, posterior_rho_init=-3.0)
    x = torch.randn(2, cin, *spatial)
    torch.manual_seed(seed + 1000)
    with_kl = not (flip and cin != cout)
    res = m(x, return_kl=with_kl)
    y = res[0] if with_kl else res
    geo = (stride, padding, output_padding, groups, dilation)
    extra = {}
    if flip:
        mean = CT[nd](x, m.mu_kernel, m.mu_bias, *geo)
        torch.manual_seed(seed + 1000)
        s_in = x.clone().uniform_(-1, 1).sign()
        s_out = mean.clone().uniform_(-1, 1).sign()
        b = sp(m.rho_bias) * m.eps_bias if bias else None
        y2 = mean + CT[nd](x * s_in, sp(m.rho_kernel) * m.eps_kernel, b, *geo) * s_out
        extra = dict(sign_in=s_in, sign_out=s_out)
    else:
        b = m.mu_bias + sp(m.rho_bias) * m.eps_bias if bias else None
        y2 = CT[nd](x, m.mu_kernel + sp(m.rho_kernel) * m.eps_kernel, b, *geo)
    assert torch.equal(y2, y), name
    put(name, x=x, y=y, mu_w=m.mu_kernel, rho_w=m.rho_kernel, eps_w=m.eps_kernel, mu_b=m.mu_bias, rho_b=m.rho_bias,
        eps_b=m.eps_bias if bias else None, kl=res[1] if with_kl else None, **extra)
    meta["cases"][name] = dict(kind="convt", nd=nd, flipout=flip, bias=bias, stride=stride, padding=padding,
                               output_padding=output_padding, dilation=dilation, groups=groups, with_kl=with_kl)


def lstm_case(name, flip, seed, fin, hid, T, batch):
    torch.manual_seed(seed)
    cls = L.LSTMFlipout if flip else L.LSTMReparameterization
    m = cls(fin, hid, prior_mean=0.0, prior_variance=1.0, posterior_mu_init=0.0, posterior_rho_init=-3.0, bias=True)
    x = torch.randn(batch, T, fin)
    torch.manual_seed(seed + 1000)
    hseq, (_, cseq), kl = m(x)
    # replay the draws: per time step ih then hh; each Linear draws eps_w, eps_b (then, Flipout, sign_in, sign_out)
    torch.manual_seed(seed + 1000)
    h = torch.zeros(batch, hid)
    c = torch.zeros(batch, hid)
    t_out = {}
    for t in range(T):
        gates = 0
        for tag, lin, inp in (("ih", m.ih, x[:, t, :]), ("hh", m.hh, h)):
            ew = torch.empty_like(lin.mu_weight).normal_()
            eb = torch.empty_like(lin.mu_bias).normal_()
            if flip:
                mean = F.linear(inp, lin.mu_weight, lin.mu_bias)
                s_in = inp.clone().uniform_(-1, 1).sign()
                s_out = mean.clone().uniform_(-1, 1).sign()
                o = mean + F.linear(inp * s_in, sp(lin.rho_weight) * ew, sp(lin.rho_bias) * eb) * s_out
                t_out[f"t{t}_{tag}_sign_in"], t_out[f"t{t}_{tag}_sign_out"] = s_in, s_out
            else:
                o = F.linear(inp, lin.mu_weight + sp(lin.rho_weight) * ew, lin.mu_bias + sp(lin.rho_bias) * eb)
            t_out[f"t{t}_{tag}_eps_w"], t_out[f"t{t}_{tag}_eps_b"] = ew, eb
            gates = gates + o
        i, f = torch.sigmoid(gates[:, :hid]), torch.sigmoid(gates[:, hid:2 * hid])
        g, o_ = torch.tanh(gates[:, 2 * hid:3 * hid]), torch.sigmoid(gates[:, 3 * hid:])
        c = f * c + i * g
        h = o_ * torch.tanh(c)
    assert torch.equal(h, hseq[:, -1]) and torch.equal(c, cseq[:, -1]), name
    put(name, x=x, hseq=hseq, cseq=cseq, kl=kl, ih_mu_w=m.ih.mu_weight, ih_rho_w=m.ih.rho_weight, ih_mu_b=m.ih.mu_bias,
        ih_rho_b=m.ih.rho_bias, hh_mu_w=m.hh.mu_weight, hh_rho_w=m.hh.rho_weight, hh_mu_b=m.hh.mu_bias,
        hh_rho_b=m.hh.rho_bias, **t_out)
    meta["cases"][name] = dict(kind="lstm", flipout=flip, T=T, hidden=hid)


seed = 900
for flip in (False, True):
    t = "flip" if flip else "rep"
    convt_case(f"convt1d_{t}", 1, flip, True, seed, 8, 8, 3, (11,), 2, 1, 1, 1, 1); seed += 1
    convt_case(f"convt2d_{t}_a", 2, flip, True, seed, 8, 12, 3, (5, 6), 2, 1, 1, 1, 1); seed += 1
    convt_case(f"convt2d_{t}_b", 2, flip, False, seed, 8, 8, 4, (6, 6), 2, 1, 0, 1, 2); seed += 1
    convt_case(f"convt3d_{t}", 3, flip, True, seed, 4, 4, 3, (3, 4, 4), 1, 1, 0, 1, 1); seed += 1
    lstm_case(f"lstm_{t}", flip, seed, 10, 12, 5, 3); seed += 1

np.savez_compressed(os.path.join(HERE, "next.npz"), **out)
with open(os.path.join(HERE, "next_meta.json"), "w") as f:
    json.dump(meta, f, indent=1)
print(f"wrote {len(out)} tensors for {len(meta['cases'])} cases")

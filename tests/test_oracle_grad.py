"""CPU: the backward restatement (oracle/bt_oracle_grad.py) against gradients minted from the
reference's own autograd (tests/golden/make_golden_grad.py -> grads.npz).  Tolerance: fp32, different summation order in
the closed-form weight / input gradients than in autograd's kernels -> |d| <= 2e-5 * max|ref| (stated)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import bt_oracle_grad as OG

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_Z = np.load(os.path.join(ROOT, "tests", "golden", "grads.npz"))
with open(os.path.join(ROOT, "tests", "golden", "grads_meta.json")) as _f:
    _META = json.load(_f)["cases"]


def _case(name):
    pre = name + "/"
    return {k[len(pre):]: torch.from_numpy(_Z[k]) for k in _Z.files if k.startswith(pre)}


@pytest.mark.parametrize("name", sorted(_META))
def test_backward_restatement_equals_reference_autograd(name):
    c, m = _case(name), _META[name]
    geo = dict(stride=m["stride"], padding=m["padding"], dilation=m["dilation"], groups=m["groups"])
    bias = dict(mu_b=c.get("mu_b"), rho_b=c.get("rho_b"), eps_b=c.get("eps_b"))
    if m["flipout"]:
        g = OG.flipout_backward(m["nd"], c["x"], c["mu_w"], c["rho_w"], c["eps_w"], c["sign_in"], c["sign_out"], c["dy"],
                                **bias, **geo)
    else:
        g = OG.reparam_backward(m["nd"], c["x"], c["mu_w"], c["rho_w"], c["eps_w"], c["dy"], **bias, **geo)
    # the loss also held kl_weight * KL (layers return it from forward): add its closed-form gradient
    ps = m["prior_variance"]          # (the reference passes prior_variance as the prior sigma, linear_variational.py:93)
    dmu, drho = OG.kl_grad(c["mu_w"], c["rho_w"], m["prior_mean"], ps)
    g["dmu_w"] = g["dmu_w"] + m["kl_weight"] * dmu
    g["drho_w"] = g["drho_w"] + m["kl_weight"] * drho
    if m["bias"]:
        dmu, drho = OG.kl_grad(c["mu_b"], c["rho_b"], m["prior_mean"], ps)
        g["dmu_b"] = g["dmu_b"] + m["kl_weight"] * dmu
        g["drho_b"] = g["drho_b"] + m["kl_weight"] * drho
    for k, v in g.items():
        ref = c[k]
        assert v.shape == ref.shape, (k, v.shape, ref.shape)
        err = float((v - ref).abs().max())
        assert err <= 2e-5 * max(1.0, float(ref.abs().max())), (name, k, err, float(ref.abs().max()))

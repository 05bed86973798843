"""-m gpu: the training path (bayesian_torch_b200/_autograd.py).

(1) gradients of  loss = sum(out * dy) + kl_weight * kl  through the fused forward + regenerated draws, against the
    gradients minted from the REFERENCE's own autograd (tests/golden/grads.npz; reference draws injected through the
    debug hooks): rel-RMS <= 2e-3 on dx (the forward / data-gradient products see tf32 operands), <= 1e-4 on the
    parameter gradients of the KL-only part and <= 2e-3 overall (stated; cuDNN TF32 is switched off in this test so the
    ATen backward products are fp32).
(2) on-chip draws: backward == the oracle backward (oracle/bt_oracle_grad.py) fed with the re-materialised eps / signs.
(3) a reference-style training loop (examples/main_bayesian_cifar_dnn2bnn.py:404-420) on a converted model learns.
(4) what is NOT differentiable says so: MC-sample context, fused inference epilogue."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

import bayesian_torch_b200 as btb
from gpu_util import build_layer, cl, errs, note, phys_eps
from oracle import bt_oracle_grad as OG

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_Z = np.load(os.path.join(ROOT, "tests", "golden", "grads.npz"))
with open(os.path.join(ROOT, "tests", "golden", "grads_meta.json")) as _f:
    _META = json.load(_f)["cases"]


def _case(name):
    pre = name + "/"
    return {k[len(pre):]: torch.from_numpy(_Z[k]) for k in _Z.files if k.startswith(pre)}


@pytest.fixture(autouse=True)
def _fp32_aten_backward():
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("name", sorted(_META))
def test_gradients_match_reference_autograd(name):
    c, m = _case(name), _META[name]
    nd, flip = m["nd"], m["flipout"]
    mu = c["mu_w"]
    if nd == 0:
        layer = build_layer("linear", 0, flip, mu.shape[1], mu.shape[0], None, bias=m["bias"],
                            prior_mean=m["prior_mean"], prior_variance=m["prior_variance"])
        w = "weight"
    else:
        layer = build_layer("conv", nd, flip, mu.shape[1] * m["groups"], mu.shape[0], tuple(mu.shape[2:]), m["stride"],
                            m["padding"], m["dilation"], m["groups"], m["bias"], m["prior_mean"], m["prior_variance"])
        w = "kernel"
    sd = {f"mu_{w}": mu, f"rho_{w}": c["rho_w"]}
    if m["bias"]:
        sd.update(mu_bias=c["mu_b"], rho_bias=c["rho_b"])
    layer.load_state_dict(sd)
    layer = layer.to(DEV)
    x = c["x"].to(DEV).requires_grad_(True)
    dbg = {"eps_w_in": phys_eps(c["eps_w"]).to(DEV)}
    if m["bias"]:
        dbg["eps_b_in"] = c["eps_b"].to(DEV)
    if flip:
        if nd == 0:
            dbg["sign_in"], dbg["sign_out"] = c["sign_in"].to(DEV).contiguous(), c["sign_out"].to(DEV).contiguous()
        else:
            dbg["sign_in"], dbg["sign_out"] = cl(c["sign_in"]).to(DEV), cl(c["sign_out"]).to(DEV)
    y, kl = layer._forward_impl(x, True, debug=dbg)
    assert y.grad_fn is not None and kl.grad_fn is not None
    ((y * c["dy"].to(DEV)).sum() + m["kl_weight"] * kl).backward()
    got = {"dx": x.grad, "dmu_w": getattr(layer, f"mu_{w}").grad, "drho_w": getattr(layer, f"rho_{w}").grad}
    if m["bias"]:
        got.update(dmu_b=layer.mu_bias.grad, drho_b=layer.rho_bias.grad)
    for k, v in got.items():
        assert v is not None and v.shape == c[k].shape, k
        rel, mx = errs(v, c[k])
        note("grad_golden", case=name, which=k, rel=rel, max_abs=mx)
        # dx goes through W (or mu / Delta) rounded by nothing here -- ATen fp32 -- so every gradient is fp32-exact
        # up to summation order; the forward's tf32 rounding does not enter the gradients of this loss
        assert rel <= 1e-4, (name, k, rel, mx)


@pytest.mark.parametrize("flip", [False, True], ids=["R", "F"])
@pytest.mark.parametrize("pdt", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_backward_regenerates_the_on_chip_draws(flip, pdt):
    torch.manual_seed(4)
    btb.manual_seed(77)
    layer = build_layer("conv", 2, flip, 16, 24, 3, 2, 1, 1, 2, True).to(DEV).to(pdt)
    x = torch.randn(5, 16, 9, 7, device=DEV).to(pdt).requires_grad_(True)
    y, kl = layer(x)
    dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(3)).to(DEV).to(pdt)
    ((y.float() * dy.float()).sum() + 0.5 * kl.float()).backward()
    layer._bt_last["sample0"] = 0
    eps_w, eps_b = layer.materialize_eps(0)
    f = lambda t: t.detach().float().cpu()
    kw = dict(mu_b=f(layer.mu_bias), rho_b=f(layer.rho_bias), eps_b=f(eps_b), stride=2, padding=1, dilation=1, groups=2)
    if flip:
        s_in, s_out = layer.materialize_signs(tuple(x.shape), tuple(y.shape), 0)
        ref = OG.flipout_backward(2, f(x), f(layer.mu_kernel), f(layer.rho_kernel), f(eps_w), f(s_in), f(s_out), f(dy), **kw)
    else:
        ref = OG.reparam_backward(2, f(x), f(layer.mu_kernel), f(layer.rho_kernel), f(eps_w), f(dy), **kw)
    dmu, drho = OG.kl_grad(f(layer.mu_kernel), f(layer.rho_kernel), 0.0, 1.0)
    ref["dmu_w"] = ref["dmu_w"] + 0.5 * dmu
    ref["drho_w"] = ref["drho_w"] + 0.5 * drho
    dmu, drho = OG.kl_grad(f(layer.mu_bias), f(layer.rho_bias), 0.0, 1.0)
    ref["dmu_b"] = ref["dmu_b"] + 0.5 * dmu
    ref["drho_b"] = ref["drho_b"] + 0.5 * drho
    got = {"dx": x.grad, "dmu_w": layer.mu_kernel.grad, "drho_w": layer.rho_kernel.grad, "dmu_b": layer.mu_bias.grad,
           "drho_b": layer.rho_bias.grad}
    tol = 1e-4 if pdt == torch.float32 else 1.5e-2      # bf16: gradients are produced and stored in bf16
    for k, v in got.items():
        rel, mx = errs(v, ref[k])
        note("grad_onchip", flip=flip, dtype=str(pdt), which=k, rel=rel)
        assert rel <= tol, (k, rel, mx)


def test_reference_style_training_loop_learns():
    """dnn_to_bnn(MLP) + the loop of examples/main_bayesian_cifar_dnn2bnn.py:404-420: output = model(x);
    kl = get_kl_loss(model); loss = CE + kl / batch; loss.backward(); optimizer.step()."""
    torch.manual_seed(0)
    btb.manual_seed(0)
    net = nn.Sequential(nn.Linear(32, 64), nn.ReLU(), nn.Linear(64, 4))
    btb.dnn_to_bnn(net, {"prior_mu": 0.0, "prior_sigma": 1.0, "posterior_mu_init": 0.0, "posterior_rho_init": -3.0,
                         "type": "Reparameterization", "moped_enable": False, "moped_delta": 0.5})
    net = net.to(DEV).train()
    xs = torch.randn(256, 32, device=DEV)
    ys = (xs[:, :4].argmax(1)).long()
    opt = torch.optim.Adam(net.parameters(), lr=1e-2)
    losses = []
    for _ in range(60):
        opt.zero_grad()
        out = net(xs)
        kl = btb.get_kl_loss(net)
        loss = nn.functional.cross_entropy(out, ys) + kl / xs.shape[0]
        loss.backward()
        for p in net.parameters():
            assert p.grad is not None and torch.isfinite(p.grad).all()
        opt.step()
        losses.append(float(loss))
    note("train_loop", first=losses[0], last=losses[-1])
    assert losses[-1] < 0.6 * losses[0], (losses[0], losses[-1])
    with torch.no_grad():
        acc = float((net(xs).argmax(1) == ys).float().mean())
    assert acc > 0.8, acc


def test_non_differentiable_modes_raise():
    layer = build_layer("conv", 2, False, 64, 64, 3, 1, 1).to(DEV)
    x = torch.randn(2, 64, 4, 4, device=DEV)
    with btb.mc_sample_context(2, 2, 0):          # the MC context switches autograd off (inference feature)
        y = layer(x, return_kl=False)
    assert y.grad_fn is None and y.shape[0] == 4
    layer._bt_ep_relu = True
    with pytest.raises(RuntimeError, match="not differentiable"):
        layer(x)
    with torch.no_grad():
        y = layer(x, return_kl=False)
    assert y.grad_fn is None

"""Helpers shared by the -m gpu parity tests (kernel vs oracle on identical mu, rho, eps, signs, x)."""
import torch

from oracle import bt_oracle as O

import bayesian_torch_b200.layers as L


def cl(t):
    """logical [N, C, *sp] -> dense physical [N, *sp, C] copy (what the C ABI consumes)."""
    return t.permute(0, *range(2, t.dim()), 1).contiguous()


def phys_eps(eps):
    """reference-layout eps [Cout, Cin/g, *k] -> physical [Cout, taps * Cin/g] (tap-major, channel-minor)."""
    if eps.dim() == 2:
        return eps.contiguous()
    return cl(eps).reshape(eps.shape[0], -1).contiguous()


def errs(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    d = (a - b)
    rel_rms = float(d.pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp_min(1e-30))
    return rel_rms, float(d.abs().max())


def build_layer(kind, nd, flip, cin, cout, ks, stride=1, padding=0, dilation=1, groups=1, bias=True,
                prior_mean=0.0, prior_variance=1.0):
    if kind == "linear":
        cls = L.LinearFlipout if flip else L.LinearReparameterization
        return cls(cin, cout, prior_mean=prior_mean, prior_variance=prior_variance, bias=bias)
    cls = getattr(L, f"Conv{nd}d" + ("Flipout" if flip else "Reparameterization"))
    if nd == 3 and not flip:
        return cls(cin, cout, ks, prior_mean, prior_variance, 0, -3.0, stride=stride, padding=padding,
                   dilation=dilation, groups=groups, bias=bias)
    return cls(cin, cout, ks, stride=stride, padding=padding, dilation=dilation, groups=groups,
               prior_mean=prior_mean, prior_variance=prior_variance, bias=bias)


def layer_params(layer):
    mu_w, rho_w = layer._mu_rho()
    return (mu_w.detach().float().cpu(), rho_w.detach().float().cpu(),
            None if layer.mu_bias is None else layer.mu_bias.detach().float().cpu(),
            None if layer.rho_bias is None else layer.rho_bias.detach().float().cpu())


def oracle_forward(layer, x, eps_w, eps_b, sign_in=None, sign_out=None, round_operands=False, path=None):
    """Oracle output for `layer` on CPU fp32.  round_operands=True applies the operand rounding of the tensor-core
    path for this dtype combination and kernel family (`path`, default: the family of the thread's last launch --
    oracle/bt_oracle.py::operand_rounding) for a tight comparison; "bf16" / "tf32" / "tf32_xtrunc" force one."""
    mu_w, rho_w, mu_b, rho_b = layer_params(layer)
    x_dtype = x.dtype
    x = x.detach().float().cpu()
    eps_w = eps_w.detach().float().cpu()
    eps_b = None if eps_b is None else eps_b.detach().float().cpu()
    flip = layer._family == "flipout"
    nd = layer._nd
    if round_operands is True:   # the rounding the kernel applies for this dtype combination
        from bayesian_torch_b200 import _native
        round_operands = O.operand_rounding(x_dtype, layer._mu_rho()[0].dtype, path or _native.last_forward_path())
    r = {False: (lambda t: t), "bf16": O.round_operand, "tf32": O.round_operand_tf32,
         "tf32_xtrunc": O.round_operand_tf32}[round_operands]
    rx = O.truncate_operand_tf32 if round_operands == "tf32_xtrunc" else r
    sig = O.sigma_of_rho(rho_w)
    if nd == 0:
        conv = lambda a, w, b: torch.nn.functional.linear(a, w, b)
    else:
        f = {1: torch.nn.functional.conv1d, 2: torch.nn.functional.conv2d, 3: torch.nn.functional.conv3d}[nd]
        conv = lambda a, w, b: f(a, w, b, layer.stride, layer.padding, layer.dilation, layer.groups)
    if not flip:
        w = r(mu_w + sig * eps_w)
        b = None if mu_b is None else mu_b + O.sigma_of_rho(rho_b) * eps_b
        return conv(rx(x), w, b)
    sign_in = sign_in.detach().float().cpu()
    sign_out = sign_out.detach().float().cpu()
    b = None if mu_b is None else O.sigma_of_rho(rho_b) * eps_b
    return conv(rx(x), r(mu_w), mu_b) + conv(rx(x) * sign_in, r(sig * eps_w), b) * sign_out


def note(name, **vals):
    """append measured values (parity errors, paths) to gpurun_out/parity.jsonl so tolerances are set from evidence"""
    import json
    import os
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "parity.jsonl"), "a") as f:
            f.write(json.dumps({"test": name, **{k: (float(v) if isinstance(v, (int, float)) else v) for k, v in vals.items()}}) + "\n")

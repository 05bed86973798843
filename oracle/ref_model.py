"""CPU port of the reference's Bayesian layers as nn.Modules, for TIMING the reference path on host
cores (bench.py `cpu_baseline` / `--impl reference`) and for whole-model checks.  TEST / BENCH
INFRASTRUCTURE ONLY (see oracle/__init__.py); never imported by the product.

Each module draws eps / signs with ATen's RNG exactly where the reference does
(/root/reference/bayesian_torch/layers/variational_layers/conv_variational.py:361-380,
 linear_variational.py:160-178, flipout_layers/conv_flipout.py:376-417, linear_flipout.py:149-174)
and evaluates the restated arithmetic of oracle/bt_oracle.py -- i.e. the same ATen op sequence the
reference executes on CPU (exp, log1p, normal_, mul, add, conv/linear [, uniform_, sign, 2nd conv]).
"""
import torch
import torch.nn as nn

from . import bt_oracle as O


class OracleBayesLayer(nn.Module):
    def __init__(self, det, flipout, posterior_rho_init=-3.0, posterior_mu_init=0.0):
        super().__init__()
        self.flipout = flipout
        self.is_conv = not isinstance(det, nn.Linear)
        w = det.weight.detach()
        self.mu_w = nn.Parameter(torch.empty_like(w).normal_(posterior_mu_init, 0.1))
        self.rho_w = nn.Parameter(torch.empty_like(w).normal_(posterior_rho_init, 0.1))
        if det.bias is not None:
            self.mu_b = nn.Parameter(torch.empty_like(det.bias).normal_(posterior_mu_init, 0.1))
            self.rho_b = nn.Parameter(torch.empty_like(det.bias).normal_(posterior_rho_init, 0.1))
        else:
            self.mu_b = self.rho_b = None
        if self.is_conv:
            self.nd = w.dim() - 2
            self.args = dict(stride=det.stride, padding=det.padding, dilation=det.dilation, groups=det.groups)

    def kl_loss(self):
        return O.kl_loss(self.mu_w, self.rho_w, 0.0, 1.0, self.mu_b, self.rho_b)

    def forward(self, x):
        if not self.flipout:
            eps_w = torch.empty_like(self.mu_w).normal_()
            eps_b = None if self.mu_b is None else torch.empty_like(self.mu_b).normal_()
            if self.is_conv:
                return O.conv_reparam(self.nd, x, self.mu_w, self.rho_w, eps_w, self.mu_b, self.rho_b, eps_b, **self.args)
            return O.linear_reparam(x, self.mu_w, self.rho_w, eps_w, self.mu_b, self.rho_b, eps_b)
        # flipout.  RNG draw order differs between the two reference files:
        #   linear_flipout.py:150,162,169-170 : eps_w, eps_b, sign_in, sign_out
        #   conv_flipout.py:385-386,390,401   : sign_in, sign_out, eps_w, eps_b
        if self.is_conv:
            conv = {1: torch.nn.functional.conv1d, 2: torch.nn.functional.conv2d, 3: torch.nn.functional.conv3d}[self.nd]
            outputs = conv(x, self.mu_w, self.mu_b, **self.args)
            sign_in = x.clone().uniform_(-1, 1).sign()
            sign_out = outputs.clone().uniform_(-1, 1).sign()
            eps_w = torch.empty_like(self.mu_w).normal_()
            eps_b = None if self.mu_b is None else torch.empty_like(self.mu_b).normal_()
            b = None if self.mu_b is None else O.sigma_of_rho(self.rho_b) * eps_b
            pert = conv(x * sign_in, O.sigma_of_rho(self.rho_w) * eps_w, b, **self.args)   # conv_flipout.py:409-417
            return outputs + pert * sign_out
        eps_w = torch.empty_like(self.mu_w).normal_()
        eps_b = None if self.mu_b is None else torch.empty_like(self.mu_b).normal_()
        outputs = torch.nn.functional.linear(x, self.mu_w, self.mu_b)
        sign_in = x.clone().uniform_(-1, 1).sign()
        sign_out = outputs.clone().uniform_(-1, 1).sign()
        b = None if self.mu_b is None else O.sigma_of_rho(self.rho_b) * eps_b
        pert = torch.nn.functional.linear(x * sign_in, O.sigma_of_rho(self.rho_w) * eps_w, b)   # linear_flipout.py:171-174
        return outputs + pert * sign_out


def oracle_dnn_to_bnn(m, flipout=False):
    """In-place surgery with the matching rule of models/dnn_to_bnn.py:127-154 (class-name substring)."""
    for name, child in list(m._modules.items()):
        if child is None:
            continue
        if child._modules:
            oracle_dnn_to_bnn(child, flipout)
        elif "Conv" in type(child).__name__ or "Linear" in type(child).__name__:
            setattr(m, name, OracleBayesLayer(child, flipout))
    return m


@torch.no_grad()
def oracle_mc_evaluate(model, x, n_mc):
    """evaluate() of examples/main_bayesian_cifar_dnn2bnn.py:541-557 for one batch: N sequential forwards,
    stack, softmax, mean over the MC dimension."""
    outs = [model(x) for _ in range(n_mc)]
    return O.mc_aggregate(torch.stack(outs))

"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the BACKWARD of the reference layers (the
product path is inference-only in round 1, this is the checker a backward kernel will be held to).

The reference gets its gradients from autograd through the op sequences of
  layers/variational_layers/linear_variational.py:157-201, conv_variational.py:183-227 / 357-402 / 530-574,
  layers/flipout_layers/linear_flipout.py:145-197,        conv_flipout.py:175-244 / 370-439 / 568-637,
  layers/base_variational_layer.py:53-68 (kl_div);
here the same derivatives are written in closed form, with eps / signs as explicit arguments (a fused backward kernel
regenerates them from the Philox key instead of storing W):

  sigma = softplus(rho),  dsigma/drho = sigmoid(rho)
  reparam:  W = mu + sigma*eps                     flipout:  D = sigma*eps
    dW   = wgrad(x, dy)                              dmu  = wgrad(x, dy)
    dmu  = dW,  drho = dW * eps * sigmoid(rho)       dD   = wgrad(x*s_in, dy*s_out),  drho = dD * eps * sigmoid(rho)
    dx   = igrad(dy, W)                              dx   = igrad(dy, mu) + igrad(dy*s_out, D) * s_in
    dmu_b = sum dy,  drho_b = sum dy * eps_b * sigmoid(rho_b)      (flipout: the rho_b term sees dy*s_out)
  KL (mean over the n elements of a tensor, prior N(pm, ps)):
    dKL/dmu = (mu - pm) / (ps^2 n),   dKL/drho = (sigma/ps^2 - 1/sigma) * sigmoid(rho) / n

Pinned on gradients minted from the reference's own autograd: tests/golden/make_golden_grad.py -> tests/golden/grads.npz,
checked by tests/test_oracle_grad.py.
"""
import torch
import torch.nn.functional as F
from torch.nn import grad as G

from .bt_oracle import sigma_of_rho


def _wgrad(nd, x, w_shape, dy, stride, padding, dilation, groups):
    if nd == 0:
        return dy.reshape(-1, dy.shape[-1]).t() @ x.reshape(-1, x.shape[-1])
    f = {1: G.conv1d_weight, 2: G.conv2d_weight, 3: G.conv3d_weight}[nd]
    return f(x, w_shape, dy, stride, padding, dilation, groups)


def _igrad(nd, x_shape, w, dy, stride, padding, dilation, groups):
    if nd == 0:
        return dy @ w
    f = {1: G.conv1d_input, 2: G.conv2d_input, 3: G.conv3d_input}[nd]
    return f(x_shape, w, dy, stride, padding, dilation, groups)


def _bsum(nd, t):
    return t.sum(tuple(i for i in range(t.dim()) if i != (t.dim() - 1 if nd == 0 else 1)))


def kl_grad(mu, rho, prior_mu, prior_sigma):
    """d mean-KL / d(mu, rho) of one tensor (base_variational_layer.py:53-68 under autograd)."""
    n = mu.numel()
    sigma = sigma_of_rho(rho)
    ps2 = prior_sigma * prior_sigma
    return (mu - prior_mu) / (ps2 * n), (sigma / ps2 - 1.0 / sigma) * torch.sigmoid(rho) / n


def reparam_backward(nd, x, mu_w, rho_w, eps_w, dy, mu_b=None, rho_b=None, eps_b=None,
                     stride=1, padding=0, dilation=1, groups=1):
    """-> dict(dx, dmu_w, drho_w[, dmu_b, drho_b]) of  sum(out * dy)  for the Reparameterization layers."""
    w = mu_w + sigma_of_rho(rho_w) * eps_w
    dw = _wgrad(nd, x, w.shape, dy, stride, padding, dilation, groups)
    out = {"dx": _igrad(nd, x.shape, w, dy, stride, padding, dilation, groups),
           "dmu_w": dw, "drho_w": dw * eps_w * torch.sigmoid(rho_w)}
    if mu_b is not None:
        db = _bsum(nd, dy)
        out["dmu_b"] = db
        out["drho_b"] = db * eps_b * torch.sigmoid(rho_b)
    return out


def flipout_backward(nd, x, mu_w, rho_w, eps_w, sign_in, sign_out, dy, mu_b=None, rho_b=None, eps_b=None,
                     stride=1, padding=0, dilation=1, groups=1):
    """-> dict(dx, dmu_w, drho_w[, dmu_b, drho_b]) of  sum(out * dy)  for the Flipout layers."""
    d = sigma_of_rho(rho_w) * eps_w
    dys = dy * sign_out
    dd = _wgrad(nd, x * sign_in, d.shape, dys, stride, padding, dilation, groups)
    out = {"dx": _igrad(nd, x.shape, mu_w, dy, stride, padding, dilation, groups) +
                 _igrad(nd, x.shape, d, dys, stride, padding, dilation, groups) * sign_in,
           "dmu_w": _wgrad(nd, x, mu_w.shape, dy, stride, padding, dilation, groups),
           "drho_w": dd * eps_w * torch.sigmoid(rho_w)}
    if mu_b is not None:
        out["dmu_b"] = _bsum(nd, dy)
        out["drho_b"] = _bsum(nd, dys) * eps_b * torch.sigmoid(rho_b)
    return out

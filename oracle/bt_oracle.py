"""Functional CPU restatement of the bayesian-torch stochastic-layer hot path.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Every function takes the
random draws (eps, signs) as explicit arguments, so that a CUDA kernel and the
reference can be compared on IDENTICAL (mu, rho, eps, signs, x).

The arithmetic itself lives in PyTorch (requirements.txt:1 pins torch>=1.7,
container has 2.11.0): F.linear / F.convNd / log1p / exp / log, exactly the
calls the reference makes.  Pinned against the reference by
tests/golden/make_golden.py -> tests/test_oracle_golden.py.

All paths relative to /root/reference/bayesian_torch/.
"""
import torch
import torch.nn.functional as F

_CONV = {1: F.conv1d, 2: F.conv2d, 3: F.conv3d}


def sigma_of_rho(rho):
    """layers/variational_layers/linear_variational.py:160  sigma = log1p(exp(rho))."""
    return torch.log1p(torch.exp(rho))


def kl_div(mu_q, sigma_q, mu_p, sigma_p):
    """layers/base_variational_layer.py:53-68 -- closed-form KL(N(mu_q,sigma_q)||N(mu_p,sigma_p)), MEAN over elements."""
    mu_p = torch.as_tensor(mu_p, dtype=mu_q.dtype)
    sigma_p = torch.as_tensor(sigma_p, dtype=mu_q.dtype)
    kl = torch.log(sigma_p) - torch.log(sigma_q) + (sigma_q ** 2 + (mu_q - mu_p) ** 2) / (2 * (sigma_p ** 2)) - 0.5
    return kl.mean()


def kl_loss(mu_w, rho_w, prior_mu, prior_sigma, mu_b=None, rho_b=None, prior_b_mu=None, prior_b_sigma=None):
    """layers/variational_layers/linear_variational.py:144-155 (same body in every layer class):
    mean-KL of the weight plus mean-KL of the bias."""
    kl = kl_div(mu_w, sigma_of_rho(rho_w), prior_mu, prior_sigma)
    if mu_b is not None:
        kl = kl + kl_div(mu_b, sigma_of_rho(rho_b),
                         prior_mu if prior_b_mu is None else prior_b_mu,
                         prior_sigma if prior_b_sigma is None else prior_b_sigma)
    return kl


def linear_reparam(x, mu_w, rho_w, eps_w, mu_b=None, rho_b=None, eps_b=None):
    """layers/variational_layers/linear_variational.py:157-201 (forward, quant observers excluded)."""
    w = mu_w + sigma_of_rho(rho_w) * eps_w
    b = None
    if mu_b is not None:
        b = mu_b + sigma_of_rho(rho_b) * eps_b
    return F.linear(x, w, b)


def linear_flipout(x, mu_w, rho_w, eps_w, sign_in, sign_out, mu_b=None, rho_b=None, eps_b=None):
    """layers/flipout_layers/linear_flipout.py:145-197."""
    delta_w = sigma_of_rho(rho_w) * eps_w
    b = None
    if mu_b is not None:
        b = sigma_of_rho(rho_b) * eps_b
    outputs = F.linear(x, mu_w, mu_b)
    pert = F.linear(x * sign_in, delta_w, b) * sign_out
    return outputs + pert


def conv_reparam(nd, x, mu_k, rho_k, eps_k, mu_b=None, rho_b=None, eps_b=None,
                 stride=1, padding=0, dilation=1, groups=1):
    """layers/variational_layers/conv_variational.py:183-227 (1d) / 357-402 (2d) / 530-574 (3d)."""
    w = mu_k + sigma_of_rho(rho_k) * eps_k
    b = None
    if mu_b is not None:
        b = mu_b + sigma_of_rho(rho_b) * eps_b
    return _CONV[nd](x, w, b, stride, padding, dilation, groups)


def conv_flipout(nd, x, mu_k, rho_k, eps_k, sign_in, sign_out, mu_b=None, rho_b=None, eps_b=None,
                 stride=1, padding=0, dilation=1, groups=1):
    """layers/flipout_layers/conv_flipout.py:175-244 (1d) / 370-439 (2d) / 568-637 (3d)."""
    outputs = _CONV[nd](x, mu_k, mu_b, stride, padding, dilation, groups)
    delta_k = sigma_of_rho(rho_k) * eps_k
    b = None
    if mu_b is not None:
        b = sigma_of_rho(rho_b) * eps_b
    pert = _CONV[nd](x * sign_in, delta_k, b, stride, padding, dilation, groups) * sign_out
    return outputs + pert


def get_rho(sigma, delta):
    """utils/util.py:63-69."""
    return torch.log(torch.expm1(delta * torch.abs(sigma)) + 1e-20)


def mc_aggregate(logits):
    """examples/main_bayesian_cifar_dnn2bnn.py:545-557 -- logits [N_mc, B, C]:
    softmax over classes, mean over the MC dimension (+ the second moment the
    B200 path all-reduces so the predictive variance can be formed)."""
    p = torch.softmax(logits.float(), dim=-1)
    mean = p.mean(0)
    var = (p * p).mean(0) - mean * mean
    return mean, var


def round_operand(t, dtype=torch.bfloat16):
    """Operand rounding the tcgen05 kind::f16 path applies (bf16 operands, fp32 accumulate)."""
    return t.to(dtype).to(torch.float32)


def round_operand_tf32(t):
    """Operand rounding of the tcgen05 kind::tf32 path (fp32 parameters with fp32 activations): the kernels apply
    cvt.rna.tf32.f32 -- round to nearest, ties away from zero, to a 10-bit mantissa -- before the tensor core."""
    bits = t.contiguous().to(torch.float32).view(torch.int32)
    mag = (bits & 0x7FFFFFFF) + 0x1000
    out = (bits & -0x80000000) | (mag & ~0x1FFF)          # (float bits are sign-magnitude: rounds |t| half-up)
    return out.view(torch.float32).view_as(t)


def truncate_operand_tf32(t):
    """What tcgen05.mma kind::tf32 does with an fp32 word that was NOT pre-rounded: the low 13 mantissa bits are
    ignored (truncation toward zero).  The TMA kernel families stage fp32 activations as they are in memory, so their
    activation operand WOULD see this (measured on B200: rel-RMS 4.6e-4 per layer instead of 2.9e-4, and a systematic
    shrink that compounds over a deep network) -- which is why their converter warps round the staged tile to nearest
    first (csrc/bt_tma.cuh::tm_round_tile_tf32).  Kept as the statement of the hardware behaviour."""
    bits = t.contiguous().to(torch.float32).view(torch.int32)
    return (bits & ~0x1FFF).view(torch.float32).view_as(t)


def operand_rounding(x_dtype, p_dtype, path=None):
    """which operand rounding bt_layer_forward applies: "bf16" (any bf16 operand: kind::f16) or "tf32" (fp32 x + fp32
    parameters: x and W rounded to nearest, cvt.rna.tf32.f32) -- include/btb200.h"""
    if x_dtype == torch.float32 and p_dtype == torch.float32:
        return "tf32"        # every kernel family rounds both operands to nearest (the TMA kernels round the staged
    return "bf16"            # activation tile in shared memory before the tensor core reads it)


# ---------------------------------------------------------------- MC-ensemble uncertainties (utils/util.py:41-60)
def entropy(prob):
    """-sum(p log(p + 1e-15)) over the last axis   (reference utils/util.py:41-42, numpy there, torch here)"""
    return -(prob * torch.log(prob + 1e-15)).sum(-1)


def predictive_entropy(mc_preds):
    """entropy of the mean over the MC axis 0   (reference utils/util.py:45-50)"""
    return entropy(mc_preds.mean(0))


def mutual_information(mc_preds):
    """H(mean_s p_s) - mean_s H(p_s)   (reference utils/util.py:53-60)"""
    return entropy(mc_preds.mean(0)) - entropy(mc_preds).mean(0)

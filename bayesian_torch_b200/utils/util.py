"""`get_rho` of the drop-in surface (reference: /root/reference/bayesian_torch/utils/util.py:63-69).
Used once per layer at conversion time by dnn_to_bnn(moped_enable=True); plain tensor math."""
import torch

__all__ = ["get_rho"]


def get_rho(sigma, delta):
    """rho such that softplus(rho) = delta * |sigma| (MOPED init); +1e-20 keeps log finite at 0."""
    return torch.log(torch.expm1(delta * torch.abs(sigma)) + 1e-20)

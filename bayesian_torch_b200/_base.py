"""BaseVariationalLayer_ of the drop-in API (reference: /root/reference/bayesian_torch/layers/
base_variational_layer.py:35-68).  Kept dependency-free so user code can subclass it."""
import collections.abc
from itertools import repeat

import torch
import torch.nn as nn


def get_kernel_size(x, n):
    """base_variational_layer.py:35-38: int -> n-tuple, iterable -> tuple."""
    if isinstance(x, collections.abc.Iterable):
        return tuple(x)
    return tuple(repeat(x, n))


class BaseVariationalLayer_(nn.Module):
    def __init__(self):
        super().__init__()
        self._dnn_to_bnn_flag = False

    @property
    def dnn_to_bnn_flag(self):
        return self._dnn_to_bnn_flag

    @dnn_to_bnn_flag.setter
    def dnn_to_bnn_flag(self, value):
        self._dnn_to_bnn_flag = value

    def kl_div(self, mu_q, sigma_q, mu_p, sigma_p):
        """KL(N(mu_q, sigma_q) || N(mu_p, sigma_p)), MEAN over elements (base_variational_layer.py:53-68).
        Generic tensor-in/tensor-out API method; the B200 layers themselves never call it --
        kl_loss() / forward(return_kl=True) run the fused rho-based kernel (csrc/bt_kl.cu)."""
        kl = torch.log(sigma_p) - torch.log(sigma_q) + (sigma_q ** 2 + (mu_q - mu_p) ** 2) / (2 * (sigma_p ** 2)) - 0.5
        return kl.mean()

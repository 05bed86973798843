"""LinearReparameterization on the fused B200 kernel.

API mirror of /root/reference/bayesian_torch/layers/variational_layers/linear_variational.py:54-201
(constructor, mu_weight/rho_weight/mu_bias/rho_bias, forward(x, return_kl) contract, kl_loss());
the forward itself is ONE sm_100a launch (csrc/bt_fused.cu) instead of the reference's 11-38
ATen launches."""
from ..._core import BayesLinearBase

__all__ = ["LinearReparameterization"]


class LinearReparameterization(BayesLinearBase):
    _family = "reparam"

    def __init__(self, in_features, out_features, prior_mean=0, prior_variance=1, posterior_mu_init=0,
                 posterior_rho_init=-3.0, bias=True):
        super().__init__()
        # the reference stores these two as 1-tuples (trailing commas, linear_variational.py:83-85)
        self.posterior_mu_init = (posterior_mu_init,)
        self.posterior_rho_init = (posterior_rho_init,)
        self._init_linear(in_features, out_features, prior_mean, prior_variance, posterior_mu_init,
                          posterior_rho_init, bias)

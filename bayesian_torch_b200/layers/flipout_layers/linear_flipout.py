"""LinearFlipout on the fused B200 kernel.

API mirror of /root/reference/bayesian_torch/layers/flipout_layers/linear_flipout.py:49-197.
Both GEMMs of Flipout (x mu^T and (x*s_in) (sigma*eps)^T), the sign draws and the combine run in
ONE launch with two TMEM accumulators (csrc/bt_fused.cu); the reference needs 46 launches."""
from ..._core import BayesLinearBase

__all__ = ["LinearFlipout"]


class LinearFlipout(BayesLinearBase):
    _family = "flipout"

    def __init__(self, in_features, out_features, prior_mean=0, prior_variance=1, posterior_mu_init=0,
                 posterior_rho_init=-3.0, bias=True):
        super().__init__()
        self.posterior_mu_init = posterior_mu_init      # scalars here (linear_flipout.py:80-81)
        self.posterior_rho_init = posterior_rho_init
        self._init_linear(in_features, out_features, prior_mean, prior_variance, posterior_mu_init,
                          posterior_rho_init, bias)

    def forward(self, x, return_kl=True):
        return self._forward_impl(x, return_kl)

"""Conv{1,2,3}dFlipout on the fused implicit-GEMM B200 kernel.

API mirror of /root/reference/bayesian_torch/layers/flipout_layers/conv_flipout.py
(Conv1d :57-244, Conv2d :247-439, Conv3d :443-637).  Unlike the Reparameterization classes the
reference Flipout constructors do not validate channels % groups (:90-105) and keep the posterior
inits as scalars; both quirks are preserved.  The mean conv, the perturbation conv, the input /
output sign draws and the final combine are ONE launch (two TMEM accumulators per tile)."""
from ..._core import BayesConvBase

__all__ = ["Conv1dFlipout", "Conv2dFlipout", "Conv3dFlipout"]


class _ConvFlipout(BayesConvBase):
    _family = "flipout"

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 prior_mean=0, prior_variance=1, posterior_mu_init=0, posterior_rho_init=-3.0, bias=True):
        super().__init__()
        self.posterior_mu_init = posterior_mu_init
        self.posterior_rho_init = posterior_rho_init
        self.kl = 0
        self._init_conv(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, prior_mean,
                        prior_variance, posterior_mu_init, posterior_rho_init, bias, validate=False)

    def forward(self, x, return_kl=True):
        return self._forward_impl(x, return_kl)


class Conv1dFlipout(_ConvFlipout):
    _nd = 1


class Conv2dFlipout(_ConvFlipout):
    _nd = 2


class Conv3dFlipout(_ConvFlipout):
    _nd = 3

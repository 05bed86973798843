"""ConvTranspose{1,2,3}dFlipout on the fused B200 kernel.

API mirror of /root/reference/bayesian_torch/layers/flipout_layers/conv_flipout.py:640-1228, including the constructor
argument ORDER quirk: ConvTranspose1dFlipout takes (..., stride, padding, dilation, groups, output_padding, prior_mean,
...) (:641-655) while the 2d / 3d classes take (..., stride, padding, output_padding, dilation, groups, prior_mean, ...)
(:835-848, :1034-1047); no channel validation in the Flipout classes."""
from ..._core import BayesConvTransposeBase

__all__ = ["ConvTranspose1dFlipout", "ConvTranspose2dFlipout", "ConvTranspose3dFlipout"]


class _ConvTransposeFlipout(BayesConvTransposeBase):
    _family = "flipout"

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, output_padding=0, dilation=1, groups=1,
                 prior_mean=0, prior_variance=1, posterior_mu_init=0, posterior_rho_init=-3.0, bias=True):
        super().__init__()
        self.posterior_mu_init = posterior_mu_init
        self.posterior_rho_init = posterior_rho_init
        self.kl = 0
        self._init_conv_transpose(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, output_padding,
                                  prior_mean, prior_variance, posterior_mu_init, posterior_rho_init, bias, validate=False)


class ConvTranspose1dFlipout(_ConvTransposeFlipout):
    _nd = 1

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, output_padding=0,
                 prior_mean=0, prior_variance=1, posterior_mu_init=0, posterior_rho_init=-3.0, bias=True):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, output_padding, dilation, groups,
                         prior_mean, prior_variance, posterior_mu_init, posterior_rho_init, bias)


class ConvTranspose2dFlipout(_ConvTransposeFlipout):
    _nd = 2


class ConvTranspose3dFlipout(_ConvTransposeFlipout):
    _nd = 3

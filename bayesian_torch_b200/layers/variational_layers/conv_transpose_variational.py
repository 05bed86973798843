"""ConvTranspose{1,2,3}dReparameterization on the fused B200 kernel (fractionally-strided implicit GEMM).

API mirror of /root/reference/bayesian_torch/layers/variational_layers/conv_variational.py:577-1094: same constructor
signature (output_padding sits between groups and prior_mean), parameters mu_kernel / rho_kernel of shape
[in_channels, out_channels // groups, *kernel_size], forward(x, return_kl=True) / kl_loss() contract."""
from ..._core import BayesConvTransposeBase

__all__ = ["ConvTranspose1dReparameterization", "ConvTranspose2dReparameterization", "ConvTranspose3dReparameterization"]


class _ConvTransposeReparam(BayesConvTransposeBase):
    _family = "reparam"

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, output_padding=0,
                 prior_mean=0, prior_variance=1, posterior_mu_init=0, posterior_rho_init=-3.0, bias=True):
        super().__init__()
        self.posterior_mu_init = (posterior_mu_init,)    # 1-tuples, conv_variational.py:627-630
        self.posterior_rho_init = (posterior_rho_init,)
        self._init_conv_transpose(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, output_padding,
                                  prior_mean, prior_variance, posterior_mu_init, posterior_rho_init, bias, validate=True)


class ConvTranspose1dReparameterization(_ConvTransposeReparam):
    _nd = 1


class ConvTranspose2dReparameterization(_ConvTransposeReparam):
    _nd = 2


class ConvTranspose3dReparameterization(_ConvTransposeReparam):
    _nd = 3

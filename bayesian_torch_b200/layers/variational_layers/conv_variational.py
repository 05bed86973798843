"""Conv{1,2,3}dReparameterization on the fused implicit-GEMM B200 kernel.

API mirror of /root/reference/bayesian_torch/layers/variational_layers/conv_variational.py
(Conv1d :64-227, Conv2d :230-402, Conv3d :405-574): same constructor signatures -- including
Conv3dReparameterization's positional order (prior/posterior arguments BEFORE stride, :406-418) --
same parameter names (mu_kernel, rho_kernel, mu_bias, rho_bias) and forward/kl_loss contract.
ConvTranspose variants are a "next" row and are not provided."""
from ..._core import BayesConvBase

__all__ = ["Conv1dReparameterization", "Conv2dReparameterization", "Conv3dReparameterization"]


class _ConvReparam(BayesConvBase):
    _family = "reparam"

    def _setup(self, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, prior_mean,
               prior_variance, posterior_mu_init, posterior_rho_init, bias):
        self.posterior_mu_init = (posterior_mu_init,)    # 1-tuples, conv_variational.py:282-284
        self.posterior_rho_init = (posterior_rho_init,)
        self._init_conv(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, prior_mean,
                        prior_variance, posterior_mu_init, posterior_rho_init, bias, validate=True)


class Conv1dReparameterization(_ConvReparam):
    _nd = 1

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 prior_mean=0, prior_variance=1, posterior_mu_init=0, posterior_rho_init=-3.0, bias=True):
        super().__init__()
        self._setup(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, prior_mean,
                    prior_variance, posterior_mu_init, posterior_rho_init, bias)


class Conv2dReparameterization(_ConvReparam):
    _nd = 2

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 prior_mean=0, prior_variance=1, posterior_mu_init=0, posterior_rho_init=-3.0, bias=True):
        super().__init__()
        self._setup(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, prior_mean,
                    prior_variance, posterior_mu_init, posterior_rho_init, bias)


class Conv3dReparameterization(_ConvReparam):
    _nd = 3

    def __init__(self, in_channels, out_channels, kernel_size, prior_mean, prior_variance, posterior_mu_init,
                 posterior_rho_init, stride=1, padding=0, dilation=1, groups=1, bias=True):
        super().__init__()
        self._setup(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, prior_mean,
                    prior_variance, posterior_mu_init, posterior_rho_init, bias)

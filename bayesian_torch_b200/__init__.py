"""bayesian_torch_b200 -- B200 (sm_100a) implementation of the bayesian-torch stochastic
variational layer forward path (Linear / Conv{1,2,3}d x Reparameterization / Flipout), its
closed-form Gaussian KL and Monte-Carlo inference, behind the reference's own
`bayesian_torch.layers` class API and `dnn_to_bnn()` surface.

Host code is Python/PyTorch (device memory, streams, torch.distributed); all arithmetic of the
path runs in hand-written CUDA kernels reached through the C ABI of libbtb200.so
(include/btb200.h).  No CPU fallback exists.
"""
from . import layers  # noqa: F401
from ._core import assign_layer_keys, manual_seed, mc_sample_context  # noqa: F401
from .fuse import fuse_inference  # noqa: F401
from .mc import mc_predict  # noqa: F401
from .models.dnn_to_bnn import dnn_to_bnn, get_kl_loss  # noqa: F401

__version__ = "0.1.0"

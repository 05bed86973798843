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
from .fuse import fuse_inference, refresh_epilogues  # noqa: F401
from .mc import mc_predict  # noqa: F401
from .models.dnn_to_bnn import dnn_to_bnn, get_kl_loss  # noqa: F401



def invalidate_caches(model):
    """Forget every cached parameter transform and captured CUDA graph of `model` -- call after editing parameters
    through `.data` (in-place ops and load_state_dict are tracked automatically)."""
    from . import mc as _mc_mod
    from ._core import BayesLayerBase
    for m in model.modules():
        if isinstance(m, BayesLayerBase):
            m.invalidate_caches()
    refresh_epilogues(model)
    _mc_mod.drop_graphs(model)
    return model


__version__ = "0.2.0"

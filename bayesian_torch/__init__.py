"""Drop-in import name.  `import bayesian_torch.layers`, `from bayesian_torch.models.dnn_to_bnn
import dnn_to_bnn, get_kl_loss` resolve to the B200 implementation in `bayesian_torch_b200`
(same module paths as IntelLabs/bayesian-torch for the hot-path surface)."""
from bayesian_torch_b200 import __version__  # noqa: F401
from . import layers, models, utils  # noqa: F401

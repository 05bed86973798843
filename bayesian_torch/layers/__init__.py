from bayesian_torch_b200.layers import *  # noqa: F401,F403
from bayesian_torch_b200.layers import base_variational_layer, flipout_layers, variational_layers  # noqa: F401

from bayesian_torch_b200.models.dnn_to_bnn import *  # noqa: F401,F403
from bayesian_torch_b200.models.dnn_to_bnn import bnn_conv_layer, bnn_linear_layer, dnn_to_bnn, get_kl_loss  # noqa: F401

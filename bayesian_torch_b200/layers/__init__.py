"""Drop-in `bayesian_torch.layers` surface (reference: /root/reference/bayesian_torch/layers/__init__.py:1-6)
for the hot-path classes.  dnn_to_bnn() resolves classes here BY NAME
(models/dnn_to_bnn.py:53-54,77-78), so the names are the contract."""
from .flipout_layers import *
from .variational_layers import *
from .base_variational_layer import *

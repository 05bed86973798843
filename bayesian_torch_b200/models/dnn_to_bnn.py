"""dnn_to_bnn() / get_kl_loss(): the conversion surface of the drop-in API.

Behavioural mirror of /root/reference/bayesian_torch/models/dnn_to_bnn.py:52-165:
  * in-place module surgery, recursing into every child that itself has children (:128-130);
  * leaf modules are matched by CLASS-NAME SUBSTRING ("Conv", "Linear") and the Bayesian class is
    looked up by name  <ClassName> + params["type"]  in the layers package (:53-54, :77-78);
  * constructor keyword mapping prior_mu -> prior_mean, prior_sigma -> prior_variance (:58-66, :79-92);
  * optional MOPED init  mu <- w, rho <- get_rho(w, delta)  (:65-71, :95-101);
  * every created layer gets dnn_to_bnn_flag = True so forward() returns `out` only (:72, :102);
  * get_kl_loss sums kl_loss() of every module that has one (:157-165).
  * nn.LSTM -> LSTM<type> built on two fused Bayesian linears (:106-122); ConvTranspose{1,2,3}d go through the
    "Conv" branch exactly as in the reference (class name + type).
"""
import bayesian_torch_b200.layers as bayesian_layers
from bayesian_torch_b200.utils.util import get_rho

__all__ = ["dnn_to_bnn", "get_kl_loss", "bnn_linear_layer", "bnn_conv_layer", "bnn_lstm_layer"]


def _prior_kwargs(params):
    return dict(prior_mean=params["prior_mu"], prior_variance=params["prior_sigma"],
                posterior_mu_init=params["posterior_mu_init"], posterior_rho_init=params["posterior_rho_init"])


def _finish(params, bnn_layer, d, wname):
    if params["moped_enable"]:
        delta = params["moped_delta"]
        getattr(bnn_layer, f"mu_{wname}").data.copy_(d.weight.data)
        getattr(bnn_layer, f"rho_{wname}").data.copy_(get_rho(d.weight.data, delta))
        if bnn_layer.mu_bias is not None:
            bnn_layer.mu_bias.data.copy_(d.bias.data)
            bnn_layer.rho_bias.data.copy_(get_rho(d.bias.data, delta))
        bnn_layer.invalidate_caches()           # (.data writes do not bump Tensor._version)
    bnn_layer.dnn_to_bnn_flag = True
    return bnn_layer


def bnn_linear_layer(params, d):
    layer_fn = getattr(bayesian_layers, d.__class__.__name__ + params["type"])
    layer = layer_fn(in_features=d.in_features, out_features=d.out_features,
                     bias=d.bias is not None, **_prior_kwargs(params))
    return _finish(params, layer, d, "weight")


def bnn_conv_layer(params, d):
    layer_fn = getattr(bayesian_layers, d.__class__.__name__ + params["type"])
    layer = layer_fn(in_channels=d.in_channels, out_channels=d.out_channels, kernel_size=d.kernel_size,
                     stride=d.stride, padding=d.padding, dilation=d.dilation, groups=d.groups,
                     bias=d.bias is not None, **_prior_kwargs(params))
    return _finish(params, layer, d, "kernel")


def bnn_lstm_layer(params, d):
    """models/dnn_to_bnn.py:106-122: nn.LSTM -> LSTM<type>(in_features=input_size, out_features=hidden_size); MOPED is
    not supported for LSTM layers (the reference prints a warning and skips it)."""
    layer_fn = getattr(bayesian_layers, d.__class__.__name__ + params["type"])
    layer = layer_fn(in_features=d.input_size, out_features=d.hidden_size, bias=d.bias is not None, **_prior_kwargs(params))
    if params["moped_enable"]:
        print("WARNING: MOPED method is not supported for LSTM layers!!!")
    layer.dnn_to_bnn_flag = True
    return layer


def dnn_to_bnn(m, bnn_prior_parameters):
    for name, child in list(m._modules.items()):
        if child is None:
            continue
        cls = child.__class__.__name__
        if child._modules:
            dnn_to_bnn(child, bnn_prior_parameters)
        elif "Conv" in cls:
            setattr(m, name, bnn_conv_layer(bnn_prior_parameters, child))
        elif "Linear" in cls:
            setattr(m, name, bnn_linear_layer(bnn_prior_parameters, child))
        elif "LSTM" in cls:
            setattr(m, name, bnn_lstm_layer(bnn_prior_parameters, child))
    return


def get_kl_loss(m):
    kl_loss = None
    for layer in m.modules():
        if hasattr(layer, "kl_loss"):
            kl = layer.kl_loss()
            kl_loss = kl if kl_loss is None else kl_loss + kl
    return kl_loss

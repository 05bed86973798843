"""`bayesian_torch.utils.util` of the drop-in surface (reference: /root/reference/bayesian_torch/utils/util.py:41-136):
get_rho (used by dnn_to_bnn(moped_enable=True)), the numpy ensemble-uncertainty helpers the example drivers import, and
the MOPED empirical-Bayes initialiser.  Host-side utilities: plain tensor / numpy math, no kernels involved (the
on-device versions of the uncertainty helpers are mc_predict(..., return_uncertainty=True))."""
import numpy as np
import torch

__all__ = ["entropy", "predictive_entropy", "mutual_information", "get_rho", "MOPED"]


def entropy(prob):
    """util.py:41-42"""
    return -1 * np.sum(prob * np.log(prob + 1e-15), axis=-1)


def predictive_entropy(mc_preds):
    """entropy of the mean of the MC predictive distribution (axis 0 = MC samples), util.py:45-50"""
    return entropy(np.mean(mc_preds, axis=0))


def mutual_information(mc_preds):
    """entropy of the mean minus the mean of the entropies, util.py:53-60"""
    return entropy(np.mean(mc_preds, axis=0)) - np.mean(entropy(mc_preds), axis=0)


def get_rho(sigma, delta):
    """rho such that softplus(rho) = delta * |sigma| (MOPED init); +1e-20 keeps log finite at 0 (util.py:63-69)."""
    return torch.log(torch.expm1(delta * torch.abs(sigma)) + 1e-20)


def MOPED(model, det_model, det_checkpoint, delta):
    """Model Priors with Empirical Bayes using a Deterministic DNN (util.py:72-136): prior mean <- deterministic
    weights, mu <- weights, rho <- get_rho(weights, delta), BatchNorm state copied.  Layers are matched by position in
    .modules(), as in the reference.  The parameter writes go through `.data` (not version-tracked), so every cached
    parameter transform of the B200 layers is dropped afterwards."""
    from .._core import BayesConvBase, BayesLayerBase, BayesLinearBase
    det_model.load_state_dict(torch.load(det_checkpoint))
    for layer, det_layer in zip(model.modules(), det_model.modules()):
        if isinstance(layer, (BayesConvBase, BayesLinearBase)):
            w = "kernel" if isinstance(layer, BayesConvBase) else "weight"
            dev, dt = getattr(layer, f"mu_{w}").device, getattr(layer, f"mu_{w}").dtype
            dw = det_layer.weight.data.to(device=dev, dtype=dt)
            layer.prior_weight_mu = dw.clone()
            if layer.prior_bias_mu is not None and det_layer.bias is not None:
                layer.prior_bias_mu = det_layer.bias.data.to(device=dev, dtype=dt).clone()
            getattr(layer, f"mu_{w}").data = dw.clone()
            getattr(layer, f"rho_{w}").data = get_rho(dw, delta)
            if layer.mu_bias is not None and det_layer.bias is not None:
                db = det_layer.bias.data.to(device=dev, dtype=dt)
                layer.mu_bias.data = db.clone()
                layer.rho_bias.data = get_rho(db, delta)
        elif str(layer).startswith('Batch'):
            layer.weight.data = det_layer.weight.data.clone()
            if layer.bias is not None:
                layer.bias.data = det_layer.bias.data.clone()
            layer.running_mean.data = det_layer.running_mean.data.clone()
            layer.running_var.data = det_layer.running_var.data.clone()
            layer.num_batches_tracked.data = det_layer.num_batches_tracked.data.clone()
    for layer in model.modules():
        if isinstance(layer, BayesLayerBase):
            layer.invalidate_caches()
    return model

"""LSTMFlipout on the fused B200 linears.

API mirror of /root/reference/bayesian_torch/layers/flipout_layers/rnn_flipout.py:46-153: two Bayesian
linears `ih` (in_features -> 4 * out_features) and `hh` (out_features -> 4 * out_features) -- same sub-module and
parameter names, so state_dicts are interchangeable -- evaluated once per time step with a FRESH weight sample each
(the reference calls self.ih(x_t) / self.hh(h_t) inside the time loop), KL accumulated per step, gates in the order
i | f | g | o.  Per step: two fused bt_layer_forward launches + one bt_lstm_cell launch (the reference: 2 x 38 + 12
ATen launches).  forward(X, hidden_states=None, return_kl=True) -> (hidden_seq, (hidden_seq, c_ts)[, kl])."""
import torch

from ... import _native
from ..._base import BaseVariationalLayer_
from .linear_flipout import LinearFlipout

__all__ = ["LSTMFlipout"]


class LSTMFlipout(BaseVariationalLayer_):
    def __init__(self, in_features, out_features, prior_mean=0, prior_variance=1, posterior_mu_init=0,
                 posterior_rho_init=-3.0, bias=True):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.prior_mean = prior_mean
        self.prior_variance = prior_variance
        self.posterior_mu_init = (posterior_mu_init,)      # 1-tuples in the reference (trailing commas)
        self.posterior_rho_init = (posterior_rho_init,)
        self.bias = bias
        kw = dict(prior_mean=prior_mean, prior_variance=prior_variance, posterior_mu_init=posterior_mu_init,
                  posterior_rho_init=posterior_rho_init, bias=bias)
        self.ih = LinearFlipout(in_features=in_features, out_features=out_features * 4, **kw)
        self.hh = LinearFlipout(in_features=out_features, out_features=out_features * 4, **kw)

    def kl_loss(self):
        return self.ih.kl_loss() + self.hh.kl_loss()

    def forward(self, X, hidden_states=None, return_kl=True):
        if self.dnn_to_bnn_flag:
            return_kl = False
        _native.require_cuda(X, "input")
        batch_size, seq_size, _ = X.size()
        HS = self.out_features
        if hidden_states is None:
            h_t = torch.zeros(batch_size, HS, dtype=X.dtype, device=X.device)
            c_t = torch.zeros(batch_size, HS, dtype=X.dtype, device=X.device)
        else:
            h_t, c_t = hidden_states
        differentiable = torch.is_grad_enabled() and (X.requires_grad or any(p.requires_grad for p in self.parameters()))
        kl = 0
        if differentiable:
            # training: the gate GEMMs are the fused (autograd-wrapped) linears, the pointwise stage stays in ATen so
            # that autograd can see it
            hs, cs = [], []
            for t in range(seq_size):
                ff_i, kl_i = self.ih(X[:, t, :])
                ff_h, kl_h = self.hh(h_t)
                gates = ff_i + ff_h
                kl = kl + kl_i + kl_h
                i_t, f_t = torch.sigmoid(gates[:, :HS]), torch.sigmoid(gates[:, HS:HS * 2])
                g_t, o_t = torch.tanh(gates[:, HS * 2:HS * 3]), torch.sigmoid(gates[:, HS * 3:])
                c_t = f_t * c_t + i_t * g_t
                h_t = o_t * torch.tanh(c_t)
                hs.append(h_t.unsqueeze(1))
                cs.append(c_t.unsqueeze(1))
            hidden_seq, c_ts = torch.cat(hs, dim=1).contiguous(), torch.cat(cs, dim=1).contiguous()
        else:
            hidden_seq = torch.empty(batch_size, seq_size, HS, dtype=X.dtype, device=X.device)
            c_ts = torch.empty_like(hidden_seq)
            h_t, c_t = h_t.to(X.dtype).contiguous(), c_t.to(X.dtype).contiguous()
            for t in range(seq_size):
                if return_kl:
                    ff_i, kl_i = self.ih(X[:, t, :])
                    ff_h, kl_h = self.hh(h_t)
                    kl = kl + kl_i + kl_h
                else:      # converted models (dnn_to_bnn_flag): nobody reads the per-step KL, skip its side output
                    ff_i, ff_h = self.ih(X[:, t, :], return_kl=False), self.hh(h_t, return_kl=False)
                h_n, c_n = torch.empty_like(h_t), torch.empty_like(c_t)
                _native.lstm_cell(ff_i.contiguous(), ff_h.contiguous(), c_t, h_n, c_n, hidden_seq, c_ts, t)
                h_t, c_t = h_n, c_n
        self.kl = kl
        if return_kl:
            return hidden_seq, (hidden_seq, c_ts), kl
        return hidden_seq, (hidden_seq, c_ts)

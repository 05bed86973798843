"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the layer families SURVEY.md 8f rank 4 lists as "next":
ConvTranspose{1,2,3}d{Reparameterization,Flipout} and LSTM{Reparameterization,Flipout}.  No product code exists for
them yet; this is the checker they will be held to, pinned on outputs of the reference itself
(tests/golden/make_golden_next.py -> next.npz, tests/test_oracle_next.py).

Reference op sequences restated (eps / signs are explicit arguments, as in bt_oracle.py):
  layers/variational_layers/conv_variational.py:698-745 (ConvTranspose1d; 2d :868-915, 3d :1043-1090)
  layers/flipout_layers/conv_flipout.py:760-832 (ConvTranspose1d; 2d / 3d alike)
  layers/variational_layers/rnn_variational.py:103-153, layers/flipout_layers/rnn_flipout.py (same loop)
"""
import torch
import torch.nn.functional as F

from .bt_oracle import linear_flipout, linear_reparam, sigma_of_rho

_CT = {1: F.conv_transpose1d, 2: F.conv_transpose2d, 3: F.conv_transpose3d}


def conv_transpose_reparam(nd, x, mu_k, rho_k, eps_k, mu_b=None, rho_b=None, eps_b=None,
                           stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    """out = conv_transposeNd(x, mu + softplus(rho) * eps, mu_b + softplus(rho_b) * eps_b)"""
    w = mu_k + sigma_of_rho(rho_k) * eps_k
    b = None if mu_b is None else mu_b + sigma_of_rho(rho_b) * eps_b
    return _CT[nd](x, w, b, stride, padding, output_padding, groups, dilation)


def conv_transpose_flipout(nd, x, mu_k, rho_k, eps_k, sign_in, sign_out, mu_b=None, rho_b=None, eps_b=None,
                           stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    """out = convT(x, mu, mu_b) + convT(x * s_in, softplus(rho) * eps, softplus(rho_b) * eps_b) * s_out"""
    d = sigma_of_rho(rho_k) * eps_k
    b = None if mu_b is None else sigma_of_rho(rho_b) * eps_b
    mean = _CT[nd](x, mu_k, mu_b, stride, padding, output_padding, groups, dilation)
    return mean + _CT[nd](x * sign_in, d, b, stride, padding, output_padding, groups, dilation) * sign_out


def lstm_forward(x, ih, hh, draws, flipout, hidden=None):
    """The reference's python time-step loop over two Bayesian Linear layers (gates = ih(x_t) + hh(h_t); i, f, g, o).
    ih / hh: dicts(mu_w, rho_w, mu_b, rho_b); draws[t] = dict(ih=..., hh=...) with eps_w, eps_b (+ sign_in, sign_out for
    Flipout) -- NEW draws at every time step, as in the reference.  -> (hidden_seq [B,T,H], c_seq [B,T,H])"""
    bsz, T, _ = x.shape
    H = hh["mu_w"].shape[1]
    h = torch.zeros(bsz, H) if hidden is None else hidden[0]
    c = torch.zeros(bsz, H) if hidden is None else hidden[1]

    def lin(p, d, inp):
        if flipout:
            return linear_flipout(inp, p["mu_w"], p["rho_w"], d["eps_w"], d["sign_in"], d["sign_out"],
                                  p.get("mu_b"), p.get("rho_b"), d.get("eps_b"))
        return linear_reparam(inp, p["mu_w"], p["rho_w"], d["eps_w"], p.get("mu_b"), p.get("rho_b"), d.get("eps_b"))

    hs, cs = [], []
    for t in range(T):
        gates = lin(ih, draws[t]["ih"], x[:, t, :]) + lin(hh, draws[t]["hh"], h)
        i, f = torch.sigmoid(gates[:, :H]), torch.sigmoid(gates[:, H:2 * H])
        g, o = torch.tanh(gates[:, 2 * H:3 * H]), torch.sigmoid(gates[:, 3 * H:])
        c = f * c + i * g
        h = o * torch.tanh(c)
        hs.append(h)
        cs.append(c)
    return torch.stack(hs, 1), torch.stack(cs, 1)

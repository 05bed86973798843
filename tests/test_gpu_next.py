"""-m gpu: the layer families of SURVEY.md 8f rank 4 -- ConvTranspose{1,2,3}d{Reparameterization,Flipout} and
LSTM{Reparameterization,Flipout} -- against the reference-minted goldens (tests/golden/next.npz) and the pinned
restatement (oracle/bt_oracle_next.py).

ConvTranspose: (1) reference draws injected through the debug hooks, kernel output vs the reference's own y (tf32 path:
rel-RMS <= 5e-4; vs the operand-rounded oracle <= 1e-4); (2) on-chip draws re-materialised and fed to the oracle, fp32
and bf16; (3) gradients through the autograd wrapper vs torch.autograd on the oracle function.
LSTM: on-chip draws of every time step re-materialised (step t of a forward uses sample index t) and fed to the
reference's loop restated in oracle/bt_oracle_next.py::lstm_forward; inference path (bt_lstm_cell) and training path."""
import json
import os

import numpy as np
import pytest
import torch

import bayesian_torch_b200 as btb
import bayesian_torch_b200.layers as L
from bayesian_torch_b200 import _native
from gpu_util import cl, errs, note
from oracle import bt_oracle as O
from oracle import bt_oracle_next as ON

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_Z = np.load(os.path.join(ROOT, "tests", "golden", "next.npz"))
with open(os.path.join(ROOT, "tests", "golden", "next_meta.json")) as _f:
    _META = json.load(_f)["cases"]


def _case(name):
    pre = name + "/"
    return {k[len(pre):]: torch.from_numpy(_Z[k]) for k in _Z.files if k.startswith(pre)}


def _convt(nd, flip, cin, cout, ks, stride, padding, output_padding, dilation, groups, bias, **kw):
    cls = getattr(L, f"ConvTranspose{nd}d" + ("Flipout" if flip else "Reparameterization"))
    return cls(cin, cout, ks, stride=stride, padding=padding, output_padding=output_padding, dilation=dilation,
               groups=groups, bias=bias, **kw)


def _oracle_convt(layer, x, eps_w, eps_b, s_in, s_out, rnd):
    f = lambda t: None if t is None else t.detach().float().cpu()
    mu, rho = f(layer.mu_kernel), f(layer.rho_kernel)
    sig = O.sigma_of_rho(rho)
    geo = (layer.stride, layer.padding, layer.output_padding, layer.groups, layer.dilation)
    ct = ON._CT[layer._nd]
    xb = rnd(f(x))
    if layer._family == "reparam":
        w = rnd(mu + sig * f(eps_w))
        b = None if layer.mu_bias is None else f(layer.mu_bias) + O.sigma_of_rho(f(layer.rho_bias)) * f(eps_b)
        return ct(xb, w, b, *geo)
    b = None if layer.mu_bias is None else O.sigma_of_rho(f(layer.rho_bias)) * f(eps_b)
    return ct(xb, rnd(mu), f(layer.mu_bias), *geo) + ct(xb * f(s_in), rnd(sig * f(eps_w)), b, *geo) * f(s_out)


@pytest.mark.parametrize("name", sorted(n for n, m in _META.items() if m["kind"] == "convt"))
def test_conv_transpose_golden_parity(name):
    c, m = _case(name), _META[name]
    mu = c["mu_w"]
    layer = _convt(m["nd"], m["flipout"], mu.shape[0], mu.shape[1] * m["groups"], tuple(mu.shape[2:]), m["stride"], m["padding"],
                   m["output_padding"], m["dilation"], m["groups"], m["bias"], prior_mean=0.0, prior_variance=1.1)   # (make_golden_next.py:43)
    sd = {"mu_kernel": mu, "rho_kernel": c["rho_w"]}
    if m["bias"]:
        sd.update(mu_bias=c["mu_b"], rho_bias=c["rho_b"])
    layer.load_state_dict(sd)
    layer = layer.to(DEV)
    assert tuple(layer.prior_weight_sigma.shape) == tuple(mu.shape)        # (not the reference's mis-shaped buffer)
    dbg = {"eps_w_in": layer.kernel_eps_layout(c["eps_w"].to(DEV))}
    if m["bias"]:
        dbg["eps_b_in"] = c["eps_b"].to(DEV)
    if m["flipout"]:
        dbg["sign_in"], dbg["sign_out"] = cl(c["sign_in"]).to(DEV), cl(c["sign_out"]).to(DEV)
    with torch.no_grad():
        y, kl = layer._forward_impl(c["x"].to(DEV), True, debug=dbg)
    torch.cuda.synchronize()
    assert y.shape == c["y"].shape
    rel, mx = errs(y, c["y"])
    yr = _oracle_convt(layer, c["x"], c["eps_w"], c.get("eps_b"), c.get("sign_in"), c.get("sign_out"), O.round_operand_tf32)
    rel2, _ = errs(y, yr)
    note("convt_golden", case=name, path=_native.last_forward_path(), rel=rel, rel_rounded=rel2)
    assert rel <= 5e-4, (name, rel, mx)
    assert rel2 <= 1e-4, (name, rel2)
    if "kl" in c:
        assert abs(float(kl) - float(c["kl"])) <= 1e-5 * max(1.0, abs(float(c["kl"]))), (float(kl), float(c["kl"]))


CT = [
    # nd, flip, cin, cout, ks, stride, pad, out_pad, dil, groups, bias, batch, spatial, dtype
    (2, False, 16, 24, 3, 2, 1, 1, 1, 1, True, 3, (5, 6), torch.float32),
    (2, True, 16, 24, 3, 2, 1, 1, 1, 2, True, 3, (5, 6), torch.float32),
    (2, False, 64, 32, 4, 2, 1, 0, 1, 1, False, 4, (8, 8), torch.bfloat16),      # GAN-style upsampling block
    (2, True, 32, 16, 3, 1, 0, 0, 2, 1, True, 2, (6, 5), torch.bfloat16),
    (1, False, 8, 12, 5, 3, 2, 2, 1, 1, True, 2, (17,), torch.float32),
    (3, True, 8, 8, 2, 2, 0, 0, 1, 1, True, 2, (3, 4, 4), torch.float32),
]


@pytest.mark.parametrize("cfg", CT, ids=lambda c: f"ct{c[0]}{'F' if c[1] else 'R'}_{c[2]}x{c[3]}_k{c[4]}s{c[5]}g{c[9]}_{str(c[13])[6:]}")
def test_conv_transpose_onchip_parity_and_grads(cfg):
    nd, flip, cin, cout, ks, st, pd, op, dl, groups, bias, batch, sp, dt = cfg
    torch.manual_seed(7)
    btb.manual_seed(5)
    layer = _convt(nd, flip, cin, cout, ks, st, pd, op, dl, groups, bias).to(DEV).to(dt)
    x = torch.randn(batch, cin, *sp).to(dt).to(DEV).requires_grad_(True)
    y, kl = layer(x)
    assert y.grad_fn is not None
    dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(2)).to(DEV).to(dt)
    ((y.float() * dy.float()).sum() + 0.3 * kl.float()).backward()
    layer._bt_last["sample0"] = 0
    eps_w, eps_b = layer.materialize_eps(0)
    assert tuple(eps_w.shape) == tuple(layer.mu_kernel.shape)
    s_in = s_out = None
    if flip:
        s_in, s_out = layer.materialize_signs(tuple(x.shape), tuple(y.shape), 0)
    tf32 = dt == torch.float32
    rnd = O.round_operand_tf32 if tf32 else O.round_operand
    yr = _oracle_convt(layer, x, eps_w, eps_b, s_in, s_out, rnd)
    yf = _oracle_convt(layer, x, eps_w, eps_b, s_in, s_out, lambda t: t)
    rel_r, _ = errs(y, yr)
    rel_f, _ = errs(y, yf)
    note("convt_onchip", cfg=str(cfg), rel_rounded=rel_r, rel_fp32=rel_f)
    assert rel_r <= (1e-4 if tf32 else 3e-3) and rel_f <= (5e-4 if tf32 else 3e-3), (rel_r, rel_f)
    # gradients: torch.autograd through the oracle function on the same draws
    f = lambda t: None if t is None else t.detach().float().cpu()
    xc = f(x).requires_grad_(True)
    mu, rho = f(layer.mu_kernel).requires_grad_(True), f(layer.rho_kernel).requires_grad_(True)
    mb = f(layer.mu_bias).requires_grad_(True) if bias else None
    rb = f(layer.rho_bias).requires_grad_(True) if bias else None
    geo = dict(stride=st, padding=pd, output_padding=op, groups=groups, dilation=dl)
    if flip:
        yo = ON.conv_transpose_flipout(nd, xc, mu, rho, f(eps_w), f(s_in), f(s_out), mb, rb, f(eps_b), **geo)
    else:
        yo = ON.conv_transpose_reparam(nd, xc, mu, rho, f(eps_w), mb, rb, f(eps_b), **geo)
    klo = O.kl_loss(mu, rho, 0.0, 1.0, mb, rb)
    ((yo * f(dy)).sum() + 0.3 * klo).backward()
    pairs = [("dx", x.grad, xc.grad), ("dmu", layer.mu_kernel.grad, mu.grad), ("drho", layer.rho_kernel.grad, rho.grad)]
    if bias:
        pairs += [("dmu_b", layer.mu_bias.grad, mb.grad), ("drho_b", layer.rho_bias.grad, rb.grad)]
    for nm, got, ref in pairs:
        rel, mx = errs(got, ref)
        note("convt_grad", cfg=str(cfg), which=nm, rel=rel)
        assert rel <= (2e-3 if tf32 else 1.5e-2), (nm, rel, mx)      # (cuDNN TF32 in the ATen backward products)


@pytest.mark.parametrize("train", [False, True], ids=["inference", "training"])
@pytest.mark.parametrize("flip", [False, True], ids=["R", "F"])
def test_lstm_equals_reference_loop_on_the_same_draws(flip, train):
    torch.manual_seed(3)
    btb.manual_seed(11)
    B, T, I, H = 5, 6, 20, 12
    cls = L.LSTMFlipout if flip else L.LSTMReparameterization
    lstm = cls(I, H).to(DEV)
    x = torch.randn(B, T, I, device=DEV)
    if train:
        hseq, (h2, cseq), kl = lstm(x)
        assert hseq.grad_fn is not None
        (hseq.sum() + 0.1 * kl).backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in lstm.parameters())
    else:
        with torch.no_grad():
            hseq, (h2, cseq), kl = lstm(x)
    torch.cuda.synchronize()
    assert hseq.shape == (B, T, H) and cseq.shape == (B, T, H) and h2 is hseq
    f = lambda t: t.detach().float().cpu()
    draws = []
    for t in range(T):                                   # step t of the first forward after manual_seed uses sample index t
        d = {}
        for tag, lin, xs in (("ih", lstm.ih, (B, I)), ("hh", lstm.hh, (B, H))):
            lin._bt_last["sample0"] = 0
            ew, eb = lin.materialize_eps(t)
            d[tag] = {"eps_w": f(ew), "eps_b": f(eb)}
            if flip:
                si, so = lin.materialize_signs(xs, (B, 4 * H), t)
                d[tag].update(sign_in=f(si), sign_out=f(so))
        draws.append(d)
    par = lambda lin: {"mu_w": f(lin.mu_weight), "rho_w": f(lin.rho_weight), "mu_b": f(lin.mu_bias), "rho_b": f(lin.rho_bias)}
    href, cref = ON.lstm_forward(f(x), par(lstm.ih), par(lstm.hh), draws, flip)
    rh, _ = errs(hseq, href)
    rc, _ = errs(cseq, cref)
    kref = sum(float(O.kl_loss(p["mu_w"].double(), p["rho_w"].double(), 0.0, 1.0, p["mu_b"].double(), p["rho_b"].double()))
               for p in (par(lstm.ih), par(lstm.hh))) * T
    note("lstm", flip=flip, train=train, rel_h=rh, rel_c=rc, kl=float(kl), kl_ref=kref)
    assert rh <= 2e-3 and rc <= 2e-3, (rh, rc)          # tf32 gate GEMMs, T recurrent steps
    assert abs(float(kl) - kref) <= 1e-4 * abs(kref)     # the reference accumulates the KL once per time step


def test_dnn_to_bnn_converts_lstm_and_conv_transpose():
    import torch.nn as nn
    prm = {"prior_mu": 0.0, "prior_sigma": 1.0, "posterior_mu_init": 0.0, "posterior_rho_init": -3.0,
           "type": "Flipout", "moped_enable": False, "moped_delta": 0.5}

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.rnn = nn.LSTM(8, 16, batch_first=True)
            self.up = nn.ConvTranspose2d(4, 6, 3, stride=2, padding=1, output_padding=1)

    net = Net()
    btb.dnn_to_bnn(net, prm)
    net = net.to(DEV)
    with torch.no_grad():
        hs, (h, c) = net.rnn(torch.randn(3, 5, 8, device=DEV))
        up = net.up(torch.randn(2, 4, 5, 5, device=DEV))
    # (the reference's bnn_conv_layer does not forward output_padding, models/dnn_to_bnn.py:79-92: 9x9, not 10x10)
    assert tuple(hs.shape) == (3, 5, 16) and tuple(up.shape) == (2, 6, 9, 9)
    assert float(btb.get_kl_loss(net)) > 0

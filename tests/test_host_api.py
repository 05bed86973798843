"""CPU: host-side logic of the drop-in surface -- names, constructors, state_dict, init parity with the
reference (same seed -> same parameters), dnn_to_bnn surgery, the no-CPU-fallback contract, and the
C-ABI library (loads, exports every symbol include/btb200.h declares; no compute without a GPU)."""
import ctypes
import os
import re

import pytest
import torch
import torch.nn as nn

import bayesian_torch_b200 as btb
import bayesian_torch_b200.layers as L
from bayesian_torch_b200 import _native
from bayesian_torch_b200.mc import shard_samples
from bayesian_torch_b200.models.dnn_to_bnn import dnn_to_bnn, get_kl_loss

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ["LinearReparameterization", "LinearFlipout", "Conv1dReparameterization", "Conv2dReparameterization",
         "Conv3dReparameterization", "Conv1dFlipout", "Conv2dFlipout", "Conv3dFlipout", "BaseVariationalLayer_"]


def test_names_and_dropin_import_path():
    for n in NAMES:
        assert hasattr(L, n), n
    import bayesian_torch.layers as RL  # drop-in module path
    from bayesian_torch.models.dnn_to_bnn import dnn_to_bnn as d2, get_kl_loss as g2
    from bayesian_torch.utils.util import get_rho
    assert RL.Conv2dFlipout is L.Conv2dFlipout and d2 is dnn_to_bnn and g2 is get_kl_loss
    assert float(get_rho(torch.tensor([0.2]), 0.5)) == pytest.approx(float(torch.log(torch.expm1(torch.tensor(0.1)))))
    assert issubclass(L.LinearFlipout, L.BaseVariationalLayer_)


def test_state_dict_keys_match_reference(golden):
    for name, m in golden.meta["cases"].items():
        c = golden.case(name)
        flip, bias = m["flipout"], m["bias"]
        if m["kind"] == "linear":
            cls = L.LinearFlipout if flip else L.LinearReparameterization
            layer = cls(c["mu_w"].shape[1], c["mu_w"].shape[0], bias=bias)
        else:
            cls = getattr(L, f"Conv{m['nd']}d" + ("Flipout" if flip else "Reparameterization"))
            ks = tuple(c["mu_w"].shape[2:])
            if m["nd"] == 3 and not flip:
                layer = cls(c["mu_w"].shape[1] * m["groups"], c["mu_w"].shape[0], ks, 0, 1, 0, -3.0,
                            groups=m["groups"], bias=bias)
            else:
                layer = cls(c["mu_w"].shape[1] * m["groups"], c["mu_w"].shape[0], ks if m["nd"] > 1 else ks[0],
                            groups=m["groups"], bias=bias)
        assert list(layer.state_dict().keys()) == m["state_dict_keys"], name
        w = "kernel" if m["kind"] == "conv" else "weight"
        assert tuple(getattr(layer, f"mu_{w}").shape) == tuple(c["mu_w"].shape)
        # a reference checkpoint loads unchanged
        sd = {f"mu_{w}": c["mu_w"], f"rho_{w}": c["rho_w"]}
        if bias:
            sd.update(mu_bias=c["mu_b"], rho_bias=c["rho_b"])
        layer.load_state_dict(sd)
        assert torch.equal(getattr(layer, f"rho_{w}").data, c["rho_w"])


def test_init_matches_reference_under_same_seed(golden):
    meta = golden.meta
    torch.manual_seed(0)
    m = L.LinearReparameterization(1024, 1024)
    assert float(m.mu_weight.double().sum()) == pytest.approx(meta["c1"]["mu_sum"], rel=1e-12)
    assert float(m.rho_weight.double().sum()) == pytest.approx(meta["c1"]["rho_sum"], rel=1e-12)
    assert m.posterior_mu_init == (0,) and m.posterior_rho_init == (-3.0,)   # 1-tuple quirk
    assert float(m.prior_weight_sigma[0, 0]) == 1.0 and m.eps_weight.shape == (1024, 1024)
    torch.manual_seed(0)
    m = L.LinearFlipout(256, 128)
    assert float(m.mu_weight.double().sum()) == pytest.approx(meta["c1_flip"]["mu_sum"], rel=1e-12)
    assert m.posterior_rho_init == -3.0
    torch.manual_seed(0)
    m = L.Conv2dReparameterization(8, 16, 3)
    assert float(m.mu_kernel.double().sum()) == pytest.approx(meta["c1_conv"]["mu_sum"], rel=1e-12)
    assert m.kernel_size == 3 and m.mu_kernel.shape == (16, 8, 3, 3)
    torch.manual_seed(0)
    m = L.Conv3dReparameterization(4, 8, 3, 0, 1, 0, -3.0)   # positional quirk (conv_variational.py:406-418)
    assert float(m.mu_kernel.double().sum()) == pytest.approx(meta["c1_conv3d"]["mu_sum"], rel=1e-12)


def test_constructor_quirks():
    with pytest.raises(ValueError, match="invalid in_channels size"):
        L.Conv2dReparameterization(6, 8, 3, groups=4)
    L.Conv2dFlipout(8, 8, 3, groups=4)                      # the Flipout classes do not validate
    a = L.Conv1dReparameterization(4, 6, 3)                 # int kernel (reference form)
    b = L.Conv1dReparameterization(4, 6, (3,))              # tuple kernel (what dnn_to_bnn passes)
    assert a.mu_kernel.shape == b.mu_kernel.shape == (6, 4, 3)
    m = L.LinearReparameterization(4, 3, bias=False)
    assert m.mu_bias is None and m.eps_bias is None and list(m.state_dict()) == ["mu_weight", "rho_weight"]
    assert m.dnn_to_bnn_flag is False
    m.dnn_to_bnn_flag = True
    assert m.dnn_to_bnn_flag is True
    s = torch.tensor([0.5, 2.0])
    kl = m.kl_div(torch.zeros(2), s, torch.zeros(2), torch.ones(2))
    assert float(kl) == pytest.approx(float((-torch.log(s) + s ** 2 / 2 - 0.5).mean()))


class TinyCNN(nn.Module):
    def __init__(self):
        super().__init__()
        self.features = nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.BatchNorm2d(8), nn.ReLU(),
                                      nn.Conv2d(8, 16, 3, stride=2, padding=1, bias=False), nn.ReLU())
        self.conv3d = nn.Conv3d(2, 4, 3)
        self.fc = nn.Linear(16, 10)

    def forward(self, x):
        return self.fc(self.features(x).mean((2, 3)))


@pytest.mark.parametrize("typ", ["Reparameterization", "Flipout"])
def test_dnn_to_bnn_surgery(golden, typ):
    prm = {"prior_mu": 0.0, "prior_sigma": 1.0, "posterior_mu_init": 0.0, "posterior_rho_init": -3.0,
           "type": typ, "moped_enable": False, "moped_delta": 0.5}
    torch.manual_seed(7)
    net = TinyCNN()
    assert dnn_to_bnn(net, prm) is None
    ref = golden.meta[f"tiny_{typ}"]
    assert {k: list(v.shape) for k, v in net.state_dict().items()} == ref["state_dict"]
    assert {n: type(m).__name__ for n, m in net.named_modules() if hasattr(m, "kl_loss")} == ref["classes"]
    assert all(m.dnn_to_bnn_flag for m in net.modules() if hasattr(m, "kl_loss"))
    # MOPED
    torch.manual_seed(7)
    net = TinyCNN()
    w = net.fc.weight.detach().clone()
    dnn_to_bnn(net, dict(prm, moped_enable=True, moped_delta=0.3))
    c = golden.case(f"moped_{typ}")
    assert torch.equal(net.fc.mu_weight.data, c["w"]) and torch.equal(w, c["w"])
    assert torch.equal(net.fc.rho_weight.data, c["rho"])


def test_dnn_to_bnn_resnet18_structure(golden):
    torchvision = pytest.importorskip("torchvision")
    prm = {"prior_mu": 0.0, "prior_sigma": 1.0, "posterior_mu_init": 0.0, "posterior_rho_init": -3.0,
           "type": "Reparameterization", "moped_enable": False, "moped_delta": 0.5}
    torch.manual_seed(11)
    net = torchvision.models.resnet18(num_classes=10)
    dnn_to_bnn(net, prm)
    ref = golden.meta["resnet18"]
    assert list(net.state_dict().keys()) == ref["state_dict_keys"]
    assert btb.mc.count_bayes_layers(net) == ref["n_bayes_layers"] == 21
    n_mu = sum(p.numel() for n, p in net.named_parameters() if n.split(".")[-1].startswith("mu_"))
    assert n_mu == ref["n_mu"]
    with pytest.raises(NotImplementedError):
        dnn_to_bnn(nn.Sequential(nn.LSTM(4, 4)), prm)


def test_no_cpu_fallback():
    m = L.LinearReparameterization(8, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(2, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.kl_loss()
    c = L.Conv2dFlipout(3, 4, 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        c(torch.randn(1, 3, 8, 8))
    with pytest.raises(RuntimeError):
        btb.mc_predict(nn.Sequential(m).eval(), torch.randn(2, 8), 4)


def test_c_abi_exports_every_declared_symbol():
    with open(os.path.join(ROOT, "include", "btb200.h")) as f:
        hdr = f.read()
    declared = sorted(set(re.findall(r"\b(bt_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 11
    assert os.path.exists(_native.LIB_PATH), "libbtb200.so not built (python -m bayesian_torch_b200.build)"
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in btb200.h but not exported"
    assert sorted(s[0] for s in _native.SYMBOLS) == declared
    lib.bt_version.restype = ctypes.c_int
    assert lib.bt_version() == int(re.search(r"#define BT_VERSION (\d+)", hdr).group(1))
    assert ctypes.sizeof(_native.BtLayerGeom) == 4 * (6 + 18 + 1) and ctypes.sizeof(_native.BtDebugIO) == 32


def test_c_abi_rejects_host_pointers_without_gpu():
    lib = _native.load()
    buf = (ctypes.c_float * 16)()
    rc = lib.bt_kl_gaussian(ctypes.addressof(buf), ctypes.addressof(buf), 16, None, None, None, None, 0, None, None,
                            0.0, 1.0, 0, ctypes.addressof(buf), 0, ctypes.addressof(buf), None)
    assert rc < 0 and len(lib.bt_last_error()) > 0
    rc = lib.bt_kl_gaussian(None, None, 0, None, None, None, None, 0, None, None, 0.0, 1.0, 0, None, 0, None, None)
    assert rc == -1 and b"n_w" in lib.bt_last_error()


def test_shard_samples_partition():
    for n in (1, 7, 32, 64):
        for w in (1, 2, 3, 4, 8):
            blocks = [shard_samples(n, w, r) for r in range(w)]
            assert sum(c for _, c in blocks) == n
            pos = 0
            for s, c in blocks:
                assert s == pos
                pos += c
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
    with pytest.raises(ValueError):
        shard_samples(4, 2, 2)

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
    # nn.LSTM -> LSTM<type> on two Bayesian linears, nn.ConvTranspose2d through the "Conv" branch
    # (/root/reference/bayesian_torch/models/dnn_to_bnn.py:106-122, :138-144)
    seq = nn.Sequential(nn.LSTM(4, 6), nn.ConvTranspose2d(4, 8, 3, stride=2))
    dnn_to_bnn(seq, prm)
    assert type(seq[0]).__name__ == "LSTM" + prm["type"] and seq[0].dnn_to_bnn_flag
    assert tuple(seq[0].ih.mu_weight.shape) == (24, 4) and tuple(seq[0].hh.mu_weight.shape) == (24, 6)
    assert sorted(k for k in seq[0].state_dict()) == sorted(
        f"{m}.{p}" for m in ("ih", "hh") for p in ("mu_weight", "rho_weight", "mu_bias", "rho_bias"))
    assert type(seq[1]).__name__ == "ConvTranspose2d" + prm["type"]
    assert tuple(seq[1].mu_kernel.shape) == (4, 8, 3, 3) and seq[1].stride == (2, 2)


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
    # 6 + 18 + 2 int32 (104 bytes) + one pointer
    assert ctypes.sizeof(_native.BtLayerGeom) == 4 * (6 + 18 + 2 + 2) + 8 and ctypes.sizeof(_native.BtDebugIO) == 32


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


# ---------------------------------------------------------------- host tiling / kernel selection (bt_layer_forward_plan)
def _geom(S, B, cin, cout, sp, k, stride=1, pad=0, dil=1, groups=1, x_shared=0):
    g = _native.BtLayerGeom()
    g.n_samples, g.x_shared, g.batch, g.c_in, g.c_out, g.groups = S, x_shared, B, cin, cout, groups
    nd = len(sp)
    for i in range(3):
        g.in_dhw[i] = g.out_dhw[i] = g.k_dhw[i] = g.stride[i] = g.dil[i] = 1
        g.pad[i] = 0
    for i, n in enumerate(sp):
        j = 3 - nd + i
        g.in_dhw[j], g.k_dhw[j], g.stride[j], g.pad[j], g.dil[j] = n, k, stride, pad, dil
        g.out_dhw[j] = (n + 2 * pad - dil * (k - 1) - 1) // stride + 1
    g.rho_is_sigma = 0
    return g


SMEM_MAX = 227 * 1024
BF, F32 = torch.bfloat16, torch.float32


def test_plan_resnet18_cifar_layers_on_148_sms():
    """The kernel family and tiling bt_layer_forward picks for the C3 layers (B=128, S=64 samples per launch)."""
    plan = lambda g, **kw: _native.plan_forward(_native.MODE_REPARAM, g, kw.pop("x", BF), kw.pop("p", BF), **kw)
    l1 = plan(_geom(64, 128, 64, 64, (8, 8), 3, pad=1))                       # layer1 3x3: in-place windows staged by TMA
    assert l1["path"] == "tma_direct" and l1["block_n"] == 64 and l1["grid"][1:] == (1, 64) and l1["threads"] == 384
    assert l1["k_blocks"] == 9 and l1["tmem_cols"] == 128 and l1["window_rows"] % 8 == 0 and l1["window_slots"] >= 3
    l2 = plan(_geom(64, 128, 128, 128, (4, 4), 3, pad=1), with_residual=True)  # layer2 3x3 (+ residual epilogue)
    assert l2["path"] == "tma_direct" and l2["k_blocks"] == 18 and l2["window_slots"] >= 2
    l1f = plan(_geom(64, 128, 64, 64, (8, 8), 3, pad=1), x=F32, p=F32)         # fp32 model: tf32 windows, 32 k per k-block
    assert l1f["path"] == "tma_direct" and l1f["k_blocks"] == 18 and l1f["threads"] == 512
    l3 = plan(_geom(64, 128, 256, 256, (2, 2), 3, pad=1))                      # layer3: 4 row tiles share every sampled tile,
    assert l3["path"] == "tma_stream" and l3["m_subtiles"] == 4                # activations staged by TMA (im2col map)
    assert l3["k_blocks"] == 36 and l3["tmem_cols"] == l3["m_subtiles"] * l3["block_n"] and l3["threads"] == 384
    assert l3["cluster_n"] == 2 and l3["grid"][1] % 2 == 0      # the two n-tile CTAs of a sample share their A tiles (multicast)
    l4 = plan(_geom(64, 128, 512, 512, (1, 1), 3, pad=1))                      # layer4 at 1x1: only the centre tap is real
    assert l4["k_blocks"] == 8 and l4["path"].startswith("tma")
    ds = plan(_geom(64, 128, 64, 128, (8, 8), 1, stride=2))                    # 1x1 stride-2 downsample: im2col map, W_s resident
    assert ds["path"] == "tma" and ds["k_blocks"] == 1
    stem = plan(_geom(64, 128 * 256, 192, 64, (), 1, x_shared=1))              # materialised-im2col stem = a linear layer
    assert stem["path"] == "tma" and stem["block_n"] == 64 and stem["k_blocks"] == 3
    assert stem["samples_per_cta"] == 4 and stem["grid"][2] == 16 and stem["tmem_cols"] == 512   # shared x: 4 samples per CTA
    assert stem["pool_fused"] == 0
    # ... with torchvision's stem max-pool inside the epilogue (16x16 output rows: two tiles per image; the CTA walks
    # whole images, the tile buffer + two carry rows per sample fit next to the resident tiles)
    gp = _geom(64, 128 * 256, 192, 64, (), 1, x_shared=1)
    gp.pool_hw[0] = gp.pool_hw[1] = 16
    sp = plan(gp)
    assert sp["pool_fused"] == 1 and sp["path"] == "tma" and sp["block_n"] == 64 and sp["smem_bytes"] <= SMEM_MAX
    assert sp["samples_per_cta"] == 4 and sp["window_slots"] >= 3 and 128 % sp["grid"][0] in range(128)
    gp32 = _geom(64, 128 * 256, 152, 64, (), 1, x_shared=1)
    gp32.pool_hw[0] = gp32.pool_hw[1] = 16
    sp32 = plan(gp32, x=F32, p=F32)
    assert sp32["pool_fused"] == 1 and sp32["smem_bytes"] <= SMEM_MAX and sp32["samples_per_cta"] >= 2
    gp.pool_hw[0] = gp.pool_hw[1] = 12                                         # 12 does not divide the 128-row tile
    with pytest.raises(ValueError):
        plan(gp)                                                               # (and 144 does not divide the row count)
    gq = _geom(4, 6 * 144, 64, 64, (), 1, x_shared=1)
    gq.pool_hw[0] = gq.pool_hw[1] = 12
    assert plan(gq)["pool_fused"] == 0
    gi = _geom(2, 8, 3, 64, (224, 224), 7, stride=2, pad=3)                    # ImageNet stem: 112 columns -> separate pool
    gi.c_in = 8
    gi.pool_hw[0] = gi.pool_hw[1] = 112
    assert plan(gi)["pool_fused"] == 0
    c1 = plan(_geom(1, 256, 1024, 1024, (), 1), x=F32, p=F32)                  # C1: fp32 -> tf32 operands, 32 k per k-block
    assert c1["path"].startswith("tma") and c1["k_blocks"] == 32
    # what forces the generic instantiation
    assert plan(_geom(1, 128, 64, 64, (8, 8), 3, pad=1), with_kl=True)["path"] == "generic"
    assert plan(_geom(1, 128, 64, 64, (8, 8), 3, pad=1), with_debug_hooks=True)["path"] == "generic"
    # fewer samples per rank (N-GPU sharding): the row tiles of a sample are spread over more CTAs
    l1s = plan(_geom(8, 128, 64, 64, (8, 8), 3, pad=1))
    assert l1s["path"] == "tma_direct" and l1s["grid"][2] == 8 and l1s["grid"][0] > 8


def test_plan_invariants_over_a_geometry_sweep():
    """Whatever the geometry: the plan fits the SM (227 KB smem, 512 TMEM columns as a power of two), covers every
    output column, and the direct kernel is only chosen where its trick is exact."""
    import itertools
    import random
    rnd = random.Random(5)
    n = 0
    for mode, xdt in itertools.product((_native.MODE_REPARAM, _native.MODE_FLIPOUT), (BF, F32)):
        for _ in range(150):
            nd = rnd.choice((0, 1, 2, 2, 2, 3))
            groups = rnd.choice((1, 1, 1, 2, 4))
            cin = groups * rnd.choice((3, 8, 16, 24, 64, 128, 192, 256))
            cout = groups * rnd.choice((5, 10, 32, 64, 100, 128, 256))
            k = rnd.choice((1, 1, 3, 3, 5)) if nd else 1
            stride, dil = (rnd.choice((1, 1, 2)), rnd.choice((1, 1, 2))) if nd else (1, 1)
            pad = rnd.choice((0, (k - 1) * dil // 2))
            sp = tuple(rnd.choice((1, 2, 4, 7, 8, 14, 28)) for _ in range(nd))
            if any(s_ + 2 * pad < dil * (k - 1) + 1 for s_ in sp):
                continue
            g = _geom(rnd.choice((1, 2, 8, 64)), rnd.choice((1, 7, 32, 128)), cin, cout, sp, k, stride, pad, dil, groups)
            p = _native.plan_forward(mode, g, xdt, rnd.choice((BF, F32)), with_residual=rnd.random() < 0.3,
                                     sm_count=rnd.choice((148, 132, 60)))
            n += 1
            assert 0 < p["smem_bytes"] <= SMEM_MAX, (p, sp)
            assert p["tmem_cols"] in (32, 64, 128, 256, 512)
            assert p["block_n"] in (32, 64, 128) and p["threads"] in (288, 384, 416, 512, 544)
            assert all(v >= 1 for v in p["grid"]) and p["grid"][2] == -(-g.n_samples // max(p["samples_per_cta"], 1))
            assert p["grid"][1] == -(-(cout // groups) // p["block_n"]) * groups
            if p["path"] == "tma_direct":
                assert groups == 1 and stride == 1 and k > 1
                assert cin % (64 if xdt == BF else 32) == 0 and all(2 * pad == dil * (k - 1) for _ in sp)
            if p["path"] == "direct":
                assert xdt == BF and groups == 1 and cin % 64 == 0 and stride == 1
                assert all(2 * pad == dil * (k - 1) for _ in sp)              # "same" output extent
                assert p["window_slots"] >= 2 and p["window_rows"] % 8 == 0 and p["threads"] == 512
            else:
                assert p["m_subtiles"] in (1, 2, 4)
            if p["path"].startswith("tma"):
                assert (mode == _native.MODE_REPARAM or p["path"] in ("tma_stream", "tma_direct")) and cin % 8 == 0
                assert (cin // groups) % (64 if xdt == BF else 32) == 0 or (k == 1 and stride == 1 and groups == 1)
    assert n > 300


def test_plan_rejects_bad_geometry():
    g = _geom(1, 4, 64, 64, (8, 8), 3, pad=1)
    g.out_dhw[2] = 5                                   # inconsistent output extent
    with pytest.raises((RuntimeError, ValueError), match="inconsistent"):
        _native.plan_forward(_native.MODE_REPARAM, g, BF, BF)
    g = _geom(1, 4, 64, 96, (8, 8), 3, pad=1, groups=5)
    with pytest.raises((RuntimeError, ValueError), match="divisible"):
        _native.plan_forward(_native.MODE_REPARAM, g, BF, BF)


@pytest.mark.parametrize("arch,res", [("resnet18", 32), ("resnet50", 224)])
def test_plan_covers_every_layer_of_the_benchmark_models(arch, res):
    """BASELINE.json configs 3 and 4: every Bayesian conv / linear of dnn_to_bnn(ResNet-18 @32, ResNet-50 @224) gets a
    plan that fits the SM, for both families and for the per-rank sample counts of 1-8 GPU runs."""
    import torchvision
    net = getattr(torchvision.models, arch)(num_classes=10).eval()
    shapes = []

    def hook(m, inp, out):
        if isinstance(m, nn.Conv2d):
            shapes.append((m.in_channels, m.out_channels, tuple(inp[0].shape[2:]), m.kernel_size[0], m.stride[0], m.padding[0]))
        elif isinstance(m, nn.Linear):
            shapes.append((m.in_features, m.out_features, (), 1, 1, 0))

    hooks = [m.register_forward_hook(hook) for m in net.modules()]
    with torch.no_grad():
        net(torch.zeros(1, 3, res, res))
    for h in hooks:
        h.remove()
    assert len(shapes) == {"resnet18": 21, "resnet50": 54}[arch]
    n_direct = 0
    for mode in (_native.MODE_REPARAM, _native.MODE_FLIPOUT):
        for S in (64, 32, 8, 4):
            for cin, cout, sp, k, st, pd in shapes:
                cin_k = (cin + 7) // 8 * 8                 # the stem's RGB input is channel-padded by the layer class
                p = _native.plan_forward(mode, _geom(S, 128, cin_k, cout, sp, k, st, pd), BF, BF)
                assert 0 < p["smem_bytes"] <= SMEM_MAX and p["tmem_cols"] <= 512 and p["grid"][2] == S
                n_direct += p["path"] in ("direct", "tma_direct")
    assert n_direct > 0


def test_fuse_inference_offers_the_stem_pool_and_skips_the_1x1_avgpool():
    """Host side of the fused stem (no GPU): fuse_inference folds bn1 + relu into conv1, offers the 3x3/2 max-pool to
    conv1's kernel (_bt_ep_pool; taken at launch time only when BtForwardPlan.pool_fused says so), keeps every
    state_dict key, and replaces the global average pool by a pass-through for 1x1 inputs."""
    import torchvision
    from bayesian_torch_b200 import fuse
    torch.manual_seed(0)
    net = torchvision.models.resnet18(num_classes=10)
    keys = None
    btb.dnn_to_bnn(net, {"prior_mu": 0.0, "prior_sigma": 1.0, "posterior_mu_init": 0.0, "posterior_rho_init": -3.0,
                         "type": "Reparameterization", "moped_enable": False, "moped_delta": 0.5})
    keys = set(net.state_dict().keys())
    net.eval()
    fuse.fuse_inference(net)
    assert set(net.state_dict().keys()) == keys
    assert net.conv1._bt_ep_pool is True and net.conv1._bt_ep_relu is True and net.conv1._bt_ep_scale is not None
    assert isinstance(net.maxpool, fuse.FusedMaxPool2d) and isinstance(net.avgpool, fuse.FusedAvgPool)
    x = torch.randn(2, 512, 1, 1)
    assert net.avgpool(x) is x                                   # 1x1: the mean of one element
    y = torch.randn(2, 512, 3, 3)
    assert torch.allclose(net.avgpool(y), y.mean((2, 3), keepdim=True))
    t = torch.randn(1, 4, 4, 4)
    t._bt_pooled = True                                          # what a pooled conv launch hands over
    assert net.maxpool(t) is t
    net.train()
    with pytest.raises(RuntimeError):
        fuse.fuse_inference(net)

"""-m gpu: parity of the BASELINE.json configurations AT FULL SIZE and in exactly the form bench.py times them.

  C2  Conv2dFlipout 64->128 k3 p1, 56x56, B=128, bf16            vs the oracle on the re-materialised eps / signs
  C5  LinearFlipout 4096->4096, B=4096, bf16                      "
  C3  dnn_to_bnn(ResNet-18) Reparameterization, 3x32x32, B=128, N=64, channels-last, fuse_inference, CUDA graph
      (bf16 model and fp32 model): per-sample logits of 3 samples and the N=64 predictive mean / variance vs the fp32
      CPU twin built from the eps the kernels used (oracle MC on identical draws)
  C4  dnn_to_bnn(ResNet-50) Flipout, 3x224x224 (small B), fused (FusedBottleneck) and unfused, vs the fp32 CPU twin
      built from the materialised eps AND signs (conv_flipout.py:370-439 through torchvision Bottleneck)

Tolerances are stated per test (layer level: the SURVEY.md 8c contract -- bf16 path rel-RMS <= 3e-3 vs fp32)."""
import copy

import pytest
import torch
import torch.nn as nn

import bayesian_torch_b200 as btb
from bayesian_torch_b200 import _native
from bayesian_torch_b200._core import BayesLayerBase
from gpu_util import build_layer, errs, note, oracle_forward
from oracle import bt_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
PRM = {"prior_mu": 0.0, "prior_sigma": 1.0, "posterior_mu_init": 0.0, "posterior_rho_init": -3.0,
       "type": "Reparameterization", "moped_enable": False, "moped_delta": 0.5}


def test_c2_conv2d_flipout_full_size():
    torch.manual_seed(0)
    btb.manual_seed(0)
    lay = build_layer("conv", 2, True, 64, 128, 3, 1, 1).to(DEV).bfloat16()
    x = torch.randn(128, 64, 56, 56, device=DEV, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        y = lay(x, return_kl=False)
    path = _native.last_forward_path()
    lay._bt_last["sample0"] = 0
    eps_w, eps_b = lay.materialize_eps(0)
    s_in, s_out = lay.materialize_signs(tuple(x.shape), tuple(y.shape), 0)
    torch.set_num_threads(max(torch.get_num_threads(), 16))
    ref = oracle_forward(lay, x, eps_w, eps_b, s_in, s_out, round_operands=False)
    rel, mx = errs(y, ref)
    note("C2_full", path=path, rel=rel, max_abs=mx)
    assert rel <= 3e-3, (path, rel, mx)


@pytest.mark.parametrize("flip", [True, False], ids=["flipout", "reparam"])
def test_c5_linear_4096_full_size(flip):
    torch.manual_seed(0)
    btb.manual_seed(0)
    lay = build_layer("linear", 0, flip, 4096, 4096, None).to(DEV).bfloat16()
    x = torch.randn(4096, 4096, device=DEV, dtype=torch.bfloat16)
    with torch.no_grad():
        y = lay(x, return_kl=False)
    path = _native.last_forward_path()
    lay._bt_last["sample0"] = 0
    eps_w, eps_b = lay.materialize_eps(0)
    s_in = s_out = None
    if flip:
        s_in, s_out = lay.materialize_signs(tuple(x.shape), tuple(y.shape), 0)
    torch.set_num_threads(max(torch.get_num_threads(), 16))
    ref = oracle_forward(lay, x, eps_w, eps_b, s_in, s_out, round_operands=False)
    rel, mx = errs(y, ref)
    note("C5_full", flip=flip, path=path, rel=rel, max_abs=mx)
    assert rel <= 3e-3, (path, rel, mx)


# ---------------------------------------------------------------------------------------------- whole models
class _Twin(nn.Module):
    """CPU fp32 twin of one Bayesian layer evaluated with EXPLICIT draws (the oracle's op sequence)."""

    def __init__(self, layer, eps_w, eps_b, s_in=None, s_out=None):
        super().__init__()
        f = lambda t: None if t is None else t.detach().float().cpu()
        self.nd, self.flip = layer._nd, layer._family == "flipout"
        mu_w, rho_w = layer._mu_rho()
        self.mu_w, self.rho_w, self.mu_b, self.rho_b = f(mu_w), f(rho_w), f(layer.mu_bias), f(layer.rho_bias)
        self.eps_w, self.eps_b, self.s_in, self.s_out = f(eps_w), f(eps_b), f(s_in), f(s_out)
        self.geo = None if self.nd == 0 else (layer.stride, layer.padding, layer.dilation, layer.groups)

    def forward(self, x):
        if self.nd == 0:
            if self.flip:
                return O.linear_flipout(x, self.mu_w, self.rho_w, self.eps_w, self.s_in, self.s_out, self.mu_b, self.rho_b, self.eps_b)
            return O.linear_reparam(x, self.mu_w, self.rho_w, self.eps_w, self.mu_b, self.rho_b, self.eps_b)
        if self.flip:
            return O.conv_flipout(self.nd, x, self.mu_w, self.rho_w, self.eps_w, self.s_in, self.s_out, self.mu_b, self.rho_b,
                                  self.eps_b, *self.geo)
        return O.conv_reparam(self.nd, x, self.mu_w, self.rho_w, self.eps_w, self.mu_b, self.rho_b, self.eps_b, *self.geo)


def _set_module(root, name, new):
    parts = name.split(".")
    for p_ in parts[:-1]:
        root = getattr(root, p_)
    setattr(root, parts[-1], new)


def _twin_net(bnn_unfused, det, sample, shapes=None):
    """deterministic torchvision net (eval) whose conv / linear modules are replaced by oracle twins carrying the draws
    the kernels used for MC sample `sample` (materialised from the Philox counters)."""
    twin = copy.deepcopy(det).eval()
    for name, m in bnn_unfused.named_modules():
        if isinstance(m, BayesLayerBase):
            eps_w, eps_b = m.materialize_eps(sample)
            s_in = s_out = None
            if m._family == "flipout":
                xs, ys = shapes[name]
                s_in, s_out = m.materialize_signs(xs, ys, sample)
            _set_module(twin, name, _Twin(m, eps_w, eps_b, s_in, s_out))
    return twin


def _resnet(arch, typ, classes, seed=11, res=32):
    """A WELL-CONDITIONED random network: the deterministic torchvision net gets BatchNorm statistics calibrated on
    random inputs (train-mode passes), and the Bayesian twin is MOPED-initialised from it (mu = w, sigma = 0.1 |w|,
    models/dnn_to_bnn.py:65-71) -- activations stay O(1) through every block and the logits are soft, so the predictive
    moments are a meaningful comparison (an uncalibrated random ResNet saturates its softmax)."""
    torchvision = pytest.importorskip("torchvision")
    torch.manual_seed(seed)
    net = getattr(torchvision.models, arch)(num_classes=classes)
    net.train()
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.momentum = None                      # cumulative average over the calibration batches
        for _ in range(3):
            net(torch.randn(16 if res <= 64 else 4, 3, res, res))
    det = copy.deepcopy(net).eval()
    btb.dnn_to_bnn(net, dict(PRM, type=typ, moped_enable=True, moped_delta=0.1))
    btb.assign_layer_keys(net)
    return net.eval(), det


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_c3_bench_configuration_vs_oracle_mc(dtype):
    """EXACTLY what bench.py times: channels-last, fuse_inference, use_graph=True, B=128, N=64."""
    B, N = 128, 64
    bnn, det = _resnet("resnet18", "Reparameterization", 10)
    bnn = bnn.to(DEV).to(dtype).to(memory_format=torch.channels_last)
    btb.fuse_inference(bnn)
    btb.manual_seed(0)
    torch.manual_seed(1234)
    x_cpu = torch.randn(B, 3, 32, 32)
    x = x_cpu.to(DEV).to(dtype).contiguous(memory_format=torch.channels_last)
    mean, var = btb.mc_predict(bnn, x, N, use_graph=True)            # capture + replay
    mean2, var2 = btb.mc_predict(bnn, x, N, use_graph=True)          # pure replay
    assert torch.equal(mean, mean2) and torch.equal(var, var2)
    torch.cuda.synchronize()
    # per-sample logits of three samples through the same fused model (eager; graph == eager is pinned bit-exactly
    # by test_gpu_model.py::test_mc_predict_cuda_graph_replay_equals_eager)
    picks = (0, 31, 63)
    with torch.no_grad():
        logits = {}
        for s in picks:
            with btb.mc_sample_context(1, B, s):
                logits[s] = bnn(x).float().cpu()
    torch.set_num_threads(max(torch.get_num_threads(), 16))
    xin = x.float().cpu()                                             # the (rounded) input the GPU model saw
    probs = []
    with torch.no_grad():
        for s in range(N):
            for m in bnn.modules():
                if isinstance(m, BayesLayerBase) and m._bt_last is not None:
                    m._bt_last["sample0"] = 0
            ref = _twin_net(bnn, det, s)(xin)
            probs.append(torch.softmax(ref, -1))
            if s in picks:
                rel, mx = errs(logits[s], ref)
                note("C3_bench_logits", dtype=str(dtype), sample=s, rel=rel, max_abs=mx)
                # 21 stacked layers: bf16 model = bf16 activations + bf16 operands; fp32 model = tf32 operands
                assert rel <= (5e-2 if dtype == torch.bfloat16 else 6e-3), (s, rel, mx)
    p = torch.stack(probs)
    ref_mean, ref_var = p.mean(0), (p * p).mean(0) - p.mean(0) ** 2
    dm = float((mean.cpu() - ref_mean).abs().max())
    dv = float((var.cpu() - ref_var).abs().max())
    note("C3_bench_moments", dtype=str(dtype), mean_max_abs=dm, var_max_abs=dv, ref_var_max=float(ref_var.max()))
    assert dm <= (6e-3 if dtype == torch.bfloat16 else 1e-3), dm       # probabilities in [0, 1]
    assert dv <= (2e-3 if dtype == torch.bfloat16 else 2e-4), dv


@pytest.mark.parametrize("fused", [False, True], ids=["unfused", "fused"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_c4_resnet50_flipout_224_vs_oracle(dtype, fused):
    B = 2
    bnn, det = _resnet("resnet50", "Flipout", 10, seed=5, res=224)
    bnn = bnn.to(DEV).to(dtype).to(memory_format=torch.channels_last)
    shapes = {}
    hooks = []
    for name, m in bnn.named_modules():
        if isinstance(m, BayesLayerBase):
            hooks.append(m.register_forward_hook(
                lambda mod, inp, out, name=name: shapes.__setitem__(name, (tuple(inp[0].shape), tuple(out.shape)))))
    btb.manual_seed(3)
    torch.manual_seed(7)
    x = torch.randn(B, 3, 224, 224).to(DEV).to(dtype).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        with btb.mc_sample_context(1, B, 0):
            y_plain = bnn(x).float().cpu()                 # unfused: also records every layer's activation shapes
    for h in hooks:
        h.remove()
    if fused:
        btb.fuse_inference(bnn)
        assert sum(1 for m in bnn.modules() if type(m).__name__ == "FusedBottleneck") == 16
        with torch.no_grad():
            with btb.mc_sample_context(1, B, 0):
                y = bnn(x).float().cpu()
    else:
        y = y_plain
    torch.set_num_threads(max(torch.get_num_threads(), 16))
    with torch.no_grad():
        ref = _twin_net(bnn, det, 0, shapes)(x.float().cpu())
    rel, mx = errs(y, ref)
    note("C4_resnet50_flipout", dtype=str(dtype), fused=fused, rel=rel, max_abs=mx)
    assert y.shape == (B, 10)
    # 54 stacked Flipout layers (two tensor-core products each); bf16 also rounds every activation to 8 bits
    assert rel <= (2e-1 if dtype == torch.bfloat16 else 3e-2), (rel, mx)
    pd = float((torch.softmax(y, -1) - torch.softmax(ref, -1)).abs().max())
    note("C4_resnet50_flipout_probs", dtype=str(dtype), fused=fused, prob_max_abs=pd)
    assert pd <= (6e-2 if dtype == torch.bfloat16 else 6e-3), pd

"""-m gpu: whole-model checks through dnn_to_bnn(torchvision ResNet-18) -- the headline workload
(BASELINE.json configs[2]) -- and the MC driver."""
import copy

import pytest
import torch
import torch.nn as nn

import bayesian_torch_b200 as btb
from bayesian_torch_b200 import _native
from bayesian_torch_b200._core import BayesLayerBase
from gpu_util import errs
from oracle import bt_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
PRM = {"prior_mu": 0.0, "prior_sigma": 1.0, "posterior_mu_init": 0.0, "posterior_rho_init": -3.0,
       "type": "Reparameterization", "moped_enable": False, "moped_delta": 0.5}


def _resnet18(typ="Reparameterization", seed=11):
    torchvision = pytest.importorskip("torchvision")
    torch.manual_seed(seed)
    net = torchvision.models.resnet18(num_classes=10)
    det = copy.deepcopy(net)
    btb.dnn_to_bnn(net, dict(PRM, type=typ))
    btb.assign_layer_keys(net)
    return net.to(DEV).eval(), det.eval()


def _oracle_net(bnn, det, sample=0):
    """CPU fp32 deterministic twin whose conv/linear weights are the oracle's W = mu + sp(rho) * eps built from the
    eps the kernels used (materialised from the Philox counters)."""
    bayes = {n: m for n, m in bnn.named_modules() if isinstance(m, BayesLayerBase)}
    for name, m in det.named_modules():
        if name in bayes:
            b = bayes[name]
            eps_w, eps_b = b.materialize_eps(sample)
            mu_w, rho_w = b._mu_rho()
            w = mu_w.detach().float().cpu() + O.sigma_of_rho(rho_w.detach().float().cpu()) * eps_w.float().cpu()
            m.weight.data.copy_(w)
            if b.mu_bias is not None:
                m.bias.data.copy_(b.mu_bias.detach().float().cpu() +
                                  O.sigma_of_rho(b.rho_bias.detach().float().cpu()) * eps_b.float().cpu())
    return det


def test_resnet18_forward_matches_oracle_network():
    bnn, det = _resnet18()
    btb.manual_seed(2024)
    x = torch.randn(8, 3, 32, 32, device=DEV)
    with torch.no_grad():
        y = bnn(x)
        torch.cuda.synchronize()
        ref = _oracle_net(bnn, det)(x.cpu())
    assert y.shape == (8, 10)
    rel, mx = errs(y, ref)
    assert rel <= 3e-2, (rel, mx)     # 21 stacked layers of bf16-operand tensor-core math vs fp32 CPU
    kl = btb.get_kl_loss(bnn)
    mu_rho = [(m._mu_rho(), m.mu_bias, m.rho_bias) for m in bnn.modules() if isinstance(m, BayesLayerBase)]
    kref = sum(float(O.kl_loss(a.detach().double().cpu(), b.detach().double().cpu(), 0.0, 1.0,
                               None if mb is None else mb.detach().double().cpu(),
                               None if rb is None else rb.detach().double().cpu())) for (a, b), mb, rb in mu_rho)
    assert abs(float(kl) - kref) <= 1e-4 * abs(kref)


@pytest.mark.parametrize("typ,dtype", [("Reparameterization", torch.bfloat16), ("Flipout", torch.float32)])
def test_mc_predict_equals_sequential_reference_style_loop(typ, dtype):
    """mc_predict (samples stacked in one pass, fused softmax/moments) == the reference's evaluate() loop
    (examples/main_bayesian_cifar_dnn2bnn.py:541-557) run sample by sample with the same sample indices."""
    bnn, _ = _resnet18(typ)
    bnn = bnn.to(dtype).to(memory_format=torch.channels_last)
    btb.manual_seed(7)
    B, N = 16, 6
    x = torch.randn(B, 3, 32, 32, device=DEV, dtype=dtype)
    mean, var = btb.mc_predict(bnn, x, N, chunk=4)
    outs = []
    with torch.no_grad():
        for s in range(N):
            with btb.mc_sample_context(1, B, s):
                outs.append(bnn(x))
    ref_mean, ref_var = O.mc_aggregate(torch.stack(outs).float().cpu())
    assert torch.allclose(mean.cpu(), ref_mean, atol=2e-5), float((mean.cpu() - ref_mean).abs().max())
    assert torch.allclose(var.cpu(), ref_var, atol=2e-5)
    assert float(var.max()) > 0          # the samples differ
    # uncertainties of the ensemble (reference utils/util.py:45-60 on the stacked per-sample probabilities), computed on
    # device from running sums; tolerance: fp32 MUFU log / exp (stated: 2e-4 absolute on entropies of O(1) nats)
    btb.manual_seed(7)
    mean2, var2, pe, mi = btb.mc_predict(bnn, x, N, chunk=4, return_uncertainty=True)
    assert torch.equal(mean2, mean) and torch.equal(var2, var)
    probs = torch.softmax(torch.stack(outs).float().cpu(), -1)           # [N, B, C]
    assert torch.allclose(pe.cpu(), O.predictive_entropy(probs), atol=2e-4), float((pe.cpu() - O.predictive_entropy(probs)).abs().max())
    assert torch.allclose(mi.cpu(), O.mutual_information(probs), atol=2e-4)
    assert float(mi.min()) > -1e-4 and float(mi.max()) > 0            # MI >= 0 (Jensen), > 0 since the samples differ
    with pytest.raises(RuntimeError, match="eval"):
        btb.mc_predict(bnn.train(), x, 2)


def test_mc_predict_cuda_graph_replay_equals_eager():
    """use_graph=True captures the rank's whole MC pass once and replays it: same numbers as the eager path, also for a
    new input batch (copied into the graph's static input) and after a parameter update (new graph)."""
    bnn, _ = _resnet18("Reparameterization")
    bnn = bnn.bfloat16().to(memory_format=torch.channels_last)
    btb.fuse_inference(bnn)
    btb.manual_seed(17)
    B, N = 8, 5
    x1 = torch.randn(B, 3, 32, 32, device=DEV).bfloat16()
    x2 = torch.randn(B, 3, 32, 32, device=DEV).bfloat16()
    e1 = btb.mc_predict(bnn, x1, N, return_uncertainty=True)
    e2 = btb.mc_predict(bnn, x2, N, return_uncertainty=True)
    l0 = _native.launch_count
    g1 = btb.mc_predict(bnn, x1, N, return_uncertainty=True, use_graph=True)     # warm-up + capture + first replay
    g2 = btb.mc_predict(bnn, x2, N, return_uncertainty=True, use_graph=True)     # replay with a new input
    g1b = btb.mc_predict(bnn, x1, N, return_uncertainty=True, use_graph=True)
    torch.cuda.synchronize()
    assert _native.launch_count - l0 > 3 * 20          # replays are counted as the launches they contain
    for a, b in zip(e1, g1):
        assert torch.equal(a, b)
    for a, b in zip(e2, g2):
        assert torch.equal(a, b)
    for a, b in zip(g1, g1b):
        assert torch.equal(a, b)
    assert not torch.equal(g1[0], g2[0])
    with torch.no_grad():                                # parameter update -> the cached graph must not be reused
        for m in bnn.modules():
            if hasattr(m, "mu_kernel"):
                m.mu_kernel.add_(0.05)
                break
    g3 = btb.mc_predict(bnn, x1, N, return_uncertainty=True, use_graph=True)
    e3 = btb.mc_predict(bnn, x1, N, return_uncertainty=True)
    assert torch.equal(g3[0], e3[0]) and not torch.equal(g3[0], g1[0])


def test_cuda_graph_replays_draw_fresh_eps_from_a_device_word():
    """The sample offset is a device word the layer kernels read at run time (BtLayerGeom.sample_offset), not part of
    the captured graph: ONE capture, replays with different offsets differ from each other and each equals the eager
    run with that offset; fresh=True advances the offset by n_samples per call (the reference draws new eps on every
    forward, conv_variational.py:362)."""
    from bayesian_torch_b200 import mc as mcmod
    bnn, _ = _resnet18("Reparameterization")
    bnn = bnn.bfloat16().to(memory_format=torch.channels_last)
    btb.fuse_inference(bnn)
    btb.manual_seed(23)
    B, N = 8, 4
    x = torch.randn(B, 3, 32, 32, device=DEV).bfloat16()
    mcmod.drop_graphs()
    outs = {}
    for off in (0, 4, 1000):
        outs[off] = btb.mc_predict(bnn, x, N, sample_offset=off, use_graph=True)
    assert len([k for k in mcmod._graphs if k[0] == id(bnn)]) == 1          # one capture served all three
    assert not torch.equal(outs[0][0], outs[4][0]) and not torch.equal(outs[4][0], outs[1000][0])
    for off in (0, 4, 1000):
        e = btb.mc_predict(bnn, x, N, sample_offset=off)
        assert torch.equal(e[0], outs[off][0]) and torch.equal(e[1], outs[off][1]), off
    # fresh=True: consecutive calls use offsets 0, N, 2N, ... (graph and eager alike)
    bnn._bt_mc_drawn = 0
    f0 = btb.mc_predict(bnn, x, N, use_graph=True, fresh=True)
    f1 = btb.mc_predict(bnn, x, N, use_graph=True, fresh=True)
    assert torch.equal(f0[0], outs[0][0]) and torch.equal(f1[0], outs[4][0])
    # the materialised eps of a layer follows the effective (offset) sample index
    conv = bnn.layer1[0].conv1
    e_a, _ = conv.materialize_eps(0)
    btb.mc_predict(bnn, x, N, sample_offset=4)
    e_b, _ = conv.materialize_eps(0)
    assert torch.equal(e_a, e_b)


# fp32 activations: the only difference is fma-vs-mul/add rounding (1e-7) on values that are then rounded to bf16
# operands by the next layer's gather; a rare flipped bf16 rounding (4e-3 of one element) is what remains.
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 5e-3), (torch.bfloat16, 3e-2)])
def test_fuse_inference_is_the_same_function(dtype, tol):
    """BatchNorm(eval) / ReLU / residual add folded into the conv epilogues (bayesian_torch_b200/fuse.py) ==
    the unfused torchvision forward, for the same weight samples."""
    bnn, _ = _resnet18()
    # non-trivial BN statistics / affine so that the folding is really exercised
    g = torch.Generator(device="cpu").manual_seed(3)
    for m in bnn.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.2)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.1)
    bnn = bnn.to(dtype).to(memory_format=torch.channels_last)
    x = torch.randn(8, 3, 32, 32, device=DEV, dtype=dtype)
    with torch.no_grad():
        btb.manual_seed(5)
        with btb.mc_sample_context(3, 8, 0):
            ref = bnn(x).float()
        n_before = sum(1 for _ in bnn.modules())
        btb.fuse_inference(bnn)
        assert sum(1 for m in bnn.modules() if type(m).__name__ == "FusedBasicBlock") == 8
        assert type(bnn.maxpool).__name__ == "FusedMaxPool2d"
        btb.manual_seed(5)
        with btb.mc_sample_context(3, 8, 0):
            out = bnn(x).float()
    assert out.shape == ref.shape == (24, 10)
    rel, mx = errs(out, ref)
    assert rel <= tol, (rel, mx)
    assert list(k for k in bnn.state_dict().keys() if "mu_" in k)   # parameters still reachable
    mean, var = btb.mc_predict(bnn, x, 4)
    assert mean.shape == (8, 10) and torch.allclose(mean.sum(-1), torch.ones(8, device=DEV), atol=1e-4)

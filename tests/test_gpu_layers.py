"""-m gpu: the fused layer forward (csrc/bt_fused.cu) through the layer classes / C ABI.

(1) golden parity: reference-generated (mu, rho, eps, signs, x) are injected (debug import hooks) and the
    kernel output is compared with the reference's own output y  -> "identical inputs" parity.
(2) on-chip RNG parity: the kernel draws eps / signs itself (Philox); the draws are re-materialised with
    bt_rng_export and fed to the oracle.
Tolerances (the contract of SURVEY.md 8c):
  * fp32 parameters + fp32 activations (the reference's default dtype) run on tcgen05 kind::tf32: rel-RMS <= 5e-4
    against the fp32 reference output, <= 1e-4 against the oracle evaluated on tf32-rounded operands;
  * any bf16 operand -> kind::f16 with bf16 operands: rel-RMS <= 3e-3 against the fp32 oracle, <= 1e-3 against the
    oracle on bf16-rounded operands when the output is fp32 (a bf16 OUTPUT adds its own 2^-9 rounding: <= 3e-3);
(remaining difference: accumulation order and the MUFU-based softplus / Box-Muller, ~1e-6 relative on W,
which can move W across a rounding boundary)."""
import json
import os

import pytest
import torch

import bayesian_torch_b200 as btb
from bayesian_torch_b200 import _native
from gpu_util import build_layer, cl, errs, layer_params, oracle_forward, phys_eps
from oracle import bt_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "tests", "golden", "meta.json")) as _f:
    _META = json.load(_f)
TOL_TF32_VS_FP32 = 5e-4      # fp32 x + fp32 parameters vs the fp32 reference
TOL_TF32_ROUNDED = 1e-4      # ... vs the oracle on tf32-rounded operands
TOL_BF16_VS_FP32 = 3e-3      # bf16 operands vs the fp32 oracle
TOL_BF16_ROUNDED = 1e-3      # ... vs the oracle on bf16-rounded operands (fp32 output)


def _golden_layer(c, m):
    flip, bias = m["flipout"], m["bias"]
    mu = c["mu_w"]
    if m["kind"] == "linear":
        layer = build_layer("linear", 0, flip, mu.shape[1], mu.shape[0], None, bias=bias,
                            prior_mean=m["prior_mean"], prior_variance=m["prior_variance"])
        w = "weight"
    else:
        ks = tuple(mu.shape[2:])
        layer = build_layer("conv", m["nd"], flip, mu.shape[1] * m["groups"], mu.shape[0], ks, m["stride"], m["padding"],
                            m["dilation"], m["groups"], bias, m["prior_mean"], m["prior_variance"])
        w = "kernel"
    sd = {f"mu_{w}": mu, f"rho_{w}": c["rho_w"]}
    if bias:
        sd.update(mu_bias=c["mu_b"], rho_bias=c["rho_b"])
    layer.load_state_dict(sd)
    return layer.to(DEV)


@pytest.mark.parametrize("name", sorted(_META["cases"].keys()))
def test_golden_parity_identical_inputs(golden, name):
    c, m = golden.case(name), golden.meta["cases"][name]
    layer = _golden_layer(c, m)
    x = c["x"].to(DEV)
    dbg = {"eps_w_in": phys_eps(c["eps_w"]).to(DEV)}
    if m["bias"]:
        dbg["eps_b_in"] = c["eps_b"].to(DEV)
    if m["flipout"]:
        if m["kind"] == "linear":
            dbg["sign_in"], dbg["sign_out"] = c["sign_in"].to(DEV).contiguous(), c["sign_out"].to(DEV).contiguous()
        else:
            dbg["sign_in"], dbg["sign_out"] = cl(c["sign_in"]).to(DEV), cl(c["sign_out"]).to(DEV)
    with torch.no_grad():          # inference call: the KL comes from the fused kernel's side output
        y, kl = layer._forward_impl(x, True, debug=dbg)
    torch.cuda.synchronize()
    assert y.shape == c["y"].shape and y.dtype == torch.float32 and y.grad_fn is None
    rel, mx = errs(y, c["y"])
    assert rel <= TOL_TF32_VS_FP32, f"{name}: rel-RMS {rel:.2e} max-abs {mx:.2e} vs the reference output"
    # tight: oracle on tf32-rounded operands (the goldens are fp32 x + fp32 parameters)
    yr = oracle_forward(layer, c["x"], c["eps_w"], c.get("eps_b"), c.get("sign_in"), c.get("sign_out"),
                        round_operands=True)
    rel2, mx2 = errs(y, yr)
    assert rel2 <= TOL_TF32_ROUNDED, f"{name}: rel-RMS {rel2:.2e} max-abs {mx2:.2e} vs operand-rounded oracle"
    # KL side output of forward(return_kl=True) and kl_loss() vs the reference's values
    assert abs(float(kl) - float(c["kl"])) <= 1e-5 * max(1.0, abs(float(c["kl"]))), (float(kl), float(c["kl"]))
    assert abs(float(layer.kl_loss()) - float(c["kl_loss"])) <= 1e-5 * max(1.0, abs(float(c["kl_loss"])))


SHAPES = [
    # kind, nd, flip, cin, cout, ks, stride, pad, dil, groups, bias, batch, spatial, xdtype, pdtype
    ("linear", 0, False, 1024, 1024, None, 1, 0, 1, 1, True, 256, (), torch.float32, torch.float32),   # C1
    ("linear", 0, True, 512, 300, None, 1, 0, 1, 1, True, 130, (), torch.bfloat16, torch.bfloat16),
    ("linear", 0, False, 100, 10, None, 1, 0, 1, 1, True, 7, (), torch.float32, torch.float32),      # K, N tails, scalar path
    ("linear", 0, True, 72, 96, None, 1, 0, 1, 1, False, 600, (), torch.float32, torch.bfloat16),     # MT > 1, mixed dtypes
    ("linear", 0, False, 512, 10, None, 1, 0, 1, 1, True, 128, (), torch.bfloat16, torch.float32),    # ResNet-18 fc
    ("conv", 2, False, 64, 64, 3, 1, 1, 1, 1, False, 4, (8, 8), torch.bfloat16, torch.float32),        # ResNet layer1
    ("conv", 2, False, 3, 64, 7, 2, 3, 1, 1, False, 3, (32, 32), torch.float32, torch.float32),        # stem, scalar gather
    ("conv", 2, True, 64, 128, 3, 1, 1, 1, 1, True, 2, (14, 14), torch.bfloat16, torch.bfloat16),      # C2-like, reduced
    ("conv", 2, False, 128, 256, 3, 2, 1, 1, 1, False, 5, (4, 4), torch.bfloat16, torch.bfloat16),
    ("conv", 2, False, 64, 128, 1, 2, 0, 1, 1, False, 6, (8, 8), torch.float32, torch.float32),        # 1x1 downsample
    ("conv", 2, True, 32, 48, (3, 2), (2, 1), (1, 0), (1, 2), 4, True, 3, (9, 11), torch.float32, torch.float32),
    ("conv", 1, False, 16, 24, 5, 2, 2, 1, 1, True, 4, (33,), torch.float32, torch.float32),
    ("conv", 1, True, 8, 8, 3, 1, 1, 2, 8, False, 2, (20,), torch.bfloat16, torch.float32),            # depthwise, scalar paths
    ("conv", 3, False, 8, 16, (2, 3, 3), 1, 1, 1, 1, True, 2, (4, 6, 6), torch.float32, torch.float32),
    ("conv", 3, True, 16, 8, 3, (1, 2, 2), 1, 1, 2, True, 2, (3, 7, 7), torch.bfloat16, torch.bfloat16),
]


@pytest.mark.parametrize("cfg", SHAPES, ids=lambda c: f"{c[0]}{c[1]}{'F' if c[2] else 'R'}_{c[3]}x{c[4]}_g{c[9]}")
def test_onchip_rng_parity(cfg):
    kind, nd, flip, cin, cout, ks, st, pd, dl, groups, bias, batch, sp, xdt, pdt = cfg
    torch.manual_seed(hash((cin, cout, batch)) % 1000)
    btb.manual_seed(4242)
    layer = build_layer(kind, nd, flip, cin, cout, ks, st, pd, dl, groups, bias, 0.0, 1.0).to(DEV).to(pdt)
    x = (torch.randn(batch, cin, *sp)).to(xdt).to(DEV)
    with torch.no_grad():
        y, kl = layer(x)
        y2, _ = layer(x)     # a second call draws a NEW sample
    torch.cuda.synchronize()
    assert not torch.equal(y, y2)
    assert y.dtype == xdt
    # re-materialise the draws of the FIRST call (sample index 0)
    layer._bt_last["sample0"] = 0
    eps_w, eps_b = layer.materialize_eps(0)
    s_in = s_out = None
    if flip:
        s_in, s_out = layer.materialize_signs(tuple(x.shape), tuple(y.shape), 0)
    yr = oracle_forward(layer, x, eps_w, eps_b, s_in, s_out, round_operands=True)
    rel, mx = errs(y, yr)
    tf32 = xdt == torch.float32 and pdt == torch.float32
    tol = TOL_TF32_ROUNDED if tf32 else (TOL_BF16_ROUNDED if xdt == torch.float32 else TOL_BF16_VS_FP32)
    assert rel <= tol, f"rel-RMS {rel:.2e} max-abs {mx:.2e}"
    yf = oracle_forward(layer, x, eps_w, eps_b, s_in, s_out, round_operands=False)
    relf = errs(y, yf)[0]
    assert relf <= (TOL_TF32_VS_FP32 if tf32 else TOL_BF16_VS_FP32), f"rel-RMS {relf:.2e} vs the fp32 oracle"
    mu_w, rho_w, mu_b, rho_b = layer_params(layer)
    kref = O.kl_loss(mu_w.double(), rho_w.double(), 0.0, 1.0, None if mu_b is None else mu_b.double(),
                     None if rho_b is None else rho_b.double())
    ktol = 2e-5 if pdt == torch.float32 else 8e-3      # kl is returned in the parameter dtype (bf16: 8 bits)
    assert kl.dtype == pdt
    assert abs(float(kl) - float(kref)) <= ktol * max(1.0, abs(float(kref))), (float(kl), float(kref))
    assert abs(float(layer.kl_loss()) - float(kref)) <= ktol * max(1.0, abs(float(kref)))
    # statistics of the on-chip normals
    if eps_w.numel() >= 1 << 16:
        assert abs(float(eps_w.float().mean())) < 2e-2 and abs(float(eps_w.float().std()) - 1) < 2e-2


@pytest.mark.parametrize("spatial,stride", [((1, 1), 1), ((2, 2), 1), ((2, 2), 2), ((3, 3), 1)])
@pytest.mark.parametrize("flip", [False, True])
def test_padding_only_taps_are_skipped_exactly(spatial, stride, flip):
    """ResNet layer3/4 at CIFAR resolution: most 3x3 taps only ever see zero padding; the kernel neither reads
    nor samples their weights.  The result must equal the full convolution."""
    torch.manual_seed(3)
    btb.manual_seed(7)
    layer = build_layer("conv", 2, flip, 128, 128, 3, stride, 1, 1, 1, False).to(DEV)
    x = torch.randn(9, 128, *spatial, device=DEV)
    y = layer(x, return_kl=False)
    layer._bt_last["sample0"] = 0
    eps_w, eps_b = layer.materialize_eps(0)
    s_in = s_out = None
    if flip:
        s_in, s_out = layer.materialize_signs(tuple(x.shape), tuple(y.shape), 0)
    yr = oracle_forward(layer, x, eps_w, eps_b, s_in, s_out, round_operands=True)
    rel, mx = errs(y, yr)
    assert rel <= TOL_TF32_ROUNDED, (rel, mx)       # fp32 x + fp32 parameters: the tf32 path


@pytest.mark.parametrize("flip", [False, True])
def test_mc_sample_dimension_equals_sequential_samples(flip):
    """S samples in one launch (sample index = grid dimension) == S single-sample launches with the same
    global sample indices; first layer shares x, deeper layers see stacked activations."""
    torch.manual_seed(5)
    btb.manual_seed(99)
    conv = build_layer("conv", 2, flip, 16, 64, 3, 1, 1, 1, 1, True).to(DEV)
    fc = build_layer("linear", 0, flip, 64, 10, None, bias=True).to(DEV)
    B, S = 6, 5
    x = torch.randn(B, 16, 5, 5, device=DEV)
    with btb.mc_sample_context(S, B, 10):
        h = conv(x, return_kl=False)                     # x shared by the S samples
        assert h.shape == (S * B, 64, 5, 5)
        o = fc(h.mean((2, 3)), return_kl=False)          # stacked input
        assert o.shape == (S * B, 10)
    for s in range(S):
        with btb.mc_sample_context(1, B, 10 + s):
            hs = conv(x, return_kl=False)
            os_ = fc(hs.mean((2, 3)), return_kl=False)
        assert torch.equal(hs, h[s * B:(s + 1) * B]), s
        assert torch.equal(os_, o[s * B:(s + 1) * B]), s
    assert not torch.equal(h[:B], h[B:2 * B])


@pytest.mark.parametrize("pdt,tol", [(torch.float32, 1e-3), (torch.bfloat16, 3e-3)])
@pytest.mark.parametrize("flip", [False, True])
def test_sigma_cache_in_mc_context_is_the_same_function(flip, pdt, tol):
    """Inside mc_sample_context the layers hand the kernels a cached sigma = softplus(rho) (geom.rho_is_sigma) instead
    of rho.  Same draws, same function: the only differences are torch's exact softplus vs the sampler's fast one
    (6e-5 relative, which moves a few sigma*eps / W values across a bf16 rounding boundary: measured 2e-4 rel-RMS, stated
    tolerance 1e-3 = the operand-rounding tolerance of DESIGN.md section 2) and, for bf16 parameters, sigma rounded to
    bf16 (2^-9 relative on sigma*eps, below the bf16 rounding of W that follows).  The cache follows Tensor._version."""
    import os
    torch.manual_seed(13)
    conv = build_layer("conv", 2, flip, 64, 96, 3, 1, 1, 1, 1, True).to(DEV).to(pdt)
    fc = build_layer("linear", 0, flip, 96, 40, None, bias=True).to(DEV).to(pdt)
    B, S = 9, 4
    x = torch.randn(B, 64, 6, 6, device=DEV)
    outs = {}
    for mode in ("cached", "plain"):
        if mode == "plain":
            os.environ["BT_DISABLE_SIGMA_CACHE"] = "1"
        try:
            btb.manual_seed(21)
            with btb.mc_sample_context(S, B, 3):
                h = conv(x, return_kl=False)
                outs[mode] = (h, fc(h.mean((2, 3)), return_kl=False))
        finally:
            os.environ.pop("BT_DISABLE_SIGMA_CACHE", None)
    assert conv._bt_sigma_cache is not None
    for a, b in zip(outs["cached"], outs["plain"]):
        rel, mx = errs(a, b)
        assert rel <= tol, (rel, mx)
    # an in-place parameter update (optimizer step) invalidates the cache
    with torch.no_grad():
        conv.rho_kernel.add_(1.0)
    btb.manual_seed(21)
    with btb.mc_sample_context(S, B, 3):
        h2 = conv(x, return_kl=False)
    assert errs(h2, outs["cached"][0])[0] > 1e-2


def test_sample_mean_converges_to_mean_weight_output():
    """E_eps[out] = conv(x, mu) + mu_b: average over many on-chip samples approaches it at 1/sqrt(S)."""
    torch.manual_seed(1)
    btb.manual_seed(1)
    for flip in (False, True):
        layer = build_layer("linear", 0, flip, 256, 128, None, bias=True).to(DEV)
        with torch.no_grad():
            layer.rho_weight.fill_(-1.0)
        x = torch.randn(32, 256, device=DEV)
        S = 256
        with btb.mc_sample_context(S, 32, 0):
            o = layer(x, return_kl=False).view(S, 32, 128)
        mean_w = torch.nn.functional.linear(x, layer.mu_weight, layer.mu_bias)
        sd = o.std(0).mean()
        err = (o.mean(0) - mean_w).abs().mean()
        assert float(sd) > 1.0                              # the samples really differ
        assert float(err) < 4.0 * float(sd) / S ** 0.5, (float(err), float(sd))


def test_layer_errors_and_contracts():
    layer = build_layer("conv", 2, False, 8, 8, 3).to(DEV)
    with pytest.raises(RuntimeError, match="channels"):
        layer(torch.randn(1, 4, 8, 8, device=DEV))
    with pytest.raises(ValueError, match="float32, bfloat16"):
        layer(torch.randn(1, 8, 8, 8, device=DEV, dtype=torch.float16))
    with pytest.raises(RuntimeError, match="Kernel size"):
        layer(torch.randn(1, 8, 2, 2, device=DEV))
    layer.dnn_to_bnn_flag = True
    out = layer(torch.randn(2, 8, 6, 6, device=DEV))
    assert torch.is_tensor(out) and out.shape == (2, 8, 4, 4)       # flag forces return_kl=False
    layer.dnn_to_bnn_flag = False
    out, kl = layer(torch.randn(2, 8, 6, 6, device=DEV))
    assert kl.dim() == 0 and kl.dtype == torch.float32
    # priors edited after init (MOPED-style) are honoured through the tensor-prior path
    layer.prior_weight_mu.normal_()
    mu_w, rho_w, mu_b, rho_b = layer_params(layer)
    ref = O.kl_div(mu_w, O.sigma_of_rho(rho_w), layer.prior_weight_mu.cpu(), layer.prior_weight_sigma.cpu()) + \
        O.kl_div(mu_b, O.sigma_of_rho(rho_b), 0.0, 1.0)
    assert abs(float(layer.kl_loss()) - float(ref)) < 1e-4 * abs(float(ref))
    _, kl2 = layer(torch.randn(2, 8, 6, 6, device=DEV))
    assert abs(float(kl2) - float(ref)) < 1e-4 * abs(float(ref))


@pytest.mark.parametrize("xdt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("flip", [False, True])
def test_fused_epilogue_scale_shift_residual_relu(flip, xdt):
    """BtEpilogue: y = relu((conv + bias) * scale + shift + residual), applied on the fp32 accumulator."""
    torch.manual_seed(2)
    btb.manual_seed(11)
    layer = build_layer("conv", 2, flip, 32, 48, 3, 1, 1, 1, 1, True).to(DEV)
    x = torch.randn(5, 32, 7, 7, device=DEV).to(xdt)
    plain = layer(x, return_kl=False)
    scale = torch.rand(48, device=DEV) + 0.5
    shift = torch.randn(48, device=DEV)
    res = torch.randn(5, 48, 7, 7, device=DEV).to(xdt).contiguous(memory_format=torch.channels_last)
    layer._bt_ep_scale, layer._bt_ep_shift, layer._bt_ep_relu = scale, shift, True
    btb.manual_seed(11)                      # same draw as `plain`
    with torch.no_grad():            # the fused epilogue is an inference feature (not differentiable)
        fused = layer._forward_impl(x, False, residual=res)
    ref = torch.relu(plain.float() * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1) + res.float())
    rel, mx = errs(fused, ref)
    assert rel <= (1e-5 if xdt == torch.float32 else 8e-3), (rel, mx)
    assert float(fused.min()) >= 0.0

"""-m gpu: bt_direct_kernel (csrc/bt_direct.cuh) -- the A operand of tcgen05.mma is read in place from a
shared-memory input window (descriptor row shift per filter tap, padded pixel numbering).

Every case is run three ways on identical (mu, rho, seed, x): forced direct mode, direct mode disabled (the
im2col kernels) and the CPU oracle fed with the re-materialised on-chip draws.  Direct vs im2col must agree to
bf16 output rounding (they form identical bf16 operands and accumulate the same k order in fp32 -- in practice
bit-exact); both must meet the stated oracle tolerance (6e-3 rel-RMS for a bf16 output)."""
import os
from contextlib import contextmanager

import pytest
import torch

import bayesian_torch_b200 as btb
from bayesian_torch_b200 import _native
from gpu_util import build_layer, errs, oracle_forward

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@contextmanager
def env(**kw):
    old = {k: os.environ.get(k) for k in kw}
    for k, v in kw.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    try:
        yield
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


CASES = [
    # kind, nd, cin, cout, ks, pad, dil, bias, batch, spatial, pdtype
    ("conv", 2, 64, 64, 3, 1, 1, False, 4, (8, 8), torch.float32),          # ResNet-18 layer1 @ CIFAR
    ("conv", 2, 64, 128, 3, 1, 1, True, 2, (14, 14), torch.bfloat16),        # C2-like, reduced
    ("conv", 2, 128, 64, 3, 2, 2, True, 3, (9, 7), torch.bfloat16),          # two slabs, dilation 2
    ("conv", 2, 128, 128, 3, 1, 1, False, 9, (2, 2), torch.float32),         # tiny images: many pad rows
    ("conv", 2, 128, 96, 3, 1, 1, False, 5, (1, 1), torch.bfloat16),         # padding-only taps skipped, N tail
    ("conv", 2, 128, 64, 1, 0, 1, True, 6, (5, 5), torch.bfloat16),          # 1x1: plain GEMM, no pad pixels
    ("conv", 1, 64, 96, 5, 2, 1, True, 4, (37,), torch.float32),             # conv1d, k=5
    ("conv", 3, 64, 32, 3, 1, 1, True, 2, (3, 5, 5), torch.bfloat16),        # conv3d
    ("linear", 0, 256, 100, None, 0, 1, True, 300, (), torch.bfloat16),      # linear = 1x1 conv on 1 pixel
]


def _run(layer, x, seed, residual=None):
    btb.manual_seed(seed)
    y = layer._forward_impl(x, False, residual=residual)
    torch.cuda.synchronize()
    return y, _native.last_forward_path()


@pytest.mark.parametrize("flip", [False, True], ids=["R", "F"])
@pytest.mark.parametrize("cfg", CASES, ids=lambda c: f"{c[0]}{c[1]}_{c[2]}x{c[3]}_k{c[4]}_{'x'.join(map(str, c[9]))}")
def test_direct_equals_im2col_and_oracle(cfg, flip):
    kind, nd, cin, cout, ks, pad, dil, bias, batch, sp, pdt = cfg
    torch.manual_seed(cin + cout + batch)
    layer = build_layer(kind, nd, flip, cin, cout, ks, 1, pad, dil, 1, bias).to(DEV).to(pdt)
    x = torch.randn(batch, cin, *sp).bfloat16().to(DEV)
    with env(BT_FORCE_DIRECT="1", BT_DISABLE_DIRECT=None):
        yd, path_d = _run(layer, x, 77)
    if path_d != "direct":
        assert flip and cin * (1 if nd < 3 else 2) >= 128 or kind == "linear" or nd == 3, (path_d, cfg)
        pytest.skip("resident sampled tiles + two input windows do not fit shared memory for this Flipout shape")
    with env(BT_FORCE_DIRECT=None, BT_DISABLE_DIRECT="1"):
        yi, path_i = _run(layer, x, 77)
    assert path_i != "direct"
    rel_di, mx_di = errs(yd, yi)
    layer._bt_last["sample0"] = 0
    eps_w, eps_b = layer.materialize_eps(0)
    s_in = s_out = None
    if flip:
        s_in, s_out = layer.materialize_signs(tuple(x.shape), tuple(yd.shape), 0)
    yr = oracle_forward(layer, x, eps_w, eps_b, s_in, s_out, round_operands=True)
    rel_do, mx_do = errs(yd, yr)
    rel_io, _ = errs(yi, yr)
    msg = f"direct-vs-im2col rel {rel_di:.2e} max {mx_di:.2e}; direct-vs-oracle rel {rel_do:.2e} max {mx_do:.2e}; " \
          f"im2col-vs-oracle rel {rel_io:.2e}"
    assert rel_do <= 6e-3, msg
    assert rel_di <= 4e-3, msg          # (both round the same fp32 accumulator to bf16; expected 0)


@pytest.mark.parametrize("flip", [False, True], ids=["R", "F"])
def test_direct_mc_samples_epilogue_and_tile_boundaries(flip):
    """S samples in one launch (shared x, then stacked x), fused BN-affine / residual / ReLU epilogue, and a batch
    large enough that one sample spans many 128-row tiles and CTAs take several tiles each."""
    torch.manual_seed(9)
    conv1 = build_layer("conv", 2, flip, 64, 64, 3, 1, 1, 1, 1, True).to(DEV).bfloat16()
    conv2 = build_layer("conv", 2, flip, 64, 128, 3, 1, 1, 1, 1, False).to(DEV).bfloat16()
    B, S = 37, 3
    x = torch.randn(B, 64, 8, 8).bfloat16().to(DEV)
    scale = (torch.rand(128, device=DEV) + 0.5)
    shift = torch.randn(128, device=DEV)
    conv2._bt_ep_scale, conv2._bt_ep_shift, conv2._bt_ep_relu = scale, shift, True
    outs = {}
    for mode, e in (("direct", dict(BT_FORCE_DIRECT="1", BT_DISABLE_DIRECT=None)),
                    ("im2col", dict(BT_FORCE_DIRECT=None, BT_DISABLE_DIRECT="1"))):
        with env(**e):
            btb.manual_seed(5)
            with btb.mc_sample_context(S, B, 100):
                h = conv1(x, return_kl=False)
                p1 = _native.last_forward_path()
                res = torch.randn(S * B, 128, 8, 8, generator=torch.Generator().manual_seed(1)).bfloat16().to(DEV) \
                    .contiguous(memory_format=torch.channels_last)
                o = conv2._forward_impl(h, False, residual=res)
                p2 = _native.last_forward_path()
            torch.cuda.synchronize()
            assert (p1 == "direct") == (mode == "direct") and (p2 == "direct") == (mode == "direct"), (mode, p1, p2)
            outs[mode] = (h, o)
    hd, od = outs["direct"]
    hi, oi = outs["im2col"]
    assert hd.shape == (S * B, 64, 8, 8) and od.shape == (S * B, 128, 8, 8)
    r1, m1 = errs(hd, hi)
    assert r1 <= 4e-3, (r1, m1)
    r2, m2 = errs(od, oi)
    assert r2 <= 8e-3, (r2, m2)          # second layer sees the (possibly 1-ulp different) bf16 h
    assert float(od.min()) >= 0.0
    assert not torch.equal(hd[:B], hd[B:2 * B])
    # single-sample launches with the same global sample indices reproduce the stacked launch bit-exactly
    with env(BT_FORCE_DIRECT="1", BT_DISABLE_DIRECT=None):
        for s in range(S):
            btb.manual_seed(5)
            with btb.mc_sample_context(1, B, 100 + s):
                hs = conv1(x, return_kl=False)
            assert torch.equal(hs, hd[s * B:(s + 1) * B]), s


@pytest.mark.parametrize("cin,cout,sp", [(128, 128, (4, 4)), (64, 64, (8, 8)), (192, 64, (1, 1))])
def test_direct_residual_epilogue_staged_and_unstaged(cin, cout, sp):
    """BatchNorm-affine + residual + ReLU epilogue on both epilogue variants: with the shared-memory staging buffer
    (64 -> 64) and without it (128 -> 128 at ResNet layer2 size: the resident tiles leave no room for it)."""
    torch.manual_seed(21)
    layer = build_layer("conv", 2, False, cin, cout, 3 if sp != (1, 1) else 1, 1, 1 if sp != (1, 1) else 0, 1, 1, True).to(DEV).bfloat16()
    B, S = 19, 2
    x = torch.randn(S * B, cin, *sp).bfloat16().to(DEV)
    res = torch.randn(S * B, cout, *sp).bfloat16().to(DEV).contiguous(memory_format=torch.channels_last)
    layer._bt_ep_scale = torch.rand(cout, device=DEV) + 0.5
    layer._bt_ep_shift = torch.randn(cout, device=DEV)
    layer._bt_ep_relu = True
    outs = {}
    for mode, e in (("direct", dict(BT_FORCE_DIRECT="1", BT_DISABLE_DIRECT=None)),
                    ("im2col", dict(BT_FORCE_DIRECT=None, BT_DISABLE_DIRECT="1"))):
        with env(**e):
            btb.manual_seed(3)
            with btb.mc_sample_context(S, B, 7):
                outs[mode] = layer._forward_impl(x, False, residual=res)
            torch.cuda.synchronize()
            assert (_native.last_forward_path() == "direct") == (mode == "direct")
    rel, mx = errs(outs["direct"], outs["im2col"])
    assert rel <= 4e-3, (rel, mx)
    assert float(outs["direct"].min()) >= 0.0


def test_direct_is_the_default_for_resnet_layer1_shapes():
    """Without any switch the tiling search must pick an in-place-window kernel for the CIFAR ResNet-18 layer1 shape:
    the TMA-window one (bt_dtma_kernel), or this file's cp.async one when the TMA families are switched off."""
    torch.manual_seed(0)
    layer = build_layer("conv", 2, False, 64, 64, 3, 1, 1, 1, 1, False).to(DEV).bfloat16()
    x = torch.randn(128, 64, 8, 8).bfloat16().to(DEV)
    with env(BT_FORCE_DIRECT=None, BT_DISABLE_DIRECT=None):
        with btb.mc_sample_context(16, 128, 0):
            layer(x, return_kl=False)
        torch.cuda.synchronize()
        assert _native.last_forward_path() == "tma_direct"
    with env(BT_FORCE_DIRECT=None, BT_DISABLE_DIRECT=None, BT_DISABLE_TMA="1"):
        with btb.mc_sample_context(16, 128, 0):
            layer(x, return_kl=False)
        torch.cuda.synchronize()
        assert _native.last_forward_path() == "direct"

"""-m gpu: the stand-alone kernels through the C ABI vs the oracle: KL (K1), Philox export vs the numpy
restatement, MC accumulate / finalize."""
import numpy as np
import pytest
import torch

from bayesian_torch_b200 import _native
from oracle import bt_oracle as O
from oracle import philox_ref as P

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_device_is_b200():
    lib = _native.load()
    assert lib.bt_device_check() == 0, lib.bt_last_error()
    assert lib.bt_sm_count() == 148


@pytest.mark.parametrize("n_w,n_b,dtype", [(5, 0, torch.float32), (1024 * 1024, 1024, torch.float32),
                                           (4097, 3, torch.float32), (1 << 22, 4096, torch.bfloat16),
                                           (777, 5, torch.bfloat16)])
def test_kl_kernel_matches_oracle(n_w, n_b, dtype):
    torch.manual_seed(n_w)
    mu = (torch.randn(n_w) * 0.1).to(dtype)
    rho = (torch.randn(n_w) * 0.5 - 3).to(dtype)
    mb = (torch.randn(n_b) * 0.1).to(dtype) if n_b else None
    rb = (torch.randn(n_b) * 0.5 - 3).to(dtype) if n_b else None
    ref = O.kl_loss(mu.double(), rho.double(), 0.1, 0.7, None if mb is None else mb.double(),
                    None if rb is None else rb.double())
    out = _native.kl_gaussian(mu.to(DEV), rho.to(DEV), None, None, None if mb is None else mb.to(DEV),
                              None if rb is None else rb.to(DEV), None, None, 0.1, 0.7)
    assert abs(float(out) - float(ref)) <= 1e-5 * max(1.0, abs(float(ref))), (float(out), float(ref))
    # deterministic: bitwise equal on re-run; workspace self-resets
    out2 = _native.kl_gaussian(mu.to(DEV), rho.to(DEV), None, None, None if mb is None else mb.to(DEV),
                               None if rb is None else rb.to(DEV), None, None, 0.1, 0.7)
    assert float(out) == float(out2)


def test_kl_kernel_tensor_priors_unaligned_and_extreme_rho():
    torch.manual_seed(1)
    n = 10007
    base = torch.randn(4, n + 3)
    mu, rho, pm, ps = (base[i, 1 + i % 2: 1 + i % 2 + n] for i in range(4))    # misaligned views
    mu, rho, pm, ps = mu * 0.1, rho * 8.0, pm * 0.05, ps.abs() * 0.3 + 0.5        # rho in [-30, 30]
    ref = O.kl_div(mu.double(), torch.nn.functional.softplus(rho.double()), pm.double(), ps.double())
    g = [t.contiguous().to(DEV) for t in (mu, rho, pm, ps)]
    buf = torch.zeros(4, n + 8, device=DEV)
    views = []
    for i, t in enumerate(g):
        buf[i, 1:1 + n] = t
        views.append(buf[i, 1:1 + n])
    out = _native.kl_gaussian(views[0], views[1], views[2], views[3])
    assert abs(float(out) - float(ref)) <= 2e-5 * max(1.0, abs(float(ref))), (float(out), float(ref))
    acc = _native.kl_gaussian(views[0], views[1], views[2], views[3], out=out.clone(), accumulate=True)
    assert float(acc) == pytest.approx(2 * float(out), rel=1e-6)


def test_rng_export_matches_numpy_restatement():
    seed, key = 0x1234567887654321, 77
    # linear weight eps
    e = torch.empty(40, 100, device=DEV)
    _native.rng_export(0, e, 40, 100, 1, 100, seed, key, 3)
    ref = P.weight_eps(40, 100, seed, key, 3)
    d = np.abs(e.cpu().numpy() - ref)
    assert d.max() < 2e-3 and d.mean() < 2e-6, (d.max(), d.mean())
    # conv weight eps: written in the reference's logical (Cout, Cin, taps) order
    cout, cin, taps = 6, 8, 9
    e = torch.empty(cout, cin * taps, device=DEV)
    _native.rng_export(0, e, cout, cin * taps, taps, cin * taps, seed, key, 0)
    phys = P.weight_eps(cout, taps * cin, seed, key, 0).reshape(cout, taps, cin)
    assert np.abs(e.cpu().numpy().reshape(cout, cin, taps) - phys.transpose(0, 2, 1)).max() < 2e-3
    # bias eps
    b = torch.empty(13, device=DEV)
    _native.rng_export(1, b, 13, 1, 1, 1, seed, key, 5)
    assert np.abs(b.cpu().numpy() - P.bias_eps(13, seed, key, 5)).max() < 2e-3
    # signs
    s = torch.empty(50, 300, device=DEV)
    _native.rng_export(2, s, 50, 300, 1, 300, seed, key, 9)
    assert np.array_equal(s.cpu().numpy(), P.sign_bits(50, 300, seed, key, 9, P.STREAM_SIGN_IN))
    s = torch.empty(50, 300, device=DEV)
    _native.rng_export(3, s, 50, 300, 1, 300, seed, key, 9)
    assert np.array_equal(s.cpu().numpy(), P.sign_bits(50, 300, seed, key, 9, P.STREAM_SIGN_OUT))
    big = torch.empty(2048, 1024, device=DEV)
    _native.rng_export(0, big, 2048, 1024, 1, 1024, 5, 1, 0)
    assert abs(float(big.mean())) < 3e-3 and abs(float(big.std()) - 1) < 3e-3


@pytest.mark.parametrize("S,B,C,dtype", [(6, 4, 10, torch.float32), (3, 130, 1000, torch.float32),
                                         (5, 7, 100, torch.bfloat16)])
def test_mc_accumulate_and_finalize(S, B, C, dtype):
    torch.manual_seed(0)
    logits = (torch.randn(S, B, C) * 3).to(dtype)
    mean_ref, var_ref = O.mc_aggregate(logits.float())
    g = logits.to(DEV).reshape(S * B, C)
    sums = torch.empty(2, B, C, device=DEV)
    half = S // 2
    _native.mc_accumulate(g[: half * B], half, B, sums, accumulate=False)
    _native.mc_accumulate(g[half * B:], S - half, B, sums, accumulate=True)
    mean, var = torch.empty(B, C, device=DEV), torch.empty(B, C, device=DEV)
    _native.mc_finalize(sums, S, mean, var)
    assert torch.allclose(mean.cpu(), mean_ref, atol=2e-6)
    assert torch.allclose(var.cpu(), var_ref, atol=2e-6)
    assert torch.allclose(mean.sum(-1).cpu(), torch.ones(B), atol=1e-5)


def test_abi_errors_are_loud():
    with pytest.raises(ValueError):
        _native.kl_gaussian(torch.zeros(4, device=DEV, dtype=torch.float16), torch.zeros(4, device=DEV, dtype=torch.float16))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _native.kl_gaussian(torch.zeros(4), torch.zeros(4))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("k,s,p,hw", [(3, 2, 1, (16, 16)), (2, 2, 0, (7, 9)), (3, 1, 1, (5, 5))])
def test_maxpool_nhwc_matches_torch(dtype, k, s, p, hw):
    torch.manual_seed(0)
    x = torch.randn(5, 64, *hw, device=DEV).to(dtype)
    ref = torch.nn.functional.max_pool2d(x.float(), k, s, p).to(dtype)
    out = _native.maxpool2d_nhwc(x.permute(0, 2, 3, 1).contiguous(), (k, k), (s, s), (p, p)).permute(0, 3, 1, 2)
    assert out.shape == ref.shape and torch.equal(out, ref)


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32], ids=["bf16", "f32"])
@pytest.mark.parametrize("cfg", [
    # C, H, W, kh, kw, stride, pad, dil, N, memory format
    (3, 32, 32, 7, 7, 2, 3, 1, 5, "nchw"),            # the CIFAR-shaped ResNet stem
    (3, 17, 23, 3, 5, 1, 2, 1, 3, "nhwc"),            # ragged sizes, channels-last input strides
    (4, 9, 9, 3, 3, 2, 1, 2, 2, "nchw"),              # dilation
    (1, 6, 40, 1, 7, 1, 3, 1, 4, "nchw"),
])
def test_im2col2d_rows_equal_unfold(cfg, dt):
    """bt_im2col2d (csrc/bt_im2col.cu, table form): rows [N*OH*OW, kpad] with column (kh, kw, c), zero-extended -- a copy,
    so BIT-EXACT against F.unfold (whose column order is (c, kh, kw)) of the same input."""
    import torch.nn.functional as F
    from bayesian_torch_b200 import _native
    C, H, W, kh, kw, st, pd, dl, N, fmt = cfg
    torch.manual_seed(2)
    x = torch.randn(N, C, H, W, device="cuda:0").to(dt)
    if fmt == "nhwc":
        x = x.contiguous(memory_format=torch.channels_last)
    vec = 8 if dt == torch.bfloat16 else 4
    ktrue = kh * kw * C
    kpad = (ktrue + 63) // 64 * 64 if dt == torch.bfloat16 else (ktrue + vec - 1) // vec * vec
    rows = _native.im2col2d(x, (kh, kw), (st, st), (pd, pd), (dl, dl), kpad)
    oh = (H + 2 * pd - dl * (kh - 1) - 1) // st + 1
    ow = (W + 2 * pd - dl * (kw - 1) - 1) // st + 1
    assert rows.shape == (N * oh * ow, kpad)
    ref = F.unfold(x.float(), (kh, kw), dilation=dl, padding=pd, stride=st)          # [N, C*kh*kw, L]
    ref = ref.view(N, C, kh * kw, oh * ow).permute(0, 3, 2, 1).reshape(N * oh * ow, ktrue).to(dt)
    assert torch.equal(rows[:, :ktrue], ref)
    assert float(rows[:, ktrue:].float().abs().max()) == 0.0 if kpad > ktrue else True

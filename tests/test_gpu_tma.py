"""-m gpu: the TMA operand path (csrc/bt_tma.cuh).

(1) bt_tma_probe: one activation tile staged by exactly the tensor map + cp.async.bulk.tensor instruction the kernels
    use (tiled 2-D map for linear-like layers, im2col map for convolutions) must equal the im2col rows the reference's
    F.convNd / F.linear implies (conv_variational.py:205,379,552): BIT-EXACT -- it is a copy.
(2) bt_tma_kernel (W_s resident) / bt_tms_kernel (streaming): on identical (mu, rho, seed, x) they must agree with the
    cp.async kernel families (bf16: same operands, same k order -> bit-exact, measured 0.0 on B200) and meet the oracle
    tolerances (tf32 path: 1e-4 vs the oracle on tf32-rounded operands -- the TMA kernels round the staged fp32
    activation tile to nearest in shared memory, the tensor core alone would truncate -- 5e-4 vs fp32; bf16 output: 3e-3)."""
import pytest
import torch
import torch.nn.functional as F

import bayesian_torch_b200 as btb
from bayesian_torch_b200 import _native
from gpu_util import build_layer, errs, note, oracle_forward
from test_gpu_direct import env

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


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
    return g


def _unswizzle(img, es):
    """[128, 128] uint8 shared-memory image -> [128 rows, 128/es elements] in logical k order"""
    rows = []
    for r in range(128):
        chunks = [img[r, ((c ^ (r & 7)) << 4):((c ^ (r & 7)) << 4) + 16] for c in range(8)]
        rows.append(torch.cat(chunks))
    raw = torch.stack(rows).contiguous()
    return raw.view(torch.bfloat16 if es == 2 else torch.float32)


PROBES = [
    # nd, cin, sp, k, stride, pad, dil, groups, B, dtype
    (2, 64, (8, 8), 3, 1, 1, 1, 1, 5, torch.bfloat16),
    (2, 128, (9, 7), 3, 2, 1, 1, 1, 4, torch.bfloat16),
    (2, 64, (11, 6), 3, 1, 2, 2, 1, 3, torch.bfloat16),
    (2, 128, (8, 8), 1, 2, 0, 1, 2, 6, torch.bfloat16),          # 1x1 stride 2, two groups
    (2, 32, (6, 6), 3, 1, 1, 1, 1, 7, torch.float32),            # tf32 path: 32 channels per k-block
    (2, 64, (5, 5), 3, 2, 1, 1, 1, 9, torch.float32),
    (1, 64, (37,), 5, 2, 2, 1, 1, 4, torch.bfloat16),
    (3, 64, (3, 5, 5), 3, 1, 1, 1, 1, 2, torch.bfloat16),
    (3, 32, (4, 4, 6), 2, 2, 0, 1, 1, 3, torch.float32),
    (0, 200, (), 1, 1, 0, 1, 1, 300, torch.bfloat16),             # linear: tiled map, K tail zero-filled
    (0, 100, (), 1, 1, 0, 1, 1, 130, torch.float32),
]


@pytest.mark.parametrize("cfg", PROBES, ids=lambda c: f"nd{c[0]}_c{c[1]}_{'x'.join(map(str, c[2]))}_k{c[3]}s{c[4]}p{c[5]}d{c[6]}g{c[7]}_{str(c[9])[6:]}")
def test_tma_probe_equals_im2col_rows(cfg):
    nd, cin, sp, k, stride, pad, dil, groups, B, dt = cfg
    torch.manual_seed(3)
    S = 2
    x = torch.randn(S * B, *sp, cin).to(dt).to(DEV)              # physical channels-last, two stacked MC samples
    g = _geom(S, B, cin, 64 * groups, sp, k, stride, pad, dil, groups)
    es = x.element_size()
    kbe = 128 // es
    cin_g = cin // groups
    osp = tuple(g.out_dhw[3 - nd + i] for i in range(nd))
    M = B
    for o in osp:
        M *= o
    # the im2col matrix of the reference: unfold the zero-padded input
    xl = x.float().cpu().view(S * B, *sp, cin)
    taps = [()]
    for _ in range(nd):
        taps = [t + (kk,) for t in taps for kk in range(k)]
    checked = 0
    for (m0, smp, grp, tap_i, slab) in [(0, 0, 0, 0, 0), (max(M - 128, 0) // 3 + 1 if M > 128 else 0, 1, groups - 1, len(taps) - 1, 0),
                                        (max(M - 40, 0), 1, 0, len(taps) // 2, (cin_g + kbe - 1) // kbe - 1)]:
        img = _native.tma_probe(g, x, m0, sample=smp, group=grp, tap=tap_i, slab=slab)
        torch.cuda.synchronize()
        got = _unswizzle(img.cpu(), es).float()
        c0 = grp * cin_g + slab * kbe
        for r in range(128):
            m = m0 + r
            if m >= M:
                break
            idx = []
            rem = m
            for o in reversed(osp):
                idx.append(rem % o)
                rem //= o
            b = rem
            idx = idx[::-1]
            pos = [idx[i] * stride - pad + taps[tap_i][i] * dil for i in range(nd)]
            exp = torch.zeros(kbe)
            if all(0 <= pos[i] < sp[i] for i in range(nd)):
                row = xl[(smp * B + b, *pos)]
                hi = min(c0 + kbe, cin if nd == 0 else (grp + 1) * cin_g)
                exp[: hi - c0] = row[c0:hi]
            assert torch.equal(got[r], exp), (cfg, m0, smp, grp, tap_i, slab, r)
            checked += 1
    assert checked > 0


CASES = [
    # kind, nd, cin, cout, ks, stride, pad, dil, groups, bias, batch, spatial, xdtype, pdtype
    ("conv", 2, 128, 256, 3, 2, 1, 1, 1, False, 5, (4, 4), torch.bfloat16, torch.bfloat16),     # ResNet layer3 entry
    ("conv", 2, 256, 256, 3, 1, 1, 1, 1, False, 9, (2, 2), torch.bfloat16, torch.bfloat16),     # layer3
    ("conv", 2, 512, 512, 3, 1, 1, 1, 1, True, 20, (1, 1), torch.bfloat16, torch.float32),      # layer4: centre tap only
    ("conv", 2, 64, 128, 1, 2, 0, 1, 1, False, 6, (8, 8), torch.bfloat16, torch.bfloat16),      # 1x1 stride-2 downsample
    ("conv", 2, 128, 96, 3, 1, 1, 1, 2, True, 3, (7, 5), torch.bfloat16, torch.bfloat16),       # groups, N tail
    ("conv", 2, 64, 64, 3, 1, 1, 1, 1, True, 4, (8, 8), torch.float32, torch.float32),          # tf32 layer1
    ("conv", 2, 128, 128, 3, 2, 1, 1, 1, False, 5, (4, 4), torch.float32, torch.float32),       # tf32 stride 2
    ("conv", 2, 64, 48, 3, 1, 2, 2, 1, True, 3, (9, 6), torch.float32, torch.float32),          # tf32 dilation
    ("conv", 1, 64, 96, 5, 2, 2, 1, 1, True, 4, (37,), torch.bfloat16, torch.float32),
    ("conv", 3, 64, 32, 3, 1, 1, 1, 1, True, 2, (3, 5, 5), torch.bfloat16, torch.bfloat16),
    ("conv", 3, 32, 40, 2, 2, 0, 1, 1, False, 3, (4, 4, 6), torch.float32, torch.float32),
    ("linear", 0, 1024, 1024, None, 1, 0, 1, 1, True, 256, (), torch.float32, torch.float32),   # C1
    ("linear", 0, 512, 10, None, 1, 0, 1, 1, True, 128, (), torch.bfloat16, torch.bfloat16),    # ResNet-18 fc
    ("linear", 0, 200, 100, None, 1, 0, 1, 1, True, 300, (), torch.bfloat16, torch.bfloat16),   # K tail, N tail, M tail
    ("linear", 0, 2048, 384, None, 1, 0, 1, 1, False, 700, (), torch.bfloat16, torch.bfloat16), # many k-blocks, M groups
]


def _run(layer, x, seed, residual=None):
    btb.manual_seed(seed)
    y = layer._forward_impl(x, False, residual=residual)
    torch.cuda.synchronize()
    return y, _native.last_forward_path()


@pytest.mark.parametrize("mode", ["1", "2"], ids=["resident", "stream"])
@pytest.mark.parametrize("cfg", CASES, ids=lambda c: f"{c[0]}{c[1]}_{c[2]}x{c[3]}_k{c[4]}s{c[5]}g{c[8]}_{str(c[12])[6:]}")
def test_tma_kernels_equal_cp_async_kernels_and_oracle(cfg, mode):
    kind, nd, cin, cout, ks, st, pad, dil, groups, bias, batch, sp, xdt, pdt = cfg
    torch.manual_seed(cin + cout + batch)
    layer = build_layer(kind, nd, False, cin, cout, ks, st, pad, dil, groups, bias).to(DEV).to(pdt)
    x = torch.randn(batch, cin, *sp).to(xdt).to(DEV)
    with env(BT_DISABLE_TMA=None, BT_TMA_PREFER="1", BT_TMA_MODE=mode):
        yt, path_t = _run(layer, x, 31)
    if not path_t.startswith("tma"):
        pytest.skip(f"no {('resident', 'streaming')[int(mode) - 1]} TMA plan fits this shape (path {path_t})")
    with env(BT_DISABLE_TMA="1", BT_TMA_PREFER=None, BT_TMA_MODE=None):
        yo, path_o = _run(layer, x, 31)
    assert not path_o.startswith("tma")
    rel_to, mx_to = errs(yt, yo)
    layer._bt_last["sample0"] = 0
    eps_w, eps_b = layer.materialize_eps(0)
    yr = oracle_forward(layer, x, eps_w, eps_b, round_operands=True, path=path_t)
    yf = oracle_forward(layer, x, eps_w, eps_b, round_operands=False)
    rel_r, mx_r = errs(yt, yr)
    rel_f, _ = errs(yt, yf)
    tf32 = xdt == torch.float32 and pdt == torch.float32
    note("tma_vs_other", cfg=str(cfg), mode=mode, path=path_t, other=path_o, rel_vs_other=rel_to, rel_rounded=rel_r, rel_fp32=rel_f)
    msg = f"{path_t} vs {path_o}: rel {rel_to:.2e} max {mx_to:.2e}; vs rounded oracle {rel_r:.2e} (max {mx_r:.2e}); vs fp32 {rel_f:.2e}"
    # same operands (bf16 / tf32: both rounded to nearest), same k order -> bit-exact (tf32: the generic path runs
    # M-subtiles of a different shape, allow fp32 summation-order noise)
    assert rel_to <= (2e-6 if tf32 else 0.0), msg
    assert rel_r <= (1e-4 if tf32 else 3e-3), msg
    assert rel_f <= (5e-4 if tf32 else 3e-3), msg


@pytest.mark.parametrize("xdt", [torch.bfloat16, torch.float32], ids=["bf16", "tf32"])
def test_tma_mc_samples_epilogue_and_tile_boundaries(xdt):
    """S samples in one launch (shared x for the first layer, stacked afterwards), fused affine / residual / ReLU
    epilogue, several row tiles per CTA and per sample; single-sample launches reproduce the stacked launch bit-exactly."""
    torch.manual_seed(9)
    conv1 = build_layer("conv", 2, False, 64, 64, 3, 2, 1, 1, 1, True).to(DEV).to(xdt)
    conv2 = build_layer("conv", 2, False, 64, 128, 1, 1, 0, 1, 1, False).to(DEV).to(xdt)
    B, S = 37, 3
    x = torch.randn(B, 64, 9, 9).to(xdt).to(DEV)
    scale = (torch.rand(128, device=DEV) + 0.5)
    shift = torch.randn(128, device=DEV)
    conv2._bt_ep_scale, conv2._bt_ep_shift, conv2._bt_ep_relu = scale, shift, True
    outs = {}
    for mode, e in (("tma", dict(BT_DISABLE_TMA=None, BT_TMA_PREFER="1")), ("other", dict(BT_DISABLE_TMA="1", BT_TMA_PREFER=None))):
        with env(**e):
            btb.manual_seed(5)
            with btb.mc_sample_context(S, B, 100):
                h = conv1(x, return_kl=False)
                p1 = _native.last_forward_path()
                res = torch.randn(S * B, 128, 5, 5, generator=torch.Generator().manual_seed(1)).to(xdt).to(DEV) \
                    .contiguous(memory_format=torch.channels_last)
                o = conv2._forward_impl(h, False, residual=res)
                p2 = _native.last_forward_path()
            torch.cuda.synchronize()
            assert p1.startswith("tma") == (mode == "tma") and p2.startswith("tma") == (mode == "tma"), (mode, p1, p2)
            outs[mode] = (h, o)
    ht, ot = outs["tma"]
    hi, oi = outs["other"]
    assert ht.shape == (S * B, 64, 5, 5) and ot.shape == (S * B, 128, 5, 5)
    r1, m1 = errs(ht, hi)
    r2, m2 = errs(ot, oi)
    note("tma_mc", dtype=str(xdt), r1=r1, r2=r2)
    assert (r1 <= 2e-6 and r2 <= 1e-4) if xdt == torch.float32 else (r1 == 0.0 and r2 == 0.0), (r1, m1, r2, m2)
    assert float(ot.min()) >= 0.0
    assert not torch.equal(ht[:B], ht[B:2 * B])
    with env(BT_DISABLE_TMA=None, BT_TMA_PREFER="1"):
        for s in range(S):
            btb.manual_seed(5)
            with btb.mc_sample_context(1, B, 100 + s):
                hs = conv1(x, return_kl=False)
            assert torch.equal(hs, ht[s * B:(s + 1) * B]), s


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32], ids=["bf16", "f32"])
def test_tma_tiled_window_rows_zero_fill_and_absolute_address_swizzle(dt):
    """What a TMA window loader for the direct kernel needs: a box {128 B of channels, W + pad pixels, 1 row} over a
    channels-last [N, H, W, C] tensor yields the padded pixel row (pixels w >= W zero-filled, rows h >= H all zero), and
    a destination that is 128-byte but not 1024-byte aligned is swizzled by ABSOLUTE shared-memory address (chunk c of
    buffer row j at j*128 + ((c ^ (j & 7)) << 4)), i.e. boxes can be stacked at arbitrary row offsets of one buffer."""
    torch.manual_seed(1)
    es = 2 if dt == torch.bfloat16 else 4
    kbe = 128 // es
    N, H, W, C = 3, 5, 7, 2 * kbe
    x = torch.randn(N, H, W, C).to(dt).to(DEV)
    Pw = W + 2
    for (c0, w0, h0, n0, row_off) in [(0, 0, 0, 0, 0), (kbe, 0, 3, 1, 5), (0, 0, H, 2, 11), (kbe, 2, 4, 2, 3)]:
        img = _native.tma_probe4d(x, (C, W, H, N), (kbe, Pw, 1, 1), (c0, w0, h0, n0), row_off * 128).cpu()
        for j in range(Pw):
            r = row_off + j                                   # buffer row
            chunks = [img[r, ((c ^ (r & 7)) << 4):((c ^ (r & 7)) << 4) + 16] for c in range(8)]
            got = torch.cat(chunks).contiguous().view(dt).float()
            w_ = w0 + j
            exp = torch.zeros(kbe)
            if w_ < W and h0 < H:
                exp = x[n0, h0, w_, c0:c0 + kbe].float().cpu()
            assert torch.equal(got, exp), (c0, w0, h0, n0, row_off, j)
        # rows outside the box are untouched (0xA5 fill)
        assert int(img[row_off + Pw, 0]) == 0xA5 and (row_off == 0 or int(img[row_off - 1, 0]) == 0xA5)


DT_CASES = [
    # nd, cin, cout, ks, pad, dil, bias, batch, spatial, xdtype, pdtype
    (2, 64, 64, 3, 1, 1, False, 4, (8, 8), torch.bfloat16, torch.float32),       # ResNet-18 layer1 @ CIFAR
    (2, 64, 64, 3, 1, 1, True, 5, (8, 8), torch.float32, torch.float32),         # ... fp32 model: tf32 windows, 2 slabs
    (2, 128, 128, 3, 1, 1, False, 9, (4, 4), torch.bfloat16, torch.bfloat16),    # layer2: 25-pixel planes
    (2, 64, 128, 3, 1, 1, True, 2, (14, 14), torch.bfloat16, torch.bfloat16),
    (2, 128, 64, 3, 2, 2, True, 3, (9, 7), torch.bfloat16, torch.bfloat16),      # dilation 2
    (2, 128, 96, 3, 1, 1, False, 20, (2, 2), torch.bfloat16, torch.float32),     # tiny images, N tail
    (2, 64, 32, 5, 2, 1, True, 2, (12, 10), torch.float32, torch.float32),       # 5x5, tf32
    (2, 64, 128, 3, 1, 1, False, 2, (56, 56), torch.bfloat16, torch.bfloat16),   # C2-sized image: 2 padded rows per tile
    (1, 64, 96, 5, 2, 1, True, 4, (37,), torch.bfloat16, torch.float32),
    (3, 64, 32, 3, 1, 1, True, 2, (3, 5, 5), torch.bfloat16, torch.bfloat16),
    (3, 32, 48, 3, 1, 1, False, 2, (4, 3, 6), torch.float32, torch.float32),
]


@pytest.mark.parametrize("cfg", DT_CASES, ids=lambda c: f"nd{c[0]}_{c[1]}x{c[2]}_k{c[3]}d{c[5]}_{'x'.join(map(str, c[8]))}_{str(c[9])[6:]}")
def test_tma_direct_kernel_equals_other_kernels_and_oracle(cfg):
    """bt_dtma_kernel (A operand read in place from a window staged by tiled TMA boxes, zero padding = out-of-range
    fill) vs the same layer with it disabled (cp.async direct kernel / im2col TMA kernels / generic) and vs the oracle."""
    nd, cin, cout, ks, pad, dil, bias, batch, sp, xdt, pdt = cfg
    torch.manual_seed(cin + cout + batch)
    layer = build_layer("conv", nd, False, cin, cout, ks, 1, pad, dil, 1, bias).to(DEV).to(pdt)
    x = torch.randn(batch, cin, *sp).to(xdt).to(DEV)
    with env(BT_DISABLE_DTMA=None):
        yt, path_t = _run(layer, x, 31)
    if path_t != "tma_direct":
        pytest.skip(f"resident tiles + two windows do not fit shared memory for this shape (path {path_t})")
    with env(BT_DISABLE_DTMA="1"):
        yo, path_o = _run(layer, x, 31)
    assert path_o != "tma_direct"
    rel_to, mx_to = errs(yt, yo)
    layer._bt_last["sample0"] = 0
    eps_w, eps_b = layer.materialize_eps(0)
    yr = oracle_forward(layer, x, eps_w, eps_b, round_operands=True)
    yf = oracle_forward(layer, x, eps_w, eps_b, round_operands=False)
    rel_r, mx_r = errs(yt, yr)
    rel_f, _ = errs(yt, yf)
    tf32 = xdt == torch.float32 and pdt == torch.float32
    note("dtma_vs_other", cfg=str(cfg), other=path_o, rel_vs_other=rel_to, rel_rounded=rel_r, rel_fp32=rel_f)
    msg = f"tma_direct vs {path_o}: rel {rel_to:.2e} max {mx_to:.2e}; vs rounded oracle {rel_r:.2e} (max {mx_r:.2e}); vs fp32 {rel_f:.2e}"
    assert rel_to <= (2e-6 if tf32 else 0.0), msg
    assert rel_r <= (1e-4 if tf32 else 3e-3), msg
    assert rel_f <= (5e-4 if tf32 else 3e-3), msg


@pytest.mark.parametrize("xdt", [torch.bfloat16, torch.float32], ids=["bf16", "tf32"])
def test_tma_direct_mc_samples_epilogue_residual(xdt):
    torch.manual_seed(9)
    conv1 = build_layer("conv", 2, False, 64, 64, 3, 1, 1, 1, 1, True).to(DEV).to(xdt)
    conv2 = build_layer("conv", 2, False, 64, 128, 3, 1, 1, 1, 1, False).to(DEV).to(xdt)
    B, S = 37, 3
    x = torch.randn(B, 64, 8, 8).to(xdt).to(DEV)
    conv2._bt_ep_scale, conv2._bt_ep_shift, conv2._bt_ep_relu = torch.rand(128, device=DEV) + 0.5, torch.randn(128, device=DEV), True
    outs = {}
    for mode, e in (("dtma", dict(BT_DISABLE_DTMA=None)), ("other", dict(BT_DISABLE_DTMA="1"))):
        with env(**e):
            btb.manual_seed(5)
            with btb.mc_sample_context(S, B, 100):
                h = conv1(x, return_kl=False)
                p1 = _native.last_forward_path()
                res = torch.randn(S * B, 128, 8, 8, generator=torch.Generator().manual_seed(1)).to(xdt).to(DEV) \
                    .contiguous(memory_format=torch.channels_last)
                o = conv2._forward_impl(h, False, residual=res)
                p2 = _native.last_forward_path()
            torch.cuda.synchronize()
            assert (p1 == "tma_direct") == (mode == "dtma") and (p2 == "tma_direct") == (mode == "dtma"), (mode, p1, p2)
            outs[mode] = (h, o)
    r1, m1 = errs(outs["dtma"][0], outs["other"][0])
    r2, m2 = errs(outs["dtma"][1], outs["other"][1])
    note("dtma_mc", dtype=str(xdt), r1=r1, r2=r2)
    assert (r1 <= 2e-6 and r2 <= 1e-4) if xdt == torch.float32 else (r1 == 0.0 and r2 == 0.0), (r1, m1, r2, m2)
    assert float(outs["dtma"][1].min()) >= 0.0
    with env(BT_DISABLE_DTMA=None):
        for s in range(S):
            btb.manual_seed(5)
            with btb.mc_sample_context(1, B, 100 + s):
                hs = conv1(x, return_kl=False)
            assert torch.equal(hs, outs["dtma"][0][s * B:(s + 1) * B]), s


FLIP_CASES = [
    # kind, nd, cin, cout, ks, stride, pad, dil, groups, bias, batch, spatial, xdtype, pdtype
    ("linear", 0, 512, 300, None, 1, 0, 1, 1, True, 130, (), torch.bfloat16, torch.bfloat16),
    ("linear", 0, 1024, 256, None, 1, 0, 1, 1, True, 700, (), torch.bfloat16, torch.bfloat16),     # several M groups
    ("linear", 0, 256, 128, None, 1, 0, 1, 1, True, 200, (), torch.float32, torch.float32),        # tf32 + Flipout
    ("conv", 2, 256, 64, 1, 1, 0, 1, 1, False, 3, (14, 14), torch.bfloat16, torch.bfloat16),       # ResNet-50 bottleneck 1x1
    ("conv", 2, 64, 128, 3, 2, 1, 1, 1, True, 4, (9, 9), torch.bfloat16, torch.bfloat16),          # strided 3x3: im2col map
    ("conv", 2, 128, 96, 3, 1, 1, 1, 2, True, 3, (7, 5), torch.bfloat16, torch.float32),           # groups, N tail
    ("conv", 2, 64, 64, 3, 2, 1, 1, 1, True, 5, (8, 8), torch.float32, torch.float32),             # tf32 conv
    ("conv", 1, 64, 96, 5, 2, 2, 1, 1, True, 4, (37,), torch.bfloat16, torch.bfloat16),
    ("conv", 3, 64, 64, 3, 2, 1, 1, 1, False, 2, (5, 6, 6), torch.bfloat16, torch.bfloat16),
]


@pytest.mark.parametrize("cfg", FLIP_CASES, ids=lambda c: f"{c[0]}{c[1]}_{c[2]}x{c[3]}_k{c[4]}s{c[5]}g{c[8]}_{str(c[12])[6:]}")
def test_tma_streaming_flipout_equals_cp_async_kernels_and_oracle(cfg):
    """bt_tms_kernel<FLIP>: mean tile + perturbation tile, the x * s_in plane built by the transform warps from the
    TMA-staged tile, two accumulators, output signs in the epilogue -- vs the cp.async Flipout kernels (same draws, same
    operands: bit-exact) and vs the oracle on the re-materialised eps / signs (linear_flipout.py:145-197,
    conv_flipout.py:370-439)."""
    kind, nd, cin, cout, ks, st, pad, dil, groups, bias, batch, sp, xdt, pdt = cfg
    torch.manual_seed(cin + cout + batch)
    layer = build_layer(kind, nd, True, cin, cout, ks, st, pad, dil, groups, bias).to(DEV).to(pdt)
    x = torch.randn(batch, cin, *sp).to(xdt).to(DEV)
    with env(BT_DISABLE_TMA=None, BT_FORCE_DIRECT=None):
        yt, path_t = _run(layer, x, 31)
    if path_t != "tma_stream":
        pytest.skip(f"no streaming TMA plan for this Flipout shape (path {path_t})")
    with env(BT_DISABLE_TMA="1"):
        yo, path_o = _run(layer, x, 31)
    assert not path_o.startswith("tma")
    rel_to, mx_to = errs(yt, yo)
    layer._bt_last["sample0"] = 0
    eps_w, eps_b = layer.materialize_eps(0)
    s_in, s_out = layer.materialize_signs(tuple(x.shape), tuple(yt.shape), 0)
    yr = oracle_forward(layer, x, eps_w, eps_b, s_in, s_out, round_operands=True)
    yf = oracle_forward(layer, x, eps_w, eps_b, s_in, s_out, round_operands=False)
    rel_r, mx_r = errs(yt, yr)
    rel_f, _ = errs(yt, yf)
    tf32 = xdt == torch.float32 and pdt == torch.float32
    note("tms_flip", cfg=str(cfg), other=path_o, rel_vs_other=rel_to, rel_rounded=rel_r, rel_fp32=rel_f)
    msg = f"tma_stream vs {path_o}: rel {rel_to:.2e} max {mx_to:.2e}; vs rounded oracle {rel_r:.2e} (max {mx_r:.2e}); vs fp32 {rel_f:.2e}"
    assert rel_to <= (2e-6 if tf32 else 0.0), msg
    assert rel_r <= (1e-4 if tf32 else 3e-3), msg
    assert rel_f <= (5e-4 if tf32 else 3e-3), msg


def test_tma_streaming_flipout_mc_samples_and_epilogue():
    torch.manual_seed(4)
    conv = build_layer("conv", 2, True, 64, 128, 1, 1, 0, 1, 1, True).to(DEV).bfloat16()
    B, S = 21, 3
    x = torch.randn(B, 64, 6, 6).bfloat16().to(DEV)
    conv._bt_ep_scale, conv._bt_ep_shift, conv._bt_ep_relu = torch.rand(128, device=DEV) + 0.5, torch.randn(128, device=DEV), True
    outs = {}
    for mode, e in (("tma", dict(BT_DISABLE_TMA=None)), ("other", dict(BT_DISABLE_TMA="1"))):
        with env(**e):
            btb.manual_seed(5)
            with btb.mc_sample_context(S, B, 40):
                res = torch.randn(S * B, 128, 6, 6, generator=torch.Generator().manual_seed(1)).bfloat16().to(DEV) \
                    .contiguous(memory_format=torch.channels_last)
                o = conv._forward_impl(x, False, residual=res)
                pth = _native.last_forward_path()
            torch.cuda.synchronize()
            assert pth.startswith("tma") == (mode == "tma"), (mode, pth)
            outs[mode] = o
    assert torch.equal(outs["tma"], outs["other"])
    assert float(outs["tma"].min()) >= 0.0


POOL_CASES = [
    # cin, cout, k, stride, pad, in_hw, B, S, shared x, dtype       -> conv output (OH, OW)
    (3, 64, 7, 2, 3, (32, 32), 6, 5, True, torch.bfloat16),         # the C3 stem: 16x16, two tiles per image, S % nsmp != 0
    (3, 64, 7, 2, 3, (32, 32), 6, 3, True, torch.float32),          # tf32 stem
    (3, 64, 3, 1, 1, (16, 8), 5, 2, True, torch.bfloat16),          # 16x8 = one tile per image (no carry)
    (64, 128, 1, 1, 0, (32, 32), 3, 2, False, torch.bfloat16),      # im2col TMA map, 8 tiles per image, stacked x, 128 outputs
    (64, 64, 3, 1, 1, (8, 64), 3, 2, False, torch.bfloat16),        # OW = 64: two output rows per tile
    (32, 64, 3, 1, 1, (32, 4), 4, 2, False, torch.float32),         # OW = 4: 32 output rows per tile
    (3, 64, 7, 2, 3, (64, 64), 2, 4, True, torch.bfloat16),         # 32x32 stem output
]


@pytest.mark.parametrize("case", POOL_CASES, ids=lambda c: f"{c[0]}x{c[1]}_k{c[2]}s{c[3]}_{c[5][0]}x{c[5][1]}_S{c[7]}_{str(c[9]).split('.')[-1]}")
def test_tma_resident_kernel_fused_stem_maxpool_is_bit_exact(case):
    """conv -> folded bn -> relu -> MaxPool2d(3, 2, 1) (torchvision resnet.py stem) with the pool inside the kernel's
    epilogue (BtLayerGeom.pool_hw) == the same launch without it followed by bt_maxpool2d_nhwc and by ATen's
    F.max_pool2d: BIT-EXACT (max commutes with the rounding to the output dtype)."""
    cin, cout, k, stride, pad, hw, B, S, shared, dt = case
    torch.manual_seed(11)
    conv = build_layer("conv", 2, False, cin, cout, k, stride, pad, 1, 1, True).to(DEV).to(dt)
    conv._bt_ep_scale, conv._bt_ep_shift, conv._bt_ep_relu = torch.rand(cout, device=DEV) + 0.5, torch.randn(cout, device=DEV), True
    x = torch.randn(B if shared else S * B, cin, *hw).to(dt).to(DEV)
    outs = {}
    with env(BT_DISABLE_DTMA="1", BT_TMA_MODE="1", BT_TMA_PREFER="1"):
        for pool in (False, True):
            conv._bt_ep_pool = pool
            btb.manual_seed(3)
            with btb.mc_sample_context(S, B, 7):
                y = conv(x, return_kl=False)
                pth = _native.last_forward_path()
            torch.cuda.synchronize()
            assert pth == "tma", (pool, pth)
            assert bool(getattr(y, "_bt_pooled", False)) == pool, "the plan refused to fuse the pool"
            outs[pool] = y
    full, pooled = outs[False], outs[True]
    oh, ow = full.shape[2:]
    assert pooled.shape == (S * B, cout, oh // 2, ow // 2)
    ref = F.max_pool2d(full, 3, 2, 1)
    mine = _native.maxpool2d_nhwc(full.permute(0, 2, 3, 1).contiguous(), (3, 3), (2, 2), (1, 1)).permute(0, 3, 1, 2)
    assert torch.equal(mine, ref)
    assert torch.equal(pooled, ref), float((pooled.float() - ref.float()).abs().max())
    assert float(pooled.min()) >= 0.0 and float(pooled.max()) > 0.0


def test_fused_pool_is_refused_where_the_tiling_cannot_hold_whole_rows():
    """OW = 12 does not divide the 128-row tile: the plan reports pool_fused == 0, the layer pools separately, and a
    direct bt_layer_forward with pool_hw set fails loudly instead of writing a wrong shape."""
    torch.manual_seed(0)
    conv = build_layer("conv", 2, False, 3, 64, 3, 1, 1, 1, 1, True).to(DEV).bfloat16()
    conv._bt_ep_relu, conv._bt_ep_pool = True, True
    x = torch.randn(4, 3, 12, 12).bfloat16().to(DEV)
    with btb.mc_sample_context(2, 4, 0):
        y = conv(x, return_kl=False)
    assert not getattr(y, "_bt_pooled", False) and y.shape == (8, 64, 12, 12)
    g = conv._bt_last["geom"]
    g.pool_hw[0] = g.pool_hw[1] = 12
    assert _native.plan_forward(_native.MODE_REPARAM, g, torch.bfloat16, torch.bfloat16)["pool_fused"] == 0


@pytest.mark.parametrize("case", [
    # kind, nd, cin, cout, ks, stride, pad, batch, spatial, dtype
    ("linear", 0, 320, 256, None, 1, 0, 300, (), torch.bfloat16),              # two n-tiles of 128, three row tiles
    ("conv", 2, 128, 256, 3, 2, 1, 6, (4, 4), torch.bfloat16),                 # ResNet layer3.0.conv1 shape: im2col map
    ("conv", 2, 128, 256, 3, 2, 1, 6, (4, 4), torch.float32),                  # tf32: converter warps round every CTA's copy
    ("conv", 2, 256, 512, 1, 1, 0, 40, (1, 1), torch.bfloat16),                # four n-tiles, one row tile: rank 1 loads nothing
], ids=lambda c: f"{c[0]}{c[1]}_{c[2]}x{c[3]}_{str(c[9]).split('.')[-1]}")
def test_tma_streaming_cluster_multicast_is_bit_exact(case):
    """bt_tms_kernel with the A tiles multicast across a 2-CTA cluster of n-tiles (BtForwardPlan.cluster_n == 2) == the same
    kernel with every CTA loading its own copy (BT_DISABLE_CLUSTER): identical operands and order -> BIT-EXACT; S MC samples
    in one launch, fused affine + ReLU epilogue."""
    kind, nd, cin, cout, ks, stride, pad, B, sp, dt = case
    torch.manual_seed(21)
    lay = build_layer(kind, nd, False, cin, cout, ks, stride, pad, 1, 1, True).to(DEV).to(dt)
    lay._bt_ep_scale, lay._bt_ep_shift, lay._bt_ep_relu = torch.rand(cout, device=DEV) + 0.5, torch.randn(cout, device=DEV), True
    S = 3
    x = torch.randn(S * B, cin, *sp).to(dt).to(DEV)
    outs = {}
    for mode, e in (("cluster", dict(BT_DISABLE_CLUSTER=None)), ("plain", dict(BT_DISABLE_CLUSTER="1"))):
        with env(BT_TMA_MODE="2", BT_TMA_PREFER="1", BT_DISABLE_DTMA="1", BT_FORCE_CLUSTER="1", **e):
            btb.manual_seed(9)
            with btb.mc_sample_context(S, B, 11):
                y = lay(x, return_kl=False)
                pth = _native.last_forward_path()
                plan = _native.plan_forward(_native.MODE_REPARAM, lay._bt_last["geom"], dt, dt)
            torch.cuda.synchronize()
            assert pth == "tma_stream", (mode, pth)
            if mode == "plain":
                assert plan["cluster_n"] == 1, plan
            clustered = plan["cluster_n"] == 2 or locals().get("clustered", False)
            outs[mode] = y
    assert torch.equal(outs["cluster"], outs["plain"])
    assert float(outs["cluster"].min()) >= 0.0 and float(outs["cluster"].max()) > 0.0
    if case[0] == "linear" or case[8] == (4, 4):      # >= 2 row tiles per k-block: the cluster form must have been taken
        assert clustered


def test_tma_direct_cluster_shared_sampling_prologue_is_bit_exact():
    """BT_CLUSTER_PROLOGUE=1 (opt-in; measured slower): the x CTAs of one (sample, n-tile) form a cluster,
    each samples a share of W_s's k-blocks and writes it into every peer's shared memory -> same W_s -> BIT-EXACT."""
    torch.manual_seed(5)
    conv = build_layer("conv", 2, False, 64, 64, 3, 1, 1, 1, 1, True).to(DEV).bfloat16()
    conv._bt_ep_scale, conv._bt_ep_shift, conv._bt_ep_relu = torch.rand(64, device=DEV) + 0.5, torch.randn(64, device=DEV), True
    S, B = 4, 64
    x = torch.randn(S * B, 64, 8, 8).bfloat16().to(DEV)
    res = torch.randn(S * B, 64, 8, 8).bfloat16().to(DEV).contiguous(memory_format=torch.channels_last)
    outs = {}
    for mode, e in (("cluster", dict(BT_CLUSTER_PROLOGUE="1")), ("plain", dict(BT_CLUSTER_PROLOGUE=None))):
        with env(**e):
            btb.manual_seed(2)
            with btb.mc_sample_context(S, B, 3):
                y = conv._forward_impl(x, False, residual=res)
                pth = _native.last_forward_path()
                plan = _native.plan_forward(_native.MODE_REPARAM, conv._bt_last["geom"], torch.bfloat16, torch.bfloat16,
                                            with_residual=True)
            torch.cuda.synchronize()
            assert pth == "tma_direct", (mode, pth)
            outs[mode] = (y, plan["cluster_n"])
    assert outs["plain"][1] == 1
    assert torch.equal(outs["cluster"][0], outs["plain"][0])
    assert outs["cluster"][1] >= 2, "the tiling search did not pick an even number of CTAs per sample: nothing was clustered"

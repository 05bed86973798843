"""Phase probe of bt_direct_kernel (needs a B200): times one layer shape under forced tilings and prints the
SM-clock stamps CTA (0,0,0) recorded (BT_DIRECT_TIMES) -- where a tile's time goes: window landing, MMA issue,
accumulator completion, epilogue."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bayesian_torch_b200 as btb  # noqa: E402
from bayesian_torch_b200 import _native  # noqa: E402
import bayesian_torch_b200.layers as L  # noqa: E402

DEV = "cuda:0"


def setenv(**kw):
    for k, v in kw.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)


def run(layer, x, S, B, iters=10):
    with btb.mc_sample_context(S, B, 0):
        for _ in range(3):
            layer(x, return_kl=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            layer(x, return_kl=False)
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000.0 / iters, _native.last_forward_path()


def stamps(layer, x, S, B):
    setenv(BT_DIRECT_TIMES=1)
    with btb.mc_sample_context(S, B, 0):
        layer(x, return_kl=False)
    torch.cuda.synchronize()
    setenv(BT_DIRECT_TIMES=None)
    ws = [w for (k, w) in _native._ws.items() if k[1] == "fwd"][0]
    t = ws.view(torch.int64)[:8 + 3 * 60 * 2].cpu().tolist()
    ws.zero_()
    t0 = t[0]
    out = {"sampled_clk": t[1] - t0}
    rows = []
    for it in range(60):
        r = [t[8 + (role * 60 + it) * 2 + w] for role in range(3) for w in range(2)]
        if r[0] == 0:
            break
        rows.append([v - t0 if v else 0 for v in r])
    out["tiles"] = rows   # [landed, next issued, mma start, mma issued, acc done, stored]
    return out


def main():
    res = []
    shapes = [
        ("layer1 64->64 3x3 8x8", False, 64, 64, 3, 1, (8, 8), 128, 64),
        ("layer2 128->128 3x3 4x4", False, 128, 128, 3, 1, (4, 4), 128, 64),
        ("C2 flipout 64->128 3x3 56x56 B=128", True, 64, 128, 3, 1, (56, 56), 128, 1),
        ("reparam 64->128 3x3 56x56 B=128", False, 64, 128, 3, 1, (56, 56), 128, 1),
    ]
    variants = [dict(), dict(BT_DIRECT_BN=64, BT_DIRECT_X=2), dict(BT_DIRECT_BN=64, BT_DIRECT_X=1)]
    shapes.append(("stem as linear 192->64 rows=32768 (x shared)", False, 192, 64, 0, 0, (), 32768, 64))
    for name, flip, cin, cout, k, pad, sp, B, S in shapes:
        torch.manual_seed(0)
        if k == 0:
            layer = L.LinearReparameterization(cin, cout, bias=False).to(DEV).bfloat16()
            x = torch.randn(B, cin, device=DEV).bfloat16()
        else:
            cls = L.Conv2dFlipout if flip else L.Conv2dReparameterization
            layer = cls(cin, cout, k, padding=pad, bias=False).to(DEV).bfloat16()
            x = torch.randn(B, cin, *sp, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
        setenv(BT_FORCE_DIRECT=None, BT_DISABLE_DIRECT=1)
        us, path = run(layer, x, S, B)
        print(f"{name}: im2col path={path} {us:.1f} us", flush=True)
        res.append(dict(shape=name, variant="im2col", path=path, us=us))
        setenv(BT_DISABLE_DIRECT=None, BT_FORCE_DIRECT=1)
        for v in variants:
            setenv(BT_DIRECT_BN=None, BT_DIRECT_X=None, BT_DIRECT_SLOTS=None)
            setenv(**v)
            try:
                us, path = run(layer, x, S, B)
            except Exception as e:  # noqa: BLE001
                print(f"  {v}: {e}")
                continue
            st = stamps(layer, x, S, B) if path == "direct" else {}
            tiles = st.get("tiles", [])
            summ = ""
            if len(tiles) >= 6:
                d = [[tiles[i + 1][c] - tiles[i][c] for c in range(6)] for i in range(2, len(tiles) - 1)]
                med = [sorted(col)[len(col) // 2] for col in zip(*d)]
                summ = f"sampled@{st['sampled_clk']} tiles={len(tiles)} per-tile clk (median deltas) landed/issued/mma0/mma1/acc/stored={med} " \
                       f"first={tiles[0]} last={tiles[-1]}"
            print(f"  {v}: path={path} {us:.1f} us {summ}", flush=True)
            res.append(dict(shape=name, variant=v, path=path, us=us, stamps=st))
        setenv(BT_DIRECT_BN=None, BT_DIRECT_X=None, BT_DIRECT_SLOTS=None, BT_FORCE_DIRECT=None)
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"))


if __name__ == "__main__":
    main()

"""CPU: restatements of the "next" layer families (oracle/bt_oracle_next.py) against outputs minted
from the reference (tests/golden/make_golden_next.py -> next.npz).  Bit-exact in fp32: the restatement issues the same ATen
ops on the same draws -- with ONE intra-op thread, as the goldens were minted (ATen's transposed convolution changes its
fp32 summation order with the thread count: 1-4 ulp)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import bt_oracle_next as ON

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_Z = np.load(os.path.join(ROOT, "tests", "golden", "next.npz"))
with open(os.path.join(ROOT, "tests", "golden", "next_meta.json")) as _f:
    _META = json.load(_f)["cases"]


@pytest.fixture(autouse=True)
def _one_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)


def _case(name):
    pre = name + "/"
    return {k[len(pre):]: torch.from_numpy(_Z[k]) for k in _Z.files if k.startswith(pre)}


@pytest.mark.parametrize("name", sorted(n for n, m in _META.items() if m["kind"] == "convt"))
def test_conv_transpose_restatement_equals_reference(name):
    c, m = _case(name), _META[name]
    geo = dict(stride=m["stride"], padding=m["padding"], output_padding=m["output_padding"], groups=m["groups"],
               dilation=m["dilation"])
    bias = dict(mu_b=c.get("mu_b"), rho_b=c.get("rho_b"), eps_b=c.get("eps_b"))
    if m["flipout"]:
        y = ON.conv_transpose_flipout(m["nd"], c["x"], c["mu_w"], c["rho_w"], c["eps_w"], c["sign_in"], c["sign_out"],
                                      **bias, **geo)
    else:
        y = ON.conv_transpose_reparam(m["nd"], c["x"], c["mu_w"], c["rho_w"], c["eps_w"], **bias, **geo)
    assert torch.equal(y, c["y"]), float((y - c["y"]).abs().max())


@pytest.mark.parametrize("name", sorted(n for n, m in _META.items() if m["kind"] == "lstm"))
def test_lstm_restatement_equals_reference(name):
    c, m = _case(name), _META[name]
    lin = lambda t: {k: c[f"{t}_{k}"] for k in ("mu_w", "rho_w", "mu_b", "rho_b")}
    draws = []
    for t in range(m["T"]):
        d = {}
        for tag in ("ih", "hh"):
            d[tag] = {k: c[f"t{t}_{tag}_{k}"] for k in ("eps_w", "eps_b", "sign_in", "sign_out") if f"t{t}_{tag}_{k}" in c}
        draws.append(d)
    hseq, cseq = ON.lstm_forward(c["x"], lin("ih"), lin("hh"), draws, m["flipout"])
    assert torch.equal(hseq, c["hseq"]) and torch.equal(cseq, c["cseq"])

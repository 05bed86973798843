import json
import os
import sys

import numpy as np
import pytest
import torch

# the parity tests flip libbtb200's A/B switches (BT_FORCE_DIRECT, BT_DISABLE_DIRECT, ...) inside one process: ask the
# library to re-read them on every call (production reads them once at the first launch)
os.environ.setdefault("BT_DYNAMIC_ENV", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a); run with -m gpu on the GPU box")


def pytest_sessionstart(session):
    """libbtb200.so is built in-tree and git-ignored: (re)build it when it is missing or older than its sources
    (nvcc cross-compiles without a GPU, ~1 min; a no-op when the digest stamp matches).  A failed build is not hidden:
    the tests that load the library then fail with the loader's message."""
    try:
        from bayesian_torch_b200 import build as _b
        _b.build()
    except Exception as e:  # noqa: BLE001
        sys.stderr.write(f"conftest: could not build libbtb200.so: {e}\n")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "layers.npz"))
    with open(os.path.join(ROOT, "tests", "golden", "meta.json")) as f:
        meta = json.load(f)

    class G:
        def case(self, name):
            pre = name + "/"
            return {k[len(pre):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(pre)}

        @property
        def meta(self):
            return meta

        def names(self, kind=None):
            return [n for n, m in meta["cases"].items() if kind is None or m["kind"] == kind]

    return G()

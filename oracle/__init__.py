"""CPU oracle for the bayesian-torch hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / baseline.
The product path (``bayesian_torch_b200``) never imports this package and
fails loudly when its CUDA library is missing.

Parity status: the reference (IntelLabs/bayesian-torch @ aa7e57b) ships no
tests, golden vectors or KATs for this path, so the restatement in
``oracle/bt_oracle.py`` is pinned against *outputs of the reference itself*:
``tests/golden/make_golden.py`` imports the reference from /root/reference in
the build container, replays its modules under a fixed seed and commits the
(params, x, eps, signs, out, kl) tuples as ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` then checks the restatement against those
fixtures (bit-exact on CPU for fp32).  ``oracle/bt_oracle_grad.py`` (backward,
for the round-2 training path) is pinned the same way on gradients minted from
the reference's autograd (``tests/golden/make_golden_grad.py``,
``tests/test_oracle_grad.py``).
"""

"""CPU oracle for the perceptual-loss training step (TEST INFRASTRUCTURE ONLY).

This package is a plain-PyTorch, CPU-only restatement of the reference hot path
(`/root/reference/cnn.py`, `/root/reference/train_cnn.py:50-107,224-244,295-334`).
It is the checker, never the product: only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import it.  The
product package (`artist_style_transfer_b200`) must never import from here.

Parity pin: the reference ships no tests or golden vectors, so
the oracle is pinned against outputs of the reference classes themselves,
generated in the build container by `oracle/make_golden.py` (which imports
`/root/reference`) and committed under `tests/golden/`.
"""

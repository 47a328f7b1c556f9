"""CPU oracle for the MLA training-step hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-PyTorch restatement of the reference's algorithm (ZhuoyangLiu2005/MLA @ 072f5d8), each function citing
the reference file:line it follows.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import this package; `mla_b200/` never does (the product path fails loudly without
its CUDA library).

Pinning: the reference ships no tests, golden vectors or known-answer fixtures for this path, so
the oracle is pinned against the reference ITSELF: `tests/golden/make_golden.py` imports the unmodified reference
from /root/reference (through `oracle/ref_shim.py`), runs it on seeded inputs and commits the boundary tensors as
`tests/golden/*.npz`; `tests/test_oracle_vs_golden.py` replays the oracle against them on every CPU test run.
Third-party arithmetic that is not in the reference tree (timm 0.9.10 Mlp/RmsNorm, flash-attn) is restated from
its published definition: parity for those two pieces is UNPINNED.
"""

"""CPU oracle for the DSNT head hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in the product package (`dsnt_pose2d_b200/`) may import this package.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs use it, and only as the checker / the timed CPU
baseline -- never as the thing shipped.

Two independent restatements live here:

* `oracle.torch_port` -- the reference's own op chain (`src/dsnt/nn.py`,
  `src/dsnt/model.py:24-63,138-145,233-246`) restated on CPU torch tensors with
  autograd, so gradients come from the same engine the reference uses.  Runs in
  fp64 (the arbiter) or fp32 ("what the reference would print").  This is also
  the CPU baseline that `bench.py` times.
* `oracle.closed_form` -- a numpy fp64 closed-form forward + analytic backward
  (SURVEY.md Appendix A).  Independent of autograd; it cross-checks the port and
  documents the exact formulas the CUDA kernels implement.

Parity is PINNED: both are checked in `tests/test_oracle_golden.py` against
(a) every known-answer vector of the reference's `tests/test_nn.py` and
(b) `tests/golden/*.npz`, produced by importing the unmodified reference from
`/root/reference/src` in the build container (`tests/golden/make_golden.py`).
"""

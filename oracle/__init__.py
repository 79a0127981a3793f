"""CPU oracle for the RDEIC relay decode path — TEST INFRASTRUCTURE ONLY.

A plain numpy / PyTorch-fp32 restatement of the reference algorithms on the hot path
(SURVEY.md §8a), each function citing the reference file:line it follows.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
this package; the product package `rdeic_b200` never does.

Pinning: the reference ships no tests or golden vectors.  The oracle is pinned
against the reference *itself*, imported from /root/reference in the build container through
import shims (tests/ref_harness.py) — see tests/golden/make_golden.py, which writes the
fixtures under tests/golden/ that travel to the GPU box; tests/test_oracle_cpu.py, which checks the
oracle against those fixtures everywhere; and tests/test_oracle_vs_reference.py, which re-runs the
generators against the live reference whenever /root/reference is present and checks that the committed
fixtures reproduce (integer / byte work bit-exact, fp32 tensors to 1e-5).
The compressai 1.2.4 arithmetic (`build_indexes`, `quantize`) is not vendored in the reference
tree and not installed here: that part is restated from the published compressai source and
is "parity unpinned" beyond the reference's own call sites (utils/ckbd.py:76-115).
"""

"""CPU oracle for the two-tower hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

This package restates, on the CPU in fp32 (torch CPU tensors / numpy), the
arithmetic of the reference's two-tower hot path.  Each
function cites the reference file:line it follows.  Paths are relative to the
reference checkout (/root/reference, which does NOT exist on the GPU box).

Who may import this package (and nobody else):
  * tests/                      -- as the checker for the CUDA path
  * __graft_entry__.smoke()     -- as the checker for one tiny invocation
  * bench.py                    -- only the `cpu_baseline` leg and
                                   `--impl reference`, as the thing timed on
                                   the host cores, never on the product path

The product package never imports it and has no CPU fallback: with the CUDA
library missing it raises.

Pinning: the reference ships no tests and no golden vectors,
so the oracle is pinned against outputs of the reference's own modules run in
the build container: tests/golden/make_golden.py imports /root/reference and
writes tests/golden/*.pt; tests/test_oracle_golden.py checks every oracle
function against those fixtures (CPU suite, `-m "not gpu"`).
Exception: the FM / DeepFM path (F1) has NO implementation in the reference
 -- oracle/fm.py restates the published deepctr-torch 0.2.9
formula and is "parity unpinned".
"""

from . import embed, losses, towers, retrieval, fm  # noqa: F401

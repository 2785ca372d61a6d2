"""CPU oracle for the batched-MCTS hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain-Python / numpy / torch-fp32 restatement of the reference's
algorithm for the hot path named in BASELINE.json (select / expand / evaluate /
rollout / backup + the four game engines + ConvolutionalNet leaf evaluation).
Every function cites the reference file:line it restates
(paths relative to /root/reference/deep_mcts/).

Who may import it: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs, and there only as the checker or
as the timed CPU baseline -- never as the product.  The product path
(``deep_mcts_b200``) never imports ``oracle`` and fails loudly when its CUDA
library is missing.

Pinning status: the reference ships no tests or golden vectors, so
the pins are outputs of the unmodified reference itself, generated in the build
container by ``tests/golden/make_golden.py`` (which imports /root/reference) and
committed under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks every
oracle function against those fixtures.
"""

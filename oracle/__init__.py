"""CPU oracle for the LineRefineNet forward hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the timed CPU baseline), never as the thing shipped.  The
product path (``pointnet_refine_b200``) never imports this package and fails
loudly when its CUDA library is missing.

Parity pinning: the reference ships no tests, golden vectors or checkpoints
for this path, so the oracle is pinned against
outputs of the reference itself: ``oracle/make_golden.py`` imports the
unmodified ``/root/reference/src/model.py`` in the build container, runs it on
CPU fp32 on deterministic synthetic weights/inputs (``oracle/synth.py``) and
commits the outputs under ``tests/golden/``; ``tests/test_oracle.py`` checks
the restatements against those fixtures.
"""

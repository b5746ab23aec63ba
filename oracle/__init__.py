"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (NumPy / torch-CPU) of the reference algorithms on the hot path of
BloodAxe/pytorch-toolbelt @ 05e07ac (v0.8.0). Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this package, and only as the
checker or the timed CPU baseline -- never as a fallback of ``pytorch_toolbelt_b200``.

Pinning: ``oracle/make_golden.py`` imports the real reference from /root/reference (it is pure Python and runs
in the build container), checks every function here against it on seeded inputs and writes the resulting
vectors to ``tests/golden/*.npz``. The CPU test-suite re-checks the oracle against those committed vectors, so
the oracle is pinned to outputs of the reference itself. The reference cannot travel to the GPU box (no
/root/reference there); the golden vectors and this restatement do.
"""

"""CPU oracle for the STCD inference hot path — TEST INFRASTRUCTURE, not product code.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  Nothing under ``stcd_b200/`` imports it, and
the product path fails loudly when the CUDA library is missing instead of falling back here.

Contents
--------
nets.py       functional fp32 restatements (``torch.nn.functional`` on CPU) of the reference
              forwards, each citing the reference file:line it follows.
metric.py     numpy restatement of ``SegmentationMetric`` (train_stcd.py:515-593).
emulate.py    executes a lowered ``stcd_b200.lowering.Program`` on the CPU with the kernel's
              exact data layout and bf16 rounding points (checks the host-side lowering).
refimport.py  import shims that make the real reference importable from /root/reference in the
              build container (used to pin the restatements and to generate tests/golden/).
make_golden.py  the script that generated tests/golden/*.npz from the real reference.

Parity pinning: the reference ships no tests or golden vectors, so the oracle
is pinned against outputs of the reference itself: ``make_golden.py`` ran the unmodified
reference modules here and committed inputs' seeds + logits under tests/golden/;
``tests/test_oracle.py`` checks nets.py against those fixtures everywhere and against the live
reference when /root/reference is present.
"""

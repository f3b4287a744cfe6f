"""ORACLE — test infrastructure only.

CPU/any-device fp32 restatement (plain torch functional ops / numpy fp64 tables) of the
reference's algorithms for the hot path. Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package, and only as the checker or the
timed CPU baseline — never from diffbir_b200 (the product fails loudly without its CUDA library).

Pinned against the reference itself: tests/golden/*.npz are produced by tests/golden/gen_golden.py
importing /root/reference; tests/test_oracle_golden.py checks every oracle function against them.
"""

"""TEST INFRASTRUCTURE ONLY.

CPU restatement ("oracle port") of the ParSeNet hot path, used as the checker in tests/,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of bench.py.
The product package (`parsenet-codebase_b200/`) never imports anything from here.

Pinned against golden vectors dumped from the *unmodified* reference by `oracle/make_golden.py`
(tests/golden/*.npz; see tests/test_oracle_golden.py).
"""

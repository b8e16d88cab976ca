"""TEST INFRASTRUCTURE ONLY — CPU (torch fp32 / numpy) restatement of the ParSeNet hot path.
Each function cites the reference file:line it follows.  Pinned by tests/test_oracle_golden.py."""

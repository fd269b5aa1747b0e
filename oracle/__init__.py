"""CPU oracle for the ArtiBoost synthesis + clasbased-network hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under `artiboost_b200/` imports this package; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may.
Each module's header states which reference lines it restates and how (or whether) it is pinned.
"""

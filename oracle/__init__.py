"""CPU oracle for the omnisafe on-policy hot path.

TEST INFRASTRUCTURE ONLY.  This package is a CPU restatement (numpy for the SafeRL-specific
arithmetic, torch-CPU for the nn/autograd pieces the reference itself delegates to torch) of the
reference algorithm on the path rollout -> dual GAE -> PPO-Lag / CPO update.  Every function cites
the reference file:line it follows.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import it -- and only as the checker or the timed CPU
baseline, never as part of the product path (`omnisafe_b200/` never imports `oracle`).

Parity pinning: the oracle is checked in `tests/test_oracle_golden.py` against
  * the reference's own known-answer vectors (tests/test_utils.py:L95-115 discount_cumsum;
    tests/test_policy.py:L68-74 CPO case selection), and
  * golden fixtures under `tests/golden/*.npz` produced by running the UNMODIFIED reference in the
    build container (`tests/golden/make_golden.py`, via `oracle/ref_shim.py`).
"""

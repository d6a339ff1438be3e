"""Test infrastructure only: CPU restatement of the reference's joint + transducer-loss path.

Nothing in `rnnt_b200/` may import this package.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs use it, and only as the checker or the
CPU baseline -- never as the thing measured or shipped.
"""

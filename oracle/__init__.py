"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

CPU oracle for the string_grouper hot path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this package; `string_grouper_b200/` never does.
"""

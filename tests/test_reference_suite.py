"""The reference's OWN test file (tests/golden/reference_test_string_grouper.py, vendored verbatim) run against
string_grouper_b200 under the reference's module names.

  not gpu : with the oracle-backed device stand-in (tests/cpu_backend.py) — the host logic (validation, option
            handling, block-guess / OverflowError flow, result shaping, mocks of _build_matches) answers like the reference
  gpu     : the same 54 tests through the CUDA path (K1, K2, K4 behind the C ABI)
"""
import importlib.util
import os
import sys
import unittest

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VENDORED = os.path.join(ROOT, "tests", "golden", "reference_test_string_grouper.py")


def _load_suite():
    import string_grouper_b200
    import string_grouper_b200.string_grouper as impl
    saved = {k: sys.modules.get(k) for k in ("string_grouper", "string_grouper.string_grouper")}
    sys.modules["string_grouper"] = string_grouper_b200
    sys.modules["string_grouper.string_grouper"] = impl
    spec = importlib.util.spec_from_file_location("reference_test_string_grouper", VENDORED)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    suite = unittest.defaultTestLoader.loadTestsFromModule(mod)
    return suite, saved


def _run():
    suite, saved = _load_suite()
    try:
        result = unittest.TestResult()
        suite.run(result)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    bad = [(str(t), tb.strip().splitlines()[-1]) for t, tb in result.failures + result.errors]
    return result.testsRun, bad


def test_reference_suite_host_logic_with_oracle_device():
    from cpu_backend import oracle_device
    with oracle_device():
        ran, bad = _run()
    assert ran >= 53, ran
    assert not bad, bad


@pytest.mark.gpu
def test_reference_suite_on_the_cuda_path():
    ran, bad = _run()
    assert ran >= 53, ran
    assert not bad, bad

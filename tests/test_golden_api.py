"""Public API against the golden answers of the UNMODIFIED reference (tests/golden/api_cases.json).

not-gpu run: host logic (ingest, options, pandas shaping) with the oracle standing in for the device;
-m gpu run: the same cases end to end through libsg_b200.so.
"""
import pytest

import string_grouper_b200 as api
from cpu_backend import oracle_device
from golden_util import assert_matches_golden, load_cases, run_case

CASES = load_cases()


@pytest.mark.parametrize("key", sorted(CASES))
def test_host_logic_against_reference_golden(key):
    with oracle_device():
        result = run_case(CASES[key], api)
    assert_matches_golden(result, CASES[key]["result"], label=key)


@pytest.mark.gpu
@pytest.mark.parametrize("key", sorted(CASES))
def test_cuda_path_against_reference_golden(key):
    result = run_case(CASES[key], api)
    assert_matches_golden(result, CASES[key]["result"], label=key)

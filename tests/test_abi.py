"""not-gpu: the C-ABI library loads and exports every symbol include/sg_b200.h declares (no compute calls)."""
import os
import re

from string_grouper_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "sg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), "libsg_b200.so does not export %s" % name
    assert sorted(_lib.SIGNATURES) == declared, set(_lib.SIGNATURES) ^ set(declared)


def test_pure_host_entry_points():
    lib = _lib.load()
    assert lib.sg_abi_version() == 3
    assert lib.sg_num_tiles(10_000, 3072) == 4 and lib.sg_num_tiles(0, 128) == 1
    assert lib.sg_tfidf_table_slots(3) == 2 ** 21 and lib.sg_tfidf_table_slots(5) == -1


def test_no_product_module_imports_the_oracle():
    pkg = os.path.join(ROOT, "string_grouper_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# oracle", ""), fn


def test_missing_device_fails_loudly():
    import pandas as pd
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import string_grouper_b200 as api
    with pytest.raises(_lib.SgB200Error):
        api.match_strings(pd.Series(["foo inc", "foo inc."]))

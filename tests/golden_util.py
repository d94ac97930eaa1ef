"""Replay of tests/golden/api_cases.json (answers of the UNMODIFIED reference, see oracle/make_golden.py)."""
import json
import os

import numpy as np
import pandas as pd

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_cases():
    with open(os.path.join(GOLDEN, "api_cases.json")) as f:
        return json.load(f)


def fixtures():
    acc = pd.read_csv(os.path.join(GOLDEN, "accounts_input.csv"))
    names = pd.Series(['Mega Enterprises Corporation', 'Hyper Startup Incorporated', 'Hyper Startup Inc.',
                       'Hyper-Startup Inc.', 'Hyper Hyper Inc.', 'Mega Enterprises Corp.'], name='Customer Name')
    ids = pd.Series(['BB016741P', 'CC082744L', 'AA098762D', 'BB099931J', 'HH072982K', 'EE059082Q'], name='Customer ID')
    names2 = pd.Series(['Mega Enterprises Corporation', 'Hyper Startup Incorporated', 'Hyper Startup Inc.',
                        'Hyper-Startup Inc.', 'HyperStartup Inc.', 'Hyper Hyper Inc.', 'Mega Enterprises Corp.'],
                       name='Customer Name')
    ids2 = pd.Series(['BB016741P', 'CC082744L', 'AA098762D', 'BB099931J', 'DD012339M', 'HH072982K', 'EE059082Q'],
                     name='Customer ID')
    multi = names.copy()
    multi.index = pd.MultiIndex.from_tuples([(1, 'a'), (1, 'b'), (2, 'a'), (2, 'b'), (3, 'a'), (3, 'b')],
                                            names=['lvl0', 'lvl1'])
    unnamed = pd.Series(names.tolist())
    shifted = names.copy()
    shifted.index = [10, 11, 12, 13, 14, 15]
    return {"acc.name": acc['name'], "acc.id": acc['id'], "names": names, "ids": ids, "names2": names2, "ids2": ids2,
            "multi": multi, "unnamed": unnamed, "shifted": shifted}


def resolve(arg, fx):
    if arg is None:
        return None
    if isinstance(arg, str):
        return fx[arg]
    return pd.Series(arg)


def run_case(case, api):
    fx = fixtures()
    args = [resolve(a, fx) for a in case["series"]]
    kw = dict(case["kwargs"])
    if kw.get("tfidf_matrix_dtype") == "float32":
        kw["tfidf_matrix_dtype"] = np.float32
    if "n_blocks" in kw:
        kw["n_blocks"] = tuple(kw["n_blocks"])
    return getattr(api, case["fn"])(*args, **kw)


def _norm(v):
    if v is None or v is pd.NA:
        return None
    if isinstance(v, (float, np.floating)):
        return None if np.isnan(v) else float(v)
    if isinstance(v, (np.integer,)):
        return int(v)
    if isinstance(v, tuple):
        return [_norm(x) for x in v]
    return v


def assert_matches_golden(result, gold, tol=1e-6, label=""):
    def close(a, b):
        if isinstance(a, float) and isinstance(b, (float, int)) and not isinstance(b, bool):
            return abs(a - float(b)) <= tol
        if isinstance(b, float) and isinstance(a, int) and not isinstance(a, bool):
            return abs(float(a) - b) <= tol
        return a == b

    if gold["kind"] == "series":
        assert isinstance(result, pd.Series), "%s: expected Series, got %r" % (label, type(result))
        assert result.name == gold["name"], "%s: name %r != %r" % (label, result.name, gold["name"])
        vals = [_norm(v) for v in result.tolist()]
        assert len(vals) == len(gold["values"]), label
        assert all(close(a, b) for a, b in zip(vals, gold["values"])), "%s: values %r != %r" % (label, vals, gold["values"])
        kind = result.dtype.kind if hasattr(result.dtype, "kind") else "O"
        assert kind == gold["dtype"] or {kind, gold["dtype"]} <= {"O", "T", "U"}, "%s: dtype kind %s != %s" % (label, kind, gold["dtype"])
    else:
        assert isinstance(result, pd.DataFrame), "%s: expected DataFrame, got %r" % (label, type(result))
        assert [str(c) for c in result.columns] == gold["columns"], "%s: columns %r != %r" % (label, list(result.columns), gold["columns"])
        rows = [[_norm(v) for v in row] for row in result.itertuples(index=False, name=None)]
        assert len(rows) == len(gold["data"]), "%s: %d rows != %d" % (label, len(rows), len(gold["data"]))
        for i, (ra, rb) in enumerate(zip(rows, gold["data"])):
            assert all(close(a, b) for a, b in zip(ra, rb)), "%s: row %d %r != %r" % (label, i, ra, rb)
        kinds = [result[c].dtype.kind if hasattr(result[c].dtype, "kind") else "O" for c in result.columns]
        for c, k, g in zip(result.columns, kinds, gold["dtypes"]):
            assert k == g or {k, g} <= {"O", "T", "U"}, "%s: column %r dtype kind %s != %s" % (label, c, k, g)
    idx = [_norm(t) for t in result.index.tolist()]
    assert idx == gold["index"], "%s: index %r != %r" % (label, idx[:10], gold["index"][:10])
    assert list(result.index.names) == gold["index_names"], "%s: index names" % label

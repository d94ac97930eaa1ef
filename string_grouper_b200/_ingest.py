"""Host side of the analyzer: pandas Series -> (uint8 bytes, int64 offsets) for the device vectoriser.

The device kernel (csrc/sg_tfidf.cu) implements the analyzer of the reference
(/root/reference/string_grouper/string_grouper.py:365-378) for ASCII text with
the default regex.  This module does what has to stay in Python:

  * validation (string_grouper.py:351-352, :988-995): every element must be a `str`;
  * zero-copy extraction of the Arrow offsets/data buffers of a pandas-3 `str` Series;
  * rows containing non-ASCII characters get `str.lower()` and
    `normalize('NFKD').encode('ASCII','ignore')` here, with the very same Python
    calls as string_grouper.py:372-375 (full Unicode semantics, length-changing);
  * a non-default `regex` is applied here with `re.sub` (string_grouper.py:376),
    the device then neither folds nor strips.
"""
import re
from unicodedata import normalize

import numpy as np
import pandas as pd
import pyarrow as pa

from ._lib import SG_FLAG_IGNORE_CASE, SG_FLAG_STRIP_DEFAULT

DEFAULT_REGEX = r'[,-./]|\s'


def is_series_of_strings(series):
    """Same verdict as StringGrouper._is_series_of_strings (string_grouper.py:988-995), vectorised."""
    if not isinstance(series, pd.Series):
        return False
    if len(series) == 0:
        return True
    if series.dtype == object:
        return pd.api.types.infer_dtype(series, skipna=False) == "string"
    if isinstance(series.dtype, pd.StringDtype) or pd.api.types.is_string_dtype(series.dtype):
        return not bool(series.isna().any())
    return False


def _arrow_buffers(series):
    """(data uint8, offsets int64) of the UTF-8 encoding of `series`, zero-copy when Arrow-backed."""
    arr = pa.array(series, from_pandas=True)
    if isinstance(arr, pa.ChunkedArray):
        arr = arr.combine_chunks() if arr.num_chunks != 1 else arr.chunk(0)
    if pa.types.is_null(arr.type) and len(arr) == 0:
        return np.zeros(0, np.uint8), np.zeros(1, np.int64)
    if not (pa.types.is_string(arr.type) or pa.types.is_large_string(arr.type)):
        if pa.types.is_string_view(arr.type):
            arr = arr.cast(pa.large_string())
        else:
            raise TypeError("input does not consist of strings only (arrow type %s)" % arr.type)
    if arr.null_count:
        raise TypeError("input contains missing values")
    n = len(arr)
    _, off_buf, data_buf = arr.buffers()
    odt = np.int64 if pa.types.is_large_string(arr.type) else np.int32
    offsets = np.frombuffer(off_buf, dtype=odt, count=n + 1, offset=arr.offset * np.dtype(odt).itemsize)
    offsets = offsets.astype(np.int64, copy=False)
    data = np.frombuffer(data_buf, dtype=np.uint8) if data_buf is not None else np.zeros(0, np.uint8)
    lo, hi = int(offsets[0]), int(offsets[-1])
    return data[lo:hi], (offsets - lo) if lo else offsets


def _encode_list(strings):
    arr = pa.array(strings, type=pa.large_string())
    _, off_buf, data_buf = arr.buffers()
    offsets = np.frombuffer(off_buf, dtype=np.int64, count=len(strings) + 1)
    data = np.frombuffer(data_buf, dtype=np.uint8) if data_buf is not None else np.zeros(0, np.uint8)
    return data[:int(offsets[-1])], offsets


def pack_strings(series_list, regex=DEFAULT_REGEX, ignore_case=True, normalize_to_ascii=True):
    """Concatenate the Series (master, then duplicates) into one ASCII byte buffer.

    Returns (data uint8, offsets int64 [n_total+1], flags for the device analyzer, pristine) where `pristine`
    says that the bytes are the callers' strings verbatim (no host normalisation happened), so the device copy can
    also serve the string gather of get_matches.
    """
    default_regex = (regex == DEFAULT_REGEX)
    flags = 0
    if default_regex:
        flags |= SG_FLAG_STRIP_DEFAULT
        if ignore_case:
            flags |= SG_FLAG_IGNORE_CASE
    datas, offs = [], []
    base = 0
    pristine = default_regex
    for s in series_list:
        data, offsets = _arrow_buffers(s)
        n = len(offsets) - 1
        mx = None               # largest byte of `data`; None = not known (after a re-encode)
        if not default_regex:
            # host runs the whole analyzer prefix (string_grouper.py:372-376) with the user's pattern
            pat = re.compile(regex)
            strings = s.tolist()
            out = []
            for x in strings:
                if ignore_case:
                    x = x.lower()
                if normalize_to_ascii:
                    x = normalize('NFKD', x).encode('ASCII', 'ignore').decode()
                out.append(pat.sub('', x))
            data, offsets = _encode_list(out)
        elif (mx := int(data.max()) if data.size else 0) >= 0x80:
            bad = np.unique(np.searchsorted(offsets, np.nonzero(data >= 0x80)[0], side='right') - 1)
            strings = s.tolist()
            for i in bad.tolist():
                x = strings[i]
                if ignore_case:
                    x = x.lower()
                if normalize_to_ascii:
                    x = normalize('NFKD', x).encode('ASCII', 'ignore').decode()
                strings[i] = x
            data, offsets = _encode_list(strings)
            pristine = False
            mx = None
        # one scan of the bytes in the common all-ASCII case: the maximum is only taken again after a re-encode
        if mx is None:
            mx = int(data.max()) if data.size else 0
        if mx >= 0x80:
            raise NotImplementedError(
                "normalize_to_ascii=False with non-ASCII characters is not supported by the device vectoriser "
                "(n-gram keys pack 7-bit characters); see DESIGN.md 'out of scope'")
        datas.append(data)
        offs.append(offsets[:-1] + base if n else np.zeros(0, np.int64))
        base += int(offsets[-1])
    offs.append(np.array([base], dtype=np.int64))
    data = np.concatenate(datas) if len(datas) > 1 else np.ascontiguousarray(datas[0])
    offsets = np.concatenate(offs)
    return data, offsets, flags, pristine


def decode_vocab_keys(keys, ngram):
    """Packed 7-bit keys -> the n-gram strings (the sorted vocabulary of the fitted vectoriser)."""
    keys = np.asarray(keys, dtype=np.uint32)
    chars = [((keys >> (7 * (ngram - 1 - q))) & 0x7f).astype(np.uint8) for q in range(ngram)]
    mat = np.stack(chars, axis=1)
    return [bytes(row).decode('ascii') for row in mat]

"""Seeded "sec__edgar-shaped" synthetic company-name corpus.

Bench / test input generator — no network, no datasets.  Statistics aimed at
(SURVEY.md Appendix B): ~23 chars per name, ~18 distinct trigrams per name,
heavy legal-suffix trigrams (INC / LLC / CORP ...) in 10-25 % of rows, ~38 %
near-duplicate rows and one mega-cluster of N/400 rows that overflows top-n.
"""
import hashlib

import numpy as np

_LETTERS = np.array(list("abcdefghijklmnopqrstuvwxyz"))
# English unigram frequencies (per mille, a..z)
_FREQ = np.array([82, 15, 28, 43, 127, 22, 20, 61, 70, 2, 8, 40, 24, 67, 75, 19, 1, 60, 63, 91,
                  28, 10, 24, 2, 20, 1], dtype=np.float64)
_SUFFIXES = ["INC", "LLC", "CORP", "LTD", "LP", "TRUST", "FUND", "CO", "HOLDINGS", "GROUP",
             "PARTNERS", "CAPITAL", "BANCORP", "PLC", "SA", "NV"]
_SUFFIX_P = np.array([.28, .20, .12, .08, .07, .05, .04, .04, .03, .03, .02, .01, .01, .01, .005, .005])


def make_names(n, seed=0, n_words=50_000):
    """Return a list of `n` upper-case company-like names."""
    rng = np.random.default_rng(seed)
    # vocabulary of pseudo-words, Zipf-ranked
    wl = rng.integers(3, 11, size=n_words)
    chars = rng.choice(_LETTERS, size=int(wl.sum()), p=_FREQ / _FREQ.sum())
    ends = np.cumsum(wl)
    flat = "".join(chars.tolist()).upper()
    words = [flat[e - l:e] for e, l in zip(ends.tolist(), wl.tolist())]
    zipf_p = 1.0 / np.arange(1, n_words + 1) ** 1.05
    zipf_p /= zipf_p.sum()

    n_mega = max(1, n // 400)
    n_rest = n - n_mega
    n_base = int(round(0.62 * n_rest))
    n_copy = n_rest - n_base

    nw = rng.choice([1, 2, 3, 4], size=n_base, p=[.15, .40, .30, .15])
    wid = rng.choice(n_words, size=int(nw.sum()), p=zipf_p)
    has_suf = rng.random(n_base) < 0.85
    suf = rng.choice(len(_SUFFIXES), size=n_base, p=_SUFFIX_P / _SUFFIX_P.sum())
    punct = rng.random(n_base)
    has_num = rng.random(n_base) < 0.08
    num = rng.integers(1, 1000, size=n_base)
    base = []
    pos = 0
    wid_l, nw_l = wid.tolist(), nw.tolist()
    for i in range(n_base):
        k = nw_l[i]
        s = " ".join([words[w] for w in wid_l[pos:pos + k]])
        pos += k
        if has_suf[i]:
            sfx = _SUFFIXES[suf[i]]
            if punct[i] < 0.15:
                s = s + ", " + sfx + "."
            elif punct[i] < 0.30:
                s = s + " " + sfx + "."
            else:
                s = s + " " + sfx
        if has_num[i]:
            s = s + " " + str(num[i])
        base.append(s)

    src = rng.integers(0, n_base, size=n_copy)
    exact = rng.random(n_copy) < 0.13
    kind = rng.integers(0, 5, size=n_copy)
    where = rng.random(n_copy)
    letter = rng.integers(0, 26, size=n_copy)
    copies = []
    for i in range(n_copy):
        s = base[src[i]]
        if not exact[i]:
            k = kind[i]
            p = int(where[i] * len(s))
            c = chr(65 + int(letter[i]))
            if k == 0:
                s = s[:p] + c + s[p + 1:]
            elif k == 1:
                s = s[:p] + s[p + 1:]
            elif k == 2:
                s = s[:p] + c + s[p:]
            elif k == 3:
                s = s.replace(" ", "-", 1)
            else:
                s = s + " /TA"
        copies.append(s)

    mega = ["ADVISORS DISCIPLINED TRUST %d" % v for v in rng.integers(1, 2000, size=n_mega)]
    names = np.array(base + copies + mega, dtype=object)
    names = names[rng.permutation(n)]
    return names.tolist()


def corpus_sha256(names):
    return hashlib.sha256("\n".join(names).encode()).hexdigest()

"""Brute-force CTC / Gram-CTC likelihood by enumerating every frame labelling (tiny cases only).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Independent of any lattice: it enumerates all
V^T frame-level id sequences, collapses each the way the reference decoder does
(``asr/error.py:39-47``: merge repeats unless separated by a blank, drop blanks), expands gram ids
to characters, and sums the probability of those whose expansion equals the target.  This is the
known-answer test the reference itself lacks.
"""
import itertools

import numpy as np


def collapse(frame_ids, blank):
    out = []
    prev = None
    for k in frame_ids:
        if k != prev and k != blank:
            out.append(k)
        prev = k
    return out


def log_likelihood(logp_tv, target_chars, gram_to_chars, blank=0):
    """log sum over frame labellings whose collapsed+expanded string equals ``target_chars``.

    gram_to_chars: dict id -> tuple of characters (1 char for a unigram id, 2 for a bigram id); ids
    missing from the dict (other than blank) can never be part of a valid labelling.
    """
    logp = np.asarray(logp_tv, np.float64)
    T, V = logp.shape
    target = tuple(target_chars)
    total = -np.inf
    usable = [k for k in range(V) if k == blank or k in gram_to_chars]
    for ids in itertools.product(usable, repeat=T):
        chars = []
        for k in collapse(ids, blank):
            chars.extend(gram_to_chars[k])
        if tuple(chars) == target:
            total = np.logaddexp(total, float(sum(logp[t, k] for t, k in enumerate(ids))))
    return total


def ctc_log_likelihood(logp_tv, labels, blank=0):
    V = np.asarray(logp_tv).shape[1]
    return log_likelihood(logp_tv, labels, {k: (k,) for k in range(V) if k != blank}, blank)

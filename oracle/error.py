"""CPU restatement of the reference's evaluation-path error computation.  TEST INFRASTRUCTURE ONLY
(see oracle/__init__.py): imported by tests/ only, never by the product path.

Follows /root/reference/asr/error.py:
  compute_character_error_rate  :7-24   (Levenshtein table in numpy.uint8, i.e. modulo 256, :10)
  compute_minibatch_error       :26-68  (drop blanks from the target :33-37; collapse with a prev_token state
                                         machine :38-47; ids -> string -> convert_sentence_to_unigram_ids :49-53;
                                         mean over the batch :55,:68)
Pinned against the reference itself (imported in the build container through oracle/ref_stub.load_error_module)
by tests/test_oracle_error.py and the committed fixtures tests/golden/cer_*.npz.
"""
import numpy as np


def edit_distance(r, h, uint8_wrap=False):
    """d[len(r)][len(h)] of asr/error.py:10-23; with uint8_wrap every stored value and every `+ 1` is taken
    modulo 256, which is what the reference's numpy.uint8 table does."""
    mask = 0xff if uint8_wrap else (1 << 62) - 1
    R, H = len(r), len(h)
    prev = [j & mask for j in range(H + 1)]                    # :13
    for i in range(1, R + 1):
        cur = [i & mask] + [0] * H                             # :14
        for j in range(1, H + 1):
            if r[i - 1] == h[j - 1]:
                cur[j] = prev[j - 1]                           # :17-18
            else:
                cur[j] = min((prev[j - 1] + 1) & mask, (cur[j - 1] + 1) & mask, (prev[j] + 1) & mask)   # :20-23
        prev = cur
    return prev[H]


def character_error_rate(r, h, uint8_wrap=False):
    if len(r) == 0:                                            # :8-9
        return len(h)
    return float(edit_distance(r, h, uint8_wrap)) / len(r)     # :24


def collapse(argmax_sequence, blank):
    """asr/error.py:38-47."""
    out, prev = [], blank
    for tok in argmax_sequence:
        tok = int(tok)
        if tok == blank:
            prev = blank
            continue
        if tok == prev:
            continue
        out.append(tok)
        prev = tok
    return out


def minibatch_error(y_batch, t_batch, blank, expansion, uint8_wrap=False, input_length=None):
    """asr/error.py:26-68 with the string round trip (:49-53) given as a table id -> unigram ids (-1 padded).
    Returns (mean, per-utterance errors, hypotheses)."""
    total = 0
    errs, hyps = [], []
    for b, (y, t) in enumerate(zip(y_batch, t_batch)):
        target = [int(v) for v in t if int(v) != blank]        # :33-37
        if input_length is not None:
            y = y[:int(input_length[b])]
        hyp = []
        for tok in collapse(y, blank):
            if 0 <= tok < len(expansion):
                hyp.extend(int(u) for u in expansion[tok] if u >= 0)
        e = character_error_rate(target, hyp, uint8_wrap)
        total += e                                             # :55
        errs.append(float(e))
        hyps.append(hyp)
    return total / len(y_batch), np.asarray(errs, dtype=np.float64), hyps     # :68

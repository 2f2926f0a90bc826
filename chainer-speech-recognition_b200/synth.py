"""Synthetic inputs of the shapes BASELINE.json names.  NumPy only, so the
same arrays feed the CUDA path, the oracle and the CPU baseline.

  activations   N(0,1) float32 drawn as (T,B,V) from numpy.random.RandomState(seed); the
                "trained-like" variant adds +8 on a plausible alignment of the labels, which keeps
                |loss| small so that fp32 parity is meaningful;
  labels        uniform in [1,V) (blank = 0 excluded), lengths uniform in [L/2, L];
  input lengths uniform in [max(2L+1, T/2), T]; utterance 0 forced to full T and L; label padding 0;
  Gram-CTC      unigram ids in [1,119), bigram id of a character pair = a fixed function of the pair,
                present with probability 0.6 (else -1, "not in the inventory"), position 0 always -1
                (asr/data/processing.py:139-146).
"""
import numpy as np

N_UNIGRAM = 119          # asr/vocab.py: 118 katakana tokens + blank


def make_lengths(rs, B, T, L, variable=True):
    if not variable:
        return np.full(B, T, np.int32), np.full(B, L, np.int32)
    lab_len = rs.randint(max(1, L // 2), L + 1, size=B).astype(np.int32) if L > 0 else np.zeros(B, np.int32)
    lo = min(T, max(2 * L + 1, T // 2))
    in_len = rs.randint(lo, T + 1, size=B).astype(np.int32)
    lab_len[0] = L
    in_len[0] = T
    return in_len, lab_len


def make_ctc_labels(rs, B, L, V, lab_len, repeat_prob=0.1):
    labels = np.zeros((B, L), np.int32)
    for b in range(B):
        n = int(lab_len[b])
        u = rs.randint(1, V, size=n)
        for i in range(1, n):
            if rs.rand() < repeat_prob:
                u[i] = u[i - 1]
        labels[b, :n] = u
    return labels


def bigram_id(a, b, V, salt=0):
    """Inventory lookup for the character pair (a,b): id in [N_UNIGRAM, V) or -1 (absent, p = 0.4)."""
    h = (int(a) * 1000003 + int(b) * 7919 + salt * 104729) & 0x7fffffff
    h = (h * 2654435761) & 0xffffffff
    if (h >> 8) % 10 >= 6:
        return -1
    return N_UNIGRAM + (h % max(1, V - N_UNIGRAM))


def make_gram_labels(rs, B, L, V, lab_len, repeat_prob=0.1, n_unigram=N_UNIGRAM):
    uni = np.zeros((B, L), np.int32)
    big = np.zeros((B, L), np.int32)
    for b in range(B):
        n = int(lab_len[b])
        u = rs.randint(1, n_unigram, size=n)
        for i in range(1, n):
            if rs.rand() < repeat_prob:
                u[i] = u[i - 1]
        uni[b, :n] = u
        g = np.full(n, -1, np.int64)
        for i in range(1, n):
            if V > n_unigram:
                h = bigram_id(u[i - 1], u[i], V)
                g[i] = h if h < 0 else n_unigram + (h - N_UNIGRAM) % (V - n_unigram)
        big[b, :n] = g
    return uni, big


def add_alignment_bump(x_tbv, labels, in_len, lab_len, blank=0, bump=8.0):
    """'trained-like' activations: +bump on an evenly spread blank/label/blank... alignment."""
    T, B, V = x_tbv.shape
    for b in range(B):
        Tb, n = int(in_len[b]), int(lab_len[b])
        path = [blank]
        for i in range(n):
            path += [int(labels[b, i]), blank]
        S = len(path)
        for t in range(Tb):
            x_tbv[t, b, path[min(S - 1, t * S // max(Tb, 1))]] += bump
    return x_tbv


def ctc_problem(B, T, V, L, seed=0, variable=True, trained=False, dtype=np.float32):
    rs = np.random.RandomState(seed)
    x = rs.standard_normal((T, B, V)).astype(dtype)
    in_len, lab_len = make_lengths(rs, B, T, L, variable)
    labels = make_ctc_labels(rs, B, L, V, lab_len)
    if trained:
        add_alignment_bump(x, labels, in_len, lab_len)
    return {"x": x, "labels": labels, "input_length": in_len, "label_length": lab_len, "blank": 0}


def gram_problem(B, T, V, L, seed=0, variable=True, trained=False, n_unigram=N_UNIGRAM, dtype=np.float32):
    rs = np.random.RandomState(seed)
    x = rs.standard_normal((T, B, V)).astype(dtype)
    in_len, lab_len = make_lengths(rs, B, T, L, variable)
    in_len = np.maximum(in_len, np.minimum(T, 3 * lab_len + 1)).astype(np.int32)
    uni, big = make_gram_labels(rs, B, L, V, lab_len, n_unigram=min(n_unigram, V))
    if trained:
        add_alignment_bump(x, uni, in_len, lab_len)
    return {"x": x, "labels": uni, "bigrams": big, "input_length": in_len, "label_length": lab_len, "blank": 0}

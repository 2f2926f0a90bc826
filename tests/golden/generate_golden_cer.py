"""Generate the character-error-rate golden vectors (tests/golden/cer/*.npz) by running the REFERENCE ITSELF.

    PYTHONPATH=/root/repo python tests/golden/generate_golden_cer.py

Loads /root/reference/asr/error.py and asr/vocab.py unmodified (oracle/ref_stub.load_error_module), builds the
reference's vocabulary (its 118 unigram tokens + seeded bigram tokens, the way load_unigram_and_bigram_ids extends
the unigram table, asr/vocab.py:74-85), runs compute_minibatch_error (asr/error.py:26-68) on seeded greedy index
batches and stores inputs, the id -> unigram-ids table obtained from the reference's own
convert_sentence_to_unigram_ids (asr/vocab.py:99-105), and the reference's outputs.  Only works where the reference
is mounted (the build container); the fixtures are committed.

The reference's distance table is numpy.uint8 (asr/error.py:10): under NumPy 2 it wraps with a RuntimeWarning as
long as the first row/column fit (sequences shorter than 256) and raises OverflowError beyond.  Below 256 ids no
entry can exceed 255, so the reference is exact wherever it runs here; the cases stay below that ("long_250" goes
close to it).  The modulo-256 behaviour of older NumPy is restated in oracle/error.py (uint8_wrap) but cannot be
pinned by running the reference in this container.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_stub  # noqa: E402

CASES = [
    # name, B, T, L, n_bigram tokens, p(blank), seed, target like prediction?
    ("unigram_small", 5, 40, 10, 0, 0.5, 1, False),
    ("bigram_small", 6, 60, 12, 300, 0.5, 2, False),
    ("bigram_close", 8, 120, 30, 500, 0.6, 3, True),
    ("empty_targets", 4, 30, 6, 100, 0.7, 4, False),
    ("long_250", 2, 250, 250, 0, 0.0, 5, False),
]


def build_vocab(vocab, n_bigram, rng):
    ids, inv = vocab.get_unigram_ids()
    toks = vocab.UNIGRAM_TOKENS
    while len(ids) < 1 + len(toks) + n_bigram:
        t = toks[rng.randint(len(toks))] + toks[rng.randint(len(toks))]
        if t not in ids:
            ids[t] = len(ids)
            inv[ids[t]] = t
    return ids, inv


def main():
    err, vocab = ref_stub.load_error_module()
    os.makedirs(os.path.join(HERE, "cer"), exist_ok=True)
    for name, B, T, L, nb, pblank, seed, close in CASES:
        rng = np.random.RandomState(seed)
        ids, inv = build_vocab(vocab, nb, rng)
        V = len(ids)
        nuni = 1 + len(vocab.UNIGRAM_TOKENS)
        y = rng.randint(1, V, size=(B, T))
        y[rng.rand(B, T) < pblank] = 0
        rep = rng.rand(B, T) < 0.3                       # repeated frames, as a greedy path has
        for t in range(1, T):
            y[rep[:, t], t] = y[rep[:, t], t - 1]
        t_batch = rng.randint(1, nuni, size=(B, L))
        for b in range(B):
            t_batch[b, rng.randint(L // 2, L + 1):] = 0  # blank padding (asr/data/processing.py:125-126)
        if name == "empty_targets":
            t_batch[1, :] = 0
            y[2, :] = 0
        if close:                                        # targets = the reference's own decoding of a noisy copy
            exp = [vocab.convert_sentence_to_unigram_ids(inv[i], ids) if i else [] for i in range(V)]
            for b in range(B):
                hyp, prev = [], 0
                for tok in y[b]:
                    if tok == 0:
                        prev = 0
                        continue
                    if tok == prev:
                        continue
                    hyp.extend(exp[tok])
                    prev = tok
                hyp = [h for h in hyp if rng.rand() > 0.1][:L]
                t_batch[b, :] = 0
                t_batch[b, :len(hyp)] = hyp
        table = [vocab.convert_sentence_to_unigram_ids(inv[i], ids) for i in range(V)]
        E = max(len(r) for r in table)
        expansion = np.full((V, E), -1, np.int32)
        for i, r in enumerate(table):
            expansion[i, :len(r)] = r
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")              # uint8 overflow warnings are the behaviour being recorded
            ref_mean = err.compute_minibatch_error(y, t_batch, 0, ids, inv)
            ref_each = np.array([err.compute_minibatch_error(y[b:b + 1], t_batch[b:b + 1], 0, ids, inv) for b in range(B)])
        np.savez_compressed(os.path.join(HERE, "cer", name + ".npz"), y=y.astype(np.int64), t=t_batch.astype(np.int32),
                            expansion=expansion, blank=0, ref_mean=np.float64(ref_mean), ref_each=ref_each.astype(np.float64))
        print(name, "V", V, "E", E, "mean", ref_mean)


if __name__ == "__main__":
    main()

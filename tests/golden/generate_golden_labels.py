"""Golden vectors for the label half of minibatch assembly (tests/golden/labels/*.npz), produced by the REFERENCE ITSELF:
/root/reference/asr/data/processing.py, Processor.features_to_minibatch (:113-171), loaded unmodified through
oracle/ref_stub.load_processing_module and fed dummy (all-zero) acoustic features of the wanted lengths.

    PYTHONPATH=/root/repo python tests/golden/generate_golden_labels.py

Stored: the tokenised transcriptions (flat token array + offsets, tokenised by the reference's
convert_sentence_to_unigram_tokens), the token inventory in id order, the feature lengths, and the reference's
t_batch / bigram_batch / t_length_batch.  Only works where the reference is mounted; the fixtures are committed.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_stub  # noqa: E402

CASES = [("plain", 8, 1, 14, 0.2, 40, 1), ("short_features", 10, 1, 16, 0.4, 12, 2), ("wide_inventory", 6, 5, 30, 0.1, 80, 3)]


def main():
    proc = ref_stub.load_processing_module()
    _, vocab = ref_stub.load_error_module()
    os.makedirs(os.path.join(HERE, "labels"), exist_ok=True)
    toks = vocab.UNIGRAM_TOKENS
    for name, B, lmin, lmax, prepeat, xmax, seed in CASES:
        rng = np.random.RandomState(seed)
        ids, _ = vocab.get_unigram_ids()
        sentences = []
        for b in range(B):
            n = int(rng.randint(lmin, lmax + 1))
            s = [toks[rng.randint(len(toks))] for _ in range(n)]
            for i in range(1, n):
                if rng.rand() < prepeat:
                    s[i] = s[i - 1]
            sentences.append("".join(s))
        tokenised = [vocab.convert_sentence_to_unigram_tokens(s) for s in sentences]
        for tk in tokenised:                                  # about half of the bigrams that occur are "in the inventory"
            for a, c in zip(tk[:-1], tk[1:]):
                if rng.rand() < 0.5 and a + c not in ids:
                    ids[a + c] = len(ids)
        x_len = [int(rng.randint(1, xmax + 1)) for _ in range(B)]
        feats = [(np.zeros((40, x), np.float32), None, None) for x in x_len]
        Lmax = max(len(tk) for tk in tokenised)
        P = proc.Processor(using_delta=False, using_delta_delta=False)
        _, xlb, tb, tlb, bb = P.features_to_minibatch(feats, sentences, max(x_len), Lmax, ids, 0)
        inventory = np.array([t for t, _ in sorted(ids.items(), key=lambda kv: kv[1])])
        flat = np.array([t for tk in tokenised for t in tk])
        offs = np.cumsum([0] + [len(tk) for tk in tokenised]).astype(np.int64)
        np.savez_compressed(os.path.join(HERE, "labels", name + ".npz"), tokens=flat, offsets=offs, inventory=inventory,
                            x_length=np.asarray(x_len, np.int32), Lmax=Lmax, ref_t=tb.astype(np.int32),
                            ref_bigram=bb.astype(np.int32), ref_t_length=np.asarray(tlb, np.int32),
                            ref_x_length=np.asarray(xlb, np.int32))
        print(name, "B", B, "Lmax", Lmax, "inventory", len(ids), "bigram hits", int((bb > 0).sum()), "cut", int(sum(len(tk) for tk in tokenised) - sum(tlb)))


if __name__ == "__main__":
    main()

"""Generate the golden vectors in this directory by running the REFERENCE ITSELF.

    PYTHONPATH=/root/repo python tests/golden/generate_golden.py

Runs /root/reference/asr/loss/gram_ctc.py unmodified (NumPy path, under the Chainer stub in
oracle/ref_stub.py) on small seeded problems and stores inputs + the reference's outputs as .npz.
Only works where the reference is mounted (the build container); the fixtures are committed so the
tests need neither the reference nor this script.  Plain-CTC fixtures use the same file with every
bigram id = -1 (asr/loss/gram_ctc.py:94-98), see SURVEY.md section 8c.
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_stub  # noqa: E402

synth = importlib.import_module("chainer-speech-recognition_b200.synth")

CASES = [
    # name, kind, B, T, V, L, seed, trained
    ("ctc_tiny", "ctc", 3, 12, 9, 3, 1, False),
    ("ctc_small", "ctc", 4, 40, 30, 8, 2, False),
    ("ctc_small_trained", "ctc", 4, 40, 30, 8, 3, True),
    ("ctc_wide", "ctc", 2, 30, 200, 6, 4, True),
    ("gram_tiny", "gram", 3, 12, 9, 3, 5, False),
    ("gram_small", "gram", 4, 40, 30, 8, 6, False),
    ("gram_small_trained", "gram", 4, 40, 30, 8, 7, True),
    ("gram_wide", "gram", 2, 45, 300, 10, 8, True),
]


def main():
    for name, kind, B, T, V, L, seed, trained in CASES:
        if kind == "ctc":
            prob = synth.ctc_problem(B, T, V, L, seed=seed, trained=trained)
            big = np.full_like(prob["labels"], -1)
        else:
            prob = synth.gram_problem(B, T, V, L, seed=seed, trained=trained, n_unigram=max(3, min(119, V // 3)))
            big = prob["bigrams"]
        loss, grad, _ = ref_stub.run_gram_ctc(prob["x"], prob["labels"], big, prob["input_length"],
                                              prob["label_length"], blank=0, reduce="no")
        loss_mean, grad_mean, _ = ref_stub.run_gram_ctc(prob["x"], prob["labels"], big, prob["input_length"],
                                                        prob["label_length"], blank=0, reduce="mean")
        out = os.path.join(HERE, name + ".npz")
        np.savez_compressed(out, kind=kind, x=prob["x"], labels=prob["labels"], bigrams=big,
                            input_length=prob["input_length"], label_length=prob["label_length"], blank=0,
                            ref_loss=np.asarray(loss, np.float32), ref_grad=grad.astype(np.float32),
                            ref_loss_mean=np.float32(loss_mean), ref_grad_mean=grad_mean.astype(np.float32))
        print("%-20s loss %s" % (name, np.round(loss, 4)))


if __name__ == "__main__":
    main()

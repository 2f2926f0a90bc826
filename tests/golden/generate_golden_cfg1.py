"""Golden vector for BASELINE configs[0] ("CTC loss fwd+bwd on CPU reference path: B=8, T=200, V=3500, L=40"),
produced by running the REFERENCE ITSELF (/root/reference/asr/loss/gram_ctc.py, unmodified, NumPy path, bigram ids
all -1 = plain CTC) under oracle/ref_stub.py.

    PYTHONPATH=/root/repo python tests/golden/generate_golden_cfg1.py

The inputs are regenerated from (shape, seed) by chainer-speech-recognition_b200/synth.py, so only the outputs are
stored: the 8 losses and a sample of the (T,B,V) gradient -- every 997th element of the flattened array plus every
element in a label or blank column of frames 0, 50, 100, 150 -- which keeps the fixture at a few tens of KB.
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


def main():
    B, T, V, L, seed, trained = 8, 200, 3500, 40, 0, True
    prob = synth.ctc_problem(B, T, V, L, seed=seed, trained=trained)
    big = np.full_like(prob["labels"], -1)
    loss, grad, _ = ref_stub.run_gram_ctc(prob["x"], prob["labels"], big, prob["input_length"], prob["label_length"],
                                          blank=0, reduce="no")
    idx = list(range(0, T * B * V, 997))
    for t in (0, 50, 100, 150):
        for b in range(B):
            cols = [0] + [int(c) for c in prob["labels"][b, :prob["label_length"][b]]]
            idx += [(t * B + b) * V + c for c in cols]
    idx = np.unique(np.asarray(idx, np.int64))
    np.savez_compressed(os.path.join(HERE, "full", "cfg1_reference.npz"), shape_seed=np.asarray([B, T, V, L, seed], np.int64),
                        trained=np.bool_(trained), ref_loss=np.asarray(loss, np.float32), sample_index=idx,
                        ref_grad_sample=grad.reshape(-1)[idx].astype(np.float32))
    print("cfg1 reference loss", np.round(loss, 4), "samples", idx.size)


if __name__ == "__main__":
    main()

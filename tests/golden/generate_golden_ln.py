"""Golden vectors for the LayerNormalization -> loss path, produced by running the REFERENCE
ITSELF: asr/nn/layernorm.py (NormalizeLayer forward/backward) and asr/loss/gram_ctc.py, both unmodified, on the NumPy
path under oracle/ref_stub.py; scale/bias by gamma/beta and the swapaxes/reshape/split of asr/model/cnn.py:41-44 are
the reference's own one-liners applied in between (oracle/ref_stub.py: run_layernorm_ctc).

    PYTHONPATH=/root/repo python tests/golden/generate_golden_ln.py

Inputs: z = the model's last convolution output (B, V, 1, T) float32, gamma / beta (V,), labels.  Outputs of the
reference: per-utterance loss, dz (B, V, 1, T), dgamma, dbeta.
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
    ("ln_ctc_small", "ctc", 3, 24, 40, 5, 21, False),
    ("ln_ctc_trained", "ctc", 4, 44, 64, 8, 22, True),        # T % 8 != 0
    ("ln_ctc_wide", "ctc", 2, 32, 300, 6, 23, True),
    ("ln_gram_small", "gram", 3, 24, 40, 5, 24, False),
    ("ln_gram_trained", "gram", 3, 48, 90, 9, 25, True),
]


def make_case(kind, B, T, V, L, seed, trained):
    """z ~ N(0.3, 1.7^2) (a convolution output is neither centred nor unit-variance), optionally +6 on a plausible
    alignment; gamma ~ 1 +- 0.2, beta ~ +-0.1; labels and lengths as in synth."""
    rs = np.random.RandomState(seed)
    if kind == "ctc":
        prob = synth.ctc_problem(B, T, V, L, seed=seed)
        big = np.full_like(prob["labels"], -1)
    else:
        prob = synth.gram_problem(B, T, V, L, seed=seed, n_unigram=max(3, min(119, V // 3)))
        big = prob["bigrams"]
    z_tbv = (rs.randn(T, B, V) * 1.7 + 0.3).astype(np.float32)
    if trained:
        synth.add_alignment_bump(z_tbv, prob["labels"], prob["input_length"], prob["label_length"], bump=6.0)
    z = np.ascontiguousarray(z_tbv.transpose(1, 2, 0))[:, :, None, :]            # (B, V, 1, T)
    gamma = (1.0 + 0.2 * rs.randn(V)).astype(np.float32)
    beta = (0.1 * rs.randn(V)).astype(np.float32)
    return prob, big, z, gamma, beta


def main():
    os.makedirs(os.path.join(HERE, "ln"), exist_ok=True)
    for name, kind, B, T, V, L, seed, trained in CASES:
        prob, big, z, gamma, beta = make_case(kind, B, T, V, L, seed, trained)
        loss, dz, dg, db, _ = ref_stub.run_layernorm_ctc(z, gamma, beta, prob["labels"], big, prob["input_length"],
                                                         prob["label_length"], blank=0, reduce="no")
        lm, dzm, dgm, dbm, _ = ref_stub.run_layernorm_ctc(z, gamma, beta, prob["labels"], big, prob["input_length"],
                                                          prob["label_length"], blank=0, reduce="mean")
        np.savez_compressed(os.path.join(HERE, "ln", name + ".npz"), kind=kind, z=z, gamma=gamma, beta=beta,
                            labels=prob["labels"], bigrams=big, input_length=prob["input_length"],
                            label_length=prob["label_length"], blank=0, ref_loss=np.asarray(loss, np.float32),
                            ref_dz=dz, ref_dgamma=dg, ref_dbeta=db, ref_loss_mean=np.float32(lm), ref_dz_mean=dzm,
                            ref_dgamma_mean=dgm, ref_dbeta_mean=dbm)
        print("%-18s loss %s  |dz| %.3f" % (name, np.round(loss, 3), np.abs(dz).max()))


if __name__ == "__main__":
    main()

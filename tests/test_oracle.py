"""CPU: the oracle (NumPy + C restatements) against the committed golden vectors, which are outputs of
the reference itself (tests/golden/generate_golden.py), against a brute-force path enumeration, and --
for plain CTC -- against torch's CPU implementation in float64.

Tolerances.  The golden gradients carry the reference's own float32 noise (it evaluates log-probabilities
of magnitude |loss| in float32): losses are compared to 1e-5 relative, gradients to
5e-5 absolute on random logits and 1e-5 on the "trained-like" fixtures where |loss| is small.
"""
import glob
import os

import numpy as np
import pytest

from oracle import lattice, c_oracle, brute_force

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def load(path):
    z = np.load(path)
    return {k: z[k] for k in z.files}


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_oracle_matches_reference_golden(path, impl):
    g = load(path)
    kind = str(g["kind"])
    if impl == "numpy":
        loss, grad = lattice.gram_ctc(g["x"], g["labels"], g["bigrams"], g["input_length"], g["label_length"], 0)
    else:
        # plain-CTC fixtures were produced with all bigrams = -1; run them through the CTC lattice proper
        r = c_oracle.run(0 if kind == "ctc" else 1, g["x"], g["labels"], g["bigrams"], g["input_length"],
                         g["label_length"], 0)
        loss, grad = r["loss"], r["grad"]
    assert np.allclose(loss, g["ref_loss"], rtol=1e-5, atol=1e-5)
    tol = 1e-5 if "trained" in path or "wide" in path else 5e-5
    assert np.abs(grad - g["ref_grad"]).max() <= tol
    # reduce='mean' = mean over the batch, gradient / B (gram_ctc.py:281,292)
    B = g["x"].shape[1]
    assert np.isclose(loss.mean(), g["ref_loss_mean"], rtol=1e-5)
    assert np.abs(grad / B - g["ref_grad_mean"]).max() <= tol


def test_ctc_is_gram_ctc_with_dead_bigrams():
    g = load([p for p in GOLDEN if p.endswith("ctc_small.npz")][0])
    a = lattice.ctc(g["x"], g["labels"], g["input_length"], g["label_length"], 0)
    b = lattice.gram_ctc(g["x"], g["labels"], np.full_like(g["labels"], -1), g["input_length"], g["label_length"], 0)
    assert np.allclose(a[0], b[0], rtol=1e-12) and np.allclose(a[1], b[1], atol=1e-12)


INVENTORY = {1: (1,), 2: (2,), 3: (3,), 4: (1, 2), 5: (2, 3), 6: (1, 1), 7: (2, 1)}
PAIR = {v: k for k, v in INVENTORY.items() if len(v) == 2}


@pytest.mark.parametrize("target", [(1, 2), (1, 2, 3), (1, 1), (1, 1, 1), (2, 1, 2), (1, 1, 1, 1), (1, 2, 1, 2), (3,), ()])
def test_brute_force_enumeration(target):
    """Known-answer test independent of any lattice: all V^T labellings, T=5, V=8."""
    rs = np.random.RandomState(len(target) * 7 + 1)
    T, V = 5, 8
    x = rs.randn(T, 1, V).astype(np.float32)
    L = len(target)
    uni = np.array([list(target)], np.int32).reshape(1, L)
    big = np.full((1, L), -1, np.int32)
    for i in range(1, L):
        big[0, i] = PAIR.get((target[i - 1], target[i]), -1)
    logp = lattice.log_softmax(x[:, 0])
    bf_gram = brute_force.log_likelihood(logp, target, INVENTORY, 0)
    bf_ctc = brute_force.ctc_log_likelihood(logp, target, 0)
    for impl in ("numpy", "c"):
        if impl == "numpy":
            lg = lattice.gram_ctc(x, uni, big, [T], [L], 0)[0][0]
            lc = lattice.ctc(x, uni, [T], [L], 0)[0][0]
        else:
            lg = c_oracle.run(1, x, uni, big, [T], [L], 0)["loss"][0]
            lc = c_oracle.run(0, x, uni, None, [T], [L], 0)["loss"][0]
        assert np.isclose(lg, -bf_gram, rtol=1e-10)
        if np.isfinite(bf_ctc):
            assert np.isclose(lc, -bf_ctc, rtol=1e-10)
        else:
            assert lc == 1e10                       # infeasible: what the reference returns


def test_ctc_oracle_matches_torch_float64():
    torch = pytest.importorskip("torch")
    import importlib
    synth = importlib.import_module("chainer-speech-recognition_b200.synth")
    prob = synth.ctc_problem(5, 60, 37, 9, seed=11)
    r = c_oracle.run(0, prob["x"], prob["labels"], None, prob["input_length"], prob["label_length"], 0)
    x = torch.tensor(prob["x"], dtype=torch.float64, requires_grad=True)
    lp = torch.log_softmax(x, dim=2)
    tl = torch.nn.functional.ctc_loss(lp, torch.tensor(prob["labels"], dtype=torch.long),
                                      torch.tensor(prob["input_length"], dtype=torch.long),
                                      torch.tensor(prob["label_length"], dtype=torch.long), blank=0, reduction="none")
    tl.sum().backward()
    assert np.allclose(r["loss"], tl.detach().numpy(), rtol=1e-10)
    assert np.abs(r["grad"] - x.grad.numpy()).max() <= 1e-6


def test_gradient_is_derivative_of_loss():
    """Central differences of the float64 oracle loss (Gram-CTC, dead nodes and repeats included)."""
    import importlib
    synth = importlib.import_module("chainer-speech-recognition_b200.synth")
    prob = synth.gram_problem(1, 9, 12, 3, seed=5, n_unigram=5)
    x = prob["x"].astype(np.float64)
    _, grad = lattice.gram_ctc(x, prob["labels"], prob["bigrams"], prob["input_length"], prob["label_length"], 0)
    rs = np.random.RandomState(0)
    for _ in range(12):
        t, v = rs.randint(0, 9), rs.randint(0, 12)
        e = np.zeros_like(x); e[t, 0, v] = 1e-5
        lp = lattice.gram_ctc(x + e, prob["labels"], prob["bigrams"], prob["input_length"], prob["label_length"], 0)[0][0]
        lm = lattice.gram_ctc(x - e, prob["labels"], prob["bigrams"], prob["input_length"], prob["label_length"], 0)[0][0]
        assert abs((lp - lm) / 2e-5 - grad[t, 0, v]) <= 1e-6


def test_greedy_argmax_semantics():
    y = np.array([[[3.0, 7.0, 7.0, 1.0], [np.nan, 1.0, 2.0, np.nan], [-np.inf] * 4]], np.float32)
    assert lattice.greedy_argmax(y).tolist() == [[1, 0, 0]]
    r = c_oracle.run(0, np.ascontiguousarray(y.transpose(1, 0, 2)), np.zeros((1, 1), np.int32), None, [3], [0], 0,
                     want_grad=False, want_argmax=True)
    assert r["argmax"].tolist() == [[1, 0, 0]]


def test_oracle_matches_reference_on_baseline_config_0():
    """BASELINE configs[0] (B=8, T=200, V=3500, L=40): the reference's own losses and a sample of its gradient
    (tests/golden/generate_golden_cfg1.py) against the C oracle on the regenerated inputs."""
    import importlib
    synth = importlib.import_module("chainer-speech-recognition_b200.synth")
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "full", "cfg1_reference.npz"))
    B, T, V, L, seed = [int(v) for v in z["shape_seed"]]
    prob = synth.ctc_problem(B, T, V, L, seed=seed, trained=bool(z["trained"]))
    r = c_oracle.run(0, prob["x"], prob["labels"], None, prob["input_length"], prob["label_length"], 0)
    assert np.allclose(r["loss"], z["ref_loss"], rtol=1e-5)
    # The reference evaluates alpha+beta-total in float32 at |log-probability| ~ 200 (ulp 1.5e-5), so ITS gradient sits
    # up to ~6.4e-5 from the float64 evaluation of the same algorithm here: 1e-4 is the
    # reference's own noise floor at this size, not slack in the oracle.
    assert np.abs(r["grad"].reshape(-1)[z["sample_index"]] - z["ref_grad_sample"]).max() <= 1e-4

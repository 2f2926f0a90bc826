"""CPU: the oracle of the LayerNormalization -> loss path (oracle/layernorm.py + oracle/lattice.py) against golden
vectors produced by the reference's own NormalizeLayer and GramCTC (tests/golden/generate_golden_ln.py), and its
backward against central differences of its forward in float64."""
import glob
import os

import numpy as np
import pytest

from oracle import layernorm, lattice, c_oracle

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ln", "*.npz")))


def ln_oracle(g, reduce="no"):
    """float64 restatement: LN forward -> loss -> LN backward.  Returns loss (B,), dz (B,V,T), dgamma, dbeta."""
    kind = str(g["kind"])
    acts, saved = layernorm.forward(g["z"][:, :, 0, :], g["gamma"], g["beta"])
    if kind == "ctc":
        loss, grad = lattice.ctc(acts, g["labels"], g["input_length"], g["label_length"], 0)
    else:
        loss, grad = lattice.gram_ctc(acts, g["labels"], g["bigrams"], g["input_length"], g["label_length"], 0)
    if reduce == "mean":
        grad = grad / acts.shape[1]
    dz, dgamma, dbeta = layernorm.backward(grad, saved)
    return loss, dz, dgamma, dbeta


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_layernorm_oracle_matches_reference_golden(path):
    z = np.load(path)
    g = {k: z[k] for k in z.files}
    loss, dz, dgamma, dbeta = ln_oracle(g)
    assert np.allclose(loss, g["ref_loss"], rtol=1e-5, atol=1e-5)
    # The reference works in float32 throughout: its loss gradient carries ~1e-5 of noise on random logits (SURVEY.md
    # 0.5), its NormalizeLayer.backward adds float32 rounding of values of magnitude ~1, and dgamma / dbeta sum B*T
    # such values in float32 (the blank channel's sums reach |150|).  The bounds below are that noise, not slack in
    # the float64 oracle -- whose backward is checked against central differences to 1e-6 further down.
    tol = 2e-5 if "trained" in path or "wide" in path else 5e-5
    assert np.abs(dz - g["ref_dz"][:, :, 0, :]).max() <= tol
    n = g["z"].shape[0] * g["z"].shape[3]
    assert np.all(np.abs(dgamma - g["ref_dgamma"]) <= tol * np.sqrt(n) * 2 + 1e-5 * np.abs(g["ref_dgamma"]))
    assert np.all(np.abs(dbeta - g["ref_dbeta"]) <= tol * np.sqrt(n) * 2 + 1e-5 * np.abs(g["ref_dbeta"]))
    lm, dzm, dgm, dbm = ln_oracle(g, "mean")
    assert np.isclose(lm.mean(), g["ref_loss_mean"], rtol=1e-5)
    assert np.abs(dzm - g["ref_dz_mean"][:, :, 0, :]).max() <= tol


def test_layernorm_backward_is_the_derivative_of_forward():
    rs = np.random.RandomState(0)
    B, V, T = 2, 7, 5
    z = rs.randn(B, V, T) * 1.3 + 0.4
    gamma = 1 + 0.3 * rs.randn(V)
    beta = 0.2 * rs.randn(V)
    w = rs.randn(T, B, V)                                  # a linear functional of the activations

    def f(z_, g_, b_):
        return (layernorm.forward(z_, g_, b_)[0] * w).sum()

    acts, saved = layernorm.forward(z, gamma, beta)
    dz, dgamma, dbeta = layernorm.backward(w, saved)
    eps = 1e-6
    for _ in range(10):
        b, v, t = rs.randint(B), rs.randint(V), rs.randint(T)
        e = np.zeros_like(z); e[b, v, t] = eps
        assert abs((f(z + e, gamma, beta) - f(z - e, gamma, beta)) / (2 * eps) - dz[b, v, t]) <= 1e-6
        e = np.zeros_like(gamma); e[v] = eps
        assert abs((f(z, gamma + e, beta) - f(z, gamma - e, beta)) / (2 * eps) - dgamma[v]) <= 1e-6
        assert abs((f(z, gamma, beta + e) - f(z, gamma, beta - e)) / (2 * eps) - dbeta[v]) <= 1e-6


def test_c_oracle_agrees_on_layernorm_activations():
    """The C oracle (what the -m gpu tests use at full size) on the LN output equals the NumPy lattice."""
    z = np.load(GOLDEN[0])
    g = {k: z[k] for k in z.files}
    acts, _ = layernorm.forward(g["z"][:, :, 0, :], g["gamma"], g["beta"])
    r = c_oracle.run(0, acts.astype(np.float32), g["labels"], None, g["input_length"], g["label_length"], 0)
    loss, grad = lattice.ctc(acts, g["labels"], g["input_length"], g["label_length"], 0)
    assert np.allclose(r["loss"], loss, rtol=1e-6)
    assert np.abs(r["grad"] - grad).max() <= 1e-6

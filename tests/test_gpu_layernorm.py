"""-m gpu: LayerNormalization fused into the loss (csrc/layernorm_loss.cu) against (1) golden vectors produced by the
reference's own NormalizeLayer + GramCTC (tests/golden/generate_golden_ln.py) and (2) the float64 oracle
(oracle/layernorm.py + the C lattice oracle) at sizes up to BASELINE configs[1]."""
import glob
import importlib
import os

import numpy as np
import pytest

from util import LOSS_RTOL, GRAD_ATOL

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ln", "*.npz")))


def synth():
    return importlib.import_module("chainer-speech-recognition_b200.synth")


def run_cuda_ln(pkg, kind, z, gamma, beta, prob, reduce="no", gy=None, z4d=True, pitch=None):
    """z: (B, V, T) numpy.  pitch: allocate rows with this many floats (>= T) and pass the strided view."""
    import torch
    dev = torch.device("cuda:0")
    B, V, T = z.shape
    if pitch is None:
        zt = torch.tensor(z, device=dev)
    else:
        buf = torch.zeros((B, V, pitch), device=dev)
        buf[:, :, :T] = torch.tensor(z, device=dev)
        zt = buf[:, :, :T]
    zin = (zt.unsqueeze(2) if z4d else zt).requires_grad_(True)
    g = torch.tensor(gamma, device=dev, requires_grad=True)
    b = torch.tensor(beta, device=dev, requires_grad=True)
    lab = torch.tensor(prob["labels"], device=dev)
    il = None if prob.get("input_length") is None else torch.tensor(prob["input_length"], device=dev)
    ll = None if prob.get("label_length") is None else torch.tensor(prob["label_length"], device=dev)
    if kind == "ctc":
        out = pkg.layernorm_ctc(zin, g, b, lab, 0, il, ll, reduce=reduce)
    else:
        out = pkg.layernorm_gram_ctc(zin, g, b, lab, torch.tensor(prob["bigrams"], device=dev), 0, il, ll, reduce=reduce)
    up = torch.ones_like(out) if gy is None else torch.tensor(np.asarray(gy, np.float32), device=dev).reshape(out.shape)
    out.backward(up)
    torch.cuda.synchronize()
    dz = zin.grad.detach().cpu().numpy().reshape(B, V, T)
    return out.detach().cpu().numpy().astype(np.float64), dz, g.grad.cpu().numpy(), b.grad.cpu().numpy()


def run_oracle_ln(kind, z, gamma, beta, prob, scale=None):
    """float64: LayerNormalization forward (oracle/layernorm.py) -> C lattice oracle on the float64-exact activations
    rounded to float32 inputs is NOT what we want: keep float64 all the way through the NumPy lattice for small
    cases, the C oracle (float64 arithmetic on float32 inputs) for large ones."""
    from oracle import layernorm, lattice, c_oracle
    acts, saved = layernorm.forward(z, gamma, beta)
    T, B, V = acts.shape
    if B * T * V <= 2e6:
        if kind == "ctc":
            loss, grad = lattice.ctc(acts, prob["labels"], prob["input_length"], prob["label_length"], 0)
        else:
            loss, grad = lattice.gram_ctc(acts, prob["labels"], prob["bigrams"], prob["input_length"], prob["label_length"], 0)
    else:
        r = c_oracle.run(0 if kind == "ctc" else 1, acts.astype(np.float32), prob["labels"], prob.get("bigrams"),
                         prob["input_length"], prob["label_length"], 0)
        loss, grad = r["loss"], r["grad"].astype(np.float64)
    if scale is not None:
        grad = grad * np.asarray(scale, np.float64).reshape(1, -1, 1)
    dz, dgamma, dbeta = layernorm.backward(grad, saved)
    return loss, dz, dgamma, dbeta


def check(got, want, B, T, what):
    loss, dz, dg, db = got
    loss_ref, dz_ref, dg_ref, db_ref = want
    rel = np.abs(loss - loss_ref) / np.maximum(np.abs(loss_ref), 1.0)
    assert rel.max() <= LOSS_RTOL, (what, rel.max())
    assert np.abs(dz - dz_ref).max() <= GRAD_ATOL, (what, np.abs(dz - dz_ref).max())
    # dgamma / dbeta sum B*T per-frame gradients, each good to GRAD_ATOL with independent rounding
    tol = GRAD_ATOL * np.sqrt(B * T) + 1e-5 * np.abs(dg_ref)
    assert np.all(np.abs(dg - dg_ref) <= tol), (what, np.abs(dg - dg_ref).max())
    assert np.all(np.abs(db - db_ref) <= GRAD_ATOL * np.sqrt(B * T) + 1e-5 * np.abs(db_ref)), (what, np.abs(db - db_ref).max())


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_fused_layernorm_matches_reference_golden(pkg, path):
    zf = np.load(path)
    g = {k: zf[k] for k in zf.files}
    kind = str(g["kind"])
    prob = {"labels": g["labels"], "bigrams": g["bigrams"], "input_length": g["input_length"], "label_length": g["label_length"]}
    z = g["z"][:, :, 0, :]
    loss, dz, dg, db = run_cuda_ln(pkg, kind, z, g["gamma"], g["beta"], prob)
    assert np.allclose(loss, g["ref_loss"], rtol=1e-5, atol=1e-5)
    # the goldens carry the reference's own float32 noise (tests/test_oracle_layernorm.py explains the bounds)
    tol = 2e-5 if "trained" in path or "wide" in path else 5e-5
    assert np.abs(dz - g["ref_dz"][:, :, 0, :]).max() <= tol
    n = z.shape[0] * z.shape[2]
    assert np.all(np.abs(dg - g["ref_dgamma"]) <= tol * np.sqrt(n) * 2 + 1e-5 * np.abs(g["ref_dgamma"]))
    assert np.all(np.abs(db - g["ref_dbeta"]) <= tol * np.sqrt(n) * 2 + 1e-5 * np.abs(g["ref_dbeta"]))
    lm, dzm, dgm, dbm = run_cuda_ln(pkg, kind, z, g["gamma"], g["beta"], prob, reduce="mean")
    assert np.isclose(lm, g["ref_loss_mean"], rtol=1e-5)
    assert np.abs(dzm - g["ref_dz_mean"][:, :, 0, :]).max() <= tol
    # and against the float64 oracle at north_star's tolerances
    check((loss, dz, dg, db), run_oracle_ln(kind, z, g["gamma"], g["beta"], prob), z.shape[0], z.shape[2], path)


def make_case(kind, B, T, V, L, seed, trained=True):
    s = synth()
    rs = np.random.RandomState(seed)
    prob = s.ctc_problem(B, T, V, L, seed=seed) if kind == "ctc" else s.gram_problem(B, T, V, L, seed=seed, n_unigram=max(3, min(119, V // 3)))
    z_tbv = (rs.standard_normal((T, B, V)) * 1.7 + 0.3).astype(np.float32)
    if trained:
        s.add_alignment_bump(z_tbv, prob["labels"], prob["input_length"], prob["label_length"], bump=6.0)
    z = np.ascontiguousarray(z_tbv.transpose(1, 2, 0))
    gamma = (1.0 + 0.2 * rs.randn(V)).astype(np.float32)
    beta = (0.1 * rs.randn(V)).astype(np.float32)
    return prob, z, gamma, beta


SHAPES = [
    # kind, B, T, V, L
    ("ctc", 3, 40, 37, 5),          # V below one box, odd V
    ("ctc", 2, 64, 241, 9),         # V = one box + 1 row
    ("ctc", 5, 100, 1213, 12),      # a real vocabulary size (119 unigrams + bigrams), T % 8 != 0
    ("gram", 4, 96, 700, 10),
    ("ctc", 8, 200, 3500, 40),      # BASELINE configs[0]
    ("ctc", 2, 48, 3601, 6),        # 16 boxes
    ("ctc", 2, 48, 4080, 6),        # the largest supported vocabulary (17 boxes)
    ("gram", 3, 120, 2000, 30),
]


@pytest.mark.parametrize("shape", SHAPES)
def test_fused_layernorm_matches_oracle(pkg, shape):
    kind, B, T, V, L = shape
    prob, z, gamma, beta = make_case(kind, B, T, V, L, seed=61)
    check(run_cuda_ln(pkg, kind, z, gamma, beta, prob), run_oracle_ln(kind, z, gamma, beta, prob), B, T, "%r" % (shape,))


def test_fused_layernorm_reduce_mean_upstream_gradient_and_padding(pkg):
    prob, z, gamma, beta = make_case("ctc", 6, 72, 300, 8, seed=62)
    B = 6
    loss_ref, dz_ref, dg_ref, db_ref = run_oracle_ln("ctc", z, gamma, beta, prob, scale=np.full(B, 2.5 / B))
    loss, dz, dg, db = run_cuda_ln(pkg, "ctc", z, gamma, beta, prob, reduce="mean", gy=2.5)
    assert abs(loss - loss_ref.mean()) <= LOSS_RTOL * abs(loss_ref.mean())
    check((loss_ref, dz, dg, db), (loss_ref, dz_ref, dg_ref, db_ref), B, 72, "mean gy=2.5")
    for b in range(B):
        assert not dz[b, :, int(prob["input_length"][b]):].any()           # padded frames: exact zeros
    gy = np.array([1.0, -2.0, 0.5, 3.0, 0.0, 1.5], np.float32)
    want = run_oracle_ln("ctc", z, gamma, beta, prob, scale=gy)
    got = run_cuda_ln(pkg, "ctc", z, gamma, beta, prob, reduce="no", gy=gy)
    assert np.abs(got[1] - want[1]).max() <= 3 * GRAD_ATOL                     # |gy| up to 3


def test_fused_layernorm_layouts(pkg):
    """(B,V,T) and (B,V,1,T) inputs, and rows with a pitch larger than T, give the same bits."""
    prob, z, gamma, beta = make_case("ctc", 3, 40, 130, 6, seed=63)
    a = run_cuda_ln(pkg, "ctc", z, gamma, beta, prob)
    b = run_cuda_ln(pkg, "ctc", z, gamma, beta, prob, z4d=False)
    c = run_cuda_ln(pkg, "ctc", z, gamma, beta, prob, pitch=44)
    for other in (b, c):
        for x, y in zip(a, other):
            assert np.array_equal(x, y)
    # a frame count that is not a multiple of 4 works with a padded row pitch
    prob, z, gamma, beta = make_case("gram", 3, 42, 130, 6, seed=66)
    check(run_cuda_ln(pkg, "gram", z, gamma, beta, prob, pitch=44), run_oracle_ln("gram", z, gamma, beta, prob), 3, 42, "T=42 pitch=44")


def test_fused_layernorm_equals_unfused_composition(pkg):
    """The same numbers as LayerNormalization written with torch ops + the transposed copy + the plain loss."""
    import torch
    prob, z, gamma, beta = make_case("ctc", 4, 80, 500, 9, seed=64)
    fused = run_cuda_ln(pkg, "ctc", z, gamma, beta, prob)
    dev = torch.device("cuda:0")
    zt = torch.tensor(z, device=dev, requires_grad=True)
    g = torch.tensor(gamma, device=dev, requires_grad=True)
    b = torch.tensor(beta, device=dev, requires_grad=True)
    mean = zt.mean(dim=1, keepdim=True)
    diff = zt - mean
    std = torch.sqrt((diff * diff).sum(dim=1, keepdim=True) / z.shape[1])
    y = diff / std * g[None, :, None] + b[None, :, None]
    acts = y.permute(2, 0, 1).contiguous()                               # the transposed copy of asr/model/cnn.py:41-44
    loss = pkg.ctc(acts, torch.tensor(prob["labels"], device=dev), 0, torch.tensor(prob["input_length"], device=dev),
                   torch.tensor(prob["label_length"], device=dev), reduce="no")
    loss.sum().backward()
    assert np.allclose(fused[0], loss.detach().cpu().numpy(), rtol=2e-6)
    assert np.abs(fused[1] - zt.grad.cpu().numpy()).max() <= 5e-6
    assert np.allclose(fused[2], g.grad.cpu().numpy(), rtol=1e-4, atol=1e-4)
    assert np.allclose(fused[3], b.grad.cpu().numpy(), rtol=1e-4, atol=1e-4)


def test_fused_layernorm_full_size(pkg):
    """BASELINE configs[1] through the fused path: B=64, T=800, V=3500."""
    prob, z, gamma, beta = make_case("ctc", 64, 800, 3500, 80, seed=65)
    got = run_cuda_ln(pkg, "ctc", z, gamma, beta, prob)
    want = run_oracle_ln("ctc", z, gamma, beta, prob)
    check(got, want, 64, 800, "cfg2 fused")
    for b in range(0, 64, 7):
        assert not got[1][b, :, int(prob["input_length"][b]):].any()


def test_fused_layernorm_unsupported_shapes_are_reported(pkg):
    import torch
    dev = torch.device("cuda:0")
    z = torch.zeros(1, 5000, 1, 16, device=dev)                          # V > 4080
    lab = torch.ones(1, 2, dtype=torch.int32, device=dev)
    with pytest.raises(NotImplementedError):
        pkg.layernorm_ctc(z, torch.ones(5000, device=dev), torch.zeros(5000, device=dev), lab, 0)
    z = torch.zeros(1, 50, 1, 18, device=dev)                            # rows not 16-byte aligned (T = 18)
    with pytest.raises(NotImplementedError):
        pkg.layernorm_ctc(z, torch.ones(50, device=dev), torch.zeros(50, device=dev), lab, 0)

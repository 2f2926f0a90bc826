"""-m gpu: the CUDA path (through the Python shim -> ctypes -> C ABI) against the float64 oracle.

Tolerances are the ones BASELINE.json's north_star states: loss 1e-5 relative, gradient 1e-5
absolute, greedy indices bit-exact.
"""
import numpy as np
import pytest

from util import run_cuda, run_oracle, assert_parity, LOSS_RTOL, GRAD_ATOL

pytestmark = pytest.mark.gpu


def synth():
    import importlib
    return importlib.import_module("chainer-speech-recognition_b200.synth")


CTC_SHAPES = [
    # B, T, V, L
    (3, 12, 9, 3),
    (4, 50, 40, 8),
    (2, 33, 37, 5),        # V not a multiple of 4: unaligned rows
    (5, 64, 130, 17),
    (2, 7, 5, 1),          # Lmax = 1 crashes the reference; must simply work here
    (8, 200, 3500, 40),    # BASELINE configs[0]
    (2, 300, 64, 100),     # N = 201: 4 recursion warps per direction
    (2, 500, 32, 150),     # N = 301
    (1, 900, 16, 250),     # N = 501
    (1, 1000, 16, 380),    # N = 761
]


@pytest.mark.parametrize("shape", CTC_SHAPES)
@pytest.mark.parametrize("trained", [False, True])
def test_ctc_matches_oracle(pkg, shape, trained):
    B, T, V, L = shape
    prob = synth().ctc_problem(B, T, V, L, seed=1, trained=trained)
    loss, grad, _ = run_cuda(pkg, prob, "ctc")
    loss_ref, grad_ref, _ = run_oracle(prob, "ctc")
    assert_parity(loss, grad, loss_ref, grad_ref, "ctc %r" % (shape,))


GRAM_SHAPES = [
    (3, 12, 9, 3),
    (4, 50, 40, 8),
    (2, 33, 37, 5),
    (2, 9, 11, 1),
    (4, 120, 300, 30),
    (2, 400, 200, 100),    # N = 301: three nodes per lane
    (1, 700, 150, 190),    # N = 571
    (1, 900, 150, 250),    # N = 751
    (1, 1200, 150, 380),   # N = 1141
]


@pytest.mark.parametrize("shape", GRAM_SHAPES)
@pytest.mark.parametrize("trained", [False, True])
def test_gram_ctc_matches_oracle(pkg, shape, trained):
    B, T, V, L = shape
    prob = synth().gram_problem(B, T, V, L, seed=2, trained=trained, n_unigram=max(3, min(119, V // 3)))
    loss, grad, _ = run_cuda(pkg, prob, "gram")
    loss_ref, grad_ref, _ = run_oracle(prob, "gram")
    assert_parity(loss, grad, loss_ref, grad_ref, "gram %r" % (shape,))


@pytest.mark.parametrize("shape", GRAM_SHAPES[:6] + [(32, 600, 8000, 60)])
@pytest.mark.parametrize("trained", [False, True])
def test_joint_gram_ctc_plus_ctc_matches_oracle(pkg, shape, trained):
    """joint_ctc=True: loss = gram_ctc + ctc (run/gram_ctc/cnn/train.py:196-198) from one pass; gradient of the sum."""
    B, T, V, L = shape
    if B * T * V > 5e7 and trained:
        pytest.skip("one full-size case is enough")
    prob = synth().gram_problem(B, T, V, L, seed=4, trained=trained, n_unigram=max(3, min(119, V // 3)))
    loss, grad, _ = run_cuda(pkg, prob, "joint")
    loss_ref, grad_ref, _ = run_oracle(prob, "joint")
    # two losses of magnitude |loss|: the absolute gradient bound doubles
    assert np.all(np.abs(loss - loss_ref) <= LOSS_RTOL * np.maximum(np.abs(loss_ref), 1.0)), (shape, loss, loss_ref)
    assert np.abs(grad - grad_ref).max() <= 2 * GRAD_ATOL, (shape, np.abs(grad - grad_ref).max())
    # and it equals the two separate calls of this library
    lg, gg, _ = run_cuda(pkg, prob, "gram")
    lc, gc, _ = run_cuda(pkg, prob, "ctc")
    assert np.allclose(loss, lg + lc, rtol=1e-6, atol=1e-6)
    assert np.abs(grad - (gg + gc)).max() <= 2e-6


def test_joint_reduce_mean_and_padding(pkg):
    prob = synth().gram_problem(6, 60, 90, 9, seed=12, n_unigram=30)
    gy = np.float32(0.75)
    loss, grad, _ = run_cuda(pkg, prob, "joint", reduce="mean", gy=gy)
    loss_ref, grad_ref, _ = run_oracle(prob, "joint")
    assert abs(loss - loss_ref.mean()) <= LOSS_RTOL * abs(loss_ref.mean())
    assert np.abs(grad - grad_ref * gy / 6).max() <= 2 * GRAD_ATOL
    for b in range(6):
        assert np.all(grad[prob["input_length"][b]:, b, :] == 0)


def test_reduce_mean_and_upstream_gradient(pkg):
    prob = synth().ctc_problem(4, 40, 30, 6, seed=3)
    loss_ref, grad_ref, _ = run_oracle(prob, "ctc")
    loss, grad, _ = run_cuda(pkg, prob, "ctc", reduce="mean", gy=2.5)
    assert abs(loss - loss_ref.mean()) <= 1e-5 * abs(loss_ref.mean())
    assert np.abs(grad - grad_ref * (2.5 / 4)).max() <= 1e-5            # gram_ctc.py:292
    gy = np.array([1.0, -2.0, 0.5, 3.0], np.float32)
    loss, grad, _ = run_cuda(pkg, prob, "ctc", reduce="no", gy=gy)
    assert np.abs(grad - grad_ref * gy[None, :, None]).max() <= 3e-5    # gram_ctc.py:294 (|gy| up to 3)


@pytest.mark.parametrize("kind", ["ctc", "gram"])
def test_layouts_agree_bitwise(pkg, kind):
    s = synth()
    prob = s.ctc_problem(3, 30, 24, 5, seed=4) if kind == "ctc" else s.gram_problem(3, 30, 24, 5, seed=4, n_unigram=8)
    base = run_cuda(pkg, prob, kind)
    as_list = run_cuda(pkg, prob, kind, as_list=True)
    btv = run_cuda(pkg, prob, kind, batch_first=True)
    for other in (as_list, btv):
        assert np.array_equal(base[0], other[0])
        assert np.array_equal(base[1], other[1])


def test_default_lengths_are_full(pkg):
    prob = synth().ctc_problem(3, 25, 12, 4, seed=5, variable=False)
    a = run_cuda(pkg, prob, "ctc")
    prob2 = dict(prob, input_length=None, label_length=None)
    b = run_cuda(pkg, prob2, "ctc")
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_padded_frames_get_exact_zero_and_batch_independence(pkg):
    prob = synth().ctc_problem(4, 60, 20, 6, seed=6)
    loss, grad, _ = run_cuda(pkg, prob, "ctc")
    for b in range(4):
        Tb = int(prob["input_length"][b])
        assert not grad[Tb:, b].any()                                   # gram_ctc.py:296
        solo = {k: (v[:, b:b + 1] if k == "x" else (v[b:b + 1] if isinstance(v, np.ndarray) else v))
                for k, v in prob.items()}
        l1, g1, _ = run_cuda(pkg, solo, "ctc")
        assert np.array_equal(l1[0], loss[b]) and np.array_equal(g1[:, 0], grad[:, b])   # bit-identical


def test_infeasible_alignment_and_empty_label(pkg):
    rs = np.random.RandomState(7)
    x = rs.randn(5, 3, 8).astype(np.float32)
    prob = {"x": x, "labels": np.array([[1, 1, 1, 1], [1, 2, 3, 0], [0, 0, 0, 0]], np.int32),
            "input_length": np.array([5, 5, 5], np.int32), "label_length": np.array([4, 3, 0], np.int32), "blank": 0}
    loss, grad, _ = run_cuda(pkg, prob, "ctc")
    loss_ref, grad_ref, _ = run_oracle(prob, "ctc")
    assert loss[0] == 1e10                                              # reference quirk: exactly 1e10
    assert np.isfinite(grad).all()
    assert_parity(loss[1:], grad[:, 1:], loss_ref[1:], grad_ref[:, 1:])
    assert np.abs(grad[:, 0] - grad_ref[:, 0]).max() <= 1e-5


def test_greedy_argmax_bit_exact(pkg):
    import torch
    rs = np.random.RandomState(8)
    y = rs.randn(6, 70, 3500).astype(np.float32)                        # (B,T,V), asr/model/cnn.py:45-47
    y[0, 3, 10] = y[0, 3, 200] = 9.0                                    # tie -> first index
    y[1, 5, :] = -np.inf
    y[2, 7, 100] = np.nan; y[2, 7, 50] = np.nan                         # NaN is maximal, first one wins
    y[3, 9, 3499] = 50.0
    out = pkg.greedy_argmax(torch.tensor(y, device="cuda:0")).cpu().numpy()
    assert out.dtype == np.int64
    assert np.array_equal(out, np.argmax(y, axis=2))
    y2 = rs.randn(3, 11, 37).astype(np.float32)                         # unaligned rows
    out2 = pkg.greedy_argmax(torch.tensor(y2, device="cuda:0")).cpu().numpy()
    assert np.array_equal(out2, np.argmax(y2, axis=2))
    # the loss kernel's fused argmax agrees with the standalone one
    prob = synth().ctc_problem(4, 50, 40, 8, seed=1)
    _, _, am = run_cuda(pkg, prob, "ctc", want_argmax=True)
    assert np.array_equal(am, np.argmax(prob["x"], axis=2).T)


# ---------------------------------------------------------------------------------------------
# golden vectors = outputs of the reference itself (tests/golden/generate_golden.py)
# ---------------------------------------------------------------------------------------------
import glob
import os

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_cuda_matches_reference_golden(pkg, path):
    z = np.load(path)
    g = {k: z[k] for k in z.files}
    kind = str(g["kind"])
    prob = {"x": g["x"], "labels": g["labels"], "bigrams": g["bigrams"], "input_length": g["input_length"],
            "label_length": g["label_length"], "blank": 0}
    loss, grad, _ = run_cuda(pkg, prob, kind)
    assert np.allclose(loss, g["ref_loss"], rtol=1e-5, atol=1e-5)
    # the golden gradients carry the reference's own float32 noise on random logits
    tol = 1e-5 if ("trained" in path or "wide" in path) else 5e-5
    assert np.abs(grad - g["ref_grad"]).max() <= tol
    lm, gm, _ = run_cuda(pkg, prob, kind, reduce="mean")
    assert np.isclose(lm, g["ref_loss_mean"], rtol=1e-5)
    assert np.abs(gm - g["ref_grad_mean"]).max() <= tol


# ---------------------------------------------------------------------------------------------
# BASELINE.json's full sizes: direct comparison with the oracle plus size-independent properties
# ---------------------------------------------------------------------------------------------
def _properties(prob, loss, grad):
    T, B, V = prob["x"].shape
    for b in range(B):
        Tb = int(prob["input_length"][b])
        assert not grad[Tb:, b].any()                                   # exact zeros on padded frames
    # softmax and the merged posteriors both sum to one over the vocabulary => every gradient row sums to ~0
    rows = np.abs(grad.astype(np.float64).sum(axis=2))
    assert rows.max() <= 2e-5, rows.max()
    assert np.isfinite(loss).all() and (loss > 0).all()


@pytest.mark.parametrize("trained", [False, True])
def test_full_size_ctc_config(pkg, trained):
    prob = synth().ctc_problem(64, 800, 3500, 80, seed=0, trained=trained)      # BASELINE configs[1]
    loss, grad, am = run_cuda(pkg, prob, "ctc", want_argmax=True)
    loss_ref, grad_ref, am_ref = run_oracle(prob, "ctc", want_argmax=True)
    assert_parity(loss, grad, loss_ref, grad_ref, "cfg2")
    assert np.array_equal(am, am_ref)                                           # bit-exact greedy indices
    _properties(prob, loss, grad)


def test_full_size_gram_config(pkg):
    prob = synth().gram_problem(32, 600, 8000, 60, seed=0)                       # BASELINE configs[2]
    loss, grad, _ = run_cuda(pkg, prob, "gram")
    loss_ref, grad_ref, _ = run_oracle(prob, "gram")
    assert_parity(loss, grad, loss_ref, grad_ref, "cfg3")
    _properties(prob, loss, grad)


@pytest.mark.parametrize("T,V", [(1600, 100), (3200, 100), (1600, 3500)])
def test_long_utterance_sweep(pkg, T, V):
    prob = synth().ctc_problem(4, T, V, T // 10, seed=1, trained=True)           # BASELINE configs[4] shapes
    loss, grad, _ = run_cuda(pkg, prob, "ctc")
    loss_ref, grad_ref, _ = run_oracle(prob, "ctc")
    assert_parity(loss, grad, loss_ref, grad_ref, "T=%d V=%d" % (T, V))      # north_star's bounds, flat in T
    _properties(prob, loss, grad)


def test_second_backward_over_the_same_graph(pkg):
    """The reference mutates its saved softmax in backward (gram_ctc.py:290-296); here a retained graph can be
    differentiated again and gives the same answer."""
    import torch
    prob = synth().ctc_problem(8, 200, 3500, 40, seed=3)
    x = torch.tensor(prob["x"], device="cuda:0", requires_grad=True)
    loss = pkg.ctc(x, torch.tensor(prob["labels"], device="cuda:0"), 0, torch.tensor(prob["input_length"], device="cuda:0"),
                   torch.tensor(prob["label_length"], device="cuda:0"), reduce="mean")
    loss.backward(retain_graph=True)
    g1 = x.grad.clone(); x.grad = None
    loss.backward()
    assert torch.equal(g1, x.grad)


def test_host_entry_point_matches_device_path(pkg):
    """asr/loss/host.py: host arrays in, host gradient out, pipelined over utterance groups -- same numbers."""
    import torch
    for kind in ("ctc", "gram"):
        s = synth()
        prob = s.ctc_problem(9, 70, 200, 12, seed=5) if kind == "ctc" else s.gram_problem(9, 70, 200, 12, seed=5, n_unigram=40)
        loss_ref, grad_ref, _ = run_oracle(prob, kind)
        xb = np.ascontiguousarray(prob["x"].transpose(1, 0, 2))                      # (B,T,V)
        if kind == "ctc":
            loss, grad = pkg.ctc_host(xb, prob["labels"], 0, prob["input_length"], prob["label_length"], reduce="no", groups=4)
            poisoned = torch.full(xb.shape, float("nan")).pin_memory()           # padded rows must come back as zeros
            lm, gm = pkg.ctc_host(torch.from_numpy(xb).pin_memory(), prob["labels"], 0, prob["input_length"],
                                  prob["label_length"], reduce="mean", groups=3, grad_out=poisoned)
            assert gm is poisoned
        else:
            loss, grad = pkg.gram_ctc_host(xb, prob["labels"], prob["bigrams"], 0, prob["input_length"],
                                           prob["label_length"], reduce="no", groups=4)
            lm, gm = pkg.gram_ctc_host(xb, prob["labels"], prob["bigrams"], 0, prob["input_length"],
                                       prob["label_length"], reduce="mean", groups=2)
        assert_parity(loss, grad.numpy().transpose(1, 0, 2), loss_ref, grad_ref, "host " + kind)
        assert abs(lm - loss_ref.mean()) <= 1e-5 * abs(loss_ref.mean())
        assert np.abs(gm.numpy().transpose(1, 0, 2) - grad_ref / 9).max() <= 1e-5


def test_cuda_array_interface_intake(pkg):
    """Arrays that only speak __cuda_array_interface__ (CuPy arrays, Chainer Variable.data) are taken without a
    copy: the reference's training scripts can hand over what they have."""
    import torch

    class Cai(object):                                   # what a cupy.ndarray looks like to a consumer
        def __init__(self, t):
            self._t = t
            self.__cuda_array_interface__ = t.__cuda_array_interface__

    class Variable(object):                              # chainer.Variable: the array is in .data
        def __init__(self, a):
            self.data = a

    prob = synth().ctc_problem(3, 30, 24, 5, seed=4)
    base = run_cuda(pkg, prob, "ctc")
    dev = torch.device("cuda:0")
    x = torch.tensor(prob["x"], device=dev)
    lab = torch.tensor(prob["labels"], device=dev)
    il = torch.tensor(prob["input_length"], device=dev)
    ll = torch.tensor(prob["label_length"], device=dev)
    for wrap in (lambda t: Cai(t), lambda t: Variable(Cai(t))):
        out = pkg.ctc([wrap(x[t]) for t in range(x.shape[0])], wrap(lab), 0, wrap(il), wrap(ll), reduce="no")
        assert np.array_equal(out.cpu().numpy().astype(np.float64), base[0])
        out = pkg.ctc(wrap(x), Cai(lab), 0, Cai(il), Cai(ll), reduce="no")
        assert np.array_equal(out.cpu().numpy().astype(np.float64), base[0])


def test_label_length_without_input_length_is_ignored_like_the_reference(pkg):
    """gram_ctc.py:310-313: when input_length is None BOTH lengths default to the full widths."""
    prob = synth().ctc_problem(3, 25, 12, 4, seed=5, variable=False)
    a = run_cuda(pkg, prob, "ctc")
    prob2 = dict(prob, input_length=None, label_length=np.array([2, 3, 1], np.int32))
    b = run_cuda(pkg, prob2, "ctc")
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("kind", ["ctc", "gram", "joint"])
def test_large_batch_runs_the_kernels_back_to_back(pkg, kind):
    """B >= #SMs: the lattice kernel is not launched next to the softmax/gather kernel (api.cu) -- same results."""
    B, T, V, L = 160, 40, 64, 6
    s = synth()
    prob = s.ctc_problem(B, T, V, L, seed=31) if kind == "ctc" else s.gram_problem(B, T, V, L, seed=31, n_unigram=20)
    loss, grad, _ = run_cuda(pkg, prob, kind)
    loss_ref, grad_ref, _ = run_oracle(prob, kind)
    assert np.all(np.abs(loss - loss_ref) <= LOSS_RTOL * np.maximum(np.abs(loss_ref), 1.0))
    assert np.abs(grad - grad_ref).max() <= (2 if kind == "joint" else 1) * GRAD_ATOL


def test_concurrent_and_serial_forward_agree_bitwise(pkg, monkeypatch):
    """The lattice kernel next to the softmax/gather kernel (default) vs behind it (B200CTC_NO_CONCURRENT): the same
    arithmetic in the same order per utterance, so identical bits."""
    prob = synth().ctc_problem(24, 300, 512, 30, seed=33)
    a = run_cuda(pkg, prob, "ctc")
    monkeypatch.setenv("B200CTC_NO_CONCURRENT", "1")
    b = run_cuda(pkg, prob, "ctc")
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.gpu
def test_graphed_step_replays_the_eager_step(pkg):
    """GraphedStep (asr/loss/graphed.py): one captured forward + backward, replayed on new data copied into the captured
    tensors, gives the eager call's loss and gradient bit for bit."""
    import importlib
    import torch
    s = importlib.import_module("chainer-speech-recognition_b200.synth")
    dev = torch.device("cuda:0")
    probs = [s.ctc_problem(6, 90, 130, 12, seed=51), s.ctc_problem(6, 90, 130, 12, seed=52)]
    x = torch.tensor(probs[0]["x"], device=dev, requires_grad=True)
    lab = torch.tensor(probs[0]["labels"], device=dev)
    il = torch.tensor(probs[0]["input_length"], device=dev)
    ll = torch.tensor(probs[0]["label_length"], device=dev)
    step = pkg.GraphedStep(lambda: pkg.ctc(x, lab, 0, il, ll, reduce="mean"), [x])
    for prob in probs + probs[:1]:
        with torch.no_grad():
            x.copy_(torch.tensor(prob["x"], device=dev)); lab.copy_(torch.tensor(prob["labels"], device=dev))
            il.copy_(torch.tensor(prob["input_length"], device=dev)); ll.copy_(torch.tensor(prob["label_length"], device=dev))
        loss = step.replay()
        torch.cuda.synchronize()
        got_loss, got_grad = float(loss.detach()), x.grad.clone()
        xe = torch.tensor(prob["x"], device=dev, requires_grad=True)
        le = pkg.ctc(xe, lab, 0, il, ll, reduce="mean")
        le.backward()
        assert got_loss == float(le.detach())
        assert torch.equal(got_grad, xe.grad)

"""-m gpu: every kernel variant the dispatcher can reach, checked against the float64 oracle.

Round-1 review: the lattice kernel's K=4 (CTC) / K=6 (Gram-CTC) instantiations, the B in [75,147] concurrent
regime, BASELINE configs[3] (B=512) and configs[4] at B=64, and real vocabulary sizes with V % 4 != 0 had never
produced a checked number.  These cases close those holes; tolerances are north_star's, flat (no allowance that
grows with T).
"""
import importlib

import numpy as np
import pytest

from util import run_cuda, run_oracle, assert_parity, LOSS_RTOL, GRAD_ATOL

pytestmark = pytest.mark.gpu


def synth():
    return importlib.import_module("chainer-speech-recognition_b200.synth")


def fast_ctc_problem(B, T, V, L, seed):
    """synth.ctc_problem with the activations drawn as float32 by a PCG64 generator (several times faster than
    RandomState.standard_normal + astype for the multi-GB cases); lengths and labels as in synth."""
    s = synth()
    rs = np.random.RandomState(seed)
    in_len, lab_len = s.make_lengths(rs, B, T, L, True)
    labels = s.make_ctc_labels(rs, B, L, V, lab_len)
    x = np.random.Generator(np.random.PCG64(seed)).standard_normal((T, B, V), dtype=np.float32)
    return {"x": x, "labels": labels, "input_length": in_len, "label_length": lab_len, "blank": 0}


# ---- long label sequences: every nodes-per-lane variant of the lattice kernel (csrc/lattice.cu, launch_lattice) ----
LONG_CTC = [
    # B, T, V, L
    (2, 1300, 16, 600),      # N = 1201: four nodes per lane
    (1, 2100, 16, 1000),     # N = 2001: near the instantiated maximum (2048 nodes)
    (2, 1100, 24, 500),      # N = 1001
    (3, 900, 32, 400),
    (2, 600, 32, 255),       # N = 511: the largest lattice with two nodes per lane (8 warps per direction)
    (2, 600, 32, 256),       # N = 513: the smallest with four
]


@pytest.mark.parametrize("shape", LONG_CTC)
@pytest.mark.parametrize("trained", [False, True])
def test_long_label_ctc(pkg, shape, trained):
    B, T, V, L = shape
    prob = synth().ctc_problem(B, T, V, L, seed=41, trained=trained)
    loss, grad, _ = run_cuda(pkg, prob, "ctc")
    loss_ref, grad_ref, _ = run_oracle(prob, "ctc")
    assert_parity(loss, grad, loss_ref, grad_ref, "long ctc %r" % (shape,))


LONG_GRAM = [
    (2, 1900, 150, 600),     # N = 1801: K=6
    (1, 3100, 150, 1000),    # N = 3001: K=6, near the instantiated maximum (3072 nodes)
    (2, 1600, 150, 500),     # N = 1501: the largest K=3 lattice
    (2, 300, 150, 85),       # N = 256: the largest K=1 lattice
    (2, 300, 150, 86),       # N = 259: the smallest K=3 lattice
]


@pytest.mark.parametrize("shape", LONG_GRAM)
@pytest.mark.parametrize("trained", [False, True])
def test_long_label_gram(pkg, shape, trained):
    B, T, V, L = shape
    prob = synth().gram_problem(B, T, V, L, seed=42, trained=trained, n_unigram=40)
    loss, grad, _ = run_cuda(pkg, prob, "gram")
    loss_ref, grad_ref, _ = run_oracle(prob, "gram")
    assert_parity(loss, grad, loss_ref, grad_ref, "long gram %r" % (shape,))


@pytest.mark.parametrize("lab_len", [[300, 1, 0, 157], [299, 2, 3, 158]])
def test_short_labels_inside_a_long_label_batch(pkg, lab_len):
    """Lmax = 300 selects the four-nodes-per-lane mapping (256-bit stores, reversed direction shifted by phantom nodes so
    that its groups are sector-aligned): label lengths 0, 1, 2, 3 and both parities of Nb + shift must come out right."""
    prob = synth().ctc_problem(4, 700, 50, 300, seed=45)
    prob["label_length"] = np.asarray(lab_len, np.int32)
    prob["input_length"] = np.asarray([700, 650, 40, 700], np.int32)
    loss, grad, _ = run_cuda(pkg, prob, "ctc")
    loss_ref, grad_ref, _ = run_oracle(prob, "ctc")
    assert_parity(loss, grad, loss_ref, grad_ref, "short labels, long Lmax %r" % (lab_len,))


def test_lattice_size_limit_is_reported_not_crashed(pkg):
    import torch
    x = torch.zeros(2100, 1, 8, device="cuda:0")
    lab = torch.ones(1, 1024, dtype=torch.int32, device="cuda:0")           # N = 2049 > 2048
    with pytest.raises(NotImplementedError):
        pkg.ctc(x, lab, 0)


# ---- batch sizes between the concurrent regime's comfortable range and the SM count (csrc/api.cu) ----
@pytest.mark.parametrize("B", [75, 96, 128, 144, 147, 148, 149])
@pytest.mark.parametrize("kind", ["ctc", "gram"])
def test_batch_sizes_around_the_sm_count(pkg, B, kind):
    s = synth()
    T, V, L = 260, 512, 40
    prob = s.ctc_problem(B, T, V, L, seed=43) if kind == "ctc" else s.gram_problem(B, T, V, L, seed=43, n_unigram=60)
    loss, grad, _ = run_cuda(pkg, prob, kind)
    loss_ref, grad_ref, _ = run_oracle(prob, kind)
    assert_parity(loss, grad, loss_ref, grad_ref, "%s B=%d" % (kind, B))


def test_concurrent_regime_with_a_large_lattice(pkg):
    """B < #SMs and a lattice whose pipeline wants more than 64 KB of shared memory next to the ring."""
    prob = synth().ctc_problem(100, 700, 256, 300, seed=44, trained=True)
    loss, grad, _ = run_cuda(pkg, prob, "ctc")
    loss_ref, grad_ref, _ = run_oracle(prob, "ctc")
    assert_parity(loss, grad, loss_ref, grad_ref, "B=100 L=300")


# ---- BASELINE configs[3]: B=512 on one GPU ----
def test_full_size_batch_512(pkg):
    prob = fast_ctc_problem(512, 800, 3500, 80, seed=45)
    loss, grad, _ = run_cuda(pkg, prob, "ctc")
    loss_ref, grad_ref, _ = run_oracle(prob, "ctc")
    assert_parity(loss, grad, loss_ref, grad_ref, "cfg4 B=512")
    for b in range(0, 512, 37):
        assert not grad[int(prob["input_length"][b]):, b].any()


# ---- BASELINE configs[4] at its own batch size, flat tolerance ----
@pytest.mark.parametrize("T,V", [(200, 100), (200, 3500), (800, 100), (1600, 100), (1600, 3500), (3200, 100), (3200, 3500)])
def test_sweep_at_batch_64(pkg, T, V):
    prob = fast_ctc_problem(64, T, V, T // 10, seed=46)
    if V <= 100:                                    # trained-like activations for the small vocabulary
        synth().add_alignment_bump(prob["x"], prob["labels"], prob["input_length"], prob["label_length"])
    loss, grad, _ = run_cuda(pkg, prob, "ctc")
    loss_ref, grad_ref, _ = run_oracle(prob, "ctc")
    assert_parity(loss, grad, loss_ref, grad_ref, "sweep T=%d V=%d" % (T, V))


# ---- real vocabularies: V = 119 unigram ids + bigrams, in general not a multiple of 4 (asr/vocab.py:62-97) ----
@pytest.mark.parametrize("V", [119, 1213, 3501, 3502, 3503])
@pytest.mark.parametrize("kind", ["ctc", "gram"])
def test_vocabulary_sizes_off_the_16_byte_grid(pkg, V, kind):
    s = synth()
    B, T, L = 6, 420, 30
    prob = s.ctc_problem(B, T, V, L, seed=47) if kind == "ctc" else s.gram_problem(B, T, V, L, seed=47, n_unigram=min(100, V // 2))
    loss, grad, am = run_cuda(pkg, prob, kind, want_argmax=True)
    loss_ref, grad_ref, am_ref = run_oracle(prob, kind, want_argmax=True)
    assert_parity(loss, grad, loss_ref, grad_ref, "%s V=%d" % (kind, V))
    assert np.array_equal(am, am_ref)
    # strided views (a row pitch that is not the vocabulary size) take the same path
    import torch
    xp = torch.zeros(T, B, V + 5, device="cuda:0")
    xp[:, :, :V] = torch.tensor(prob["x"], device="cuda:0")
    view = xp[:, :, :V].requires_grad_(True)
    lab = torch.tensor(prob["labels"], device="cuda:0")
    il = torch.tensor(prob["input_length"], device="cuda:0")
    ll = torch.tensor(prob["label_length"], device="cuda:0")
    if kind == "ctc":
        out = pkg.ctc(view, lab, 0, il, ll, reduce="no")
    else:
        out = pkg.gram_ctc(view, lab, torch.tensor(prob["bigrams"], device="cuda:0"), 0, il, ll, reduce="no")
    assert np.array_equal(out.detach().cpu().numpy().astype(np.float64), loss)


# ---- wide rows: V = 8000 (BASELINE configs[2]) and beyond what a single ring slot holds ----
@pytest.mark.parametrize("V", [8000, 20000, 60000])
def test_wide_rows(pkg, V):
    prob = synth().ctc_problem(4, 120, V, 12, seed=48, trained=True)
    loss, grad, am = run_cuda(pkg, prob, "ctc", want_argmax=True)
    loss_ref, grad_ref, am_ref = run_oracle(prob, "ctc", want_argmax=True)
    assert_parity(loss, grad, loss_ref, grad_ref, "V=%d" % V)
    assert np.array_equal(am, am_ref)


# ---- BASELINE configs[0] against the reference's own output (tests/golden/generate_golden_cfg1.py) ----
def test_cfg1_matches_the_reference_itself(pkg):
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "full", "cfg1_reference.npz"))
    B, T, V, L, seed = [int(v) for v in z["shape_seed"]]
    prob = synth().ctc_problem(B, T, V, L, seed=seed, trained=bool(z["trained"]))
    loss, grad, _ = run_cuda(pkg, prob, "ctc")
    assert np.allclose(loss, z["ref_loss"], rtol=1e-5)
    # the reference's gradient, sampled: every `stride`-th element of the flattened (T,B,V) array plus the label columns
    flat = grad.reshape(-1)
    got = flat[z["sample_index"]]
    # the reference's own float32 noise floor at |log-probability| ~ 200 is 6.4e-5 against float64 (tests/test_oracle.py,
    # SURVEY.md 0.5); the CUDA path is held to 1e-5 against the float64 oracle in test_gpu_parity.py (same shape)
    assert np.abs(got - z["ref_grad_sample"]).max() <= 1e-4


# ---- batch shards of different sizes through the package's own group= path (two processes on one GPU, gloo) ----
def _uneven_worker(rank, world, port, out):
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    import b200ctc
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob = b200ctc.synth.ctc_problem(5, 30, 20, 6, seed=9)                     # B = 5: shards of 3 and 2
    x, lab, il, ll = b200ctc.distributed.shard_batch(rank, world, (prob["x"], 1), prob["labels"], prob["input_length"],
                                                     prob["label_length"])
    dev = torch.device("cuda:0")
    xt = torch.tensor(np.ascontiguousarray(x), device=dev, requires_grad=True)
    loss = b200ctc.ctc(xt, torch.tensor(lab, device=dev), 0, torch.tensor(il, device=dev), torch.tensor(ll, device=dev),
                       reduce="mean", group=dist.group.WORLD)                 # no batch_global: agreed on by all-reduce
    loss.backward()
    torch.cuda.synchronize()
    out[rank] = (float(loss.item()), xt.grad.cpu().numpy())
    dist.destroy_process_group()


def test_uneven_shards_through_the_group_argument(pkg):
    import os
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_uneven_worker, args=(world, 29600 + os.getpid() % 2000, out), nprocs=world, join=True)
    prob = synth().ctc_problem(5, 30, 20, 6, seed=9)
    loss_ref, grad_ref, _ = run_oracle(prob, "ctc")
    for rank in range(world):
        s, e = pkg.distributed.shard_range(5, rank, world)
        loss, grad = out[rank]
        assert abs(loss - loss_ref.mean()) <= LOSS_RTOL * abs(loss_ref.mean())        # the mean over all 5, on every rank
        assert np.abs(grad - grad_ref[:, s:e] / 5.0).max() <= GRAD_ATOL               # scaled by 1/5, not 1/6 or 1/4

"""CPU: the C-ABI library loads and exports every symbol include/b200ctc.h declares, its host-only entry
points behave, and the Python shim validates arguments the way the reference does -- no GPU needed, no
compute call made."""
import ctypes
import importlib
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(pkg):
    header = open(os.path.join(ROOT, "include", "b200ctc.h")).read()
    declared = set(re.findall(r"\b(b200ctc_[a-z_0-9]+)\s*\(", header))
    assert {"b200ctc_forward", "b200ctc_backward", "b200ctc_greedy_argmax", "b200ctc_workspace_bytes",
            "b200ctc_last_error", "b200ctc_version"} <= declared
    lib = ctypes.CDLL(pkg._lib.library_path())
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.b200ctc_version() == int(re.search(r"#define B200CTC_VERSION (\d+)", header).group(1))


def test_workspace_bytes_and_status_codes(pkg):
    L = pkg._lib
    n = L.workspace_bytes(L.KIND_CTC, 64, 800, 3500, 80)
    assert 100e6 < n < 400e6 and n % 16 == 0
    assert L.workspace_bytes(L.KIND_GRAM, 32, 600, 8000, 60) > 0
    # joint Gram-CTC + CTC: the Gram-CTC workspace plus the plain-CTC lattice's alpha/beta rows
    assert L.workspace_bytes(L.KIND_JOINT, 32, 600, 8000, 60) > L.workspace_bytes(L.KIND_GRAM, 32, 600, 8000, 60)
    assert L.workspace_bytes(L.KIND_CTC, 0, 10, 5, 2) >= 0
    with pytest.raises(ValueError):
        L.workspace_bytes(7, 1, 1, 1, 1)                      # bad kind -> INVALID_ARGUMENT
    with pytest.raises(ValueError):
        L.workspace_bytes(L.KIND_CTC, 1, 1, 0, 1)             # V = 0
    with pytest.raises(NotImplementedError):
        L.workspace_bytes(L.KIND_CTC, 1, 10, 5, 100000)       # lattice larger than any instantiation
    out = ctypes.c_size_t(0)
    lib = L.load()
    assert lib.b200ctc_workspace_bytes(0, 1, 1, 1, 1, None) == L.INVALID_ARGUMENT
    assert b"NULL" in lib.b200ctc_last_error()
    assert lib.b200ctc_workspace_bytes(0, 1, 1, 1, 1, ctypes.byref(out)) == L.OK


def test_argument_checks_mirror_the_reference(pkg):
    torch = pytest.importorskip("torch")
    x = [torch.zeros(2, 5) for _ in range(4)]
    lab = torch.zeros(2, 2, dtype=torch.int32)
    with pytest.raises(ValueError):                           # gram_ctc.py:224-227
        pkg.gram_ctc(x, lab, lab, 0, reduce="sum")
    with pytest.raises(TypeError):                            # gram_ctc.py:303-304
        pkg.gram_ctc(x, lab, lab, 0.0)
    with pytest.raises(TypeError):                            # gram_ctc.py:301-302
        pkg.gram_ctc(3, lab, lab, 0)
    with pytest.raises(TypeError):                            # gram_ctc.py:241-242: float32 only
        pkg.ctc([t.double() for t in x], lab, 0)
    with pytest.raises(RuntimeError, match="no CPU path"):    # never a silent CPU fallback
        pkg.ctc(x, lab, 0)
    with pytest.raises(ValueError):
        pkg.GramCTC(0, reduce="median")


def test_stack_frames_recovers_views_without_copy(pkg):
    torch = pytest.importorskip("torch")
    F = importlib.import_module("chainer-speech-recognition_b200.asr.loss._function")
    base = torch.arange(4 * 3 * 5, dtype=torch.float32).reshape(4, 3, 5)
    views = [base[t] for t in range(4)]
    s = F.stack_frames(views)
    assert s.data_ptr() == base.data_ptr() and s.shape == (4, 3, 5) and torch.equal(s, base)
    btv = base.transpose(0, 1).contiguous()                   # (B,T,V) storage, frames strided
    s2 = F.stack_frames([btv[:, t] for t in range(4)])
    assert s2.data_ptr() == btv.data_ptr() and torch.equal(s2, base)
    s3 = F.stack_frames([v.clone() for v in views])           # unrelated storages: one copy, as xp.vstack does
    assert torch.equal(s3, base)


def test_synthetic_inputs_are_deterministic():
    synth = importlib.import_module("chainer-speech-recognition_b200.synth")
    a = synth.ctc_problem(4, 30, 20, 6, seed=3)
    b = synth.ctc_problem(4, 30, 20, 6, seed=3)
    assert all(np.array_equal(a[k], b[k]) for k in ("x", "labels", "input_length", "label_length"))
    assert a["input_length"][0] == 30 and a["label_length"][0] == 6
    g = synth.gram_problem(3, 40, 300, 8, seed=1)
    assert (g["bigrams"][:, 0] == -1).all()                   # asr/data/processing.py:139
    for b in range(3):                                        # ids within the label length; padding is the blank id
        n = g["label_length"][b]
        assert ((g["bigrams"][b, :n] == -1) | (g["bigrams"][b, :n] >= 119)).all()
        assert (g["bigrams"][b, n:] == 0).all()                # asr/data/processing.py:125-126
    assert (g["input_length"] >= 3 * g["label_length"] + 1).all()


def test_expansion_table_is_the_reference_string_round_trip(pkg):
    """asr/error.py:49-53: predicted ids -> token strings -> convert_sentence_to_unigram_ids, tabulated per id."""
    ids = {"_": 0, "a": 1, "b": 2, "c": 3, "ab": 4, "ca": 5}
    inv = {v: k for k, v in ids.items()}
    t = pkg.build_expansion_table(ids, inv).numpy()
    assert t.shape == (6, 2) and t.dtype == np.int32
    assert t.tolist() == [[0, -1], [1, -1], [2, -1], [3, -1], [1, 2], [3, 1]]
    # with the reference's own vocabulary and tokeniser, where the reference is mounted
    from oracle import ref_stub
    if ref_stub.available():
        _, vocab = ref_stub.load_error_module()
        rids, rinv = vocab.get_unigram_ids()
        first, second = vocab.UNIGRAM_TOKENS[70], vocab.UNIGRAM_TOKENS[3]      # a two-character mora + a plain one
        rids[first + second] = len(rids)
        rinv[rids[first + second]] = first + second
        rt = pkg.build_expansion_table(rids, rinv, convert=vocab.convert_sentence_to_unigram_ids).numpy()
        assert rt.shape[1] == 2
        assert rt[rids[first + second]].tolist() == [rids[first], rids[second]]
        assert all(rt[i, 0] == i and rt[i, 1] == -1 for i in range(1, 1 + len(vocab.UNIGRAM_TOKENS)))


def test_error_functions_refuse_to_run_without_a_gpu(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        pkg.compute_minibatch_error(np.zeros((1, 3), np.int64), np.zeros((1, 2), np.int32), 0, {"_": 0}, {0: "_"})

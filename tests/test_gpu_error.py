"""GPU: the device error-rate path (greedy collapse, gram expansion, edit distance, batch mean) through the C ABI
against the oracle (oracle/error.py) and the reference's committed golden outputs.  Integer work: bit-exact; the
float64 rates and their mean are formed with the reference's operations in the reference's order, so they are
compared for equality too."""
import glob
import os

import numpy as np
import pytest

from oracle import error as oerr

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "cer", "*.npz")))


def load(path):
    z = np.load(path)
    return {k: z[k] for k in z.files}


def run_device(pkg, y, t, blank, expansion, wrap=False, input_length=None):
    import torch
    out = pkg.minibatch_error_details(torch.tensor(y, device="cuda:0"), torch.tensor(t, device="cuda:0"), blank,
                                      torch.tensor(expansion, device="cuda:0"),
                                      None if input_length is None else torch.tensor(input_length, device="cuda:0"),
                                      wrap)
    torch.cuda.synchronize()
    hl = out["hyp_len"].cpu().numpy()
    hyp = out["hyp"].cpu().numpy()
    return (float(out["error"].item()), out["errors"].cpu().numpy(), [list(hyp[b, :hl[b]]) for b in range(len(hl))],
            out["distance"].cpu().numpy(), out["ref_len"].cpu().numpy())


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_device_matches_reference_golden(pkg, path):
    g = load(path)
    mean, errs, hyps, _, _ = run_device(pkg, g["y"], g["t"], int(g["blank"]), g["expansion"])
    assert mean == float(g["ref_mean"])
    assert np.array_equal(errs, g["ref_each"])
    _, _, ohyps = oerr.minibatch_error(g["y"], g["t"], int(g["blank"]), g["expansion"])
    assert hyps == ohyps


@pytest.mark.parametrize("B,T,L,V,E,pblank,seed", [
    (1, 1, 1, 5, 1, 0.0, 0), (3, 33, 7, 40, 2, 0.5, 1), (16, 200, 40, 500, 3, 0.6, 2), (4, 800, 300, 3500, 2, 0.1, 3),
    (2, 64, 0, 10, 1, 0.3, 4), (5, 31, 33, 8, 1, 0.9, 5)])
@pytest.mark.parametrize("wrap", [False, True])
def test_device_matches_oracle_random(pkg, B, T, L, V, E, pblank, seed, wrap):
    rng = np.random.RandomState(seed)
    y = rng.randint(0, V, size=(B, T)).astype(np.int64)
    y[rng.rand(B, T) < pblank] = 0
    rep = rng.rand(B, T) < 0.3
    for t in range(1, T):
        y[rep[:, t], t] = y[rep[:, t], t - 1]
    tb = rng.randint(0, V, size=(B, L)).astype(np.int32)
    expansion = rng.randint(1, V, size=(V, E)).astype(np.int32)
    expansion[rng.rand(V, E) < 0.3] = -1
    expansion[:, 0] = np.where(expansion[:, 0] < 0, 1, expansion[:, 0])
    expansion = np.sort(expansion, axis=1)[:, ::-1].copy()           # valid ids first, -1 padding last
    il = rng.randint(0, T + 1, size=B).astype(np.int32) if seed % 2 else None
    mean, errs, hyps, dist, rl = run_device(pkg, y, tb, 0, expansion, wrap, il)
    omean, oerrs, ohyps = oerr.minibatch_error(y, tb, 0, expansion, wrap, il)
    assert hyps == ohyps
    assert np.array_equal(errs, oerrs) and mean == omean
    for b in range(B):
        target = [int(v) for v in tb[b] if v != 0]
        assert rl[b] == len(target)
        assert dist[b] == oerr.edit_distance(target, ohyps[b], wrap)


def test_character_error_rate_pairs(pkg):
    assert pkg.compute_character_error_rate([1, 2, 3], [1, 3]) == 1.0 / 3
    assert pkg.compute_character_error_rate([], [4, 4, 4]) == 3                  # asr/error.py:8-9
    assert pkg.compute_character_error_rate([7], []) == 1.0
    rng = np.random.RandomState(9)
    for _ in range(5):
        r, h = list(rng.randint(0, 6, size=rng.randint(1, 400))), list(rng.randint(0, 6, size=rng.randint(0, 500)))
        assert pkg.compute_character_error_rate(r, h) == oerr.character_error_rate(r, h)
        assert pkg.compute_character_error_rate(r, h, uint8_wrap=True) == oerr.character_error_rate(r, h, True)


def test_evaluation_path_from_activations(pkg):
    """run/ctc/cnn/train.py:231-233: argmax over the vocabulary, then compute_minibatch_error with the vocabulary
    dicts -- here with a toy vocabulary whose bigram tokens are two characters."""
    import torch
    ids = {"_": 0}
    for ch in "abcdefg":
        ids[ch] = len(ids)
    for a in "abc":
        for b in "defg":
            ids[a + b] = len(ids)
    inv = {v: k for k, v in ids.items()}
    V = len(ids)
    rng = np.random.RandomState(11)
    B, T, L = 6, 50, 9
    x = rng.randn(B, T, V).astype(np.float32)
    tb = rng.randint(1, 8, size=(B, L)).astype(np.int32)
    tb[:, 6:] = 0
    y = pkg.greedy_argmax(torch.tensor(x, device="cuda:0"))
    got = pkg.compute_minibatch_error(y, torch.tensor(tb, device="cuda:0"), 0, ids, inv)
    table = pkg.build_expansion_table(ids, inv).numpy()
    assert table.shape == (V, 2) and list(table[ids["ad"]]) == [ids["a"], ids["d"]]
    want, _, _ = oerr.minibatch_error(np.argmax(x, axis=2), tb, 0, table)
    assert got == want

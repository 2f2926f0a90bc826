"""CPU, world_size 2, gloo: the multi-GPU protocol of SURVEY.md section 8e -- contiguous utterance shards,
ONE scalar all-reduce for reduce='mean', gradient scale 1/B_global without communication.  The per-rank
compute is the oracle here (no GPU in this container); on the GPU box the same protocol runs in bench.py
--gpus N over NCCL."""
import importlib
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist          # noqa: E402
import torch.multiprocessing as mp        # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    D = importlib.import_module("chainer-speech-recognition_b200.distributed")
    synth = importlib.import_module("chainer-speech-recognition_b200.synth")
    from oracle import c_oracle
    prob = synth.ctc_problem(5, 30, 20, 6, seed=9)            # B = 5: uneven shards (3 + 2)
    x, lab, il, ll = D.shard_batch(rank, world, (prob["x"], 1), prob["labels"], prob["input_length"], prob["label_length"])
    B_global = prob["x"].shape[1]
    r = c_oracle.run(0, np.ascontiguousarray(x), lab, None, il, ll, 0, grad_scale=np.full(len(il), 1.0 / B_global), nthreads=1)
    local = torch.tensor(r["loss"].sum() / B_global, dtype=torch.float64)
    D.reduce_mean_loss(local)                                 # the one collective of the path
    s, e = D.shard_range(B_global, rank, world)
    out[rank] = (float(local), s, e, r["grad"])
    dist.destroy_process_group()


def test_two_rank_mean_loss_and_local_gradients():
    world = 2
    port = 29500 + os.getpid() % 2000
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    synth = importlib.import_module("chainer-speech-recognition_b200.synth")
    from oracle import c_oracle
    prob = synth.ctc_problem(5, 30, 20, 6, seed=9)
    full = c_oracle.run(0, prob["x"], prob["labels"], None, prob["input_length"], prob["label_length"], 0,
                        grad_scale=np.full(5, 0.2), nthreads=1)
    assert (out[0][1], out[0][2], out[1][1], out[1][2]) == (0, 3, 3, 5)
    for rank in range(world):
        loss, s, e, grad = out[rank]
        assert np.isclose(loss, full["loss"].mean(), rtol=1e-12)          # every rank holds the global mean
        assert np.array_equal(grad, full["grad"][:, s:e])                # gradients never leave the owning rank


def test_shard_ranges_cover_the_batch():
    D = importlib.import_module("chainer-speech-recognition_b200.distributed")
    for B in (0, 1, 7, 64, 513):
        for G in (1, 2, 4, 8):
            ranges = [D.shard_range(B, r, G) for r in range(G)]
            assert ranges[0][0] == 0 and ranges[-1][1] == B
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(G - 1))
            sizes = [e - s for s, e in ranges]
            assert max(sizes) - min(sizes) <= 1

"""Device-time table for every workload BASELINE.json names, beyond the one bench.py reports:

  cfg1  CTC B=8,T=200,V=3500,L=40            cfg2  CTC B=64,T=800,V=3500,L=80 (the bench workload)
  cfg3  Gram-CTC B=32,T=600,V=8000,L=60      cfg4  CTC B=512,T=800,V=3500,L=80 on ONE GPU (its per-GPU share at 8 GPUs is cfg2)
  cfg5  CTC sweep T in {200,800,1600,3200} x V in {100,3500}, B=64, L=T/10
  plus  cfg2 at B = 96, 128, 144 (the softmax/gather and lattice kernels still run side by side, two lattice CTAs per SM)

For each: forward / backward / step time (CUDA events, inputs resident in HBM, 5 warm-up + 20 timed steps), padded
frames per second and the fraction of the HBM roofline of the step's algorithmic bytes
(8*V*sum(T_b) + 4*V*B*T) against MEASURED_PEAKS.json.  Activations are generated on the device
(torch.randn) -- this is a timing tool, parity is the tests' job.

    python tools/sweep.py [--only cfg3] > profiles/rN_sweep.txt
"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np      # noqa: E402
import torch            # noqa: E402
import b200ctc          # noqa: E402

synth = importlib.import_module("chainer-speech-recognition_b200.synth")


def peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6549.8


def configs():
    yield "cfg1", "ctc", 8, 200, 3500, 40
    yield "cfg2", "ctc", 64, 800, 3500, 80
    yield "cfg3", "gram", 32, 600, 8000, 60
    yield "cfg3 +ctc, 2 calls", "gram+ctc", 32, 600, 8000, 60      # joint training, run/gram_ctc/cnn/train.py:196-198
    yield "cfg3 +ctc, joint", "joint", 32, 600, 8000, 60           # the same objective from one pass (joint_ctc=True)
    for B in (96, 128, 144):                                       # between "every lattice CTA pair has an SM" and the serial schedule
        yield "cfg2 B=%d" % B, "ctc", B, 800, 3500, 80
    yield "cfg4", "ctc", 512, 800, 3500, 80
    yield "cfg4 length-sorted", "ctc", 512, 800, 3500, 80          # asr/data/processing.py sort_by_length=True
    for T in (200, 800, 1600, 3200):
        for V in (100, 3500):
            yield "cfg5 T=%d V=%d" % (T, V), "ctc", 64, T, V, T // 10


def run(name, kind, B, T, V, L, steps=20, warmup=5):
    dev = torch.device("cuda:0")
    rs = np.random.RandomState(0)
    in_len, lab_len = synth.make_lengths(rs, B, T, L)
    if "sorted" in name:
        order = np.argsort(-in_len, kind="stable")
        in_len, lab_len = np.ascontiguousarray(in_len[order]), np.ascontiguousarray(lab_len[order])
    if kind == "ctc":
        labels = synth.make_ctc_labels(rs, B, L, V, lab_len)
        big = None
    else:
        labels, bigs = synth.make_gram_labels(rs, B, L, V, lab_len)
        big = torch.tensor(bigs, device=dev)
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    x = torch.randn((T, B, V), device=dev, generator=g).requires_grad_(True)
    lab = torch.tensor(labels, device=dev)
    il, ll = torch.tensor(in_len, device=dev), torch.tensor(lab_len, device=dev)

    def fwd():
        if kind == "ctc":
            return b200ctc.connectionist_temporal_classification(x, lab, 0, il, ll, reduce="mean")
        if kind == "gram+ctc":
            return (b200ctc.gram_ctc(x, lab, big, 0, il, ll, reduce="mean") +
                    b200ctc.connectionist_temporal_classification(x, lab, 0, il, ll, reduce="mean"))
        return b200ctc.gram_ctc(x, lab, big, 0, il, ll, reduce="mean", joint_ctc=(kind == "joint"))

    for _ in range(warmup):
        x.grad = None
        fwd().backward()
    torch.cuda.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    for k in range(steps):
        x.grad = None
        ev[k][0].record()
        loss = fwd()
        ev[k][1].record()
        loss.backward()
        ev[k][2].record()
    torch.cuda.synchronize()
    f = float(np.median([ev[k][0].elapsed_time(ev[k][1]) for k in range(steps)]))
    b = float(np.median([ev[k][1].elapsed_time(ev[k][2]) for k in range(steps)]))
    s = ev[0][0].elapsed_time(ev[-1][2]) / steps
    nbytes = 8.0 * V * float(in_len.sum()) + 4.0 * V * B * T          # one loss; the joint objective needs no more
    frac = nbytes / (s * 1e-3) / 1e9 / peak_gbs()
    print("%-20s %-8s B=%-3d T=%-4d V=%-4d L=%-3d  fwd %7.3f ms  bwd %7.3f ms  step %7.3f ms  %8.1f M padded frames/s  "
          "%5.1f %% of HBM roofline  loss %.4f" % (name, kind, B, T, V, L, f, b, s, B * T / s / 1e3, 100 * frac, float(loss.detach())),
          flush=True)
    del x


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    print("# tools/sweep.py on %s, HBM peak %.1f GB/s (MEASURED_PEAKS.json)" % (torch.cuda.get_device_name(0), peak_gbs()))
    for cfg in configs():
        if a.only and not cfg[0].startswith(a.only):
            continue
        run(*cfg)

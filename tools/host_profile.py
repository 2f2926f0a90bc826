"""Where does the host time of one eager step go?  (~0.2 ms of enqueue per step.)

Times, per call and without waiting for the GPU (BASELINE configs[0], whose kernels are shorter than the enqueue):
  raw C-ABI forward / backward through ctypes      -- the library's own host work + the CUDA launches
  the public API forward (autograd Function)       -- + argument checks, allocations, autograd bookkeeping
  loss.backward()                                  -- + the autograd engine
then prints the cProfile top list of the public-API loop.

    python tools/host_profile.py [--steps 300]
"""
import argparse
import cProfile
import importlib
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np      # noqa: E402
import torch            # noqa: E402
import b200ctc          # noqa: E402

pkg = importlib.import_module("chainer-speech-recognition_b200")
synth = importlib.import_module("chainer-speech-recognition_b200.synth")
_lib = importlib.import_module("chainer-speech-recognition_b200._lib")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--shape", default="8,200,3500,40")
    a = ap.parse_args()
    B, T, V, L = (int(v) for v in a.shape.split(","))
    dev = torch.device("cuda:0")
    rs = np.random.RandomState(0)
    in_len, lab_len = synth.make_lengths(rs, B, T, L)
    labels = torch.tensor(synth.make_ctc_labels(rs, B, L, V, lab_len), device=dev)
    il, ll = torch.tensor(in_len, device=dev), torch.tensor(lab_len, device=dev)
    x = torch.randn((T, B, V), device=dev).requires_grad_(True)
    n = a.steps

    def api_step():
        x.grad = None
        loss = b200ctc.connectionist_temporal_classification(x, labels, 0, il, ll, reduce="mean")
        loss.backward()

    for _ in range(20):
        api_step()
    torch.cuda.synchronize()

    # --- raw C ABI ---
    lib = _lib.load()
    ws = torch.empty(_lib.workspace_bytes(_lib.KIND_CTC, B, T, V, L), dtype=torch.uint8, device=dev)
    losses = torch.empty(B + 1, device=dev)
    grad = torch.empty_like(x)
    gy = torch.ones((), device=dev)
    xd = x.detach()
    stream = torch.cuda.current_stream().cuda_stream

    def raw_fwd():
        _lib.check(lib.b200ctc_forward(_lib.KIND_CTC, xd.data_ptr(), xd.stride(0), xd.stride(1), labels.data_ptr(), None,
                                       il.data_ptr(), ll.data_ptr(), 0, B, T, V, L, losses.data_ptr(),
                                       losses[B:].data_ptr(), 1.0 / B, None, ws.data_ptr(), ws.numel(), 0, stream))

    def raw_bwd():
        _lib.check(lib.b200ctc_backward(_lib.KIND_CTC, xd.data_ptr(), xd.stride(0), xd.stride(1), labels.data_ptr(), None,
                                        0, B, T, V, L, gy.data_ptr(), 0, 1.0 / B, grad.data_ptr(), grad.stride(0),
                                        grad.stride(1), ws.data_ptr(), ws.numel(), stream))

    def timed(fn, label):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print("%-44s %7.1f us per call on the host   (%7.1f us incl. waiting for the GPU)" %
              (label, (t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6), flush=True)

    raw_fwd(); raw_bwd()
    timed(raw_fwd, "C ABI b200ctc_forward (ctypes)")
    timed(raw_bwd, "C ABI b200ctc_backward (ctypes)")
    timed(lambda: (raw_fwd(), raw_bwd()), "C ABI forward + backward")

    def api_fwd_only():
        return b200ctc.connectionist_temporal_classification(x, labels, 0, il, ll, reduce="mean")

    xn = x.detach()
    timed(lambda: b200ctc.connectionist_temporal_classification(xn, labels, 0, il, ll, reduce="mean"),
          "public API forward, no autograd graph")
    timed(api_fwd_only, "public API forward, autograd")
    timed(api_step, "public API forward + loss.backward()")

    pr = cProfile.Profile()
    torch.cuda.synchronize()
    pr.enable()
    for _ in range(n):
        api_step()
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(22)


if __name__ == "__main__":
    main()

"""Where one training step's device time goes: first-CTA-start / last-CTA-end of each kernel on the GPU's
globaltimer (needs the b200ctc_debug_timeline hook), for the bench workload.

    python tools/step_timeline.py
"""
import ctypes, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import b200ctc
synth = importlib.import_module("chainer-speech-recognition_b200.synth")
lib = b200ctc._lib.load()
lib.b200ctc_debug_timeline.argtypes = [ctypes.c_void_p]
kind = sys.argv[1] if len(sys.argv) > 1 else "ctc"          # ctc | gram | joint  [B T V L]
dims = [int(v) for v in sys.argv[2:6]] if len(sys.argv) >= 6 else None
if kind == "ctc":
    prob = synth.ctc_problem(*(dims or [64, 800, 3500, 80]), seed=0)
else:
    prob = synth.gram_problem(*(dims or [32, 600, 8000, 60]), seed=0)
dev = torch.device("cuda:0")
x = torch.tensor(prob["x"], device=dev, requires_grad=True)
lab = torch.tensor(prob["labels"], device=dev); il = torch.tensor(prob["input_length"], device=dev); ll = torch.tensor(prob["label_length"], device=dev)
big = torch.tensor(prob["bigrams"], device=dev) if kind != "ctc" else None
def step():
    x.grad = None
    if kind == "ctc":
        b200ctc.ctc(x, lab, 0, il, ll, reduce="mean").backward()
    else:
        b200ctc.gram_ctc(x, lab, big, 0, il, ll, reduce="mean", joint_ctc=(kind == "joint")).backward()
for _ in range(5): step()
torch.cuda.synchronize()
BIG = np.iinfo(np.int64).max
rows = []
for it in range(6):
    tl = torch.tensor([BIG, 0] * 4, dtype=torch.int64, device=dev)
    lib.b200ctc_debug_timeline(tl.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(); step(); step(); e1.record()            # two steps; the second one's marks overwrite nothing (min/max): use gaps
    torch.cuda.synchronize()
    lib.b200ctc_debug_timeline(None)
    rows.append((tl.cpu().numpy().reshape(4, 2), e0.elapsed_time(e1)))
# single-step timeline (marks accumulate min/max over both steps, so rerun with one step)
for it in range(4):
    tl = torch.tensor([BIG, 0] * 4, dtype=torch.int64, device=dev)
    lib.b200ctc_debug_timeline(tl.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(); step(); e1.record()
    torch.cuda.synchronize()
    lib.b200ctc_debug_timeline(None)
    t = tl.cpu().numpy().reshape(4, 2).astype(np.float64)
    t0 = t[0, 0]
    names = ["softmax/gather", "lattice alpha", "lattice beta", "gradient"]
    print("step %d: event time %.1f us" % (it, e0.elapsed_time(e1) * 1e3))
    for n, (a, b) in zip(names, t):
        print("   %-16s start %7.1f us   end %7.1f us   (%.1f us)" % (n, (a - t0) / 1e3, (b - t0) / 1e3, (b - a) / 1e3))

"""Profiling aid: per-warp cycle breakdown of the lattice kernel (needs the debug hook in lattice.cu).

For every lattice CTA (one per utterance and direction) and warp: cycles per chunk (16 frames) spent waiting for
the previous warp's boundary values / the emission rows (wait), in the recursion steps (work) and in the
hand-over."""
import ctypes, importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import b200ctc
synth = importlib.import_module("chainer-speech-recognition_b200.synth")
lib = b200ctc._lib.load()
kind = sys.argv[1] if len(sys.argv) > 1 else "ctc"
dims = [int(v) for v in sys.argv[2:6]] if len(sys.argv) >= 6 else None        # B T V L
if kind == "ctc":
    prob = synth.ctc_problem(*(dims or [64, 800, 3500, 80]), seed=0)
else:
    prob = synth.gram_problem(*(dims or [32, 600, 8000, 60]), seed=0)
dev = torch.device("cuda:0")
B = prob["x"].shape[1]
x = torch.tensor(prob["x"], device=dev)
lab = torch.tensor(prob["labels"], device=dev); il = torch.tensor(prob["input_length"], device=dev); ll = torch.tensor(prob["label_length"], device=dev)
big = torch.tensor(prob["bigrams"], device=dev) if kind != "ctc" else None
dbg = torch.zeros(64 * 32 * 8, dtype=torch.int64, device=dev)
def run():
    if kind == "ctc": return b200ctc.ctc(x, lab, 0, il, ll, reduce="no")
    return b200ctc.gram_ctc(x, lab, big, 0, il, ll, reduce="no")
for it in range(5):
    if it == 4:
        lib.b200ctc_debug_lattice.argtypes = [ctypes.c_void_p]
        lib.b200ctc_debug_lattice(dbg.data_ptr())
    run()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
lib.b200ctc_debug_lattice(None)
e0.record(); [run() for _ in range(10)]; e1.record(); torch.cuda.synchronize()
print("forward (K1+K2) %.1f us" % (e0.elapsed_time(e1) * 100))
a = dbg.cpu().numpy().reshape(64, 32, 8)       # [lattice CTA = 2*b + direction][warp][counter]
t_end = a[:, :, 6]; t0 = t_end[t_end > 0].min()
for b in (0, 5, min(B - 1, 31)):
    for d, nm in ((0, "alpha"), (1, "beta")):
        for wi in range(16):
            r = a[2 * b + d, wi]
            if r[5] == 0: continue
            print("utt %2d T %d %s warp %d per-chunk cycles: wait %.0f work %.0f hand-over %.0f (chunks %d)" % (
                b, prob["input_length"][b], nm, wi, r[0] / r[5], r[3] / r[5], r[4] / r[5], r[5]))
print("spread of warp end times over the first 32 utterances: %.1f us" % ((t_end.max() - t0) / 1e3))

"""Where the softmax/gather kernel's warps spend their time (experiment build only: B200CTC_EXPERIMENT=1).

    B200CTC_EXPERIMENT=1 python tools/k1_roles.py [serial]

Per CTA and warp: producer = [cycles waiting for a ticket, cycles waiting for free slots, rows, total];
consumer = [cycles waiting for a row, cycles processing, rows, total].  Prints averages over the CTAs.
"""
import ctypes, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import b200ctc
synth = importlib.import_module("chainer-speech-recognition_b200.synth")
lib = b200ctc._lib.load()
lib.b200ctc_debug_k1_roles.argtypes = [ctypes.c_void_p]
if len(sys.argv) > 1 and sys.argv[1] == "serial":
    os.environ["B200CTC_NO_CONCURRENT"] = "1"
prob = synth.ctc_problem(64, 800, 3500, 80, seed=0)
dev = torch.device("cuda:0")
x = torch.tensor(prob["x"], device=dev)
lab = torch.tensor(prob["labels"], device=dev); il = torch.tensor(prob["input_length"], device=dev); ll = torch.tensor(prob["label_length"], device=dev)
for _ in range(3):
    b200ctc.ctc(x, lab, 0, il, ll, reduce="mean")
torch.cuda.synchronize()
buf = torch.zeros(148 * 16 * 4, dtype=torch.int64, device=dev)
lib.b200ctc_debug_k1_roles(buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); b200ctc.ctc(x, lab, 0, il, ll, reduce="mean"); e1.record()
torch.cuda.synchronize()
lib.b200ctc_debug_k1_roles(None)
r = buf.cpu().numpy().reshape(148, 16, 4).astype(np.float64)
print("forward %.1f us" % (e0.elapsed_time(e1) * 1e3))
p = r[:, 0]
print("producer : ticket wait %7.0f  slot wait %7.0f  rows %5.0f  total %7.0f cycles   (per row: %5.0f cycles)" %
      (p[:, 0].mean(), p[:, 1].mean(), p[:, 2].mean(), p[:, 3].mean(), p[:, 3].mean() / max(p[:, 2].mean(), 1)))
c = r[:, 1:11]
act = c[:, :, 2] > 0
print("consumers: row wait %7.0f  processing %7.0f  rows %5.1f  total %7.0f cycles   (per row: wait %5.0f, proc %5.0f)" %
      (c[:, :, 0][act].mean(), c[:, :, 1][act].mean(), c[:, :, 2][act].mean(), c[:, :, 3][act].mean(),
       c[:, :, 0][act].sum() / c[:, :, 2][act].sum(), c[:, :, 1][act].sum() / c[:, :, 2][act].sum()))

"""Cycles per phase of the fused LayerNormalization kernels' tile loop (experiment build: B200CTC_EXPERIMENT=1).
forward : 0 wait first box, 1 load tile + symbol rows + release, 2 moments (regs + shuffles), 3 barrier, 4 cross-warp merge + barrier,
          5 transform/max/sum + shuffles, 6 barrier + merge + barrier, 7 emission + barrier + signal
backward: 0 wait first box, 1 padded tile, 2 constants + wait alpha/beta, 3 posteriors + barrier, 4 sweep 1, 5 reduce + 2 barriers,
          6 sweep 2, 7 end barrier"""
import ctypes, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import b200ctc
synth = importlib.import_module("chainer-speech-recognition_b200.synth")
lib = b200ctc._lib.load()
lib.b200ctc_debug_ln.argtypes = [ctypes.c_void_p]
B, T, V, L = 64, 800, 3500, 80
dev = torch.device("cuda:0")
rs = np.random.RandomState(0)
in_len, lab_len = synth.make_lengths(rs, B, T, L)
labels = synth.make_ctc_labels(rs, B, L, V, lab_len)
z = (torch.randn((B, V, 1, T), device=dev) * 1.7 + 0.3).requires_grad_(True)
gamma = torch.ones(V, device=dev, requires_grad=True); beta = torch.zeros(V, device=dev, requires_grad=True)
lab = torch.tensor(labels, device=dev); il = torch.tensor(in_len, device=dev); ll = torch.tensor(lab_len, device=dev)
for _ in range(3):
    z.grad = None
    b200ctc.layernorm_ctc(z, gamma, beta, lab, 0, il, ll).backward()
torch.cuda.synchronize()
for name in (sys.argv[1:2] or ["forward"]):
    buf = torch.zeros(148 * 16 * 8, dtype=torch.int64, device=dev)
    z.grad = None
    if name == "forward":
        lib.b200ctc_debug_ln(buf.data_ptr())
    loss = b200ctc.layernorm_ctc(z, gamma, beta, lab, 0, il, ll)
    torch.cuda.synchronize()
    lib.b200ctc_debug_ln(buf.data_ptr() if name == "backward" else None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); loss.backward(); e1.record()
    torch.cuda.synchronize()
    lib.b200ctc_debug_ln(None)
    r = buf.cpu().numpy().reshape(148, 16, 8).astype(np.float64)[:, :15]
    tot = r.sum(axis=2).mean()
    print("%-8s total %8.0f cycles per warp (%.1f us); phases (avg over warps, %% of total):" % (name, tot, tot / 1.93e3))
    print("   " + "  ".join("%d:%4.1f%%" % (i, 100 * r[:, :, i].mean() / tot) for i in range(8)))
    print("   warp 0 :" + "  ".join("%d:%4.1f%%" % (i, 100 * r[:, 0, i].mean() / tot) for i in range(8)))
    print("   warp 14:" + "  ".join("%d:%4.1f%%" % (i, 100 * r[:, 14, i].mean() / tot) for i in range(8)))

import sys, importlib, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch, b200ctc
synth = importlib.import_module("chainer-speech-recognition_b200.synth")
B, T, V, L = [int(a) for a in sys.argv[1:5]]
dev = torch.device("cuda:0")
rs = np.random.RandomState(0)
in_len, lab_len = synth.make_lengths(rs, B, T, L)
labels = synth.make_ctc_labels(rs, B, L, V, lab_len)
z = (torch.randn((B, V, 1, T), device=dev) * 1.7 + 0.3).requires_grad_(True)
gamma = torch.ones(V, device=dev, requires_grad=True); beta = torch.zeros(V, device=dev, requires_grad=True)
lab = torch.tensor(labels, device=dev); il = torch.tensor(in_len, device=dev); ll = torch.tensor(lab_len, device=dev)
loss = b200ctc.layernorm_ctc(z, gamma, beta, lab, 0, il, ll)
torch.cuda.synchronize(); print("forward ok", float(loss), flush=True)
loss.backward(); torch.cuda.synchronize(); print("backward ok", flush=True)

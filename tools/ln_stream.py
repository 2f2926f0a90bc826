"""Experiment build: forward pass of the LayerNormalization-fused loss with the lattice BEHIND the softmax/gather kernel
(B200CTC_NO_CONCURRENT=1), normally and with B200CTC_DBG_PROGRESS=64 (compute warps only pull the tiles into registers and
free the boxes: the rate the tile loads alone sustain).  Results of the second run are garbage by construction."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import b200ctc
synth = importlib.import_module("chainer-speech-recognition_b200.synth")
B, T, V, L = 64, 800, 3500, 80
dev = torch.device("cuda:0")
rs = np.random.RandomState(0)
in_len, lab_len = synth.make_lengths(rs, B, T, L)
labels = synth.make_ctc_labels(rs, B, L, V, lab_len)
z = (torch.randn((B, V, 1, T), device=dev) * 1.7 + 0.3)
gamma = torch.ones(V, device=dev); beta = torch.zeros(V, device=dev)
lab = torch.tensor(labels, device=dev); il = torch.tensor(in_len, device=dev); ll = torch.tensor(lab_len, device=dev)
for _ in range(3):
    b200ctc.layernorm_ctc(z, gamma, beta, lab, 0, il, ll)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
e[0].record()
for k in range(10):
    b200ctc.layernorm_ctc(z, gamma, beta, lab, 0, il, ll)
    e[k + 1].record()
torch.cuda.synchronize()
print("forward (softmax/gather kernel + lattice behind it): %.1f us per call" % (e[0].elapsed_time(e[10]) * 100.0),
      " valid bytes of z: %.0f MB" % (4.0 * V * in_len.sum() / 1e6))

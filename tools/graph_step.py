"""One training step (loss forward + backward) captured in a CUDA graph and replayed: the library's fork/join onto
its side stream is capturable, and replay removes the launch/dependency gaps of the eager step.

    python tools/graph_step.py
"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import b200ctc
synth = importlib.import_module("chainer-speech-recognition_b200.synth")
prob = synth.ctc_problem(64, 800, 3500, 80, seed=0)
dev = torch.device("cuda:0")
x = torch.tensor(prob["x"], device=dev, requires_grad=True)
lab = torch.tensor(prob["labels"], device=dev); il = torch.tensor(prob["input_length"], device=dev); ll = torch.tensor(prob["label_length"], device=dev)

def step():
    x.grad = None
    loss = b200ctc.ctc(x, lab, 0, il, ll, reduce="mean")
    loss.backward()
    return loss

s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        step()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
eager_loss = float(step().detach()); eager_grad = x.grad.clone()
g = torch.cuda.CUDAGraph()
x.grad = None
with torch.cuda.graph(g):
    loss = b200ctc.ctc(x, lab, 0, il, ll, reduce="mean")
    loss.backward()
torch.cuda.synchronize()
g.replay(); torch.cuda.synchronize()
print("graph loss %.6f eager loss %.6f  grad max diff %.3g" % (float(loss.detach()), eager_loss, float((x.grad - eager_grad).abs().max())))
for name, fn in (("eager", step), ("graph replay", g.replay)):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print("%-13s %.1f us per step" % (name, e0.elapsed_time(e1) * 1e3 / 20))

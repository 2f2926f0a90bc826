import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import b200ctc
synth = importlib.import_module("chainer-speech-recognition_b200.synth")
prob = synth.ctc_problem(64, 800, 3500, 80, seed=0)
xh = torch.from_numpy(np.ascontiguousarray(prob["x"].transpose(1, 0, 2))).pin_memory()
gh = torch.empty_like(xh).pin_memory()
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    l, g = b200ctc.ctc_host(xh, prob["labels"], 0, prob["input_length"], prob["label_length"], reduce="mean", grad_out=gh)
    print(it, l, flush=True)
print("done host")

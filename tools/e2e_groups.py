import sys, importlib, numpy as np, torch, time
sys.path.insert(0, '/root/repo')
import b200ctc
synth = importlib.import_module("chainer-speech-recognition_b200.synth")
prob = synth.ctc_problem(64, 800, 3500, 80, seed=0)
dev = torch.device("cuda:0")
labels = torch.tensor(prob["labels"], device=dev); il = torch.tensor(prob["input_length"], device=dev); ll = torch.tensor(prob["label_length"], device=dev)
x_host = torch.from_numpy(np.ascontiguousarray(prob["x"].transpose(1, 0, 2))).pin_memory()
g_host = torch.empty_like(x_host).pin_memory()
for groups in (4, 8, 16, 32, 64):
    for rep in range(2):
        b200ctc.ctc_host(x_host, labels, 0, il, ll, reduce="mean", grad_out=g_host, groups=groups)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for rep in range(5):
        l, _ = b200ctc.ctc_host(x_host, labels, 0, il, ll, reduce="mean", grad_out=g_host, groups=groups)
    torch.cuda.synchronize()
    print("groups", groups, "ms/step %.2f" % ((time.perf_counter() - t0) * 1e3 / 5), "loss", l)
# raw PCIe numbers
xd = torch.empty_like(x_host, device=dev)
for name, fn in (("H2D", lambda: xd.copy_(x_host, non_blocking=True)), ("D2H", lambda: g_host.copy_(xd, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(name, "%.2f ms  %.1f GB/s" % (dt * 1e3, x_host.numel() * 4 / dt / 1e9))

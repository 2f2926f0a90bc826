"""End-to-end step (asr/loss/host.py) timed alone, with the host-side phases split out (development tool).
Run one copy per GPU at the same time to see what the ranks share:  tools/e2e_pair.sh"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, b200ctc
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
synth = importlib.import_module("chainer-speech-recognition_b200.synth")
B, T, V, L = 64, 800, 3500, 80
rs = np.random.RandomState(0)
in_len, lab_len = synth.make_lengths(rs, B, T, L)
labels = synth.make_ctc_labels(rs, B, L, V, lab_len)
x_host = torch.randn((B, T, V)).pin_memory(); g_host = torch.empty_like(x_host).pin_memory()
lab = torch.tensor(labels, device="cuda"); il = torch.tensor(in_len, device="cuda"); ll = torch.tensor(lab_len, device="cuda")
try:
    print("affinity", len(os.sched_getaffinity(0)), "OMP_NUM_THREADS", os.environ.get("OMP_NUM_THREADS"), "torch threads", torch.get_num_threads(), flush=True)
    print("cpu.max", open("/sys/fs/cgroup/cpu.max").read().strip(), flush=True)
except Exception as e:
    print("cgroup:", e)
for groups in (16, 8):
    b200ctc.ctc_host(x_host, lab, 0, il, ll, reduce="mean", grad_out=g_host, groups=groups)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        b200ctc.ctc_host(x_host, lab, 0, il, ll, reduce="mean", grad_out=g_host, groups=groups)
        ts.append(time.perf_counter() - t0)
    print("groups=%2d  e2e step wall %s ms" % (groups, " ".join("%.2f" % (t * 1e3) for t in ts)), flush=True)

"""LayerNormalization -> loss: the fused path (csrc/layernorm_loss.cu) against the unfused pipelines it replaces, on the
bench workload (B=64, T=800, V=3500, L<=80, variable lengths).  Device time, CUDA events, inputs resident in HBM.

  fused            b200ctc.layernorm_ctc(z, gamma, beta, ...): z (B,V,1,T) read in place, dz written in place
  reference-style  what the reference's model tail does, op by op (asr/nn/layernorm.py:33-46, asr/nn/nn.py:265,
                   asr/model/cnn.py:41-44) as torch ops, then b200ctc.ctc on the transposed copy
  torch-native     torch.nn.functional.layer_norm over a (B,T,V) view (one transposing copy + one fused LN kernel),
                   then b200ctc.ctc(batch_first=True)

    python tools/ln_bench.py > profiles/r2_layernorm_fusion.txt
"""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import b200ctc
synth = importlib.import_module("chainer-speech-recognition_b200.synth")


def measure(B=64, T=800, V=3500, L=80, steps=10, warmup=3):
    dev = torch.device("cuda:0")
    rs = np.random.RandomState(0)
    in_len, lab_len = synth.make_lengths(rs, B, T, L)
    labels = synth.make_ctc_labels(rs, B, L, V, lab_len)
    g = torch.Generator(device=dev); g.manual_seed(0)
    z = (torch.randn((B, V, 1, T), device=dev, generator=g) * 1.7 + 0.3).requires_grad_(True)
    gamma = (1 + 0.2 * torch.randn(V, device=dev, generator=g)).requires_grad_(True)
    beta = (0.1 * torch.randn(V, device=dev, generator=g)).requires_grad_(True)
    lab = torch.tensor(labels, device=dev); il = torch.tensor(in_len, device=dev); ll = torch.tensor(lab_len, device=dev)

    def fused():
        return b200ctc.layernorm_ctc(z, gamma, beta, lab, 0, il, ll, reduce="mean")

    def reference_style():
        x = z
        mean = x.mean(dim=(1, 2), keepdim=True)
        diff = x - mean
        std = torch.sqrt((diff * diff).sum(dim=(1, 2), keepdim=True) / V)
        y = diff / std * gamma[None, :, None, None] + beta[None, :, None, None]
        out = y.swapaxes(1, 3).reshape(B, -1)                               # the transposed copy
        xs = out.view(B, T, V).transpose(0, 1)                                # split_axis: T views (B,V) = a (T,B,V) view
        return b200ctc.ctc(xs, lab, 0, il, ll, reduce="mean")

    def torch_native():
        y = torch.nn.functional.layer_norm(z.squeeze(2).transpose(1, 2), (V,), gamma, beta, eps=0.0)   # (B,T,V)
        return b200ctc.ctc(y, lab, 0, il, ll, reduce="mean", batch_first=True)

    out = {}
    for name, fn in (("fused", fused), ("reference_style", reference_style), ("torch_native", torch_native)):
        for _ in range(warmup):
            z.grad = gamma.grad = beta.grad = None
            fn().backward()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            z.grad = gamma.grad = beta.grad = None
            loss = fn()
            loss.backward()
        e1.record()
        torch.cuda.synchronize()
        out[name] = {"ms_per_step": e0.elapsed_time(e1) / steps, "loss": float(loss)}
        if name == "fused":                       # forward and backward separately
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            z.grad = gamma.grad = beta.grad = None
            ev[0].record(); loss = fn(); ev[1].record(); loss.backward(); ev[2].record()
            torch.cuda.synchronize()
            out[name]["fwd_ms"] = ev[0].elapsed_time(ev[1]); out[name]["bwd_ms"] = ev[1].elapsed_time(ev[2])
    valid = float(in_len.sum())
    # algorithmic bytes: fused = read z twice (valid frames) + write dz once; the unfused pipelines add LN forward
    # (read z, write y), the transposed copy (read + write), its backward (read + write) and LN backward (read g, read z, write dz)
    out["fused"]["algorithmic_bytes"] = 8 * V * valid + 4 * V * B * T
    out["reference_style"]["algorithmic_bytes"] = out["fused"]["algorithmic_bytes"] + 4 * V * B * T * (2 + 2 + 2 + 3)
    out["frames"] = B * T
    return out


if __name__ == "__main__":
    r = measure()
    print(json.dumps(r))
    f = r["fused"]["ms_per_step"]
    for k in ("fused", "reference_style", "torch_native"):
        print("# %-16s %8.3f ms per step   %6.2fx the fused path   loss %.4f" % (k, r[k]["ms_per_step"], r[k]["ms_per_step"] / f, r[k]["loss"]))

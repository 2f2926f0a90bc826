"""Helpers shared by the parity tests: run the CUDA path and the oracle on the same NumPy inputs."""
import numpy as np

LOSS_RTOL = 1e-5      # BASELINE.json north_star: per-utterance loss within 1e-5 relative
GRAD_ATOL = 1e-5      # gradients within 1e-5 absolute (against the float64 oracle)


def run_cuda(pkg, prob, kind, reduce="no", gy=None, as_list=False, batch_first=False, want_argmax=False,
             batch_global=None):
    import torch
    dev = torch.device("cuda:0")
    x = torch.tensor(prob["x"], device=dev)                      # (T,B,V)
    if batch_first:
        xin = x.transpose(0, 1).contiguous().requires_grad_(True)    # (B,T,V)
        leaf = xin
    elif as_list:
        leaf = x.requires_grad_(True)
        xin = [leaf[t] for t in range(leaf.shape[0])]
    else:
        xin = x.requires_grad_(True)
        leaf = xin
    labels = torch.tensor(prob["labels"], device=dev)
    il = None if prob.get("input_length") is None else torch.tensor(prob["input_length"], device=dev)
    ll = None if prob.get("label_length") is None else torch.tensor(prob["label_length"], device=dev)
    kw = {}
    if batch_first:
        kw["batch_first"] = True
    if want_argmax:
        kw["return_argmax"] = True
    if batch_global is not None:
        kw["batch_global"] = batch_global
    if kind == "ctc":
        out = pkg.connectionist_temporal_classification(xin, labels, prob["blank"], il, ll, reduce=reduce, **kw)
    else:
        big = torch.tensor(prob["bigrams"], device=dev)
        if kind == "joint":
            kw["joint_ctc"] = True
        out = pkg.gram_ctc(xin, labels, big, prob["blank"], il, ll, reduce=reduce, **kw)
    amax = None
    if want_argmax:
        out, amax = out
    if gy is None:
        g = torch.ones_like(out)
    else:
        g = torch.tensor(np.asarray(gy, np.float32), device=dev).reshape(out.shape)
    out.backward(g)
    grad = leaf.grad
    if batch_first:
        grad = grad.transpose(0, 1)
    torch.cuda.synchronize()
    return (out.detach().cpu().numpy().astype(np.float64), grad.detach().cpu().numpy(),
            None if amax is None else amax.cpu().numpy())


def run_oracle(prob, kind, want_argmax=False, nthreads=0):
    from oracle import c_oracle
    if kind == "joint":        # run/gram_ctc/cnn/train.py:196-198: Gram-CTC loss + plain CTC loss on the same activations
        lg, gg, am = run_oracle(prob, "gram", want_argmax, nthreads)
        lc, gc, _ = run_oracle(prob, "ctc", False, nthreads)
        return lg + lc, gg + gc, am
    r = c_oracle.run(0 if kind == "ctc" else 1, prob["x"], prob["labels"], prob.get("bigrams"),
                     prob.get("input_length"), prob.get("label_length"), prob["blank"],
                     want_grad=True, want_argmax=want_argmax, nthreads=nthreads)
    return r["loss"], r["grad"], r["argmax"]


def assert_parity(loss, grad, loss_ref, grad_ref, what=""):
    loss = np.asarray(loss, np.float64)
    # relative 1e-5; losses below 1 nat are held to 1e-5 absolute (no fp32 forward pass -- the reference's
    # included -- resolves a sum of T log-probabilities finer than that)
    rel = np.abs(loss - loss_ref) / np.maximum(np.abs(loss_ref), 1.0)
    assert np.all(np.isfinite(loss)), "%s non-finite loss %r" % (what, loss)
    assert rel.max() <= LOSS_RTOL, "%s loss rel err %.3e (%r vs %r)" % (what, rel.max(), loss, loss_ref)
    err = np.abs(grad.astype(np.float64) - grad_ref.astype(np.float64)).max()
    assert err <= GRAD_ATOL, "%s grad abs err %.3e" % (what, err)
    return rel.max(), err

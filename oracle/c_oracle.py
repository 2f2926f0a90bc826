"""ctypes front end of oracle/ctc_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libctc_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "ctc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libctc_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        l = ctypes.CDLL(_SO)
        l.ctc_oracle_run.restype = ctypes.c_int
        l.ctc_oracle_run.argtypes = [
            ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        l.ctc_oracle_max_threads.restype = ctypes.c_int
        _lib = l
    return _lib


def max_threads():
    return int(lib().ctc_oracle_max_threads())


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def run(kind, x_tbv, labels, bigrams, input_length, label_length, blank=0, want_grad=True,
        grad_scale=None, want_gamma=False, want_argmax=False, nthreads=0):
    """x_tbv: (T,B,V) float32 C-contiguous.  Returns dict(loss (B,) f64, grad (T,B,V) f32, gamma, argmax)."""
    x = np.ascontiguousarray(x_tbv, np.float32)
    T, B, V = x.shape
    labels = np.ascontiguousarray(labels, np.int32).reshape(B, -1)
    Lmax = labels.shape[1]
    if kind == 1:
        bigrams = np.ascontiguousarray(bigrams, np.int32).reshape(B, Lmax)
    else:
        bigrams = None
    il = None if input_length is None else np.ascontiguousarray(input_length, np.int32)
    ll = None if label_length is None else np.ascontiguousarray(label_length, np.int32)
    loss = np.zeros(B, np.float64)
    grad = np.empty((T, B, V), np.float32) if want_grad else None
    gs = None if grad_scale is None else np.ascontiguousarray(grad_scale, np.float64)
    Nmax = (2 if kind == 0 else 3) * Lmax + 1
    gamma = np.empty((B, T, Nmax), np.float64) if want_gamma else None
    amax = np.empty((B, T), np.int64) if want_argmax else None
    rc = lib().ctc_oracle_run(kind, _ptr(x), B * V, V, B, T, V, _ptr(labels), _ptr(bigrams), Lmax,
                              _ptr(il), _ptr(ll), int(blank), _ptr(loss), _ptr(grad), _ptr(gs),
                              _ptr(gamma), _ptr(amax), int(nthreads))
    if rc != 0:
        raise ValueError("ctc_oracle_run: invalid argument (status %d)" % rc)
    return {"loss": loss, "grad": grad, "gamma": gamma, "argmax": amax}

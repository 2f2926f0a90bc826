"""``asr.error`` -- drop-in for the reference's ``asr/error.py``: the character error rate of a development
batch, computed on the device (csrc/greedy_error.cu through include/b200ctc.h).

Same names and argument order as the reference:

    compute_minibatch_error(y_batch, t_batch, BLANK, vocab_token_to_id, vocab_id_to_token, print_sequences=False)
        (asr/error.py:26-68; called from run/ctc/cnn/train.py:233 with y_batch = argmax over the vocabulary)
    compute_character_error_rate(r, h)                                   (asr/error.py:7-24)

``y_batch`` is the (B,T) greedy index array (``greedy_argmax``), ``t_batch`` the (B,L) padded targets; both may
be CUDA tensors (nothing leaves the device except the final scalar) or host arrays (copied over).  The
reference's string round trip -- predicted ids -> token strings -> ``convert_sentence_to_unigram_ids``
(asr/error.py:49-53) -- becomes a table id -> unigram ids built once per vocabulary (``build_expansion_table``).
There is no CPU path: without a CUDA device these functions raise.
"""
import numpy as np
import torch

from .. import _lib

_tables = {}


def _default_convert(sentence, vocab_token_to_id):
    """The reference's own tokeniser when this package sits in the reference tree (asr/vocab.py:99-126);
    otherwise one id per character, which is the same thing for vocabularies without small kana."""
    try:
        from asr.vocab import convert_sentence_to_unigram_ids          # the reference's
        return convert_sentence_to_unigram_ids(sentence, vocab_token_to_id)
    except ImportError:
        return [vocab_token_to_id[ch] for ch in sentence]


def build_expansion_table(vocab_token_to_id, vocab_id_to_token, convert=None, device=None):
    """(V, E) int32 tensor: row i = the unigram ids token i expands to, padded with -1.

    Follows asr/error.py:49-53 per token: ``vocab_id_to_token[i]`` is re-tokenised with ``convert`` (default: the
    reference's ``convert_sentence_to_unigram_ids``).  Doing this per token equals doing it per sentence because no
    vocabulary token starts with a small kana (asr/vocab.py:3-36), so nothing attaches across a token boundary
    (:117-124).  Ids missing from ``vocab_id_to_token`` expand to nothing."""
    convert = convert or _default_convert
    V = max(vocab_id_to_token.keys()) + 1 if len(vocab_id_to_token) else 1
    rows = [[] for _ in range(V)]
    for i, token in vocab_id_to_token.items():
        if 0 <= i < V:
            rows[i] = [int(u) for u in convert(token, vocab_token_to_id)]
    E = max(1, max(len(r) for r in rows))
    table = np.full((V, E), -1, dtype=np.int32)
    for i, r in enumerate(rows):
        table[i, :len(r)] = r
    t = torch.from_numpy(table)
    return t.to(device) if device is not None else t


def _cached_table(vocab_token_to_id, vocab_id_to_token, device):
    key = (id(vocab_token_to_id), id(vocab_id_to_token), len(vocab_id_to_token), str(device))
    t = _tables.get(key)
    if t is None:
        t = _tables[key] = build_expansion_table(vocab_token_to_id, vocab_id_to_token, device=device)
    return t


def _device_of(*xs):
    for x in xs:
        if isinstance(x, torch.Tensor) and x.is_cuda:
            return x.device
    if not torch.cuda.is_available():
        raise RuntimeError("b200ctc has no CPU path: compute_minibatch_error needs a CUDA device")
    return torch.device("cuda", torch.cuda.current_device())


def _to_device(a, dtype, dev):
    if isinstance(a, torch.Tensor):
        return a.to(device=dev, dtype=dtype).contiguous()
    return torch.as_tensor(np.asarray(a), dtype=dtype, device=dev).contiguous()


def minibatch_error_details(y_batch, t_batch, BLANK, expansion, input_length=None, uint8_wrap=False):
    """Device-side result of one batch: dict with ``error`` (0-d float64 tensor, the batch mean), ``errors`` (B),
    ``distance``, ``ref_len``, ``hyp_len`` (B, int32) and ``hyp`` (B, T*E int32, valid up to hyp_len)."""
    dev = _device_of(y_batch, t_batch, expansion)
    y = _to_device(y_batch, torch.int64, dev)
    t = _to_device(t_batch, torch.int32, dev)
    exp = _to_device(expansion, torch.int32, dev)
    if y.dim() != 2 or t.dim() != 2 or y.shape[0] != t.shape[0] or exp.dim() != 2:
        raise ValueError("y_batch must be (B,T), t_batch (B,L) and expansion (V,E)")
    B, T = y.shape
    Lmax = t.shape[1]
    V, E = exp.shape
    il = _to_device(input_length, torch.int32, dev) if input_length is not None else None
    hyp = torch.empty((B, max(T * E, 1)), dtype=torch.int32, device=dev)
    hyp_len = torch.empty(B, dtype=torch.int32, device=dev)
    ref_len = torch.empty(B, dtype=torch.int32, device=dev)
    dist = torch.empty(B, dtype=torch.int32, device=dev)
    errs = torch.empty(B, dtype=torch.float64, device=dev)
    mean = torch.zeros((), dtype=torch.float64, device=dev)
    ws = torch.empty(_lib.ERROR_WORKSPACE_BYTES, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().b200ctc_greedy_error(
            y.data_ptr(), il.data_ptr() if il is not None else None, B, T, t.data_ptr(), Lmax, int(BLANK),
            exp.data_ptr(), V, E, 1 if uint8_wrap else 0, hyp.data_ptr(), hyp_len.data_ptr(), ref_len.data_ptr(),
            dist.data_ptr(), errs.data_ptr(), mean.data_ptr(), ws.data_ptr(), ws.numel(),
            torch.cuda.current_stream(dev).cuda_stream))
    return {"error": mean, "errors": errs, "distance": dist, "ref_len": ref_len, "hyp_len": hyp_len, "hyp": hyp}


def compute_minibatch_error(y_batch, t_batch, BLANK, vocab_token_to_id, vocab_id_to_token, print_sequences=False,
                            expansion=None, input_length=None, uint8_wrap=False):
    """Mean character error rate of the batch as a Python float (asr/error.py:26-68).

    ``expansion`` may be passed instead of the two vocabulary dicts (see ``build_expansion_table``).
    ``uint8_wrap=True`` reproduces the reference's numpy.uint8 distance table (asr/error.py:10), which wraps at 256;
    the default is the true edit distance (identical whenever every distance is below 256)."""
    if len(y_batch) == 0:
        raise ZeroDivisionError("division by zero")            # what `sum_error / len(y_batch)` raises (:68)
    dev = _device_of(y_batch, t_batch)
    if expansion is None:
        expansion = _cached_table(vocab_token_to_id, vocab_id_to_token, dev)
    out = minibatch_error_details(y_batch, t_batch, BLANK, expansion, input_length, uint8_wrap)
    if print_sequences and vocab_id_to_token is not None:       # :57-66
        hyp, hl = out["hyp"].cpu().numpy(), out["hyp_len"].cpu().numpy()
        tb = t_batch.cpu().numpy() if isinstance(t_batch, torch.Tensor) else np.asarray(t_batch)
        for b in range(len(hl)):
            print("#{}".format(b + 1))
            print("pred:\t" + "".join(vocab_id_to_token[int(i)] for i in hyp[b, :hl[b]]))
            print("true:\t" + "".join(vocab_id_to_token[int(i)] for i in tb[b] if int(i) != BLANK))
    return float(out["error"].item())


def compute_character_error_rate(r, h, uint8_wrap=False):
    """Edit distance of two id sequences divided by len(r); len(h) when r is empty (asr/error.py:7-24)."""
    dev = _device_of()
    R, H = len(r), len(h)
    ref = torch.as_tensor(np.asarray(list(r), dtype=np.int32).reshape(1, R), device=dev)
    hyp = torch.as_tensor(np.asarray(list(h), dtype=np.int32).reshape(1, H), device=dev)
    rl = torch.tensor([R], dtype=torch.int32, device=dev)
    hl = torch.tensor([H], dtype=torch.int32, device=dev)
    dist = torch.empty(1, dtype=torch.int32, device=dev)
    err = torch.empty(1, dtype=torch.float64, device=dev)
    ws = torch.empty(_lib.ERROR_WORKSPACE_BYTES, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().b200ctc_edit_distance(
            ref.data_ptr() if R else None, rl.data_ptr(), R, hyp.data_ptr() if H else None, hl.data_ptr(), H, 1,
            1 if uint8_wrap else 0, dist.data_ptr(), err.data_ptr(), None, ws.data_ptr(), ws.numel(),
            torch.cuda.current_stream(dev).cuda_stream))
    v = float(err.item())
    return int(v) if R == 0 else v                              # the reference returns the int len(h) here (:9)

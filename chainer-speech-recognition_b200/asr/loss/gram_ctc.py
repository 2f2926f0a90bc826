"""Gram-CTC loss -- drop-in for the reference's ``asr/loss/gram_ctc.py``.

Same call surface as ``gram_ctc`` (gram_ctc.py:300): activations as a sequence of T (B,V) arrays
(or one (T,B,V) / (B,T,V) tensor), padded unigram and bigram label ids, blank id, per-utterance
input and label lengths, ``reduce`` in {'mean','no'}.  The arithmetic runs in the sm_100a kernels of
``csrc/`` through the C ABI; there is no NumPy/CPU branch.
"""
from ... import _lib
from ._function import lattice_loss


def gram_ctc(xs, label_unigram, label_bigram, blank_symbol, input_length=None, length_unigram=None,
             reduce='mean', joint_ctc=False, **kw):
    """Reference: asr/loss/gram_ctc.py:300-315.  Returns a 0-d tensor ('mean') or a (B,) tensor ('no').

    Extra keyword arguments (not in the reference): ``batch_first`` for a single (B,T,V) tensor,
    ``batch_global``/``group`` for batch-sharded multi-GPU use, ``return_argmax``, and ``joint_ctc``:

    ``joint_ctc=True`` returns ``gram_ctc(...) + connectionist_temporal_classification(xs, label_unigram, ...)``,
    the joint-training objective of run/gram_ctc/cnn/train.py:196-198 (``args.joint_training``), from ONE pass:
    both lattices run on the same softmax statistics and emission rows, and one gradient kernel writes the
    gradient of the sum -- 12 instead of 24 bytes of HBM traffic per activation element.
    """
    kind = _lib.KIND_JOINT if joint_ctc else _lib.KIND_GRAM
    return lattice_loss(kind, xs, label_unigram, label_bigram, blank_symbol, input_length,
                        length_unigram, reduce, **kw)


def joint_gram_ctc(xs, label_unigram, label_bigram, blank_symbol, input_length=None, length_unigram=None,
                   reduce='mean', **kw):
    """``gram_ctc(...) + connectionist_temporal_classification(...)`` in one pass (see ``gram_ctc``)."""
    return gram_ctc(xs, label_unigram, label_bigram, blank_symbol, input_length, length_unigram, reduce,
                    joint_ctc=True, **kw)


class GramCTC(object):
    """Object form mirroring ``GramCTC(blank_symbol, reduce)(input_length, length_unigram,
    label_unigram, label_bigram, *xs)`` (gram_ctc.py:219-228, :315)."""

    def __init__(self, blank_symbol, reduce='mean'):
        if reduce not in ('mean', 'no'):
            raise ValueError("only 'mean' and 'no' are valid for 'reduce', but '%s' is given" % reduce)
        self.blank_symbol = blank_symbol
        self.reduce = reduce
        self.zero_padding = -10000000000.0

    def __call__(self, input_length, length_unigram, label_unigram, label_bigram, *xs):
        return gram_ctc(list(xs), label_unigram, label_bigram, self.blank_symbol, input_length,
                        length_unigram, self.reduce)

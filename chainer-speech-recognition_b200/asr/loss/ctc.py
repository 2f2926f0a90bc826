"""Plain CTC loss -- drop-in for ``chainer.functions.connectionist_temporal_classification`` as the
reference calls it (run/ctc/cnn/train.py:162,191; run/ctc/sru/train.py:161,191;
run/gram_ctc/cnn/train.py:166,198): ``F.connectionist_temporal_classification(y_batch, t_batch,
ID_BLANK, x_length_batch, t_length_batch)``.

The arithmetic is the 2L+1 blank-interleaved lattice, i.e. the reference's own Gram-CTC lattice with
every bigram node disconnected (asr/loss/gram_ctc.py:94-98); see SURVEY.md section 8c for why that
identity is the parity anchor.
"""
from ... import _lib
from ._function import lattice_loss


def connectionist_temporal_classification(x, t, blank_symbol, input_length=None, label_length=None,
                                          reduce='mean', **kw):
    """x: sequence of T (B,V) float32 CUDA tensors (or one (T,B,V) tensor); t: (B,Lmax) int labels."""
    return lattice_loss(_lib.KIND_CTC, x, t, None, blank_symbol, input_length, label_length, reduce, **kw)


ctc = connectionist_temporal_classification


class ConnectionistTemporalClassification(object):
    def __init__(self, blank_symbol, reduce='mean'):
        if reduce not in ('mean', 'no'):
            raise ValueError("only 'mean' and 'no' are valid for 'reduce', but '%s' is given" % reduce)
        self.blank_symbol = blank_symbol
        self.reduce = reduce

    def __call__(self, input_length, label_length, t, *xs):
        return connectionist_temporal_classification(list(xs), t, self.blank_symbol, input_length,
                                                     label_length, self.reduce)

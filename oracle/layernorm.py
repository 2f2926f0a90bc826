"""NumPy restatement of the model tail in front of the loss: LayerNormalization -> swapaxes/reshape/split.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  float64 by default.

What it follows in the reference (``/root/reference``):
  * the last layer of every CTC / Gram-CTC model is ``nn.Convolution2D(ndim_dense, vocab_size, ksize=1)`` followed by
    ``nn.LayerNormalization(None)``                                           run/ctc/cnn/model.py:85-88
  * ``LayerNormalization.__call__``: ``bias(scale(normalize_layer(x), gamma), beta)`` with gamma, beta of size
    ``x.shape[1]`` (the vocabulary axis), broadcast from axis 1                asr/nn/nn.py:240-265
  * ``NormalizeLayer.forward``: x is (B, V, 1, T); mean and std over axes (1, 2), i.e. PER FRAME over the vocabulary;
    ``std = sqrt(sum(diff**2) / size)`` -- no epsilon is added (the ``eps`` argument is stored and never used);
    returns ``diff / std``                                                     asr/nn/layernorm.py:33-46
  * ``NormalizeLayer.backward``                                                asr/nn/layernorm.py:48-60
  * ``AcousticModel.__call__`` hands the loss a tuple of T arrays (B, V): ``swapaxes(1, 3)``, ``reshape``,
    ``split_axis`` -- a transposed copy of the whole tensor                    asr/model/cnn.py:41-44

Here z is (B, V, T) (the height axis of size 1 dropped).  ``forward`` returns the activations as the loss sees them,
(T, B, V), plus what backward needs; ``backward`` maps d loss / d activations (T, B, V) to (dz, dgamma, dbeta).
"""
import numpy as np


def forward(z, gamma, beta, dtype=np.float64):
    z = np.asarray(z, dtype)
    gamma = np.asarray(gamma, dtype)
    beta = np.asarray(beta, dtype)
    B, V, T = z.shape
    mean = z.mean(axis=1, keepdims=True)                              # layernorm.py:41  (axes (1,2) of (B,V,1,T))
    diff = z - mean                                                   # :43
    std = np.sqrt((diff ** 2).sum(axis=1, keepdims=True) / V)         # :44
    n = diff / std                                                    # :46
    y = n * gamma[None, :, None] + beta[None, :, None]                # nn.py:265  scale, then bias, along axis 1
    acts = np.ascontiguousarray(y.transpose(2, 0, 1))                 # cnn.py:42-44: frame t = (B, V) slice
    return acts, (n, std, gamma)


def backward(grad_acts, saved, dtype=np.float64):
    """grad_acts: (T, B, V) = d loss / d activations.  Returns dz (B, V, T), dgamma (V,), dbeta (V,)."""
    n, std, gamma = saved
    gy = np.asarray(grad_acts, dtype).transpose(1, 2, 0)              # back through split/reshape/swapaxes: (B, V, T)
    dbeta = gy.sum(axis=(0, 2))                                       # bias backward: sum over the broadcast axes
    dgamma = (gy * n).sum(axis=(0, 2))                                # scale backward
    dn = gy * gamma[None, :, None]
    # NormalizeLayer.backward (:48-60), written out: with s = std, d = diff = n * s
    #   std_grad  = sum_v(-dn * d / s^2)                      :51
    #   var_grad  = std_grad * 0.5 / s / V * 2 * d + dn / s   :52-54   = (dn - n * mean_v(dn * n)) / s
    #   mean_grad = sum_v(-var_grad) / V                      :56-57
    #   result    = var_grad + mean_grad                      :59
    var_grad = (dn - n * (dn * n).mean(axis=1, keepdims=True)) / std
    dz = var_grad - var_grad.mean(axis=1, keepdims=True)
    return dz, dgamma, dbeta

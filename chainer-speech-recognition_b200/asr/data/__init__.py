"""``asr.data`` -- only the label half of minibatch assembly (reference: asr/data/processing.py:113-171)."""
from .processing import labels_to_minibatch                                   # noqa: F401

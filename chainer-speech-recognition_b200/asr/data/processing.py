"""Label half of the reference's minibatch assembly, ``Processor.features_to_minibatch``
(asr/data/processing.py:113-171): the int32 arrays the loss consumes.

    t_batch        (B, Lmax) unigram ids, padded with the blank id                      (:125, :168)
    bigram_batch   (B, Lmax) per-position bigram ids: position 0 and bigrams that are not in the inventory are -1,
                   padding is the blank id                                              (:126, :139-146, :169)
    t_length_batch per-utterance label lengths after the CTC feasibility cut            (:159-166, :170)

The feasibility cut is the reference's: a transcription that cannot be aligned to its ``x_length`` frames
(2*len + 1 + repeated neighbours, the neighbour test being the reference's circular ``np.roll`` comparison) is
truncated to ``(x_length - repeats - 1) // 2`` ids.  Everything is assembled straight into ONE pinned host block,
so that a single asynchronous copy moves the labels, bigrams and both length vectors to the device (the
reference issues four ``cuda.to_gpu`` calls per batch, asr/data/loaders/base.py:28-31).  Host-side code: there is
nothing here for the GPU to do.

``sort_by_length=True`` lays the batch out longest utterance first (stable) and also returns the permutation, which
the caller applies to the feature batch: the lattice kernel starts its CTAs in batch order, so when a batch has
more utterances than the GPU can hold lattice CTAs at once (B > ~2 x #SMs) the long ones start first and the short
ones fill the tail (longest-processing-time-first).  The node -> vocabulary-slot tables of the gradient kernel are
built on the device inside the forward call (csrc/prep.cuh), not here.
"""
import numpy as np
import torch


def _default_tokenizer(sentence):
    """The reference's own tokeniser when this package sits in the reference tree (asr/vocab.py:107-126);
    otherwise one token per character."""
    try:
        from asr.vocab import convert_sentence_to_unigram_tokens      # the reference's
        return convert_sentence_to_unigram_tokens(sentence)
    except ImportError:
        return list(sentence)


def labels_to_minibatch(sentences, x_length_batch, max_sentence_length, token_ids, id_blank, tokenizer=None,
                        device=None, pin=True, sort_by_length=False):
    """Returns ``(t_batch, bigram_batch, x_length_batch, t_length_batch)`` as int32 torch tensors -- views of one
    (pinned) host block, or of its device copy when ``device`` is given.  ``sentences``: transcriptions (strings, run
    through ``tokenizer``) or ready lists of unigram tokens.  With ``sort_by_length`` a fifth value follows: ``order``
    (int64, host), row i of every output belongs to ``sentences[order[i]]``."""
    assert isinstance(token_ids, dict)                                 # :114
    assert isinstance(id_blank, int)                                   # :115
    tokenizer = tokenizer or _default_tokenizer
    B, Lmax = len(sentences), int(max_sentence_length)
    assert len(x_length_batch) == B
    use_pin = pin and torch.cuda.is_available()
    block = torch.empty(2 * B * Lmax + 2 * B, dtype=torch.int32, pin_memory=use_pin)
    host = block.numpy()
    t_batch = host[:B * Lmax].reshape(B, Lmax)
    bigram_batch = host[B * Lmax:2 * B * Lmax].reshape(B, Lmax)
    x_len = host[2 * B * Lmax:2 * B * Lmax + B]
    t_len = host[2 * B * Lmax + B:]
    t_batch[...] = id_blank                                            # :125
    bigram_batch[...] = id_blank                                       # :126
    order = np.argsort(-np.asarray(x_length_batch, dtype=np.int64), kind="stable") if sort_by_length else np.arange(B)
    for row, src in enumerate(order):
        sentence = sentences[src]
        tokens = tokenizer(sentence) if isinstance(sentence, str) else list(sentence)      # :132
        unigram_ids = [token_ids[tok] for tok in tokens]               # :140-141
        bigram_ids = [-1] + [token_ids.get(first + second, -1) for first, second in zip(tokens[:-1], tokens[1:])]   # :139-146
        x_length = int(x_length_batch[src])
        t_length = len(unigram_ids)
        # CTC feasibility (:159-166).  np.roll makes the neighbour test circular: id 0 is compared with the last id.
        repeats = int(np.count_nonzero(np.asarray(unigram_ids) == np.roll(unigram_ids, 1))) if t_length else 0
        if x_length < t_length * 2 + 1 + repeats:
            possible = (x_length - repeats - 1) // 2
            unigram_ids = unigram_ids[:possible]                       # Python slice semantics, negative values included
            bigram_ids = bigram_ids[:possible]
            t_length = len(unigram_ids)
        t_batch[row, :t_length] = unigram_ids                            # :168
        bigram_batch[row, :t_length] = bigram_ids                        # :169
        x_len[row] = x_length
        t_len[row] = t_length
    out = block if device is None else block.to(device, non_blocking=True)
    res = (out[:B * Lmax].view(B, Lmax), out[B * Lmax:2 * B * Lmax].view(B, Lmax),
           out[2 * B * Lmax:2 * B * Lmax + B], out[2 * B * Lmax + B:])
    return res + (torch.from_numpy(np.ascontiguousarray(order, dtype=np.int64)),) if sort_by_length else res

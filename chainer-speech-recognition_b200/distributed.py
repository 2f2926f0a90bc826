"""Batch sharding across the GPUs of one box.

Utterances are independent (every reference op is per-b: asr/loss/gram_ctc.py:155,173,191), so rank r of G
owns the contiguous utterance range ``shard_range(B, r, G)`` -- activations, labels, workspace and
gradients stay rank-local.  The only exchange is ONE float for ``reduce='mean'`` (gram_ctc.py:281): each
rank's kernel already scales its partial sum by 1/B_global, the ranks all-reduce (sum) that scalar over
NCCL.  The backward pass needs no collective: its scale 1/B_global (:292) is known without communication.
"""


def shard_range(batch, rank, world):
    """[start, stop) of the utterances rank ``rank`` owns; sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world: %r/%r" % (rank, world))
    base, extra = divmod(int(batch), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(rank, world, *arrays, batch_axis=0):
    """Slice every array/tensor along its batch axis to this rank's shard (a tuple of (array, axis) selects an axis)."""
    out = []
    for a in arrays:
        axis = batch_axis
        if isinstance(a, tuple):
            a, axis = a
        if a is None:
            out.append(None)
            continue
        s, e = shard_range(a.shape[axis], rank, world)
        index = [slice(None)] * len(a.shape)
        index[axis] = slice(s, e)
        out.append(a[tuple(index)])
    return out


def reduce_mean_loss(local_scaled_sum, group=None):
    """all-reduce (sum) of the per-rank ``sum_b loss_b / B_global`` scalars -> the global batch mean."""
    import torch.distributed as dist
    dist.all_reduce(local_scaled_sum, op=dist.ReduceOp.SUM, group=group)
    return local_scaled_sum

"""Multi-GPU plumbing for the GA3C loop (one process per GPU, torch.distributed; NCCL on GPUs, gloo in CPU tests).

The env step itself needs no communication: worlds are independent, each rank owns the contiguous world range
`shard_range(W, rank, world_size)`.  The learner is the only exchange.  Two equivalent forms:
  * `gather_rows` + `broadcast_weights`: every rank's emitted rows (x_, r_, a_) are all-gathered (padded to the
    longest, then trimmed) so that the trainer rank sees exactly the concatenation the reference's training_q would
    deliver, and the updated weights are broadcast back;
  * `allreduce_flat`: each rank back-propagates its own rows and the sum-loss gradients are summed — the same
    update (the A3C loss is a sum over rows, GA3C/NetworkVPCore.py:71-98) with 0.68 MB instead of the rows on the wire.
"""
import torch
import torch.distributed as dist


def shard_range(num_worlds, rank, world_size):
    """Contiguous, balanced partition of [0, num_worlds): returns (start, stop) of this rank."""
    base, rem = divmod(int(num_worlds), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_rows(x, r, a, group=None):
    """All-gather variable-length row batches; returns the concatenation over ranks in rank order (on every rank)."""
    ws = dist.get_world_size(group)
    n = torch.tensor([x.shape[0]], dtype=torch.int64, device=x.device)
    counts = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)

    def padded(t):
        out = t.new_zeros((cap,) + tuple(t.shape[1:]))
        out[:t.shape[0]] = t
        return out

    outs = []
    for t in (x, r, a):
        bufs = [torch.empty((cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for _ in range(ws)]
        dist.all_gather(bufs, padded(t), group=group)
        outs.append(torch.cat([b[:c] for b, c in zip(bufs, counts)]))
    return tuple(outs)


def broadcast_weights(parameters, src=0, group=None):
    flat = torch.cat([p.data.reshape(-1) for p in parameters])
    dist.broadcast(flat, src=src, group=group)
    off = 0
    for p in parameters:
        n = p.numel()
        p.data.copy_(flat[off:off + n].view_as(p))
        off += n


def allreduce_flat(flat_grad, group=None):
    """Sum one flat gradient buffer over ranks in place (the parameters' .grad are views into it)."""
    dist.all_reduce(flat_grad, group=group)


def max_over_ranks(value, device, group=None):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())

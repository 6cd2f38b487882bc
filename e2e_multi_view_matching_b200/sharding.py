"""Host-side logic of the multi-GPU path: tuples are independent units, every rank works on its own
contiguous shard (weak scaling, no data-path collective); the only exchanges are one scalar all-reduce per step
(mirrors the reference's validation-loss all-reduce, train.py:104-106) and the max-over-ranks of the device time
for reporting.  Backend-agnostic (NCCL on the GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def tuple_shard(rank, world, tuples_per_rank, base=1000):
    """Tuple ids (= synthetic-input seeds) of this rank: [base + rank*B, base + (rank+1)*B)."""
    assert 0 <= rank < world and tuples_per_rank >= 1
    return list(range(base + rank * tuples_per_rank, base + (rank + 1) * tuples_per_rank))


def all_reduce_step_loss(loss):
    """Sum of the per-rank scalar step loss, in place; no-op without a process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(loss)
    return loss


def max_over_ranks(value, device='cpu'):
    """Largest `value` (e.g. milliseconds of the timed region) over all ranks."""
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_throughput(tuples_per_rank, steps, world, max_ms):
    """Tuples all ranks processed divided by the slowest rank's time."""
    return tuples_per_rank * steps * world / (max_ms * 1e-3)


def all_reduce_gradients(params, bucket_bytes=64 << 20):
    """Data-parallel TRAINING (cfg5): average the gradients of `params` over the ranks -- what the reference gets from
    DistributedDataParallel (train.py:350-352).  The gradients are packed into flat buckets (one all-reduce each: NCCL's
    ring / NVLS cost is launch latency + bytes, so a few large buckets instead of 539 small tensors) and averaged in
    place.  Parameters without a gradient on this rank contribute zeros, so every rank issues the same collectives.
    No-op without a process group.  -> number of all-reduce calls issued."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return 0
    world = dist.get_world_size()
    params = [p for p in params if p.requires_grad]
    calls, bucket, size = 0, [], 0

    def flush():
        nonlocal calls, bucket, size
        if not bucket:
            return
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in bucket])
        dist.all_reduce(flat)
        flat /= world
        o = 0
        for p in bucket:
            n = p.numel()
            if p.grad is None:
                p.grad = flat[o:o + n].view_as(p).clone()
            else:
                p.grad.copy_(flat[o:o + n].view_as(p))
            o += n
        calls += 1
        bucket, size = [], 0

    for p in params:
        bucket.append(p)
        size += p.numel() * p.element_size()
        if size >= bucket_bytes:
            flush()
    flush()
    return calls

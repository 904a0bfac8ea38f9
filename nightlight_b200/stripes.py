"""Row-stripe sharding of a stack across GPUs.

Every output pixel depends only on its own column of N samples, so GPU g of G owns the contiguous
rows [row0, row0+rows) of ALL frames and stacks them without any exchange.  The only collective is the
reassembly of the final image (one all-gather of the stripes) and a sum of the two clip counters.
The reference has no counterpart (single process, goroutines over pixel ranges, stack.go:134-147).
"""
from typing import List, Tuple


def stripe_rows(height: int, world: int, rank: int) -> Tuple[int, int]:
    """(row0, rows) of `rank`: the first height % world ranks own one extra row."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, extra = divmod(height, world)
    rows = base + (1 if rank < extra else 0)
    row0 = rank * base + min(rank, extra)
    return row0, rows


def all_stripes(height: int, world: int) -> List[Tuple[int, int]]:
    return [stripe_rows(height, world, r) for r in range(world)]


def allgather_image(local_stripe, width: int, height: int, group=None):
    """All-gathers the per-rank stripes (torch tensors, 1-D, rows*width floats) into the full image on
    every rank.  Stripes may differ by one row, so the gather is padded to the largest stripe."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    stripes = all_stripes(height, world)
    max_rows = max(r for _, r in stripes)
    pad = torch.empty(max_rows * width, dtype=local_stripe.dtype, device=local_stripe.device)
    pad[: local_stripe.numel()] = local_stripe
    gathered = torch.empty(world * max_rows * width, dtype=local_stripe.dtype, device=local_stripe.device)
    dist.all_gather_into_tensor(gathered, pad, group=group)
    if all(r == max_rows for _, r in stripes):
        return gathered[: height * width]
    out = torch.empty(height * width, dtype=local_stripe.dtype, device=local_stripe.device)
    for r, (row0, rows) in enumerate(stripes):
        out[row0 * width:(row0 + rows) * width] = gathered[r * max_rows * width: r * max_rows * width + rows * width]
    return out


def allreduce_clip_counts(clip_low: int, clip_high: int, device, group=None):
    import torch
    import torch.distributed as dist

    t = torch.tensor([clip_low, clip_high], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t[0].item()), int(t[1].item())

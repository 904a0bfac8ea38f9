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


class PeerGather:
    """The gathered image, allocated on every rank and peer-mapped on every other rank (CUDA IPC over
    NVLink / NVSwitch), so that the stack kernel's epilogue can store rank r's stripe straight into
    everybody's image (nl_stack_run_dev_bcast) -- the all-gather is fused into the producing kernel.
    Equal stripes of `stripe_px` pixels; rank r's stripe lives at element offset r*stripe_px."""

    def __init__(self, ctx, stripe_px, group=None):
        import torch.distributed as dist
        self.ctx, self.stripe_px, self.group = ctx, int(stripe_px), group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.buf = ctx.dev_alloc(4 * self.stripe_px * self.world)
        handles = [None] * self.world
        dist.all_gather_object(handles, ctx.ipc_handle(self.buf), group=group)
        self.peers = [self.buf if r == self.rank else ctx.ipc_open(h) for r, h in enumerate(handles)]
        off = 4 * self.rank * self.stripe_px
        self.local_out = self.buf + off
        self.peer_outs = [p + off for r, p in enumerate(self.peers) if r != self.rank]

    def to_host(self):
        import numpy as np
        out = np.empty(self.stripe_px * self.world, dtype=np.float32)
        self.ctx.d2h(out, self.buf)
        return out

    def close(self):
        import torch.distributed as dist
        self.ctx.sync()
        dist.barrier(group=self.group)            # nobody may still be storing into a buffer that goes away
        for r, p in enumerate(self.peers):
            if r != self.rank:
                self.ctx.ipc_close(p)
        self.ctx.dev_free(self.buf)

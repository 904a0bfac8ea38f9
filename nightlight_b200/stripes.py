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


def frame_shard(n_frames: int, world: int, rank: int) -> List[int]:
    """Frames that `rank` detects / aligns / resamples: round robin, like the reference hands frames to its worker
    goroutines (operator.go:80-98), so that early frames of every rank are in flight together."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    return list(range(rank, n_frames, world))


def alltoall_frames_to_stripes(local_frames, frame_ids, n_frames: int, width: int, height: int, group=None):
    """The plain-collective form of the exchange between the two shardings (the baseline PeerScatter is checked
    against): every rank holds whole resampled frames `frame_ids` (torch tensor [len(frame_ids), height*width]) and
    ends up with rows stripe_rows(height, world, rank) of ALL frames, frame-major: tensor [n_frames, rows*width].
    Works on gloo (CPU) and NCCL."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    stripes = all_stripes(height, world)
    sends = []
    for r, (row0, rows) in enumerate(stripes):
        sends.append(local_frames[:, row0 * width:(row0 + rows) * width].contiguous().reshape(-1))
    my_rows = stripes[rank][1]
    recvs = [torch.empty(len(frame_shard(n_frames, world, r)) * my_rows * width, dtype=local_frames.dtype,
                         device=local_frames.device) for r in range(world)]
    if dist.get_backend(group) == "nccl":
        dist.all_to_all(recvs, sends, group=group)
    else:                                   # gloo has no all-to-all: pairwise exchange
        recvs[rank].copy_(sends[rank])
        ops = []
        for r in range(world):
            if r != rank:
                if sends[r].numel():
                    ops.append(dist.P2POp(dist.isend, sends[r], r, group))
                if recvs[r].numel():
                    ops.append(dist.P2POp(dist.irecv, recvs[r], r, group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
    out = torch.empty(n_frames, my_rows * width, dtype=local_frames.dtype, device=local_frames.device)
    for r in range(world):
        ids = frame_shard(n_frames, world, r)
        if ids:
            out[ids] = recvs[r].reshape(len(ids), my_rows * width)
    return out


class PeerScatter:
    """Frame-sharded resample -> row-sharded stack without a staging image or an all-to-all pass.
    Every rank owns a stack job for its row stripe of all `n_frames` frames; the jobs' frame buffers are peer-mapped
    on every rank (CUDA IPC over NVLink / NVSwitch), and `project` lets the resample kernel store each destination
    row of a frame into the job that owns it (nl_project_scatter_dev).  After `finish` every job is complete."""

    def __init__(self, ctx, job_factory, n_frames, width, height, group=None):
        import torch.distributed as dist
        self.ctx, self.group = ctx, group
        self.n_frames, self.width, self.height = int(n_frames), int(width), int(height)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.stripes = all_stripes(self.height, self.world)
        self.row0, self.rows = self.stripes[self.rank]
        self.job = job_factory(ctx, self.n_frames, self.rows * self.width)
        base, stride = self.job.frames_dev
        assert stride == self.rows * self.width
        handles = [None] * self.world
        dist.all_gather_object(handles, ctx.ipc_handle(base), group=group)
        self.bases = [base if r == self.rank else ctx.ipc_open(h) for r, h in enumerate(handles)]

    def project(self, dev_src, src_w, src_h, frame_index, trans, out_of_bounds=float("nan"), multiplier=1.0, offset=0.0):
        """resample the device frame `dev_src` as frame `frame_index` of every rank's job (asynchronous)"""
        scatter_project(self.ctx, dev_src, src_w, src_h, self.width, self.height, trans, frame_index, self.bases,
                        [s[0] for s in self.stripes] + [self.height], out_of_bounds, multiplier, offset)

    def finish(self):
        """all stores of all ranks have landed in every job"""
        import torch.distributed as dist
        self.ctx.sync()
        dist.barrier(group=self.group)

    def close(self):
        import torch.distributed as dist
        self.ctx.sync()
        dist.barrier(group=self.group)
        for r, p in enumerate(self.bases):
            if r != self.rank:
                self.ctx.ipc_close(p)
        self.job.close()


def scatter_project(ctx, dev_src, src_w, src_h, dst_w, dst_h, trans, frame_index, stripe_bases, stripe_row0,
                    out_of_bounds=float("nan"), multiplier=1.0, offset=0.0):
    """nl_project_scatter_dev: stripe_bases[g] = frame-major device buffer of stripe g, stripe_row0 = G+1 row bounds"""
    import ctypes as C
    import numpy as np
    from .binding import check, load_library
    n = len(stripe_bases)
    bases = (C.c_void_p * n)(*[C.c_void_p(b) for b in stripe_bases])
    row0 = (C.c_int32 * (n + 1))(*[int(r) for r in stripe_row0])
    t = np.ascontiguousarray(trans, dtype=np.float32)
    check(load_library().nl_project_scatter_dev(ctx.handle, C.c_void_p(dev_src), int(src_w), int(src_h), int(dst_w), int(dst_h),
                                                t.ctypes.data_as(C.POINTER(C.c_float)), float(out_of_bounds), float(multiplier),
                                                float(offset), int(frame_index), bases, row0, n))

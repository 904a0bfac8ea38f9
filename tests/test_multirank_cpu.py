"""world_size-2 gloo test of the multi-GPU host logic: row-stripe partition, stripe all-gather with
ragged stripes, clip-counter all-reduce.  The per-stripe stack itself is played by the oracle here
(CPU box); on the GPU box tests/test_gpu_stack.py::test_stripes_equal_whole checks the CUDA path."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _worker(rank, world, port, width, height, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nightlight_b200.stripes import stripe_rows, allgather_image, allreduce_clip_counts
    from oracle import oracle as O
    row0, rows = stripe_rows(height, world, rank)
    frames = O.synth_frames(n, row0 * width, rows * width)
    res, cl, ch = O.stack(frames, "sigma")
    full = allgather_image(torch.from_numpy(res), width, height)
    tl, th = allreduce_clip_counts(cl, ch, "cpu")
    if rank == 0:
        q.put((full.numpy().copy(), tl, th))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_stripes_equal_whole():
    from oracle import oracle as O
    width, height, n, world = 37, 13, 12, 2          # ragged: 7 + 6 rows
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, width, height, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    full, tl, th = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want, cl, ch = O.stack(O.synth_frames(n, 0, width * height), "sigma")
    assert np.array_equal(full.view(np.uint32), want.view(np.uint32))
    assert (tl, th) == (cl, ch)


def _a2a_worker(rank, world, port, width, height, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nightlight_b200.stripes import alltoall_frames_to_stripes, frame_shard, stripe_rows
    from oracle import oracle as O
    ids = frame_shard(n, world, rank)
    # "resampled" frames of this rank's frame shard: the synthetic frames themselves
    local = torch.from_numpy(np.stack([O.synth_frame(0, width * height, k) for k in ids]))
    mine = alltoall_frames_to_stripes(local, ids, n, width, height)
    row0, rows = stripe_rows(height, world, rank)
    res, cl, ch = O.stack(mine.numpy(), "sigma")
    q.put((rank, row0, rows, res.copy(), cl, ch))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_frame_shards_to_row_stripes():
    """N4 host logic: frames resampled frame-sharded (rank r owns frames r, r+G, ...) reach the row-sharded
    stack through the all-to-all baseline; stripes stacked per rank equal the whole-image stack"""
    from oracle import oracle as O
    width, height, n, world = 29, 11, 9, 2            # ragged stripes (6 + 5 rows) and ragged frame shards (5 + 4)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_a2a_worker, args=(r, world, port, width, height, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want, cl, ch = O.stack(O.synth_frames(n, 0, width * height), "sigma")
    full = np.empty(width * height, np.float32)
    tl = th = 0
    for rank, row0, rows, res, l, h in got:
        full[row0 * width:(row0 + rows) * width] = res
        tl, th = tl + l, th + h
    assert np.array_equal(full.view(np.uint32), want.view(np.uint32))
    assert (tl, th) == (cl, ch)


def _c5_worker(rank, world, port, width, height, n, bs, q):
    """the multi-rank host logic of bench.py --config c5 with the oracle playing the per-stripe stack: batches from one
    permutation, the sigma goal-seek state machine of the C ABI stepped with clip totals SUMMED over the ranks, stack of
    stacks per stripe, one all-gather"""
    import ctypes as C
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import nightlight_b200 as nl
    from nightlight_b200 import binding
    from nightlight_b200.stripes import stripe_rows, allgather_image
    from oracle import oracle as O
    lib = nl.load_library()
    row0, rows = stripe_rows(height, world, rank)
    px = rows * width
    perm = [int(x) for x in np.random.default_rng(5).permutation(n)]
    nb = (n + bs - 1) // bs
    order = list(perm)
    for i in range(nb):
        order[i * bs:(i + 1) * bs] = sorted(order[i * bs:(i + 1) * bs])
    batches = [order[i * bs:(i + 1) * bs] for i in range(nb)]
    frames = {k: O.synth_frame(row0 * width, px, k) for k in range(n)}
    seek = binding.SigmaSeek()
    binding.check(lib.nl_sigma_seek_begin(C.byref(seek), 2, len(batches[0]), width * height, 2.0, 3.0))
    first = np.stack([frames[k] for k in batches[0]])
    trials = 0
    while not seek.done:
        _, cl, ch = O.stack(first, "sigma", seek.trial_low, seek.trial_high)
        t = torch.tensor([cl, ch], dtype=torch.int64)
        dist.all_reduce(t)
        assert lib.nl_sigma_seek_step(C.byref(seek), int(t[0]), int(t[1])) >= 0
        trials += 1
    acc = np.zeros(px, np.float32)
    fp = C.POINTER(C.c_float)
    for b, batch in enumerate(batches):
        res, _, _ = O.stack(np.stack([frames[k] for k in batch]), "sigma", seek.result_low, seek.result_high)
        O.lib().nlo_stack_incremental(acc.ctypes.data_as(fp), res.ctypes.data_as(fp), px, float(len(batch)), 1 if b == 0 else 0)
    O.lib().nlo_stack_incremental_finalize(acc.ctypes.data_as(fp), px, float(n))
    full = allgather_image(torch.from_numpy(acc), width, height)
    if rank == 0:
        q.put((full.numpy().copy(), float(seek.result_low), float(seek.result_high), trials, batches))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_batched_goal_seek_equals_one_rank():
    """c5's host logic on two ranks (row stripes) gives the image, the sigmas and the trial count of the same job on the
    whole image: the goal-seek sees the summed clip totals, the stack of stacks is per stripe"""
    import ctypes as C
    import nightlight_b200 as nl
    from nightlight_b200 import binding
    from oracle import oracle as O
    width, height, n, bs, world = 31, 9, 30, 12, 2     # 3 batches (12 + 12 + 6), ragged stripes (5 + 4 rows)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_c5_worker, args=(r, world, port, width, height, n, bs, q)) for r in range(world)]
    for p in procs:
        p.start()
    full, sl, sh, trials, batches = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # one rank, whole image
    lib = nl.load_library()
    frames = O.synth_frames(n, 0, width * height)
    seek = binding.SigmaSeek()
    binding.check(lib.nl_sigma_seek_begin(C.byref(seek), 2, len(batches[0]), width * height, 2.0, 3.0))
    t1 = 0
    while not seek.done:
        _, cl, ch = O.stack(frames[batches[0]], "sigma", seek.trial_low, seek.trial_high)
        lib.nl_sigma_seek_step(C.byref(seek), cl, ch)
        t1 += 1
    assert (sl, sh, trials) == (float(seek.result_low), float(seek.result_high), t1)
    fp = C.POINTER(C.c_float)
    acc = np.zeros(width * height, np.float32)
    for b, batch in enumerate(batches):
        res, _, _ = O.stack(frames[batch], "sigma", sl, sh)
        O.lib().nlo_stack_incremental(acc.ctypes.data_as(fp), res.ctypes.data_as(fp), acc.size, float(len(batch)), 1 if b == 0 else 0)
    O.lib().nlo_stack_incremental_finalize(acc.ctypes.data_as(fp), acc.size, float(n))
    assert np.array_equal(full.view(np.uint32), acc.view(np.uint32))

"""CPU tier, world_size 2 over gloo: the N>1 host logic (tile ownership, ray-set slicing, frame reduce) with the CPU
oracle standing in for the per-rank renderer."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from foundation_b200 import distributed as fdist


def test_tiles_partition_the_frame_exactly_once():
    for world in (2, 3, 8):
        total = np.zeros((90, 160), np.int32)
        for r in range(world):
            total += fdist.owned_mask(160, 90, r, world, 16)
        assert (total == 1).all()
    counts = [fdist.owned_mask(1920, 1080, r, 8, 32).sum() for r in range(8)]
    assert max(counts) / min(counts) < 1.06            # load balance of the interleave at 1080p / 8 GPUs


def test_packed_index_is_the_scanline_rank_of_an_owned_pixel():
    """Host model of the root's scatter (k_unpack_gathered): position of pixel (x, y) in its owner's packed block, ragged frames included."""
    for (W, H, T, N) in [(160, 90, 16, 3), (1920, 1080, 32, 8), (100, 70, 32, 2), (64, 48, 16, 5), (37, 29, 8, 5)]:
        for r in range(N):
            ys, xs = np.nonzero(fdist.owned_mask(W, H, r, N, T))
            assert np.array_equal(fdist.packed_index(xs, ys, r, N, W, H, T), np.arange(len(xs))), (W, H, T, N, r)


def test_ray_slices_cover_the_set():
    for n, w in ((1 << 20, 8), (1000, 3), (5, 8)):
        sl = [fdist.ray_slice(n, r, w) for r in range(w)]
        assert sl[0][0] == 0 and sl[-1][1] == n and all(a[1] == b[0] for a, b in zip(sl, sl[1:]))


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from foundation_b200 import scenes
    from oracle.pt_oracle import OracleScene
    r, w, _ = fdist.init_process_group("gloo")
    sc = scenes.cornell_box(64, 48)
    o = OracleScene(sc)
    part = o.render(64, 48, 9, 0, 2, 3, rank=r, count=w, tile=16, threads=1)
    assert (part[..., 3] > 0).sum() == fdist.owned_mask(64, 48, r, w, 16).sum()
    full = fdist.reduce_frames(torch.from_numpy(part), 0)
    if r == 0:
        ref = o.render(64, 48, 9, 0, 2, 3, threads=1)
        out.put(bool(np.array_equal(full.numpy(), ref)))
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_reduce_equals_single_rank_frame():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=180)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert ok

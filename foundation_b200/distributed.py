"""Multi-GPU sharding of the path-tracing pass for the one-process-per-GPU host shape (torchrun).

SURVEY.md §8e: pixels and samples are independent, the scene is replicated per GPU (every rank runs the same
deterministic build), the frame is split into interleaved 32x32 tiles (tile (tx,ty) -> rank (tx+ty) % world,
the same rule as `foundation_pt_partition_set`), and the only exchange step is the gather of the owned tiles
into rank 0's accumulation buffer.  That exchange lives INSIDE the C ABI (`foundation_pt_comm_init` /
`foundation_pt_gather`: packed ncclSend/ncclRecv, or — COMM_DIRECT — the accumulate kernel storing straight into
rank 0's frame over NVLink peer memory); `torch.distributed` is used here only to hand the 128-byte communicator
id from rank 0 to the other ranks and for the bench's barriers.  The gathered frame is bit-identical to the
single-GPU frame.  The explicit-ray-set metric needs no collective at all: each rank traces its own slice.

`reduce_frames` (a torch sum-reduce of whole zero-padded frames, round 1's gather) is kept as the host-side
model of the exchange for the CPU tier (gloo, world size 2) — the product path no longer calls it.

The reference is single-device (mos9527/Foundation src/Editor/Editor.cpp:18 takes EnumerateDevices()[0]), so
nothing here replaces a reference interface.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend: Optional[str] = None):
    """Initialise torch.distributed from the torchrun environment (RANK / WORLD_SIZE / MASTER_*).  No-op for world 1."""
    import torch
    import torch.distributed as dist

    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def tile_owner(x, y, world: int, tile: int = 32):
    """Rank that owns pixel (x, y): interleaved tiles with a one-tile rotation per tile row."""
    return ((np.asarray(x) // tile) + (np.asarray(y) // tile)) % world


def owned_mask(width: int, height: int, rank: int, world: int, tile: int = 32) -> np.ndarray:
    ys, xs = np.mgrid[0:height, 0:width]
    return tile_owner(xs, ys, world, tile) == rank


def ray_slice(num_rays: int, rank: int, world: int):
    """Contiguous 1/world slice of an explicit ray set."""
    per = (num_rays + world - 1) // world
    lo = min(num_rays, rank * per)
    return lo, min(num_rays, lo + per)


class _DevicePtr:
    """Exposes a raw device allocation of the C ABI to torch through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr: int, nfloats: int):
        self.__cuda_array_interface__ = {"shape": (nfloats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def accum_as_tensor(tracer):
    """The context's accumulation buffer (height x width x float4) as a torch CUDA tensor aliasing the same memory."""
    import torch

    ptr, size = tracer.accum_device_ptr()
    t = torch.as_tensor(_DevicePtr(ptr, size // 4), device=f"cuda:{tracer.device}")
    return t.view(tracer.height, tracer.width, 4)


def reduce_frames(local_frame, dst: int = 0):
    """Sum-reduce per-rank frames to `dst`.  Works for CUDA tensors (NCCL) and CPU tensors (gloo).
    Returns the full frame on dst, None elsewhere.  The input is not modified."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_frame.clone()
    out = local_frame.clone()
    dist.reduce(out, dst=dst, op=dist.ReduceOp.SUM)
    return out if dist.get_rank() == dst else None


def packed_index(x, y, rank: int, world: int, width: int, height: int, tile: int = 32):
    """Position of owned pixel (x, y) inside rank `rank`'s packed block (scanline order of its owned pixels) — the host-side
    statement of what k_unpack_gathered computes on the root (row base from the rows above + full tiles to the left)."""
    x = np.asarray(x, np.int64); y = np.asarray(y, np.int64)
    tiles_x = (width + tile - 1) // tile
    tx_all = np.arange(tiles_x)
    tw = np.minimum(tile, width - tx_all * tile)
    row_count = np.asarray([tw[(tx_all + ty) % world == rank].sum() for ty in range((height + tile - 1) // tile)], np.int64)   # per tile row
    rows = row_count[np.arange(height) // tile]
    base = np.concatenate([[0], np.cumsum(rows)[:-1]])
    tx, ty = x // tile, y // tile
    a = (rank - ty) % world
    before = np.where(tx > a, (tx - a - 1) // world + 1, 0)
    return base[y] + before * tile + (x - tx * tile)


def share_comm_id(make_id):
    """Rank 0 calls make_id() (PathTracer.comm_unique_id); every rank returns the same 128 bytes."""
    import torch.distributed as dist

    rank, world, _ = env_rank_world()
    if world == 1 or not dist.is_initialized():
        return make_id()
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]


class DistributedRenderer:
    """Tile-sharded progressive render over all ranks of the process group; the gather runs inside the C ABI."""

    def __init__(self, tracer, rank: int, world: int, tile: int = 32, direct: bool = False):
        from . import pt

        self.tracer, self.rank, self.world = tracer, rank, world
        if world > 1:
            cid = share_comm_id(pt.PathTracer.comm_unique_id)
            tracer.comm_init(cid, rank, world, tile, pt.COMM_DIRECT if direct else 0)
        else:
            tracer.partition_set(0, 1, tile)

    def render(self, sample_begin: int, sample_count: int, max_bounces: int, gather: bool = True):
        """Renders this rank's tiles; with gather=True rank 0's accumulation buffer then holds the whole frame."""
        self.tracer.render(sample_begin, sample_count, max_bounces)
        if gather and self.world > 1:
            self.tracer.gather(0)

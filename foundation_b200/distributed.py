"""Multi-GPU sharding of the path-tracing pass: one process per GPU, `torch.distributed` for the plumbing.

SURVEY.md §8e: pixels and samples are independent, the scene is replicated per GPU (every rank runs the same
deterministic build), the frame is split into interleaved 32x32 tiles (tile (tx,ty) -> rank (tx+ty) % world,
the same rule as `foundation_pt_partition_set`), and the only exchange step is a sum-reduce of the float4
accumulation buffer to rank 0 over NCCL / NVLink.  Ranks write disjoint pixels into zero-initialised frames,
so the reduce adds exact zeros and the result is bit-identical to the single-GPU frame.
The explicit-ray-set metric needs no collective at all: each rank traces its own slice.

The reference is single-device (mos9527/Foundation src/Editor/Editor.cpp:18 takes EnumerateDevices()[0]), so
nothing here replaces a reference interface; torch is used only for process-group plumbing and as the owner of
the reduce's output tensor.
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


class DistributedRenderer:
    """Tile-sharded progressive render over all ranks of the process group."""

    def __init__(self, tracer, rank: int, world: int, tile: int = 32):
        self.tracer, self.rank, self.world = tracer, rank, world
        tracer.partition_set(rank, world, tile)

    def render(self, sample_begin: int, sample_count: int, max_bounces: int, gather: bool = True):
        self.tracer.render(sample_begin, sample_count, max_bounces)
        if not gather:
            return None
        return reduce_frames(accum_as_tensor(self.tracer), 0)

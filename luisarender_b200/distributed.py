"""Multi-GPU plumbing: one process per GPU, scene replicated, pixels sharded by interleaved
tiles (lrk_tile_owner(tile_id, world) == rank), and exactly ONE collective — a sum-reduce of the raw float4 film to rank 0
after the last pass.  Because tiles are disjoint the reduce only adds zeros, so the reduced film is
bit-identical to the single-GPU film.  The reference has no multi-device path at all (one `-d` index,
src/apps/cli.cpp:62,167-181); this module is new.

torch.distributed is used for the rendezvous and the NCCL (GPU) / gloo (CPU tests) reduce only.
"""
from __future__ import annotations

import os

import numpy as np

TILE_SIZE = 32


def env_world() -> tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment; (0, 1, 0) when not launched by torchrun."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend: str):
    import torch.distributed as dist

    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group(backend=backend)
    return dist


def owned_pixel_mask(width: int, height: int, rank: int, world: int, tile_size: int = TILE_SIZE) -> np.ndarray:
    """Boolean [H, W] mask of the pixels rank `rank` renders (same rule as lrk_set_shard / oracle_render)."""
    tiles_x = (width + tile_size - 1) // tile_size
    ty, tx = np.meshgrid(np.arange(height) // tile_size, np.arange(width) // tile_size, indexing="ij")
    return tile_owner(ty * tiles_x + tx, world) == rank


def tile_owner(tile_id, world: int):
    """include/lrk.h lrk_tile_owner, vectorised: each run of `world` consecutive tiles gives one tile to each rank, in an
    order rotated by a hash of the run's index."""
    tile_id = np.asarray(tile_id, dtype=np.uint64)
    mask = np.uint64(0xFFFFFFFF)
    h = ((tile_id // np.uint64(world)) * np.uint64(0x9E3779B1)) & mask
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(0x85EBCA77)) & mask
    h ^= h >> np.uint64(13)
    return ((tile_id % np.uint64(world) + h % np.uint64(world)) % np.uint64(world)).astype(np.int64)


def device_film_tensor(renderer, height: int, width: int):
    """Zero-copy torch view of the renderer's raw device film (float32 [H, W, 4]) for the NCCL reduce."""
    import torch

    ptr, nbytes = renderer.film_device_ptr()
    assert nbytes == height * width * 16

    class _Film:
        __cuda_array_interface__ = {"shape": (height, width, 4), "typestr": "<f4", "data": (ptr, False), "version": 3,
                                    "strides": None}

    return torch.as_tensor(_Film(), device=f"cuda:{torch.cuda.current_device()}")


def reduce_film(film, dst: int = 0):
    """Sum-reduce the raw film to rank `dst` (ncclReduce under the NCCL backend). In place; returns `film`."""
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM)
    return film


def init_film_comm(renderer, rank: int, world: int):
    """Create the library's own film-reduce communicator (lrk_comm_init) for `renderer`: rank 0 makes the NCCL unique id and
    the process group (any backend) carries it to the other ranks.  After this, renderer.reduce_film(root) is the multi-GPU
    path's one collective, issued by the C-ABI on the renderer's own stream - no torch tensor involved."""
    import torch.distributed as dist

    box = [renderer.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    renderer.comm_init(box[0], rank, world)

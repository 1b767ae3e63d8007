"""Host-side logic of the chunk-ownership sharded map: who owns a chunk, how a frame travels
from the ingest rank to the others, and how per-rank results are merged. torch.distributed is only plumbing here; the
same code runs over NCCL on the GPU box and over gloo in the CPU tests.

  owner(id)   = chs_owner(id) % world          (the device kernels apply the same rule: chunk_candidates_kernel)
  frame       = one contiguous byte buffer [depth float32 W*H | colour uint8 W*H*C] -> ONE broadcast per frame
  results     = per-rank maps are disjoint; their union is the map (tests: equal to the 1-rank map bit for bit)
"""
from __future__ import annotations

import numpy as np

from . import capi


def owners(ids: np.ndarray, world: int) -> np.ndarray:
    """Owning rank of every chunk ID (n x 3 int array)."""
    ids = np.asarray(ids, np.int32).reshape(-1, 3)
    return np.array([capi.owner(int(x), int(y), int(z)) % world for x, y, z in ids], dtype=np.int64)


def filter_owned(state, rank: int, world: int):
    """Restrict a (ids, sdf, weight, rgbw) state to the chunks `rank` owns."""
    ids = state[0]
    keep = owners(ids, world) == rank
    return tuple(a[keep] for a in state)


def merge_states(parts):
    """Union of disjoint per-rank states, sorted by chunk ID."""
    ids = np.concatenate([p[0] for p in parts])
    if len(np.unique(ids, axis=0)) != len(ids):
        raise ValueError("shards overlap: a chunk ID is present on more than one rank")
    order = np.lexsort((ids[:, 2], ids[:, 1], ids[:, 0]))
    return tuple(np.concatenate([p[k] for p in parts])[order] for k in range(len(parts[0])))


def merge_dirty(parts):
    """Union of the per-rank dirty sets (a rank marks the 27-neighbourhood of ITS updated chunks, whoever owns them)."""
    ids = np.concatenate([np.asarray(p, np.int32).reshape(-1, 3) for p in parts])
    return np.unique(ids, axis=0) if len(ids) else ids


def frame_nbytes(width: int, height: int, channels: int) -> int:
    return 4 * width * height + channels * width * height


def pack_frame(depth: np.ndarray, color: np.ndarray | None) -> np.ndarray:
    d = np.ascontiguousarray(depth, np.float32).view(np.uint8).reshape(-1)
    if color is None:
        return d.copy()
    return np.concatenate([d, np.ascontiguousarray(color, np.uint8).reshape(-1)])


def unpack_frame(buf: np.ndarray, width: int, height: int, channels: int):
    n = 4 * width * height
    depth = buf[:n].view(np.float32).reshape(height, width)
    color = buf[n:n + channels * width * height].reshape(height, width, channels) if channels else None
    return depth, color


def broadcast_frame(buf, src: int = 0):
    """One collective per frame: the ingest rank's packed frame to every rank (torch tensor, any backend)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(buf, src)
    return buf


def gather_to_root(obj, dst: int = 0):
    """Gather picklable per-rank results (small: IDs, counters, digests) on the root."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(obj, out, dst=dst)
    return out

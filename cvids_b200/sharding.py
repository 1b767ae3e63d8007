"""Host-side logic of the chunk-ownership sharded map: who owns a chunk, how a frame travels
from the ingest rank to the others, and how per-rank results are merged. torch.distributed is only plumbing here; the
same code runs over NCCL on the GPU box and over gloo in the CPU tests.

  owner(id)   = chs_owner(id) % world          (the device kernels apply the same rule: chunk_candidates_kernel)
  frame       = one contiguous byte buffer [depth float32 W*H | colour uint8 W*H*C] -> ONE broadcast per frame
  step block  = the frames of a step, contiguous; ingested in `world` equal byte ranges (rank r ingests range r) and
                replicated by ONE all-gather: every rank has the whole block when the collective ends
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


def ingest_share(total_bytes: int, world: int) -> int:
    """Bytes per rank of a step block split into `world` equal ranges (16-byte aligned; the last range may be short or empty)."""
    return ((total_bytes + world - 1) // world + 15) // 16 * 16


def ingest_range(total_bytes: int, rank: int, world: int):
    """(offset, valid bytes) of rank's range inside the step block."""
    share = ingest_share(total_bytes, world)
    lo = min(rank * share, total_bytes)
    return lo, min((rank + 1) * share, total_bytes) - lo


def all_gather_block(out, mine):
    """Replicate a step block: `mine` = this rank's range padded to ingest_share bytes, `out` = world * share bytes (the block
    followed by padding). One collective; NCCL all_gather_into_tensor on the device, list all_gather elsewhere (gloo)."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        out[:mine.numel()].copy_(mine)
        return out
    if out.is_cuda:
        dist.all_gather_into_tensor(out, mine)
    else:
        dist.all_gather(list(out.view(world, -1).unbind(0)), mine)
    return out


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


# ---------------------------------------------------------------------------------------------------------------
# Sharded meshing: marching cubes reads the +x/+y/+z neighbour chunks and the gradient normals read
# +-1 voxel, i.e. chunks owned by other ranks. Each re-mesh therefore runs in three phases:
#   1. all-gather the per-rank dirty ID lists -> the global dirty set D (a rank marks the 27-neighbourhood of ITS updated
#      chunks, whoever owns those IDs);
#   2. every rank exports the chunks it owns inside the 27-neighbourhood of D; the exports are all-gathered, and each
#      rank imports them into a scratch "ghost" map (a plain 1-rank map, reset per re-mesh);
#   3. every rank meshes the dirty IDs IT owns on its ghost map; meshes are gathered on the root.
# The result is the single-GPU mesh bit for bit, with one documented exception: the reference's colour lookup samples
# voxel INDICES as if they were metres (quirk Q9); if the map happens to hold chunks at those aliased positions outside
# the exchanged neighbourhood, their colours are not seen by the ghost map.

def neighbourhood27(ids: np.ndarray) -> np.ndarray:
    ids = np.asarray(ids, np.int32).reshape(-1, 3)
    if not len(ids):
        return ids
    off = np.array([(dx, dy, dz) for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1)], np.int32)
    return np.unique((ids[:, None, :] + off[None, :, :]).reshape(-1, 3), axis=0)


def phase1_dirty(shard) -> np.ndarray:
    return shard.get_meshes_to_update()


def phase2_export(shard, dirty_all: np.ndarray):
    """This rank's contribution: (ids, sdf, weight, rgbw) of the chunks it holds in the 27-neighbourhood of D."""
    need = neighbourhood27(dirty_all)
    found, sdf, w, rgbw = shard.export_chunks(need)
    return need[found], sdf[found], w[found], rgbw[found]


def phase3_mesh(ghost, contributions, dirty_all: np.ndarray, rank: int, world: int) -> dict:
    """Mesh the dirty IDs `rank` owns on the ghost map; returns {id: mesh} for every re-meshed chunk (incl. empty ones)."""
    ghost.reset()
    for ids, sdf, w, rgbw in contributions:
        ghost.import_chunks(ids, sdf, w, rgbw)
    mine = dirty_all[owners(dirty_all, world) == rank] if len(dirty_all) else dirty_all
    ghost.set_dirty(mine)
    ghost._lib.chs_update_meshes(ghost._h)
    return ghost.download_last_meshes()


def merge_remeshed(all_meshes: dict, remeshed: dict) -> None:
    """The reference's publication rule (ChunkManager.cpp:101-127, quirk Q10) on the root's MeshMap."""
    for cid, mesh in remeshed.items():
        if cid in all_meshes or len(mesh["grids"]) > 0:
            all_meshes[cid] = mesh


def sharded_remesh(shard, ghost, rank: int, world: int, all_meshes: dict | None = None, root: int = 0):
    """One distributed re-mesh over torch.distributed (any backend; payloads travel as pickled NumPy arrays). Returns the
    root's MeshMap (None elsewhere)."""
    import torch.distributed as dist

    def all_gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    dirty_all = merge_dirty(all_gather(phase1_dirty(shard)))
    contributions = all_gather(phase2_export(shard, dirty_all))
    remeshed = phase3_mesh(ghost, contributions, dirty_all, rank, world)
    shard.set_dirty(np.zeros((0, 3), np.int32))                      # meshesToUpdate.clear() (Chisel.cpp:57)
    gathered = gather_to_root(remeshed, root)
    if rank != root:
        return None
    if all_meshes is None:
        all_meshes = {}
    for part in gathered:
        merge_remeshed(all_meshes, part)
    return all_meshes

"""CPU suite: the N > 1 host-side path over a real process group (gloo, world_size 2): one-collective frame broadcast,
ownership partition, and the merge of per-rank maps. The per-rank map here comes from the C oracle restricted to the
chunks the rank owns (a chunk's voxels depend on nothing but the frames, so the restriction of the full map IS what a
shard must hold); the GPU suite checks the same property on the device (test_sharded_union_equals_single_map)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from cvids_b200 import sharding
    from tests import common
    from tests.common import Setup
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cam = common.SMALL_CAM
    setup = Setup(16, 0.05, True)
    frames = list(common.orbit_stream(cam, 4, total=30, color=True, seed=9)) if rank == 0 else None
    drv = common.Driver(setup, "oracle")
    nbytes = sharding.frame_nbytes(cam.width, cam.height, 3)
    poses = torch.zeros((4, 12), dtype=torch.float32)
    if rank == 0:
        poses = torch.from_numpy(np.stack([f[2].reshape(12) for f in frames]))
    dist.broadcast(poses, 0)
    for i in range(4):
        buf = torch.from_numpy(sharding.pack_frame(frames[i][0], frames[i][1])) if rank == 0 else torch.empty(nbytes, dtype=torch.uint8)
        sharding.broadcast_frame(buf, 0)
        depth, col = sharding.unpack_frame(buf.numpy(), cam.width, cam.height, 3)
        drv.integrate(depth, poses[i].numpy().reshape(3, 4), cam.as_array(), col)
    full = drv.state()
    mine = sharding.filter_owned(full, rank, world)
    # dirty IDs a rank would mark: the 27-neighbourhood of its own updated chunks -- emulate with the full dirty set split by owner of the source
    parts = sharding.gather_to_root((mine, drv.dirty()), 0)
    if rank == 0:
        merged = sharding.merge_states([p[0] for p in parts])
        common.assert_state_equal(merged, full, "gloo shards")
        assert sum(len(p[0][0]) for p in parts) == len(full[0])
        assert all(len(p[0][0]) > 0 for p in parts)
        assert np.array_equal(sharding.merge_dirty([p[1] for p in parts]), drv.dirty())
        np.save(os.path.join(tmpdir, "ok.npy"), np.array([len(full[0])]))
    dist.barrier()
    dist.destroy_process_group()


def _ingest_worker(rank, world, port, tmpdir):
    """Sharded ingest of a step block (bench.py, N > 1): every rank contributes ITS byte range, one all-gather, and every rank
    ends up with the whole block of frames."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from cvids_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fbytes = sharding.frame_nbytes(40, 30, 3)
    for b in (1, 3, 10):
        total = b * fbytes
        block = torch.from_numpy(np.random.RandomState(b).randint(0, 256, total).astype(np.uint8))     # the same stream on every rank
        share = sharding.ingest_share(total, world)
        lo, n = sharding.ingest_range(total, rank, world)
        mine = torch.full((share,), 255, dtype=torch.uint8)
        mine[:n] = block[lo:lo + n]                                                                     # only this rank's range is read
        out = torch.zeros(share * world, dtype=torch.uint8)
        sharding.all_gather_block(out, mine)
        assert torch.equal(out[:total], block), "rank %d: block of %d frames not reconstructed" % (rank, b)
    spans = [sharding.ingest_range(1000, r, 3) for r in range(3)]
    assert spans[0][0] == 0 and sum(s[1] for s in spans) == 1000 and all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(2))
    if rank == 0:
        np.save(os.path.join(tmpdir, "ingest_ok.npy"), np.array([1]))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_sharded_ingest_all_gather(tmp_path):
    import torch.multiprocessing as mp
    port = 29900 + os.getpid() % 90
    mp.spawn(_ingest_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ingest_ok.npy")


def test_gloo_world2_broadcast_partition_merge(tmp_path):
    import torch.multiprocessing as mp
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok.npy")


def test_pack_unpack_roundtrip():
    sys.path.insert(0, ROOT)
    from cvids_b200 import sharding
    rng = np.random.RandomState(0)
    d = rng.uniform(0, 5, (12, 16)).astype(np.float32)
    d[3, 4] = np.nan
    c = rng.randint(0, 255, (12, 16, 3)).astype(np.uint8)
    buf = sharding.pack_frame(d, c)
    assert buf.nbytes == sharding.frame_nbytes(16, 12, 3)
    d2, c2 = sharding.unpack_frame(buf, 16, 12, 3)
    assert np.array_equal(d.view(np.uint32), d2.view(np.uint32)) and np.array_equal(c, c2)
    with pytest.raises(ValueError):
        sharding.merge_states([(np.zeros((1, 3), np.int32), np.zeros((1, 8))), (np.zeros((1, 3), np.int32), np.zeros((1, 8)))])

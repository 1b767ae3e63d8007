"""Worker of tests/test_multi_gpu.py (one process per GPU, launched by torch.distributed.run): chunk-ownership sharded
integration with NCCL frame broadcast, then the distributed re-mesh, checked on rank 0 against the CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cvids_b200 import sharding  # noqa: E402
from tests import common  # noqa: E402
from tests.common import Setup  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cam = common.SMALL_CAM
    setup = Setup(16, 0.05, True)
    n = 8
    frames = list(common.orbit_stream(cam, n, total=30, color=True, seed=12)) if rank == 0 else None
    poses = torch.zeros((n, 12), dtype=torch.float32, device=dev)
    if rank == 0:
        poses.copy_(torch.from_numpy(np.stack([f[2].reshape(12) for f in frames])))
    dist.broadcast(poses, 0)
    poses = poses.cpu().numpy()
    shard = common.Driver(setup, "cuda", device=local, rank=rank, world=world)
    ghost = common.Driver(setup, "cuda", device=local)
    nbytes = sharding.frame_nbytes(cam.width, cam.height, 3)
    buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    root_meshes = {}
    oracle = common.Driver(setup, "oracle") if rank == 0 else None
    for i in range(n):
        if rank == 0:
            buf.copy_(torch.from_numpy(sharding.pack_frame(frames[i][0], frames[i][1])))
        sharding.broadcast_frame(buf, 0)
        torch.cuda.synchronize(dev)
        p = buf.data_ptr()
        shard.m.integrate_depth_scan_color(shard.integ, None, poses[i], cam.as_array(), None, device_ptrs=(p, p + 4 * cam.width * cam.height), channels=3)
        shard.m.synchronize()
        if rank == 0:
            oracle.integrate(frames[i][0], frames[i][2], cam.as_array(), frames[i][1])
        if i in (3, 7):
            got = sharding.sharded_remesh(shard.m, ghost.m, rank, world, root_meshes, root=0)
            if rank == 0:
                oracle.remesh()
                common.assert_meshes_equal(got, oracle.meshes(), "distributed re-mesh at frame %d" % i)
    parts = sharding.gather_to_root((shard.state(), shard.dirty()), 0)
    if rank == 0:
        merged = sharding.merge_states([p[0] for p in parts])
        common.assert_state_equal(merged, oracle.state(), "union of %d NCCL shards" % world)
        assert all(len(p[0][0]) > 0 for p in parts)
        print("MULTI_GPU_OK world=%d chunks=%d triangles=%d" % (world, len(merged[0]), sum(len(m["vertices"]) for m in root_meshes.values()) // 3), flush=True)
    dist.barrier()

    # ---- the fused multi-frame path, fed the way bench.py feeds it at N > 1: every rank ingests ITS byte range of the step's frame
    # block, ONE all-gather replicates the block, one chs_integrate_batch per rank. Union of the shards == the oracle's map.
    fbytes = nbytes
    block_host = torch.empty(n * fbytes, dtype=torch.uint8)
    if rank == 0:
        for i in range(n):
            block_host[i * fbytes:(i + 1) * fbytes].copy_(torch.from_numpy(sharding.pack_frame(frames[i][0], frames[i][1])))
    stream_all = block_host.to(dev)
    dist.broadcast(stream_all, 0)                                   # set-up: every rank can play the ingest rank of its range
    shard2 = common.Driver(setup, "cuda", device=local, rank=rank, world=world)
    for first, cnt in ((0, 3), (3, 5)):                            # two steps: 3 frames, then 5
        total = cnt * fbytes
        share = sharding.ingest_share(total, world)
        lo, nb = sharding.ingest_range(total, rank, world)
        mine = torch.zeros(share, dtype=torch.uint8, device=dev)
        mine[:nb].copy_(stream_all[first * fbytes + lo:first * fbytes + lo + nb])
        out = torch.zeros(share * world, dtype=torch.uint8, device=dev)
        sharding.all_gather_block(out, mine)
        torch.cuda.synchronize(dev)
        base = out.data_ptr()
        ptrs = [(base + j * fbytes, base + j * fbytes + 4 * cam.width * cam.height) for j in range(cnt)]
        shard2.m.integrate_batch(shard2.integ, None, [poses[first + j] for j in range(cnt)], cam.as_array(), device_ptrs=ptrs, channels=3)
        assert len(shard2.m.batch_stats()) == cnt
    parts = sharding.gather_to_root((shard2.state(), shard2.dirty()), 0)
    if rank == 0:
        merged = sharding.merge_states([p[0] for p in parts])
        common.assert_state_equal(merged, oracle.state(), "union of %d NCCL shards, fused batches" % world)
        print("MULTI_GPU_BATCH_OK world=%d chunks=%d" % (world, len(merged[0])), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

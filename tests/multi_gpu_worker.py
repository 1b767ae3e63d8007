"""Worker of tests/test_multi_gpu.py (one process per GPU, launched by torch.distributed.run).

The DATA PLANE is the C ABI only: chs_comm_init (NCCL communicator inside the library), chs_integrate_batch_distributed (every
rank ingests its share of a step's frames, in-place all-gather over NVLink, fused integration of the chunks it owns),
chs_update_meshes_distributed (dirty-set union, device-side ghost-chunk exchange, per-rank meshing, mesh gather on the root).
torch.distributed (gloo) is only the side channel for the NCCL unique id and for collecting results to compare: rank 0 checks
the union of the shards and the root's meshes against the CPU oracle fed the same frames in arrival order."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cvids_b200 import scenes, sharding  # noqa: E402
from tests import common  # noqa: E402
from tests.common import Setup  # noqa: E402


def gather(obj, dst=0):
    out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(obj, out, dst=dst)
    return out


def run_case(name, setup, cam, steps, rank, world, local, device_frames=False, remesh_after=()):
    """steps: list of lists of (depth, colour|None, pose); every step has a multiple of `world` frames."""
    shard = common.Driver(setup, "cuda", device=local, rank=rank, world=world)
    shard.m.comm_init_torch()
    oracle = common.Driver(setup, "oracle") if rank == 0 else None
    camv = cam.as_array()
    dev = torch.device("cuda", local)
    keep = []
    for si, frames in enumerate(steps):
        n = len(frames)
        per = n // world
        lo = rank * per
        color = frames[0][1] is not None
        poses = [f[2] for f in frames]
        if device_frames:
            ptrs = [None] * n
            for i in range(lo, lo + per):
                d = torch.from_numpy(frames[i][0]).to(dev)
                c = torch.from_numpy(frames[i][1]).to(dev) if color else None
                keep += [d, c]
                ptrs[i] = (d.data_ptr(), c.data_ptr() if color else None)
            torch.cuda.synchronize(dev)
            fill = ptrs[lo]
            shard.m.integrate_batch_distributed(shard.integ, None, poses, camv, device_ptrs=[p or fill for p in ptrs], channels=3 if color else 0)
        else:
            depths = [f[0] if lo <= i < lo + per else None for i, f in enumerate(frames)]
            cols = [f[1] if lo <= i < lo + per else None for i, f in enumerate(frames)] if color else None
            shard.m.integrate_batch_distributed(shard.integ, depths, poses, camv, cols)
        got = shard.m.batch_stats()
        assert len(got) == n
        if rank == 0:
            want = []
            for depth, col, pose in frames:
                oracle.integrate(depth, pose, camv, col)
                want.append(oracle.counters())
        parts = gather([[g[k] for k in ("candidates", "n_upd", "n_carve", "n_col", "n_new", "updated_chunks")] for g in got])
        if rank == 0:
            tot = np.sum(np.asarray(parts, np.int64), axis=0)
            for j, w in enumerate(want):
                for k, key in enumerate(("candidates", "n_upd", "n_carve", "n_col", "n_new", "updated_chunks")):
                    assert tot[j][k] == w[key], "%s step %d frame %d %s: shards %d oracle %d" % (name, si, j, key, tot[j][k], w[key])
        if si in remesh_after:
            shard.m.recompute_meshes_distributed(root=0)
            if rank == 0:
                oracle.remesh()
                common.assert_meshes_equal(shard.meshes(), oracle.meshes(), "%s: distributed re-mesh after step %d" % (name, si))
            assert len(shard.dirty()) == 0
    parts = gather((shard.state(), len(shard.state()[0])))
    if rank == 0:
        merged = sharding.merge_states([p[0] for p in parts])
        common.assert_state_equal(merged, oracle.state(), "%s: union of %d shards" % (name, world))
        assert all(p[1] > 0 for p in parts), "a rank holds no chunk"
        tris = sum(len(m["vertices"]) for m in shard.meshes().values()) // 3
        print("MULTI_GPU_CASE_OK %s world=%d chunks=%d triangles=%d" % (name, world, len(merged[0]), tris), flush=True)
    dist.barrier()
    shard.m.comm_destroy()
    shard.m.close()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    cam = common.SMALL_CAM
    agents = 8
    # eight agents (phase-shifted orbits), one step = the eight frames of a time step in arrival order; colour + NaN pixels
    cfg = scenes.StreamConfig("8agent-small", scenes.ROOM, cam, 0.05, 30, True, nan_frac=0.02, agents=agents)
    steps = [[scenes.stream_frame(cfg, t, agent=a) for a in range(agents)] for t in range(4)]
    run_case("eight-agents-colour-host", Setup(16, 0.05, True), cam, steps, rank, world, local, remesh_after=(1, 3))
    # the same shape, depth only, 2 cm voxels, frames resident on the device
    cfg2 = scenes.StreamConfig("8agent-small-2cm", scenes.ROOM, cam, 0.02, 30, False, agents=agents)
    steps2 = [[scenes.stream_frame(cfg2, t, agent=a) for a in range(agents)] for t in range(3)]
    run_case("eight-agents-depth-device", Setup(16, 0.02, False), cam, steps2, rank, world, local, device_frames=True, remesh_after=(2,))
    # carving inside a distributed step
    carve = list(common.carve_stream(cam, 4, 4, color=True))
    run_case("carving", Setup(16, 0.05, True, weight=2.0), cam, [carve[:8]], rank, world, local, remesh_after=(0,))
    if rank == 0:
        print("MULTI_GPU_OK world=%d" % world, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""GPU suite for the fused multi-frame path (chs_integrate_batch, cvids_b200/csrc/integrate_batch_impl.cuh): K frames in one pass
must leave the map, the dirty set, the meshes and the per-frame counters bit-identical to K single-frame integrations -- checked
against the CPU oracle (which integrates frame by frame, like the reference) and against the single-frame CUDA path."""
import numpy as np
import pytest

from cvids_b200 import scenes
from tests import common
from tests.common import Setup

pytestmark = pytest.mark.gpu

COUNTERS = ("candidates", "n_upd", "n_carve", "n_col", "n_new", "updated_chunks")


def _run_batched(setup, frames, cam, batch, remesh=True, truncs=False, **cuda_kw):
    """Feed `frames` to the CUDA map in groups of `batch` (chs_integrate_batch) and frame by frame to the oracle."""
    a, b = common.Driver(setup, "cuda", **cuda_kw), common.Driver(setup, "oracle")
    camv = cam.as_array()
    frames = list(frames)
    i = 0
    while i < len(frames):
        grp = frames[i:i + batch]
        color = grp[0][1] is not None
        a.m.integrate_batch(a.integ, [g[0] for g in grp], [g[2] for g in grp], camv, [g[1] for g in grp] if color else None)
        want = []
        for depth, col, pose in grp:
            b.integrate(depth, pose, camv, col)
            want.append(b.counters())
        got = a.m.batch_stats()
        assert len(got) == len(grp)
        for j, (g, w) in enumerate(zip(got, want)):
            for k in COUNTERS:
                assert g[k] == w[k], "frame %d counter %s: cuda batch %d oracle %d" % (i + j, k, g[k], w[k])
        assert np.array_equal(a.dirty(), b.dirty()), "dirty sets differ after the batch ending at frame %d" % (i + len(grp) - 1)
        if remesh:
            a.remesh()
            b.remesh()
        i += len(grp)
    common.assert_state_equal(a.state(), b.state())
    assert np.array_equal(a.dirty(), b.dirty())
    if remesh:
        common.assert_meshes_equal(a.meshes(), b.meshes())
    return a, b


@pytest.mark.parametrize("batch", [2, 5, 16])
def test_batch_color_room_2cm(batch):
    _run_batched(Setup(16, 0.02, True), common.orbit_stream(common.SMALL_CAM, 10, total=30, color=True, nan_frac=0.03, seed=7, noise=0.004),
                 common.SMALL_CAM, batch)


def test_batch_depth_path_5cm():
    _run_batched(Setup(16, 0.05, False), common.orbit_stream(common.MID_CAM, 12, total=30, seed=3), common.MID_CAM, 6)


@pytest.mark.parametrize("color", [False, True])
def test_batch_carving(color):
    """The obstacle appears and disappears INSIDE one batch: band updates and carving of the same voxels in one pass."""
    _run_batched(Setup(16, 0.05, color, weight=2.0), common.carve_stream(common.SMALL_CAM, color=color), common.SMALL_CAM, 10)
    _run_batched(Setup(16, 0.05, color, weight=2.0), common.carve_stream(common.SMALL_CAM, color=color), common.SMALL_CAM, 3)


def test_batch_carving_disabled():
    _run_batched(Setup(16, 0.05, False, carve=False), common.carve_stream(common.SMALL_CAM), common.SMALL_CAM, 5)


@pytest.mark.parametrize("kind,param", [(common.TRUNC_INVERSE, 2.0), (common.TRUNC_QUADRATIC, 4.0)])
def test_batch_truncators(kind, param):
    _run_batched(Setup(16, 0.05, True, trunc_kind=kind, trunc_param=param, carve_dist=0.0),
                 common.orbit_stream(common.SMALL_CAM, 6, total=30, color=True, seed=2), common.SMALL_CAM, 3)


@pytest.mark.parametrize("chunk,res", [(8, 0.1), (8, 0.04), (32, 0.03)])
def test_batch_chunk_sizes(chunk, res):
    _run_batched(Setup(chunk, res, True), common.orbit_stream(common.SMALL_CAM, 6, total=30, color=True, seed=4), common.SMALL_CAM, 3)


def test_batch_multi_agent_interleave():
    """Frames of four agents with different poses, interleaved round-robin (config 3 shape), eight per batch."""
    def frames():
        for f in range(4):
            for agent in range(4):
                pose = scenes.orbit_pose(f, 30, phase=agent * np.pi / 2)
                depth, col = scenes.render(scenes.ROOM, common.SMALL_CAM, pose, color=True, seed=100 * agent + f)
                yield depth, col, pose
    _run_batched(Setup(16, 0.05, True), frames(), common.SMALL_CAM, 8)


def test_batch_equals_single_frame_cuda_and_mixed_calls():
    """Batch calls, single-frame calls and a batch of one may be mixed freely on one map."""
    setup = Setup(16, 0.04, True)
    frames = list(common.orbit_stream(common.SMALL_CAM, 9, total=30, color=True, nan_frac=0.02, seed=11))
    camv = common.SMALL_CAM.as_array()
    a, b = common.Driver(setup, "cuda"), common.Driver(setup, "cuda")
    for depth, col, pose in frames:
        a.integrate(depth, pose, camv, col)
    grp = frames[:4]
    b.m.integrate_batch(b.integ, [g[0] for g in grp], [g[2] for g in grp], camv, [g[1] for g in grp])
    b.integrate(*[frames[4][k] for k in (0, 2)], camv, frames[4][1])
    grp = frames[5:6]
    b.m.integrate_batch(b.integ, [g[0] for g in grp], [g[2] for g in grp], camv, [g[1] for g in grp])
    assert len(b.m.batch_stats()) == 1
    grp = frames[6:]
    b.m.integrate_batch(b.integ, [g[0] for g in grp], [g[2] for g in grp], camv, [g[1] for g in grp])
    common.assert_state_equal(a.state(), b.state())
    assert np.array_equal(a.dirty(), b.dirty())
    a.remesh()
    b.remesh()
    common.assert_meshes_equal(a.meshes(), b.meshes())


def test_batch_separate_color_camera_runs_frame_by_frame():
    setup = Setup(16, 0.05, True)
    cam, ccam = common.SMALL_CAM, scenes.Camera(140.0, 139.0, 80.0, 60.0, 160, 120)
    a, b = common.Driver(setup, "cuda"), common.Driver(setup, "oracle")
    depths, cols, poses, cposes = [], [], [], []
    for f in range(3):
        pose = scenes.orbit_pose(f, 30)
        cpose = scenes.orbit_pose(f, 30, phase=0.02)
        depth, _ = scenes.render(scenes.ROOM, cam, pose)
        _, col = scenes.render(scenes.ROOM, ccam, cpose, color=True)
        depths.append(depth); cols.append(col); poses.append(pose); cposes.append(cpose)
        b.integrate(depth, pose, cam.as_array(), col, cpose, ccam.as_array())
    a.m.integrate_batch(a.integ, depths, poses, cam.as_array(), cols, cposes, ccam.as_array())
    assert len(a.m.batch_stats()) == 3
    common.assert_state_equal(a.state(), b.state())


def test_batch_larger_than_sixteen_and_device_memory():
    import torch
    setup = Setup(16, 0.05, True)
    cam = common.SMALL_CAM
    frames = list(common.orbit_stream(cam, 20, total=40, color=True, seed=5))
    a, b = common.Driver(setup, "cuda"), common.Driver(setup, "cuda")
    for depth, col, pose in frames:
        a.integrate(depth, pose, cam.as_array(), col)
    dd = [torch.from_numpy(f[0]).cuda() for f in frames]
    dc = [torch.from_numpy(f[1]).cuda() for f in frames]
    torch.cuda.synchronize()
    b.m.integrate_batch(b.integ, None, [f[2] for f in frames], cam.as_array(), device_ptrs=[(d.data_ptr(), c.data_ptr()) for d, c in zip(dd, dc)],
                        channels=3)
    st = b.m.batch_stats()
    assert len(st) == 20 and all(s["n_upd"] > 0 for s in st)
    common.assert_state_equal(a.state(), b.state())
    assert np.array_equal(a.dirty(), b.dirty())


def test_batch_sharded_union_equals_single_map():
    """Ownership filter in the fused candidates kernel: the union of three shards equals the unsharded map."""
    setup = Setup(16, 0.05, True)
    cam = common.SMALL_CAM
    frames = list(common.orbit_stream(cam, 6, total=30, color=True, seed=9))
    whole = common.Driver(setup, "cuda")
    whole.m.integrate_batch(whole.integ, [f[0] for f in frames], [f[2] for f in frames], cam.as_array(), [f[1] for f in frames])
    ids_w, sdf_w, w_w, c_w = whole.state()
    parts = []
    for r in range(3):
        d = common.Driver(setup, "cuda", rank=r, world=3)
        d.m.integrate_batch(d.integ, [f[0] for f in frames], [f[2] for f in frames], cam.as_array(), [f[1] for f in frames])
        parts.append(d.state())
    ids = np.concatenate([p[0] for p in parts])
    order = np.lexsort((ids[:, 2], ids[:, 1], ids[:, 0]))
    merged = tuple(np.concatenate([p[k] for p in parts])[order] for k in range(4))
    common.assert_state_equal(merged, (ids_w, sdf_w, w_w, c_w))


def test_fast_reciprocal_and_quotient_are_correctly_rounded():
    """The straight-line voxel update replaces __frcp_rn / __fdiv_rn by their range-check-free fast paths inside a guarded
    operand range: exhaustive comparison for the reciprocal, 2^30 operand pairs for the quotient."""
    from cvids_b200 import capi
    r = capi.selftest_arithmetic(1 << 30)
    assert r["rcp_tested"] > (1 << 30) and r["rcp_mismatches"] == 0, r
    assert r["div_tested"] > (1 << 27) and r["div_mismatches"] == 0, r


def test_pipelined_batches_with_tickets():
    """Issue batch k + 1 before reading batch k's counters (chs_wait_batch): same map and counters as the blocking way."""
    setup = Setup(16, 0.05, True)
    cam = common.SMALL_CAM
    frames = list(common.orbit_stream(cam, 12, total=30, color=True, seed=21))
    a, b = common.Driver(setup, "cuda"), common.Driver(setup, "cuda")
    blocking, piped = [], []
    for i in range(0, 12, 3):
        grp = frames[i:i + 3]
        a.m.integrate_batch(a.integ, [g[0] for g in grp], [g[2] for g in grp], cam.as_array(), [g[1] for g in grp])
        blocking += a.m.batch_stats()
    prev = None
    for i in range(0, 12, 3):
        grp = frames[i:i + 3]
        b.m.integrate_batch(b.integ, [g[0] for g in grp], [g[2] for g in grp], cam.as_array(), [g[1] for g in grp])
        t = b.m.last_batch_ticket()
        if prev is not None:
            piped += b.m.wait_batch(prev)
        prev = t
    piped += b.m.wait_batch(prev)
    assert len(piped) == 12
    for x, y in zip(blocking, piped):
        for k in COUNTERS:
            assert x[k] == y[k]
    common.assert_state_equal(a.state(), b.state())


def test_batch_inverse_truncator_8cube_one_then_ten():
    """The facade's pattern for an 11-frame chisel_ros stream: one frame alone (flushed by the first UpdateMeshes), then ten
    in one batch; InverseTruncator, 8^3 chunks (a chunk is one brick)."""
    setup = Setup(8, 0.1, True, trunc_kind=common.TRUNC_INVERSE, trunc_param=2.0, carve_dist=0.0)
    frames = list(common.orbit_stream(common.SMALL_CAM, 11, total=30, color=True, seed=3))
    a, b = common.Driver(setup, "cuda"), common.Driver(setup, "oracle")
    camv = common.SMALL_CAM.as_array()
    for grp in (frames[:1], frames[1:]):
        a.m.integrate_batch(a.integ, [g[0] for g in grp], [g[2] for g in grp], camv, [g[1] for g in grp])
        want = []
        for depth, col, pose in grp:
            b.integrate(depth, pose, camv, col)
            want.append(b.counters())
        for j, (g, w) in enumerate(zip(a.m.batch_stats(), want)):
            for k in COUNTERS:
                assert g[k] == w[k], "frame %d of the group counter %s: cuda batch %d oracle %d" % (j, k, g[k], w[k])
    common.assert_state_equal(a.state(), b.state())


def test_batch_depth_in_millimetres_converted_on_device():
    """16UC1 depth (ROS): the device conversion equals chisel_ros's host loop, depth = (1.0f / 1000.0f) * mm
    (CR Conversions.h:141-152); zero (no reading) stays 0.0 m like there."""
    setup = Setup(16, 0.04, True)
    cam = common.SMALL_CAM
    frames = list(common.orbit_stream(cam, 6, total=30, color=True, seed=13))
    mm = []
    for depth, _, _ in frames:
        q = np.clip(np.nan_to_num(depth, nan=0.0) * 1000.0, 0, 65535).astype(np.uint16)
        q[::17, ::13] = 0
        mm.append(q)
    k = np.float32(1.0) / np.float32(1000.0)
    a, b = common.Driver(setup, "cuda"), common.Driver(setup, "oracle")
    camv = cam.as_array()
    want = []
    for q, (_, col, pose) in zip(mm, frames):
        b.integrate((k * q.astype(np.float32)).astype(np.float32), pose, camv, col)
        want.append(b.counters())
    a.m.integrate_batch(a.integ, mm[:1], [frames[0][2]], camv, [frames[0][1]])          # a single frame also goes through the fused kernels
    got = a.m.batch_stats()
    a.m.integrate_batch(a.integ, mm[1:], [f[2] for f in frames[1:]], camv, [f[1] for f in frames[1:]])
    got += a.m.batch_stats()
    for j, (g, w) in enumerate(zip(got, want)):
        for c in COUNTERS:
            assert g[c] == w[c], "frame %d counter %s: %d vs %d" % (j, c, g[c], w[c])
    common.assert_state_equal(a.state(), b.state())
    assert np.array_equal(a.dirty(), b.dirty())


def test_full_size_config2_batch_equals_frame_by_frame():
    """BASELINE configs[1] at full size (752x480, 2 cm, colour, NaN pixels): 20 frames as two fused batches of 10 leave exactly
    the map, dirty set, per-frame counters and meshes that 20 single-frame calls leave (the oracle is too slow at this size; the
    single-frame path is pinned to it at small sizes)."""
    cfg = scenes.CONFIG2
    setup = Setup(cfg.chunk, cfg.resolution, True)
    a, b = common.Driver(setup, "cuda"), common.Driver(setup, "cuda")
    camv = cfg.cam.as_array()
    frames = [scenes.stream_frame(cfg, f) for f in range(20)]
    single = []
    for depth, col, pose in frames:
        a.integrate(depth, pose, camv, col)
        single.append(a.counters())
    got = []
    for i in (0, 10):
        grp = frames[i:i + 10]
        b.m.integrate_batch(b.integ, [g[0] for g in grp], [g[2] for g in grp], camv, [g[1] for g in grp])
        got += b.m.batch_stats()
    for j, (g, w) in enumerate(zip(got, single)):
        for k in COUNTERS:
            assert g[k] == w[k], "frame %d counter %s: batch %d single %d" % (j, k, g[k], w[k])
    common.assert_state_equal(a.state(), b.state())
    assert np.array_equal(a.dirty(), b.dirty())
    a.remesh()
    b.remesh()
    common.assert_meshes_equal(a.meshes(), b.meshes())


def test_full_size_config5_eight_agents_one_batch_per_time_step():
    """BASELINE configs[4] shape at full size: 8 agents, 640x480 depth only, 2 cm; the eight frames of a time step form one batch
    (different poses, one union candidate box). Equal to frame-by-frame integration in arrival order."""
    cfg = scenes.CONFIG5
    setup = Setup(cfg.chunk, cfg.resolution, False)
    a, b = common.Driver(setup, "cuda"), common.Driver(setup, "cuda")
    camv = cfg.cam.as_array()
    for t in range(3):
        grp = [scenes.stream_frame(cfg, t, agent=ag) for ag in range(8)]
        for depth, _, pose in grp:
            a.integrate(depth, pose, camv)
        b.m.integrate_batch(b.integ, [g[0] for g in grp], [g[2] for g in grp], camv)
        st = b.m.batch_stats()
        assert len(st) == 8 and all(s["error_flags"] == 0 for s in st)
    common.assert_state_equal(a.state(), b.state())
    assert np.array_equal(a.dirty(), b.dirty())


def _axis_pose(t):
    m = np.zeros((3, 4), np.float32)
    m[0, 0] = m[1, 1] = m[2, 2] = 1.0
    m[:, 3] = t
    return m


def test_batch_exact_fallback_paths():
    """Inputs built to leave the ranges of the fast brick kernel's guard-free arithmetic, so that its exact fallback runs:
    a power-of-two voxel size and an axis-aligned camera on a voxel centre make (a) camera-space z of a voxel layer exactly +0
    (reciprocal out of range), (b) depth - z exactly 0 on a layer inside the band (numerator of DistVoxel::Integrate exactly 0).
    Part of the image sees a surface 6 cm in front of the camera so that the z = 0 layer lies inside the band."""
    cam = common.SMALL_CAM
    frames = []
    for i in range(6):
        pose = _axis_pose((0.03125, 0.03125 * (1 + 2 * i), 0.03125 + 0.0625 * (i % 3)))
        d = np.full((cam.height, cam.width), 1.0 + 0.0625 * (i % 2), np.float32)
        d[:60, :80] = 0.0625
        d[::7, ::5] = np.nan
        col = np.full((cam.height, cam.width, 3), 100 + 10 * i, np.uint8)
        frames.append((d, col, pose))
    a, b = _run_batched(Setup(16, 0.0625, True), frames, cam, 3)
    ids, sdf, w, _ = b.state()
    assert int(((sdf == 0) & (w > 0)).sum()) > 50                   # the zero-numerator case really occurred


def test_batch_state_outside_fast_preconditions():
    """Voxel state outside the fast kernel's per-task preconditions (weight > 2^20, |sdf| > 2^17, negative weight) arrives through
    chs_import_chunks; the batch that follows must treat it exactly like the oracle does."""
    setup = Setup(16, 0.05, True)
    cam = common.SMALL_CAM
    frames = list(common.orbit_stream(cam, 6, total=30, color=True, seed=5))
    a, b = common.Driver(setup, "cuda"), common.Driver(setup, "oracle")
    camv = cam.as_array()
    for depth, col, pose in frames[:2]:
        a.integrate(depth, pose, camv, col)
        b.integrate(depth, pose, camv, col)
    ids, sdf, w, rgbw = b.state()
    rng = np.random.RandomState(11)
    sdf, w = sdf.copy(), w.copy()
    obs = np.argwhere(w > 0)
    pick = obs[rng.choice(len(obs), 600, replace=False)]
    for j, (c, v) in enumerate(pick):
        if j % 3 == 0:
            w[c, v] = np.float32(3.0e6)
        elif j % 3 == 1:
            sdf[c, v] = np.float32(-2.5e5)
        else:
            w[c, v] = np.float32(-1.5)
    a.m.import_chunks(ids, sdf, w, rgbw)
    b.m.import_chunks(ids, sdf, w, rgbw)
    grp = frames[2:]
    a.m.integrate_batch(a.integ, [g[0] for g in grp], [g[2] for g in grp], camv, [g[1] for g in grp])
    for depth, col, pose in grp:
        b.integrate(depth, pose, camv, col)
    common.assert_state_equal(a.state(), b.state())

"""GPU suite, BASELINE.json configs at FULL size against the CPU oracle (not against another CUDA path): the headline
configs[1] stream as fused 10-frame batches, the configs[4] eight-agent shape (one batch per time step), the configs[2]
four-agent interleave and the configs[3] 1 cm hall with a re-mesh. Bit-exact voxel state, chunk-ID set, dirty set, per-frame
counters and meshes. The C oracle integrates a 752x480 / 2 cm frame in ~0.4 s, so these stay in the tens of seconds.

The Hi-Z levels >= 128 px, the capacity planner's deferred pool sizing and the multi-slab pool only come into play at these
sizes -- a size-dependent culling bug would drop updates here and nowhere in the small-image tests."""
import numpy as np
import pytest

from cvids_b200 import scenes
from tests import common
from tests.common import Setup
from tests.test_batch_gpu import COUNTERS, _run_batched

pytestmark = pytest.mark.gpu


def test_config2_headline_fused_batches_vs_oracle():
    """configs[1]: 752x480 depth+colour, 2 cm, 2 % NaN pixels; frames 0..19 as two fused batches of 10 (the bench's step)."""
    cfg = scenes.CONFIG2
    frames = [scenes.stream_frame(cfg, f) for f in range(20)]
    a, b = _run_batched(Setup(cfg.chunk, cfg.resolution, True), frames, cfg.cam, 10)
    ids, _, w, rgbw = a.state()
    assert len(ids) > 400 and int((w > 0).sum()) > 500_000 and rgbw[..., 3].max() == 8


def test_config2_headline_single_frame_calls_vs_oracle():
    """configs[1] through the reference's call granularity (one chs_integrate_depth_color per frame) at full size."""
    cfg = scenes.CONFIG2
    a, b = common.Driver(Setup(cfg.chunk, cfg.resolution, True), "cuda"), common.Driver(Setup(cfg.chunk, cfg.resolution, True), "oracle")
    camv = cfg.cam.as_array()
    for f in range(4):
        depth, col, pose = scenes.stream_frame(cfg, f * 7)
        a.integrate(depth, pose, camv, col)
        b.integrate(depth, pose, camv, col)
        ca, cb = a.counters(), b.counters()
        for k in COUNTERS:
            assert ca[k] == cb[k], "frame %d counter %s: cuda %d oracle %d" % (f, k, ca[k], cb[k])
    common.assert_state_equal(a.state(), b.state())
    assert np.array_equal(a.dirty(), b.dirty())
    a.remesh()
    b.remesh()
    common.assert_meshes_equal(a.meshes(), b.meshes())


def test_config2_mid_stream_batches_vs_oracle():
    """The same stream far from the start (frames 95..124: the orbit has turned by more than half a revolution, the candidate
    boxes move through negative chunk IDs) in batches of 10, map seeded by frame-by-frame integration of every 5th earlier frame."""
    cfg = scenes.CONFIG2
    setup = Setup(cfg.chunk, cfg.resolution, True)
    a, b = common.Driver(setup, "cuda"), common.Driver(setup, "oracle")
    camv = cfg.cam.as_array()
    for f in range(0, 95, 5):
        depth, col, pose = scenes.stream_frame(cfg, f)
        a.integrate(depth, pose, camv, col)
        b.integrate(depth, pose, camv, col)
    for lo in (95, 105, 115):
        grp = [scenes.stream_frame(cfg, f) for f in range(lo, lo + 10)]
        a.m.integrate_batch(a.integ, [g[0] for g in grp], [g[2] for g in grp], camv, [g[1] for g in grp])
        want = []
        for depth, col, pose in grp:
            b.integrate(depth, pose, camv, col)
            want.append(b.counters())
        for j, (g, w) in enumerate(zip(a.m.batch_stats(), want)):
            for k in COUNTERS:
                assert g[k] == w[k], "frame %d counter %s: cuda batch %d oracle %d" % (lo + j, k, g[k], w[k])
    common.assert_state_equal(a.state(), b.state())
    assert np.array_equal(a.dirty(), b.dirty())
    a.remesh()
    b.remesh()
    common.assert_meshes_equal(a.meshes(), b.meshes())


def test_config5_eight_agents_vs_oracle():
    """configs[4] shape: 8 agents x 640x480 depth, 2 cm; the eight frames of a time step are ONE batch (eight different poses,
    one union candidate box), arrival order = agent order (CR ChiselServer.cpp:297-367: every agent's frame lands in the same
    map in callback order)."""
    cfg = scenes.CONFIG5
    frames = [scenes.stream_frame(cfg, t, agent=ag) for t in range(3) for ag in range(cfg.agents)]
    _run_batched(Setup(cfg.chunk, cfg.resolution, False), frames, cfg.cam, cfg.agents)


def test_config3_four_agents_interleaved_vs_oracle():
    """configs[2] shape: four trajectories (phase offsets pi/2) fused into one map, round-robin per time step, 640x480, 5 cm."""
    cfg = scenes.CONFIG3
    frames = [scenes.stream_frame(cfg, t, agent=ag) for t in range(0, 40, 4) for ag in range(cfg.agents)]
    _run_batched(Setup(cfg.chunk, cfg.resolution, False), frames, cfg.cam, 8)


def test_hall_1cm_frames_and_remesh_vs_oracle():
    """configs[3] shape: 50 x 50 x 5 m pillar hall at 1 cm, 640x480: three frames in one batch (candidate boxes of ~50 k chunks: the
    deferred pool sizing of the fused path), then a re-mesh of everything dirty."""
    hall = scenes.hall(seed=3)
    cam = scenes.Camera(525.0, 525.0, 319.5, 239.5, 640, 480, near=0.05, far=5.0)
    # yaw 0 exactly makes the reference's frustum predicate reject every chunk (quirk Q5): kept as the first frame on purpose
    poses = [scenes.yaw_pose(0.35 * i, (-20.0 + 0.15 * i, -20.0 + 0.9 * i, 0.0)) for i in range(3)]
    frames = [(scenes.render(hall, cam, p)[0], None, p) for p in poses]
    a, b = _run_batched(Setup(16, 0.01, False), frames, cam, 3)
    tris = sum(len(m["vertices"]) for m in a.meshes().values()) // 3
    assert tris > 100_000, tris

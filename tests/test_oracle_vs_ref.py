"""CPU suite, build container only: the C restatement against the compiled reference (oracle/_ref), live and bit for
bit, on streams that are NOT in the golden set. Skipped where /root/reference (and hence oracle/_ref) is absent."""
import numpy as np
import pytest

from cvids_b200 import scenes
from oracle import pyoracle
from tests import common
from tests.common import Setup

pytestmark = pytest.mark.skipif(not pyoracle.ref_available(), reason="oracle/_ref not built (no /root/reference here)")


def _run_pair(setup, frames, cam, remesh_every=3):
    a, b = common.Driver(setup, "ref"), common.Driver(setup, "oracle")
    for i, (depth, col, pose) in enumerate(frames):
        a.integrate(depth, pose, cam.as_array(), col)
        b.integrate(depth, pose, cam.as_array(), col)
        if (i + 1) % remesh_every == 0:
            a.remesh()
            b.remesh()
    common.assert_state_equal(a.state(), b.state())
    assert np.array_equal(a.dirty(), b.dirty())
    common.assert_meshes_equal(a.meshes(), b.meshes())
    return a, b


def test_depth_path_mid_camera():
    _run_pair(Setup(16, 0.05, False), common.orbit_stream(common.MID_CAM, 6, total=30, seed=11, phase=0.3), common.MID_CAM)


def test_color_path_noise_and_nan():
    _run_pair(Setup(16, 0.04, True, weight=3.0),
              common.orbit_stream(common.SMALL_CAM, 7, total=30, color=True, nan_frac=0.05, seed=5, noise=0.01), common.SMALL_CAM)


def test_camera_far_from_origin_negative_ids():
    """Chunk IDs far from zero and negative: exercises GetIDAt / floor for negative coordinates."""
    off = (-37.3, 12.9, -3.4)
    scene = scenes.Scene(tuple(np.add(scenes.ROOM.lo, off)), tuple(np.add(scenes.ROOM.hi, off)))

    def frames():
        for f in range(4):
            pose = scenes.yaw_pose(0.4 * f + 2.0, tuple(np.add((0.2 * f, -0.1 * f, 0.05), off)))
            depth, col = scenes.render(scene, common.SMALL_CAM, pose, color=True, channels=4)
            yield depth, col, pose
    a, _ = _run_pair(Setup(8, 0.05, True), frames(), common.SMALL_CAM, remesh_every=2)
    assert a.state()[0].min() < -40


def test_frustum_and_candidates():
    from oracle.pyoracle import OracleChisel, RefChisel
    r, o = RefChisel(16, 0.05, False), OracleChisel(16, 0.05, False)
    for f in range(0, 40, 7):
        pose = scenes.orbit_pose(f, 40, 0.1)
        cam = scenes.EUROC_752.as_array()
        for x, y in zip(r.frustum(pose, cam), o.frustum(pose, cam)):
            assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
        assert np.array_equal(r.candidate_ids(pose, cam), o.candidate_ids(pose, cam))


def test_truncators_bitwise():
    from oracle.pyoracle import OracleChisel, RefChisel
    r, o = RefChisel(16, 0.05, False), OracleChisel(16, 0.05, False)
    rng = np.random.RandomState(0)
    for d in np.concatenate([rng.uniform(0.05, 20, 200), [0.0, 1e-3, 50.0, 99.9]]).astype(np.float32):
        for kind, p in ((0, 0.2), (1, 2.0), (1, 8.0), (2, 2.0), (2, 8.0)):
            x, y = np.float32(r.truncation(kind, p, float(d))), np.float32(o.truncation(kind, p, float(d)))
            assert x.view(np.uint32) == y.view(np.uint32) or (np.isnan(x) and np.isnan(y)), (kind, p, d, x, y)

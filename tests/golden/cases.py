"""The streams behind the golden fixtures; shared by make_golden.py (reference side) and the tests."""
from __future__ import annotations

from cvids_b200 import scenes
from tests import common
from tests.common import Setup

CASES = {
    # BASELINE config 1 shape at a quarter of the image size: depth only, 5 cm, 16^3, constant truncation
    "room_depth_5cm": dict(setup=Setup(16, 0.05, False), stream=("orbit", dict(cam="small", n_frames=12, total=24)),
                           remesh_every=5),
    # BASELINE config 2 shape, reduced: colour path with NaN pixels, 4 cm voxels
    "room_color_4cm_nan": dict(setup=Setup(16, 0.04, True, weight=1.0),
                               stream=("orbit", dict(cam="small", n_frames=8, total=24, color=True, nan_frac=0.02, seed=1)),
                               remesh_every=4),
    # obstacle removed mid-stream: exercises both carving variants
    "carve_depth": dict(setup=Setup(16, 0.05, False), stream=("carve", dict(cam="small", color=False)), remesh_every=3),
    "carve_color": dict(setup=Setup(16, 0.05, True, weight=2.0), stream=("carve", dict(cam="small", color=True)), remesh_every=3),
    # the truncator chisel_ros actually instantiates (CR ChiselNode.cpp:98) and 8^3 chunks (CR launch/sample.launch:6-8)
    "inverse_trunc_8cube": dict(setup=Setup(8, 0.1, True, trunc_kind=common.TRUNC_INVERSE, trunc_param=2.0, carve_dist=0.0),
                                stream=("orbit", dict(cam="small", n_frames=6, total=24, color=True)), remesh_every=3),
    "quadratic_trunc": dict(setup=Setup(16, 0.05, False, trunc_kind=common.TRUNC_QUADRATIC, trunc_param=2.0),
                            stream=("orbit", dict(cam="small", n_frames=5, total=24)), remesh_every=5),
}
FULL_STATE_CASE = "inverse_trunc_8cube"

CAMS = {"small": common.SMALL_CAM, "mid": common.MID_CAM}


def frames_of(case):
    kind, kw = case["stream"]
    kw = dict(kw)
    cam = CAMS[kw.pop("cam")]
    if kind == "orbit":
        return cam, common.orbit_stream(cam, **kw)
    if kind == "carve":
        return cam, common.carve_stream(cam, **kw)
    raise ValueError(kind)


def run_case(drv, case):
    cam, frames = frames_of(case)
    camv = cam.as_array()
    for i, (depth, col, pose) in enumerate(frames):
        drv.integrate(depth, pose, camv, col)
        if (i + 1) % case["remesh_every"] == 0:
            drv.remesh()

"""CPU suite: the C restatement (oracle/chisel_oracle.c) against the golden fixtures that the compiled, unmodified
reference produced (tests/golden/make_golden.py). This is what pins the oracle on machines without /root/reference."""
import json
import os

import numpy as np
import pytest

from tests import common
from tests.golden import cases

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
DIGESTS = json.load(open(os.path.join(GOLDEN, "digests.json")))


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_oracle_matches_reference_digests(name):
    case = cases.CASES[name]
    drv = common.Driver(case["setup"], "oracle")
    cases.run_case(drv, case)
    want = DIGESTS[name]
    assert common.digest_state(drv.state()) == want["state"]
    assert common.digest_meshes(drv.meshes()) == want["meshes"]
    assert len(drv.dirty()) == want["dirty"]


def test_oracle_matches_reference_full_state():
    name = cases.FULL_STATE_CASE
    case = cases.CASES[name]
    drv = common.Driver(case["setup"], "oracle")
    cases.run_case(drv, case)
    g = np.load(os.path.join(GOLDEN, "small_state.npz"))
    common.assert_state_equal(drv.state(), (g["ids"], g["sdf"], g["weight"], g["rgbw"]), name)
    assert np.array_equal(drv.dirty(), g["dirty"].reshape(-1, 3))
    meshes = drv.meshes()
    gold = {}
    for i, k in enumerate(g["mesh_ids"]):
        gold[tuple(int(x) for x in k)] = {f: g["mesh%d_%s" % (i, f)] for f in ("vertices", "normals", "colors", "grids")}
    common.assert_meshes_equal(meshes, gold, name)


def test_carving_is_exercised():
    """The carve fixtures must actually carve, otherwise they pin nothing (ProjectionIntegrator.h:88-95, 166-178)."""
    for name in ("carve_depth", "carve_color"):
        case = cases.CASES[name]
        drv = common.Driver(case["setup"], "oracle")
        cam, frames = cases.frames_of(case)
        carved = 0
        for depth, col, pose in frames:
            drv.integrate(depth, pose, cam.as_array(), col)
            carved += drv.counters()["n_carve"]
        assert carved > 1000, (name, carved)


def test_primitive_known_answers():
    """Hand-derived known answers."""
    from oracle.pyoracle import OracleChisel
    f32 = np.float32
    # ConstantTruncator returns its value; InverseTruncator: (1/(0.10*471.27)) / (1/d)^2 * scale in float
    assert OracleChisel(16, 0.05, False).truncation(0, 0.2, 3.0) == f32(0.2)
    dep = f32(1.0) / (f32(0.10) * f32(471.27))
    inv = f32(1.0 / float(f32(2.0)))
    assert OracleChisel(16, 0.05, False).truncation(2, 8.0, 2.0) == f32(f32(dep / f32(inv * inv)) * f32(8.0))
    # QuadraticTruncator at d = 1, scale 1: |0.019 + 0.0152 + 0.01504| evaluated in double from float constants
    q = abs(float(f32(0.019)) * 1.0 + float(f32(f32(0.0152) * f32(1.0))) + float(f32(0.01504)))
    assert OracleChisel(16, 0.05, False).truncation(1, 1.0, 1.0) == f32(q)

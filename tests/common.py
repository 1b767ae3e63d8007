"""Shared harness for the parity tests: drive the CUDA library (through the C ABI), the C oracle and the compiled
reference through one interface on the same seeded synthetic streams, and compare their state."""
from __future__ import annotations

import dataclasses
import hashlib

import numpy as np

from cvids_b200 import scenes

TRUNC_CONSTANT, TRUNC_QUADRATIC, TRUNC_INVERSE = 0, 1, 2


@dataclasses.dataclass
class Setup:
    """Map + integrator configuration of one test stream."""
    chunk: int = 16
    resolution: float = 0.05
    color: bool = False
    trunc_kind: int = TRUNC_CONSTANT
    trunc_param: float | None = None     # None: 4 voxels
    weight: float = 1.0
    carve: bool = True
    carve_dist: float = 0.05

    @property
    def trunc(self):
        if self.trunc_param is not None:
            return self.trunc_param
        return float(np.float32(4.0) * np.float32(self.resolution))


def make_oracle(setup: Setup, cls=None):
    from oracle.pyoracle import OracleChisel
    cls = cls or OracleChisel
    o = cls(setup.chunk, setup.resolution, setup.color)
    o.setup_integrator(setup.trunc_kind, setup.trunc, setup.weight, setup.carve, setup.carve_dist)
    return o


def make_cuda(setup: Setup, **kw):
    from cvids_b200 import capi
    m = capi.Chisel(setup.chunk, setup.resolution, setup.color, **kw)
    integ = capi.ProjectionIntegrator(setup.trunc_kind, setup.trunc, setup.weight, setup.carve, setup.carve_dist)
    return m, integ


class Driver:
    """Uniform front-end: .integrate(depth, pose, cam, color=None, ...) / .remesh() / .state() / .dirty() / .meshes()"""

    def __init__(self, setup: Setup, impl: str, **kw):
        self.setup, self.impl = setup, impl
        if impl == "cuda":
            self.m, self.integ = make_cuda(setup, **kw)
        elif impl == "oracle":
            self.m = make_oracle(setup)
        elif impl == "ref":
            from oracle.pyoracle import RefChisel
            self.m = make_oracle(setup, RefChisel)
        else:
            raise ValueError(impl)

    def integrate(self, depth, pose, cam, color=None, color_pose=None, color_cam=None, force_color_path=None):
        use_color_path = (color is not None) if force_color_path is None else force_color_path
        if self.impl == "cuda":
            if use_color_path:
                self.m.integrate_depth_scan_color(self.integ, depth, pose, cam, color, color_pose, color_cam)
            else:
                self.m.integrate_depth_scan(self.integ, depth, pose, cam)
        else:
            if use_color_path:
                self.m.integrate_color(depth, pose, cam, color, color_pose, color_cam)
            else:
                self.m.integrate_depth(depth, pose, cam)

    def remesh(self):
        if self.impl == "cuda":
            self.m.recompute_meshes()
        else:
            self.m.update_meshes()

    def state(self):
        return self.m.state()

    def dirty(self):
        return self.m.dirty_ids()

    def meshes(self):
        if self.impl == "cuda":
            return self.m.chunk_manager.get_all_meshes()
        return self.m.all_meshes()

    def counters(self):
        if self.impl == "cuda":
            return self.m.frame_stats()
        return self.m.frame_counters()


def bits(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a).view(np.uint32) if a.dtype == np.float32 else a


def assert_state_equal(a, b, what=""):
    """Bit-exact voxel state: chunk-ID set, sdf, weight, colour (stronger than the north star's tolerances:
    SDF within 1e-4 x truncation, colours within 1 LSB, integer state exact)."""
    ia, sa, wa, ca = a
    ib, sb, wb, cb = b
    assert ia.shape == ib.shape and np.array_equal(ia, ib), "%s: allocated-chunk sets differ (%d vs %d)" % (what, len(ia), len(ib))
    assert np.array_equal(bits(wa), bits(wb)), "%s: voxel weights differ in %d voxels" % (what, int((bits(wa) != bits(wb)).sum()))
    assert np.array_equal(bits(sa), bits(sb)), "%s: SDF differs in %d voxels (max abs %g)" % (
        what, int((bits(sa) != bits(sb)).sum()), float(np.nanmax(np.abs(sa - sb))))
    assert np.array_equal(ca, cb), "%s: colour voxels differ in %d entries" % (what, int((ca != cb).sum()))


def assert_meshes_equal(ma: dict, mb: dict, what="", with_colors=True):
    """Index-for-index identical triangle soup per chunk: vertices, normals, colours, grids (north star: identical
    triangle counts, vertices within 1e-5 m; we require bit equality)."""
    assert sorted(ma) == sorted(mb), "%s: meshed chunk sets differ (%d vs %d)" % (what, len(ma), len(mb))
    fields = ("vertices", "normals", "grids") + (("colors",) if with_colors else ())
    for k in ma:
        for f in fields:
            x, y = ma[k][f], mb[k][f]
            assert x.shape == y.shape, "%s: chunk %s %s count %s vs %s" % (what, k, f, x.shape, y.shape)
            if not np.array_equal(bits(x), bits(y)):
                bad = np.argwhere(bits(x) != bits(y))
                raise AssertionError("%s: chunk %s %s differs at %d entries, first %s: %r vs %r" % (
                    what, k, f, len(bad), bad[0], x[tuple(bad[0])], y[tuple(bad[0])]))


def digest_state(state) -> dict:
    ids, sdf, w, rgbw = state
    h = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    return dict(n_chunks=int(len(ids)), ids=h(ids), sdf=h(sdf), weight=h(w), rgbw=h(rgbw),
                observed=int((w > 0).sum()), weight_sum=float(w.astype(np.float64).sum()))


def digest_meshes(meshes: dict) -> dict:
    h = hashlib.sha256()
    tris = grids = 0
    for k in sorted(meshes):
        m = meshes[k]
        h.update(np.asarray(k, np.int32).tobytes())
        for f in ("vertices", "normals", "colors", "grids"):
            h.update(np.ascontiguousarray(m[f]).tobytes())
        tris += len(m["vertices"]) // 3
        grids += len(m["grids"])
    return dict(n_meshes=len(meshes), triangles=tris, grids=grids, sha=h.hexdigest())


# ---------------------------------------------------------------------------------------------------------------
# test streams (small enough for the oracle to finish in seconds)

SMALL_CAM = scenes.Camera(131.25, 131.25, 79.5, 59.5, 160, 120)          # KINECT_640 scaled by 1/4
MID_CAM = scenes.Camera(262.5, 262.5, 159.5, 119.5, 320, 240)


def orbit_stream(cam, n_frames, total=40, color=False, channels=3, nan_frac=0.0, scene=scenes.ROOM, seed=0, phase=0.0,
                 noise=0.0):
    for f in range(n_frames):
        pose = scenes.orbit_pose(f, total, phase)
        depth, col = scenes.render(scene, cam, pose, color=color, channels=channels, nan_frac=nan_frac, seed=seed + f,
                                   noise_sigma=noise)
        yield depth, col, pose


BOX_SCENE = scenes.Scene(scenes.ROOM.lo, scenes.ROOM.hi, (((1.0, -0.5, -1.5), (1.6, 0.5, 0.2)),))


def carve_stream(cam, n_before=4, n_after=6, color=False):
    """An obstacle is observed, then removed: the later frames carve the voxels it left behind."""
    pose = scenes.yaw_pose(0.0, (-1.0, 0.0, 0.0))
    for f in range(n_before + n_after):
        scene = BOX_SCENE if f < n_before else scenes.ROOM
        p = scenes.yaw_pose(0.02 * f, (-1.0 + 0.01 * f, 0.0, 0.0))
        depth, col = scenes.render(scene, cam, p, color=color)
        yield depth, col, p

"""ctypes front-ends for the two CPU checkers. TEST INFRASTRUCTURE, not a product component.

  RefChisel    -> oracle/_ref/libchisel_ref.so   (the unmodified reference sources + ref_driver.cpp)
  OracleChisel -> oracle/libchisel_oracle.so     (the plain-C restatement, chisel_oracle.c)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module. Both classes expose the same methods so that tests can swap them.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libchisel_ref.so")
ORACLE_SO = os.path.join(HERE, "libchisel_oracle.so")
REFERENCE_ROOT = "/root/reference/OpenChisel/open_chisel"

TRUNC_CONSTANT, TRUNC_QUADRATIC, TRUNC_INVERSE = 0, 1, 2

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def build(target: str) -> None:
    subprocess.check_call(["make", "-s", "-C", HERE, target])


def ref_available() -> bool:
    return os.path.exists(REF_SO)


def _as_pose(p) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(p, dtype=np.float32).reshape(3, 4))
    return a


class _Base:
    """Shared result accessors; subclasses set self._lib, self._h and the symbol prefix."""
    _prefix = ""

    def _fn(self, name):
        return getattr(self._lib, self._prefix + name)

    # ---- state ----
    def chunk_ids(self) -> np.ndarray:
        n = self._fn("num_chunks")(self._h)
        out = np.zeros((n, 3), dtype=np.int32)
        if n:
            self._fn("chunk_ids")(self._h, out)
        return out

    def chunk_voxels(self, cid):
        V = self.chunk ** 3
        sdf = np.zeros(V, np.float32)
        w = np.zeros(V, np.float32)
        rgbw = np.zeros((V, 4), np.uint8)
        ok = self._fn("chunk_voxels")(self._h, np.asarray(cid, np.int32), sdf, w, rgbw)
        if not ok:
            raise KeyError(tuple(cid))
        return sdf, w, rgbw

    def dirty_ids(self) -> np.ndarray:
        n = self._fn("num_dirty")(self._h)
        out = np.zeros((n, 3), dtype=np.int32)
        if n:
            self._fn("dirty_ids")(self._h, out)
        return out

    def mesh_ids(self) -> np.ndarray:
        n = self._fn("num_meshes")(self._h)
        out = np.zeros((n, 3), dtype=np.int32)
        if n:
            self._fn("mesh_ids")(self._h, out)
        return out

    def mesh(self, cid):
        """dict(vertices [n,3], normals [n,3], colors [m,3], grids [g,3]) for one chunk."""
        cid = np.asarray(cid, np.int32)
        sizes = np.zeros(5, np.int64)
        if not self._fn("mesh_sizes")(self._h, cid, sizes):
            raise KeyError(tuple(cid))
        v = np.zeros((sizes[0], 3), np.float32)
        nrm = np.zeros((sizes[1], 3), np.float32)
        col = np.zeros((sizes[2], 3), np.float32)
        g = np.zeros((sizes[3], 3), np.float32)
        idx = np.zeros(sizes[4], np.int64)
        self._fn("mesh_data")(self._h, cid, v, nrm, col, g, idx)
        return dict(vertices=v, normals=nrm, colors=col, grids=g, indices=idx)

    def state(self):
        """Full voxel state: ids [n,3] sorted, sdf [n,V], weight [n,V], rgbw [n,V,4]."""
        ids = self.chunk_ids()
        V = self.chunk ** 3
        sdf = np.zeros((len(ids), V), np.float32)
        w = np.zeros((len(ids), V), np.float32)
        rgbw = np.zeros((len(ids), V, 4), np.uint8)
        for i, cid in enumerate(ids):
            sdf[i], w[i], rgbw[i] = self.chunk_voxels(cid)
        return ids, sdf, w, rgbw

    def all_meshes(self):
        return {tuple(int(x) for x in cid): self.mesh(cid) for cid in self.mesh_ids()}


_ref_copies = {}


def _load_ref_copy(key):
    """One private copy of the reference library per (chunk, resolution): ChunkManager::GetIDAt keeps
    function-local statics (ChunkManager.h:138-140, quirk Q1), so one loaded image must only ever see
    one resolution."""
    if key in _ref_copies:
        return _ref_copies[key]
    if not ref_available():
        raise RuntimeError("oracle/_ref/libchisel_ref.so is missing: run `make -C oracle ref` where "
                           "/root/reference exists")
    tmpdir = tempfile.mkdtemp(prefix="chisel_ref_")
    path = os.path.join(tmpdir, "libchisel_ref_%d.so" % len(_ref_copies))
    shutil.copy(REF_SO, path)
    lib = C.CDLL(path)
    _declare(lib, "ref_")
    _ref_copies[key] = lib
    return lib


def _declare(lib, p):
    g = lambda n: getattr(lib, p + n)
    g("create").restype = C.c_void_p
    g("create").argtypes = [C.c_int, C.c_float, C.c_int]
    g("destroy").argtypes = [C.c_void_p]
    g("reset").argtypes = [C.c_void_p]
    g("setup_integrator").argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_float]
    g("truncation").restype = C.c_float
    g("truncation").argtypes = [C.c_int, C.c_float, C.c_float]
    g("integrate_depth").argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, _f32p, _f32p]
    g("integrate_color").argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, _f32p, _f32p,
                                     _u8p, C.c_int, C.c_int, C.c_int, _f32p, _f32p, C.c_int]
    g("candidate_ids").restype = C.c_int
    g("candidate_ids").argtypes = [C.c_void_p, _f32p, _f32p, _i32p, C.c_int]
    g("frustum").argtypes = [_f32p, _f32p, _f32p, _f32p, _f32p]
    g("update_meshes").argtypes = [C.c_void_p, C.c_int]
    for n in ("num_chunks", "num_dirty", "num_meshes"):
        g(n).restype = C.c_int
        g(n).argtypes = [C.c_void_p]
    for n in ("chunk_ids", "dirty_ids", "mesh_ids"):
        g(n).argtypes = [C.c_void_p, _i32p]
    g("chunk_voxels").restype = C.c_int
    g("chunk_voxels").argtypes = [C.c_void_p, _i32p, _f32p, _f32p, _u8p]
    g("mesh_sizes").restype = C.c_int
    g("mesh_sizes").argtypes = [C.c_void_p, _i32p, _i64p]
    g("mesh_data").restype = C.c_int
    g("mesh_data").argtypes = [C.c_void_p, _i32p, _f32p, _f32p, _f32p, _f32p, _i64p]
    g("last_counts").argtypes = [C.c_void_p, _i64p]


class RefChisel(_Base):
    """The reference's own OpenChisel, driven through oracle/ref_driver.cpp."""
    _prefix = "ref_"
    kind = "reference"

    def __init__(self, chunk: int, resolution: float, use_color: bool):
        self.chunk, self.resolution, self.use_color = chunk, float(np.float32(resolution)), use_color
        self._lib = _load_ref_copy((chunk, self.resolution))
        self._h = self._lib.ref_create(chunk, resolution, int(use_color))

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.ref_destroy(self._h)
            self._h = None

    def setup_integrator(self, trunc_kind, trunc_param, weight, carve, carve_dist):
        self._lib.ref_setup_integrator(self._h, trunc_kind, trunc_param, weight, int(carve), carve_dist)

    def truncation(self, kind, param, depth):
        return self._lib.ref_truncation(kind, param, depth)

    def integrate_depth(self, depth, pose, cam):
        H, W = depth.shape
        self._lib.ref_integrate_depth(self._h, np.ascontiguousarray(depth, np.float32), W, H, _as_pose(pose),
                                      np.ascontiguousarray(cam, np.float32))

    def integrate_color(self, depth, pose, cam, color, cpose=None, ccam=None, as_is=False):
        H, W = depth.shape
        cH, cW, ch = color.shape
        cpose = pose if cpose is None else cpose
        ccam = cam if ccam is None else ccam
        self._lib.ref_integrate_color(self._h, np.ascontiguousarray(depth, np.float32), W, H, _as_pose(pose),
                                      np.ascontiguousarray(cam, np.float32), np.ascontiguousarray(color), cW, cH, ch,
                                      _as_pose(cpose), np.ascontiguousarray(ccam, np.float32), int(as_is))

    def candidate_ids(self, pose, cam) -> np.ndarray:
        cap = 1 << 16
        while True:
            out = np.zeros((cap, 3), np.int32)
            n = self._lib.ref_candidate_ids(self._h, _as_pose(pose), np.ascontiguousarray(cam, np.float32), out, cap)
            if n <= cap:
                return out[:n]
            cap = n

    def frustum(self, pose, cam):
        corners = np.zeros((8, 3), np.float32)
        lines = np.zeros((24, 3), np.float32)
        planes = np.zeros((6, 4), np.float32)
        self._lib.ref_frustum(_as_pose(pose), np.ascontiguousarray(cam, np.float32), corners, lines, planes)
        return corners, lines, planes

    def update_meshes(self, as_is=False):
        self._lib.ref_update_meshes(self._h, int(as_is))

    def reset(self):
        self._lib.ref_reset(self._h)

    def last_counts(self):
        out = np.zeros(3, np.int64)
        self._lib.ref_last_counts(self._h, out)
        return dict(candidates=int(out[0]), new=int(out[1]), garbage=int(out[2]))


_oracle_lib = None


def _load_oracle():
    global _oracle_lib
    if _oracle_lib is None:
        src = os.path.join(HERE, "chisel_oracle.c")
        if (not os.path.exists(ORACLE_SO)) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
            build("oracle")
        lib = C.CDLL(ORACLE_SO)
        _declare(lib, "orc_")
        lib.orc_frame_counters.argtypes = [C.c_void_p, _i64p]
        _oracle_lib = lib
    return _oracle_lib


COUNTER_NAMES = ("candidates", "visited", "n_upd", "n_carve", "n_col", "n_new", "updated_chunks", "garbage")


class OracleChisel(RefChisel):
    """The plain-C restatement (oracle/chisel_oracle.c). Same methods as RefChisel."""
    _prefix = "orc_"
    kind = "port"

    def __init__(self, chunk: int, resolution: float, use_color: bool):
        self.chunk, self.resolution, self.use_color = chunk, float(np.float32(resolution)), use_color
        self._lib = _load_oracle()
        self._h = self._lib.orc_create(chunk, resolution, int(use_color))

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.orc_destroy(self._h)
            self._h = None

    def setup_integrator(self, trunc_kind, trunc_param, weight, carve, carve_dist):
        self._lib.orc_setup_integrator(self._h, trunc_kind, trunc_param, weight, int(carve), carve_dist)

    def truncation(self, kind, param, depth):
        return self._lib.orc_truncation(kind, param, depth)

    def integrate_depth(self, depth, pose, cam):
        H, W = depth.shape
        self._lib.orc_integrate_depth(self._h, np.ascontiguousarray(depth, np.float32), W, H, _as_pose(pose),
                                      np.ascontiguousarray(cam, np.float32))

    def integrate_color(self, depth, pose, cam, color, cpose=None, ccam=None, as_is=False):
        H, W = depth.shape
        cH, cW, ch = color.shape
        cpose = pose if cpose is None else cpose
        ccam = cam if ccam is None else ccam
        self._lib.orc_integrate_color(self._h, np.ascontiguousarray(depth, np.float32), W, H, _as_pose(pose),
                                      np.ascontiguousarray(cam, np.float32), np.ascontiguousarray(color), cW, cH, ch,
                                      _as_pose(cpose), np.ascontiguousarray(ccam, np.float32), 0)

    def import_chunks(self, ids, sdf, weight, rgbw=None):
        """Test-only state injection (same signature as capi.Chisel.import_chunks): overwrite / create the chunks `ids`."""
        ids = np.ascontiguousarray(np.asarray(ids, np.int32).reshape(-1, 3))
        fn = self._lib.orc_set_chunk_voxels
        fn.argtypes = [C.c_void_p, _i32p, _f32p, _f32p, C.c_void_p]
        for i, cid in enumerate(ids):
            c = np.ascontiguousarray(rgbw[i], np.uint8) if rgbw is not None else None
            fn(self._h, np.ascontiguousarray(cid), np.ascontiguousarray(sdf[i], np.float32), np.ascontiguousarray(weight[i], np.float32),
               c.ctypes.data if c is not None else None)

    def candidate_ids(self, pose, cam) -> np.ndarray:
        cap = 1 << 16
        while True:
            out = np.zeros((cap, 3), np.int32)
            n = self._lib.orc_candidate_ids(self._h, _as_pose(pose), np.ascontiguousarray(cam, np.float32), out, cap)
            if n <= cap:
                return out[:n]
            cap = n

    def frustum(self, pose, cam):
        corners = np.zeros((8, 3), np.float32)
        lines = np.zeros((24, 3), np.float32)
        planes = np.zeros((6, 4), np.float32)
        self._lib.orc_frustum(_as_pose(pose), np.ascontiguousarray(cam, np.float32), corners, lines, planes)
        return corners, lines, planes

    def update_meshes(self, as_is=False):
        self._lib.orc_update_meshes(self._h, 0)

    def reset(self):
        self._lib.orc_reset(self._h)

    def last_counts(self):
        out = np.zeros(3, np.int64)
        self._lib.orc_last_counts(self._h, out)
        return dict(candidates=int(out[0]), new=int(out[1]), garbage=int(out[2]))

    def frame_counters(self):
        out = np.zeros(8, np.int64)
        self._lib.orc_frame_counters(self._h, out)
        return dict(zip(COUNTER_NAMES, (int(x) for x in out)))

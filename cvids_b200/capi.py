"""ctypes binding of the C ABI (include/chisel_b200.h) and a thin host-side mirror of the reference's
`chisel::Chisel` call surface used by the Python harness (tests/, bench.py).

This module is the product path for Python callers. It never imports anything under oracle/ and it has no
CPU fallback: if libchisel_b200.so is missing it is built with nvcc, and every compute call fails loudly
(ChiselError) when there is no CUDA device.

Method names follow the reference's API one for one:
    Chisel.integrate_depth_scan        chisel::Chisel::IntegrateDepthScan<float>             OC Chisel.h:59-112
    Chisel.integrate_depth_scan_color  chisel::Chisel::IntegrateDepthScanColor<float,uint8>   OC Chisel.h:114-213
    Chisel.update_meshes               chisel::Chisel::UpdateMeshes (every-10th gate kept)    OC Chisel.cpp:50-59
    Chisel.get_meshes_to_update        chisel::Chisel::GetMeshesToUpdate                      OC Chisel.h:221-224
    Chisel.reset                       chisel::Chisel::Reset                                  OC Chisel.cpp:44-48
    Chisel.chunk_manager.*             ChunkManager::GetChunks/HasChunk/GetAllMeshes ...      OC ChunkManager.h:69-182
    ProjectionIntegrator               setters of OC ProjectionIntegrator.h:185-222
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libchisel_b200.so")

CHS_OK, CHS_ERR_INVALID, CHS_ERR_CUDA, CHS_ERR_CAPACITY, CHS_ERR_NOT_FOUND = range(5)
TRUNC_CONSTANT, TRUNC_QUADRATIC, TRUNC_INVERSE, TRUNC_PER_PIXEL = range(4)
MEM_HOST, MEM_DEVICE, MEM_HOST_ASYNC, MEM_DEVICE_ASYNC = 0, 1, 2, 3


class ChiselError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("chisel_b200 error %d: %s" % (code, msg))
        self.code = code


class chs_config(C.Structure):
    _fields_ = [("chunk_size", C.c_int), ("resolution", C.c_float), ("use_color", C.c_int), ("device", C.c_int),
                ("rank", C.c_int), ("world", C.c_int), ("initial_chunks", C.c_int64), ("stream", C.c_void_p)]


class chs_camera(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("width", C.c_int), ("height", C.c_int), ("near_plane", C.c_float), ("far_plane", C.c_float)]


class chs_integrator(C.Structure):
    _fields_ = [("trunc_kind", C.c_int), ("trunc_param", C.c_float), ("trunc_per_pixel", C.c_void_p),
                ("weight", C.c_float), ("carving_enabled", C.c_int), ("carving_dist", C.c_float)]


class chs_frame_stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("candidates", "new_candidates", "brick_units", "n_upd", "n_carve", "n_col", "n_new",
                                         "updated_chunks", "total_chunks", "dirty_chunks", "error_flags")]


class chs_frame(C.Structure):
    _fields_ = [("depth", C.c_void_p), ("depth_mm", C.c_void_p), ("color", C.c_void_p), ("trunc_per_pixel", C.c_void_p),
                ("pose", C.c_float * 12), ("color_pose", C.c_float * 12)]


class chs_mesh_counts(C.Structure):
    _fields_ = [("n_chunks", C.c_int64), ("n_vertices", C.c_int64), ("n_grids", C.c_int64), ("has_colors", C.c_int)]


class chs_timings(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("prepare_ms", "candidates_ms", "new_chunks_ms", "integrate_ms", "frame_ms",
                                         "mesh_count_ms", "mesh_emit_ms", "mesh_ms", "bricks_span_ms")]


EXPORTS = (
    "chs_last_error_string", "chs_abi_version", "chs_create", "chs_destroy", "chs_reset", "chs_synchronize",
    "chs_set_stream", "chs_set_profiling", "chs_integrate_depth", "chs_integrate_depth_color", "chs_integrate_batch",
    "chs_get_batch_stats", "chs_last_batch_ticket", "chs_wait_batch", "chs_get_frame_stats",
    "chs_get_timings", "chs_update_meshes", "chs_mesh_counts_last", "chs_download_meshes", "chs_num_chunks",
    "chs_chunk_ids", "chs_has_chunk", "chs_download_chunk", "chs_download_all", "chs_export_chunks", "chs_import_chunks", "chs_set_dirty", "chs_num_dirty", "chs_dirty_ids", "chs_frustum",
    "chs_candidate_ids", "chs_truncation", "chs_owner", "chs_selftest_arithmetic", "chs_host_alloc", "chs_host_free",
    "chs_device_alloc", "chs_device_free", "chs_upload", "chs_save_map", "chs_load_map", "chs_ingest_depth",
    "chs_comm_unique_id", "chs_comm_init", "chs_comm_attach", "chs_comm_destroy", "chs_integrate_batch_distributed",
    "chs_comm_sync_dirty", "chs_update_meshes_distributed", "chs_get_device_timeline",
)

_lib = None


def load_library(build_if_missing: bool = True):
    """dlopen libchisel_b200.so (building it with nvcc if absent). No fallback of any kind."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("CHS_LIB_PATH") or LIB_PATH       # CHS_LIB_PATH: A/B runs of differently tuned builds (tools/)
    if build_if_missing and path == LIB_PATH:
        from . import build as _build
        _build.build()
    if not os.path.exists(path):
        raise ChiselError(CHS_ERR_CUDA, "libchisel_b200.so is missing and could not be built")
    lib = C.CDLL(path)
    vp, i32, i64, f32p = C.c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_float)
    lib.chs_last_error_string.restype = C.c_char_p
    lib.chs_create.argtypes = [C.POINTER(chs_config), C.POINTER(vp)]
    for n in ("chs_destroy", "chs_reset", "chs_synchronize", "chs_update_meshes"):
        getattr(lib, n).argtypes = [vp]
    lib.chs_set_stream.argtypes = [vp, vp]
    lib.chs_set_profiling.argtypes = [vp, i32]
    lib.chs_get_device_timeline.argtypes = [vp, C.POINTER(C.c_longlong), i32, C.POINTER(i32)]
    lib.chs_integrate_depth.argtypes = [vp, C.POINTER(chs_integrator), vp, i32, vp, C.POINTER(chs_camera)]
    lib.chs_integrate_depth_color.argtypes = [vp, C.POINTER(chs_integrator), vp, i32, vp, C.POINTER(chs_camera),
                                              vp, i32, vp, C.POINTER(chs_camera)]
    lib.chs_integrate_batch.argtypes = [vp, C.POINTER(chs_integrator), i32, C.POINTER(chs_frame), i32, C.POINTER(chs_camera), i32,
                                        C.POINTER(chs_camera)]
    lib.chs_get_batch_stats.argtypes = [vp, C.POINTER(chs_frame_stats), i32, C.POINTER(i32)]
    lib.chs_last_batch_ticket.argtypes = [vp, C.POINTER(i64)]
    lib.chs_wait_batch.argtypes = [vp, i64, C.POINTER(chs_frame_stats), i32, C.POINTER(i32)]
    lib.chs_get_frame_stats.argtypes = [vp, C.POINTER(chs_frame_stats)]
    lib.chs_get_timings.argtypes = [vp, C.POINTER(chs_timings)]
    lib.chs_mesh_counts_last.argtypes = [vp, C.POINTER(chs_mesh_counts)]
    lib.chs_download_meshes.argtypes = [vp] + [vp] * 7
    lib.chs_num_chunks.argtypes = [vp, C.POINTER(i64)]
    lib.chs_chunk_ids.argtypes = [vp, vp, i64]
    lib.chs_has_chunk.argtypes = [vp, vp, C.POINTER(i32)]
    lib.chs_download_chunk.argtypes = [vp, vp, vp, vp, vp]
    lib.chs_download_all.argtypes = [vp, i64, vp, vp, vp, vp]
    lib.chs_export_chunks.argtypes = [vp, i64, vp, vp, vp, vp, vp]
    lib.chs_import_chunks.argtypes = [vp, i64, vp, vp, vp, vp]
    lib.chs_set_dirty.argtypes = [vp, i64, vp]
    lib.chs_num_dirty.argtypes = [vp, C.POINTER(i64)]
    lib.chs_dirty_ids.argtypes = [vp, vp, i64]
    lib.chs_ingest_depth.argtypes = [vp, vp, i32, i32, vp, i32, i32, C.c_float, C.c_float, i32]
    lib.chs_save_map.argtypes = [vp, C.c_char_p]
    lib.chs_load_map.argtypes = [vp, C.c_char_p]
    lib.chs_device_alloc.restype = vp
    lib.chs_device_alloc.argtypes = [vp, C.c_size_t]
    lib.chs_device_free.restype = None
    lib.chs_device_free.argtypes = [vp, vp]
    lib.chs_upload.argtypes = [vp, vp, vp, C.c_size_t]
    lib.chs_comm_unique_id.argtypes = [vp]
    lib.chs_comm_init.argtypes = [vp, vp]
    lib.chs_comm_attach.argtypes = [vp, vp]
    lib.chs_comm_destroy.argtypes = [vp]
    lib.chs_integrate_batch_distributed.argtypes = [vp, C.POINTER(chs_integrator), i32, C.POINTER(chs_frame), i32, C.POINTER(chs_camera), i32, C.POINTER(chs_camera)]
    lib.chs_comm_sync_dirty.argtypes = [vp]
    lib.chs_update_meshes_distributed.argtypes = [vp, i32]
    lib.chs_frustum.argtypes = [vp, C.POINTER(chs_camera), vp, vp, vp]
    lib.chs_candidate_ids.argtypes = [i32, C.c_float, vp, C.POINTER(chs_camera), vp, i64, C.POINTER(i64)]
    lib.chs_selftest_arithmetic.argtypes = [i64, C.POINTER(i64 * 4)]
    lib.chs_host_alloc.restype = C.c_void_p
    lib.chs_host_alloc.argtypes = [C.c_size_t]
    lib.chs_host_free.argtypes = [vp]
    lib.chs_truncation.restype = C.c_float
    lib.chs_truncation.argtypes = [i32, C.c_float, C.c_float]
    lib.chs_owner.restype = C.c_uint32
    lib.chs_owner.argtypes = [i32, i32, i32]
    _lib = lib
    return lib


def _check(rc):
    if rc != CHS_OK:
        raise ChiselError(rc, load_library().chs_last_error_string().decode())


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_camera(cam) -> chs_camera:
    """cam: [fx, fy, cx, cy, W, H, near, far] (scenes.Camera.as_array()) or a chs_camera."""
    if isinstance(cam, chs_camera):
        return cam
    c = np.asarray(cam, dtype=np.float32)
    return chs_camera(float(c[0]), float(c[1]), float(c[2]), float(c[3]), int(c[4]), int(c[5]), float(c[6]), float(c[7]))


def _pose(p) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(p, dtype=np.float32).reshape(12))


@dataclasses.dataclass
class ProjectionIntegrator:
    """Integrator state (OC ProjectionIntegrator.h:224-229) with the reference's setter names."""
    trunc_kind: int = TRUNC_CONSTANT
    trunc_param: float = 0.2
    weight: float = 1.0
    carving_enabled: bool = True
    carving_dist: float = 0.05
    trunc_per_pixel: np.ndarray | None = None

    def set_truncator(self, kind, param):
        self.trunc_kind, self.trunc_param = kind, param

    def set_weighter(self, weight):
        self.weight = weight

    def set_carving_dist(self, d):
        self.carving_dist = d

    def set_carving_enabled(self, e):
        self.carving_enabled = e

    def as_struct(self, device_ptr=None) -> chs_integrator:
        tp = None
        if self.trunc_kind == TRUNC_PER_PIXEL:
            tp = device_ptr if device_ptr is not None else _ptr(np.ascontiguousarray(self.trunc_per_pixel, np.float32))
        return chs_integrator(self.trunc_kind, self.trunc_param, tp, self.weight, int(self.carving_enabled), self.carving_dist)


class ChunkManager:
    """Lazily synchronised host view of the device map (OC ChunkManager.h:58-223)."""

    def __init__(self, owner: "Chisel"):
        self._o = owner
        self.all_meshes = {}          # ChunkManager::allMeshes: {(x,y,z): dict(vertices, normals, colors, grids)}

    def get_chunk_size(self):
        return (self._o.chunk,) * 3

    def get_resolution(self):
        return self._o.resolution

    def num_chunks(self) -> int:
        n = C.c_int64()
        _check(self._o._lib.chs_num_chunks(self._o._h, C.byref(n)))
        return n.value

    def chunk_ids(self) -> np.ndarray:
        n = self.num_chunks()
        out = np.zeros((n, 3), np.int32)
        if n:
            _check(self._o._lib.chs_chunk_ids(self._o._h, _ptr(out), n))
        return out

    def has_chunk(self, cid) -> bool:
        found = C.c_int()
        cid = np.ascontiguousarray(cid, np.int32)
        _check(self._o._lib.chs_has_chunk(self._o._h, _ptr(cid), C.byref(found)))
        return bool(found.value)

    def get_chunk(self, cid, want_color=True):
        V = self._o.chunk ** 3
        sdf = np.zeros(V, np.float32)
        w = np.zeros(V, np.float32)
        rgbw = np.zeros((V, 4), np.uint8)
        cid = np.ascontiguousarray(cid, np.int32)
        _check(self._o._lib.chs_download_chunk(self._o._h, _ptr(cid), _ptr(sdf), _ptr(w),
                                               _ptr(rgbw) if (want_color and self._o.use_color) else None))
        return sdf, w, rgbw

    def get_chunks(self):
        """Whole map: ids [n,3] (pool order), sdf [n,V], weight [n,V], rgbw [n,V,4]."""
        n = self.num_chunks()
        V = self._o.chunk ** 3
        ids = np.zeros((n, 3), np.int32)
        sdf = np.zeros((n, V), np.float32)
        w = np.zeros((n, V), np.float32)
        rgbw = np.zeros((n, V, 4), np.uint8)
        if n:
            _check(self._o._lib.chs_download_all(self._o._h, n, _ptr(ids), _ptr(sdf), _ptr(w),
                                                 _ptr(rgbw) if self._o.use_color else None))
        return ids, sdf, w, rgbw

    def get_all_meshes(self):
        return self.all_meshes


class Chisel:
    """Host-side mirror of chisel::Chisel over the C ABI."""
    kind = "cuda"

    def __init__(self, chunk: int, resolution: float, use_color: bool, device: int = -1, rank: int = 0, world: int = 1,
                 stream: int | None = None, initial_chunks: int = 0):
        self._lib = load_library()
        self.chunk, self.resolution, self.use_color = chunk, float(np.float32(resolution)), bool(use_color)
        self.rank, self.world = int(rank), max(int(world), 1)
        cfg = chs_config(chunk, resolution, int(use_color), device, rank, world, initial_chunks, stream)
        h = C.c_void_p()
        _check(self._lib.chs_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self.chunk_manager = ChunkManager(self)
        self._update_calls = 0        # per-instance counterpart of the static counter in Chisel::UpdateMeshes (Q4)
        self._keep = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.chs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- integration ----
    def integrate_depth_scan(self, integrator: ProjectionIntegrator, depth, pose, cam, device_ptrs=None):
        """depth: float32 [H,W] host array, or with device_ptrs=(depth_ptr, trunc_ptr|None) device memory."""
        cam = make_camera(cam)
        p = _pose(pose)
        if device_ptrs is None:
            d = np.ascontiguousarray(depth, np.float32)
            integ = integrator.as_struct()
            _check(self._lib.chs_integrate_depth(self._h, C.byref(integ), _ptr(d), MEM_HOST, _ptr(p), C.byref(cam)))
        else:
            integ = integrator.as_struct(device_ptrs[1])
            _check(self._lib.chs_integrate_depth(self._h, C.byref(integ), device_ptrs[0], MEM_DEVICE, _ptr(p), C.byref(cam)))

    def integrate_depth_scan_color(self, integrator: ProjectionIntegrator, depth, pose, cam, color, color_pose=None,
                                   color_cam=None, device_ptrs=None, channels=None):
        cam = make_camera(cam)
        ccam = cam if color_cam is None else make_camera(color_cam)
        p = _pose(pose)
        cp = p if color_pose is None else _pose(color_pose)
        if device_ptrs is None:
            d = np.ascontiguousarray(depth, np.float32)
            c = np.ascontiguousarray(color, np.uint8)
            ch = c.shape[2] if c.ndim == 3 else 1
            integ = integrator.as_struct()
            _check(self._lib.chs_integrate_depth_color(self._h, C.byref(integ), _ptr(d), MEM_HOST, _ptr(p), C.byref(cam),
                                                       _ptr(c), ch, _ptr(cp), C.byref(ccam)))
        else:
            integ = integrator.as_struct(device_ptrs[2] if len(device_ptrs) > 2 else None)
            _check(self._lib.chs_integrate_depth_color(self._h, C.byref(integ), device_ptrs[0], MEM_DEVICE, _ptr(p),
                                                       C.byref(cam), device_ptrs[1], channels, _ptr(cp), C.byref(ccam)))

    def prepare_batch(self, integrator: ProjectionIntegrator, depths, poses, cam, colors=None, color_poses=None, color_cam=None,
                      device_ptrs=None, channels=None, truncs=None, host_async=False, device_async=False):
        """Marshal the arguments of one chs_integrate_batch call once (a streaming caller that owns a ring of frame buffers
        does this at start-up); integrate_prepared() then only makes the call. See integrate_batch for the arguments."""
        cam = make_camera(cam)
        n = len(poses)
        use_color = (colors is not None) or (device_ptrs is not None and len(device_ptrs) and device_ptrs[0][1] is not None)
        ccam = (cam if color_cam is None else make_camera(color_cam)) if use_color else None
        arr = (chs_frame * max(n, 1))()
        keep = []
        for i in range(n):
            p = _pose(poses[i])
            cp = p if (color_poses is None or color_poses[i] is None) else _pose(color_poses[i])
            C.memmove(arr[i].pose, p.ctypes.data, 48)
            C.memmove(arr[i].color_pose, cp.ctypes.data, 48)
            if device_ptrs is None:
                if np.asarray(depths[i]).dtype == np.uint16:           # millimetres (ROS 16UC1): converted on the device
                    d = np.ascontiguousarray(depths[i], np.uint16)
                    arr[i].depth_mm = d.ctypes.data
                else:
                    d = np.ascontiguousarray(depths[i], np.float32)
                    arr[i].depth = d.ctypes.data
                keep.append(d)
                if use_color:
                    c = np.ascontiguousarray(colors[i], np.uint8)
                    keep.append(c)
                    arr[i].color = c.ctypes.data
                    channels = c.shape[2] if c.ndim == 3 else 1
                if truncs is not None:
                    t = np.ascontiguousarray(truncs[i], np.float32)
                    keep.append(t)
                    arr[i].trunc_per_pixel = t.ctypes.data
            else:
                arr[i].depth = device_ptrs[i][0]
                arr[i].color = device_ptrs[i][1]
                arr[i].trunc_per_pixel = device_ptrs[i][2] if len(device_ptrs[i]) > 2 else None
        integ = integrator.as_struct(device_ptr=0) if integrator.trunc_kind == TRUNC_PER_PIXEL else integrator.as_struct()
        mem = (MEM_DEVICE_ASYNC if device_async else MEM_DEVICE) if device_ptrs is not None else (MEM_HOST_ASYNC if host_async else MEM_HOST)
        return (integ, n, arr, mem, cam, int(channels or 0), ccam, keep)

    def integrate_prepared(self, prepared):
        integ, n, arr, mem, cam, channels, ccam, _keep = prepared
        _check(self._lib.chs_integrate_batch(self._h, C.byref(integ), n, arr, mem, C.byref(cam), channels,
                                             C.byref(ccam) if ccam is not None else None))

    def integrate_batch(self, integrator: ProjectionIntegrator, depths, poses, cam, colors=None, color_poses=None, color_cam=None,
                        device_ptrs=None, channels=None, truncs=None, host_async=False):
        """n consecutive frames in one call (chs_integrate_batch): the same result as n calls of integrate_depth_scan[_color]
        in order. depths/colors: lists of host arrays (float32 metres, or uint16 millimetres), or with
        device_ptrs=[(depth_ptr, color_ptr|None, trunc_ptr|None), ...] device memory. colors=None (and no colour device
        pointers): the depth path. host_async: the host buffers stay untouched until the call's ticket has been waited for."""
        prepared = self.prepare_batch(integrator, depths, poses, cam, colors, color_poses, color_cam, device_ptrs, channels, truncs, host_async)
        if host_async:
            self._keep = (self._keep or [])[-64:] + [prepared]      # the buffers must outlive the call
        self.integrate_prepared(prepared)

    # ---- multi-GPU (one process per GPU; every call below is collective over the map's world) ----
    @staticmethod
    def comm_unique_id() -> bytes:
        """ncclGetUniqueId through the library: call on ONE rank and hand the 128 bytes to every rank (any side channel)."""
        buf = (C.c_uint8 * 128)()
        _check(load_library().chs_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, unique_id: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _check(self._lib.chs_comm_init(self._h, buf))

    def comm_init_torch(self, group=None):
        """Convenience for harnesses that already run torch.distributed: rank 0 makes the NCCL unique id, the process group
        (any backend) carries its 128 bytes to the other ranks, every rank joins the library's own communicator."""
        import torch.distributed as dist
        box = [self.comm_unique_id() if dist.get_rank(group) == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        self.comm_init(box[0])

    def comm_destroy(self):
        _check(self._lib.chs_comm_destroy(self._h))

    def integrate_batch_distributed(self, integrator: ProjectionIntegrator, depths, poses, cam, colors=None, device_ptrs=None, channels=None,
                                    host_async=False):
        """One step of len(poses) frames spread over the ranks (chs_integrate_batch_distributed): poses of ALL frames; depths /
        colors / device_ptrs entries only for the frames this rank ingests (None elsewhere)."""
        self.integrate_prepared_distributed(self.prepare_batch_distributed(integrator, depths, poses, cam, colors, device_ptrs, channels, host_async))

    def integrate_prepared_distributed(self, prepared):
        integ, n, arr, mem, cam_s, channels, use_color, _keep = prepared
        _check(self._lib.chs_integrate_batch_distributed(self._h, C.byref(integ), n, arr, mem, C.byref(cam_s), channels,
                                                         C.byref(cam_s) if use_color else None))

    def prepare_batch_distributed(self, integrator: ProjectionIntegrator, depths, poses, cam, colors=None, device_ptrs=None, channels=None,
                                  host_async=False, device_async=False):
        """Marshal one chs_integrate_batch_distributed call (see integrate_batch_distributed)."""
        n = len(poses)
        per = n // self.world
        lo = self.rank * per
        cam_s = make_camera(cam)
        use_color = (colors is not None) or (device_ptrs is not None and device_ptrs[lo][1] is not None)
        arr = (chs_frame * max(n, 1))()
        keep = []
        for i in range(n):
            p = _pose(poses[i])
            C.memmove(arr[i].pose, p.ctypes.data, 48)
            C.memmove(arr[i].color_pose, p.ctypes.data, 48)
            if not (lo <= i < lo + per):
                continue
            if device_ptrs is None:
                if np.asarray(depths[i]).dtype == np.uint16:
                    d = np.ascontiguousarray(depths[i], np.uint16)
                    arr[i].depth_mm = d.ctypes.data
                else:
                    d = np.ascontiguousarray(depths[i], np.float32)
                    arr[i].depth = d.ctypes.data
                keep.append(d)
                if use_color:
                    c = np.ascontiguousarray(colors[i], np.uint8)
                    keep.append(c)
                    arr[i].color = c.ctypes.data
                    channels = c.shape[2] if c.ndim == 3 else 1
            else:
                arr[i].depth = device_ptrs[i][0]
                arr[i].color = device_ptrs[i][1]
        integ = integrator.as_struct()
        mem = (MEM_DEVICE_ASYNC if device_async else MEM_DEVICE) if device_ptrs is not None else (MEM_HOST_ASYNC if host_async else MEM_HOST)
        if host_async:
            self._keep = (self._keep or [])[-64:] + [keep]
        return (integ, n, arr, mem, cam_s, int(channels or 0), use_color, keep)

    def sync_dirty(self):
        _check(self._lib.chs_comm_sync_dirty(self._h))

    def recompute_meshes_distributed(self, root: int = 0):
        """Distributed Chisel::UpdateMeshes (chs_update_meshes_distributed). On the root the gathered meshes of ALL ranks are merged
        into the host MeshMap with the reference's publication rule; returns the re-meshed chunks this rank received."""
        _check(self._lib.chs_update_meshes_distributed(self._h, root))
        got = self.download_last_meshes()
        if self.rank == root:
            for cid, mesh in got.items():
                if cid in self.chunk_manager.all_meshes or len(mesh["grids"]) > 0:
                    self.chunk_manager.all_meshes[cid] = mesh
        return got

    def last_batch_ticket(self) -> int:
        t = C.c_int64()
        _check(self._lib.chs_last_batch_ticket(self._h, C.byref(t)))
        return t.value

    def wait_batch(self, ticket: int) -> list:
        """Counters of ONE integrate_batch call; calls issued after it stay in flight."""
        arr = (chs_frame_stats * 16)()
        n = C.c_int()
        _check(self._lib.chs_wait_batch(self._h, ticket, arr, 16, C.byref(n)))
        if n.value > 16:
            arr = (chs_frame_stats * n.value)()
            _check(self._lib.chs_wait_batch(self._h, ticket, arr, n.value, C.byref(n)))
        return [{k: getattr(arr[i], k) for k, _ in chs_frame_stats._fields_} for i in range(n.value)]

    def batch_stats(self) -> list:
        n = C.c_int()
        _check(self._lib.chs_get_batch_stats(self._h, None, 0, C.byref(n)))
        arr = (chs_frame_stats * max(n.value, 1))()
        _check(self._lib.chs_get_batch_stats(self._h, arr, n.value, C.byref(n)))
        return [{k: getattr(arr[i], k) for k, _ in chs_frame_stats._fields_} for i in range(n.value)]

    def frame_stats(self) -> dict:
        s = chs_frame_stats()
        _check(self._lib.chs_get_frame_stats(self._h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in chs_frame_stats._fields_}

    def timings(self) -> dict:
        t = chs_timings()
        _check(self._lib.chs_get_timings(self._h, C.byref(t)))
        return {n: getattr(t, n) for n, _ in chs_timings._fields_}

    def set_profiling(self, on):
        """True / 1: event timings between the kernels (serialises them); 2: device timeline only (kernels stamp %globaltimer)."""
        _check(self._lib.chs_set_profiling(self._h, int(on)))

    TIMELINE = ("push_start", "push_end", "wait_end", "hiz_start", "hiz_end", "cand_start", "cand_end", "bricks_start", "bricks_end")

    def device_timeline(self, max_batches: int = 256) -> list:
        """Absolute device times (ns) of the most recent fused batches, oldest first; 0 = stamp not recorded."""
        n = len(self.TIMELINE)
        out = (C.c_longlong * (n * max_batches))()
        got = C.c_int32(0)
        _check(self._lib.chs_get_device_timeline(self._h, out, max_batches, C.byref(got)))
        return [dict(zip(self.TIMELINE, [int(out[i * n + k]) for k in range(n)])) for i in range(got.value)]

    def set_stream(self, stream: int | None):
        _check(self._lib.chs_set_stream(self._h, stream))

    def synchronize(self):
        _check(self._lib.chs_synchronize(self._h))

    def reset(self):
        _check(self._lib.chs_reset(self._h))
        self.chunk_manager.all_meshes.clear()

    # ---- meshing ----
    def get_meshes_to_update(self) -> np.ndarray:
        n = C.c_int64()
        _check(self._lib.chs_num_dirty(self._h, C.byref(n)))
        out = np.zeros((n.value, 3), np.int32)
        if n.value:
            _check(self._lib.chs_dirty_ids(self._h, _ptr(out), n.value))
        return out

    def update_meshes(self, force: bool = False) -> bool:
        """Chisel::UpdateMeshes: re-meshes on calls 1, 11, 21, ... (Chisel.cpp:50-59); force=True bypasses the gate.
        Returns True if a re-mesh ran."""
        gate = (self._update_calls % 10) == 0
        self._update_calls += 1
        if not (gate or force):
            return False
        self.recompute_meshes()
        return True

    def recompute_meshes(self):
        """ChunkManager::RecomputeMeshes(meshesToUpdate) + clear, then merge into the host MeshMap with the
        reference's publication rule (ChunkManager.cpp:101-127, quirk Q10)."""
        _check(self._lib.chs_update_meshes(self._h))
        for cid, mesh in self.download_last_meshes().items():
            if cid in self.chunk_manager.all_meshes or len(mesh["grids"]) > 0:
                self.chunk_manager.all_meshes[cid] = mesh

    def last_mesh_counts(self) -> dict:
        mc = chs_mesh_counts()
        _check(self._lib.chs_mesh_counts_last(self._h, C.byref(mc)))
        return dict(n_chunks=mc.n_chunks, n_vertices=mc.n_vertices, n_grids=mc.n_grids, has_colors=mc.has_colors)

    def download_last_meshes(self) -> dict:
        mc = self.last_mesh_counts()
        n, nv, ng = mc["n_chunks"], mc["n_vertices"], mc["n_grids"]
        ids = np.zeros((n, 3), np.int32)
        voff = np.zeros(n + 1, np.int64)
        goff = np.zeros(n + 1, np.int64)
        v = np.zeros((nv, 3), np.float32)
        nr = np.zeros((nv, 3), np.float32)
        col = np.zeros((nv, 3), np.float32) if mc["has_colors"] else None
        g = np.zeros((ng, 3), np.float32)
        _check(self._lib.chs_download_meshes(self._h, _ptr(ids), _ptr(voff), _ptr(goff), _ptr(v), _ptr(nr), _ptr(col), _ptr(g)))
        out = {}
        for i in range(n):
            a, b = voff[i], voff[i + 1]
            out[tuple(int(x) for x in ids[i])] = dict(
                vertices=v[a:b].copy(), normals=nr[a:b].copy(),
                colors=(col[a:b].copy() if col is not None else np.zeros((0, 3), np.float32)),
                grids=g[goff[i]:goff[i + 1]].copy())
        return out

    # ---- chunk export / import / explicit dirty set (ghost chunks for sharded meshing, checkpoints) ----
    def export_chunks(self, ids):
        """(found [n] bool, sdf [n,V], weight [n,V], rgbw [n,V,4]) for the listed IDs; rows of absent chunks are zero."""
        ids = np.ascontiguousarray(np.asarray(ids, np.int32).reshape(-1, 3))
        n, V = len(ids), self.chunk ** 3
        found = np.zeros(n, np.uint8)
        sdf = np.zeros((n, V), np.float32)
        w = np.zeros((n, V), np.float32)
        rgbw = np.zeros((n, V, 4), np.uint8)
        if n:
            _check(self._lib.chs_export_chunks(self._h, n, _ptr(ids), _ptr(found), _ptr(sdf), _ptr(w), _ptr(rgbw) if self.use_color else None))
        return found.astype(bool), sdf, w, rgbw

    def import_chunks(self, ids, sdf, weight, rgbw=None):
        ids = np.ascontiguousarray(np.asarray(ids, np.int32).reshape(-1, 3))
        if len(ids):
            sdf = np.ascontiguousarray(sdf, np.float32)
            weight = np.ascontiguousarray(weight, np.float32)
            c = np.ascontiguousarray(rgbw, np.uint8) if (rgbw is not None and self.use_color) else None
            _check(self._lib.chs_import_chunks(self._h, len(ids), _ptr(ids), _ptr(sdf), _ptr(weight), _ptr(c)))

    def ingest_depth(self, depth: np.ndarray, width: int, height: int, valid_min: float = 0.1, valid_max: float = 20.0) -> np.ndarray:
        """The collaborative server's frame ingestion (chs_ingest_depth): bilinear resize to width x height + NaN outside the valid range."""
        src = np.ascontiguousarray(depth, np.float32)
        out = np.empty((height, width), np.float32)
        _check(self._lib.chs_ingest_depth(self._h, _ptr(src), src.shape[1], src.shape[0], _ptr(out), width, height, valid_min, valid_max, MEM_HOST))
        return out

    def save_map(self, path: str):
        """Checkpoint of the whole map (voxels + dirty set) in one file (chs_save_map)."""
        _check(self._lib.chs_save_map(self._h, path.encode()))

    def load_map(self, path: str):
        """Replace the map's contents by a checkpoint; the host MeshMap mirror is emptied (meshes are not part of the checkpoint)."""
        _check(self._lib.chs_load_map(self._h, path.encode()))
        self.chunk_manager.all_meshes.clear()

    def set_dirty(self, ids):
        ids = np.ascontiguousarray(np.asarray(ids, np.int32).reshape(-1, 3))
        _check(self._lib.chs_set_dirty(self._h, len(ids), _ptr(ids) if len(ids) else None))

    # ---- helpers mirroring the oracle front-ends (tests swap implementations) ----
    def state(self):
        ids, sdf, w, rgbw = self.chunk_manager.get_chunks()
        order = np.lexsort((ids[:, 2], ids[:, 1], ids[:, 0])) if len(ids) else np.zeros(0, np.int64)
        return ids[order], sdf[order], w[order], rgbw[order]

    def dirty_ids(self) -> np.ndarray:
        ids = self.get_meshes_to_update()
        order = np.lexsort((ids[:, 2], ids[:, 1], ids[:, 0])) if len(ids) else np.zeros(0, np.int64)
        return ids[order]


def frustum(pose, cam):
    lib = load_library()
    cam = make_camera(cam)
    p = _pose(pose)
    corners = np.zeros((8, 3), np.float32)
    lines = np.zeros((24, 3), np.float32)
    planes = np.zeros((6, 4), np.float32)
    _check(lib.chs_frustum(_ptr(p), C.byref(cam), _ptr(corners), _ptr(lines), _ptr(planes)))
    return corners, lines, planes


def candidate_ids(chunk, resolution, pose, cam) -> np.ndarray:
    lib = load_library()
    cam = make_camera(cam)
    p = _pose(pose)
    n = C.c_int64()
    _check(lib.chs_candidate_ids(chunk, resolution, _ptr(p), C.byref(cam), None, 0, C.byref(n)))
    out = np.zeros((n.value, 3), np.int32)
    _check(lib.chs_candidate_ids(chunk, resolution, _ptr(p), C.byref(cam), _ptr(out), n.value, C.byref(n)))
    return out


def selftest_arithmetic(div_pairs: int = 1 << 30) -> dict:
    out = (C.c_int64 * 4)()
    _check(load_library().chs_selftest_arithmetic(div_pairs, C.byref(out)))
    return dict(rcp_mismatches=out[0], rcp_tested=out[1], div_mismatches=out[2], div_tested=out[3])


def truncation(kind, param, depth) -> float:
    return load_library().chs_truncation(kind, param, depth)


def owner(x, y, z) -> int:
    return load_library().chs_owner(x, y, z)

"""Build and drive tests/cpp/chisel_client.cpp against the facade (and, in the build container, the reference)."""
from __future__ import annotations

import glob
import os
import struct
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "cpp", "_build")
SRC = os.path.join(ROOT, "tests", "cpp", "chisel_client.cpp")
REF = "/root/reference/OpenChisel/open_chisel"
HEADER_FMT = "<9i10f"


def build_facade_client() -> str:
    from cvids_b200 import build as cuda_build
    cuda_build.build()
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, "chisel_client_b200")
    deps = [SRC] + glob.glob(os.path.join(ROOT, "cvids_b200", "include", "open_chisel", "**", "*.h"), recursive=True) + \
        [os.path.join(ROOT, "include", "chisel_b200.h")]
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    lib = os.path.join(ROOT, "cvids_b200")
    subprocess.check_call(["g++", "-std=c++11", "-O2", "-pthread", "-Wall", "-Wno-unused-parameter", "-I", os.path.join(ROOT, "oracle", "eigen_shim"),
                           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "cvids_b200", "include"), SRC, "-o", out,
                           "-L", lib, "-lchisel_b200", "-Wl,-rpath," + lib])
    return out


def build_reference_client() -> str:
    """The SAME client source against the reference's own headers and sources (build container only)."""
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, "chisel_client_ref")
    if os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(SRC):
        return out
    srcs = sorted(glob.glob(os.path.join(REF, "src", "*.cpp")) + glob.glob(os.path.join(REF, "src", "*", "*.cpp")))
    subprocess.check_call(["g++", "-std=c++11", "-O2", "-w", "-ffp-contract=off", "-pthread", "-include", os.path.join(ROOT, "oracle", "quiet_stdio.h"),
                           "-I", os.path.join(ROOT, "oracle", "eigen_shim"), "-I", os.path.join(REF, "include"), SRC] + srcs + ["-o", out])
    return out


def write_stream(path, setup, cam, frames, channels, update_every=1):
    """frames: iterable of (depth, color|None, pose)."""
    frames = list(frames)
    with open(path, "wb") as f:
        f.write(struct.pack(HEADER_FMT, cam.width, cam.height, channels, len(frames), setup.chunk, int(setup.color), setup.trunc_kind,
                            int(setup.carve), update_every, setup.resolution, setup.trunc, setup.weight, setup.carve_dist, cam.near, cam.far,
                            cam.fx, cam.fy, cam.cx, cam.cy))
        for depth, col, pose in frames:
            f.write(np.ascontiguousarray(pose, np.float32).tobytes())
            f.write(np.ascontiguousarray(depth, np.float32).tobytes())
            if channels:
                f.write(np.ascontiguousarray(col, np.uint8).tobytes())
    return frames


def read_dump(path):
    buf = open(path, "rb").read()
    off = 0

    def take(fmt):
        nonlocal off
        v = struct.unpack_from(fmt, buf, off)
        off += struct.calcsize(fmt)
        return v

    def arr(dtype, n):
        nonlocal off
        a = np.frombuffer(buf, dtype=dtype, count=n, offset=off).copy()
        off += a.nbytes
        return a

    def vecs():
        (n,) = take("<q")
        return arr(np.float32, 3 * n).reshape(n, 3)

    n, remeshes = take("<2q")
    ids, centers, sdf, w, rgbw = [], [], [], [], []
    for _ in range(n):
        ids.append(take("<3i"))
        centers.append(take("<3f"))
        V, hc = take("<2i")
        dv = arr(np.float32, 2 * V).reshape(V, 2)
        sdf.append(dv[:, 0].copy())
        w.append(dv[:, 1].copy())
        rgbw.append(arr(np.uint8, 4 * V).reshape(V, 4) if hc else np.zeros((V, 4), np.uint8))
    (nd,) = take("<q")
    dirty = arr(np.int32, 3 * nd).reshape(nd, 3)
    (nm,) = take("<q")
    meshes = {}
    for _ in range(nm):
        mid = take("<3i")
        v, nr, col, g = vecs(), vecs(), vecs(), vecs()
        take("<2i")
        meshes[tuple(mid)] = dict(vertices=v, normals=nr, colors=col, grids=g)
    lines = arr(np.float32, 72).reshape(24, 3)
    V = len(sdf[0]) if n else 0
    state = (np.asarray(ids, np.int32).reshape(n, 3), np.asarray(sdf, np.float32).reshape(n, V), np.asarray(w, np.float32).reshape(n, V),
             np.asarray(rgbw, np.uint8).reshape(n, V, 4))
    return dict(state=state, centers=np.asarray(centers, np.float32), dirty=dirty, meshes=meshes, lines=lines, remeshes=remeshes)

"""The drop-in C++ facade (cvids_b200/include/open_chisel) driven by a chisel_ros-style client.

CPU suite: the client compiles and links against the facade; in the build container the SAME source also compiles against
the reference's own headers and sources, and that binary's dump equals the C oracle's (so the client and the oracle agree
on what the reference does through its public API).
GPU suite: the facade binary runs on the device and its dump equals the oracle's on the same stream, bit for bit."""
import os
import subprocess

import numpy as np
import pytest

from cvids_b200 import scenes
from tests import common, facade_util
from tests.common import Setup

HAVE_REF = os.path.isdir(facade_util.REF)
CASES = {
    # depth-only path: serial in the reference, hence deterministic
    "depth": dict(setup=Setup(16, 0.05, False), color=False, n=12),
    # colour path with the truncator chisel_ros instantiates; the reference's threaded path is run pinned to one core (quirk Q2)
    "color_inverse": dict(setup=Setup(8, 0.1, True, trunc_kind=common.TRUNC_INVERSE, trunc_param=2.0, carve_dist=0.0), color=True, n=11),
}


def _stream(tmp_path, case):
    c = CASES[case]
    path = str(tmp_path / (case + ".stream"))
    frames = facade_util.write_stream(path, c["setup"], common.SMALL_CAM,
                                      common.orbit_stream(common.SMALL_CAM, c["n"], total=30, color=c["color"], seed=3),
                                      3 if c["color"] else 0)
    return path, frames


def _oracle_expectation(case, frames):
    """What the reference does for this client: integrate every frame, UpdateMeshes after each; the gate re-meshes on
    calls 1, 11, 21, ... (Chisel.cpp:50-59)."""
    c = CASES[case]
    drv = common.Driver(c["setup"], "oracle")
    for i, (depth, col, pose) in enumerate(frames):
        drv.integrate(depth, pose, common.SMALL_CAM.as_array(), col)
        if i % 10 == 0:
            drv.remesh()
    return drv


def _compare(dump, drv, with_colors):
    common.assert_state_equal(dump["state"], drv.state())
    assert np.array_equal(dump["dirty"], drv.dirty())
    common.assert_meshes_equal(dump["meshes"], drv.meshes(), with_colors=with_colors)
    ids = dump["state"][0]
    res, cs = drv.setup.resolution, drv.setup.chunk
    assert np.allclose(dump["centers"], (ids.astype(np.float32) + 0.5) * np.float32(cs * res), atol=1e-5)


def test_facade_client_compiles_and_links():
    exe = facade_util.build_facade_client()
    assert os.path.exists(exe)
    out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libchisel_b200.so" in out


@pytest.mark.skipif(not HAVE_REF, reason="reference sources not present")
@pytest.mark.parametrize("case", sorted(CASES))
def test_same_client_against_the_reference(tmp_path, case):
    exe = facade_util.build_reference_client()
    stream, frames = _stream(tmp_path, case)
    dump = str(tmp_path / "ref.dump")
    subprocess.run(["taskset", "-c", "0", exe, stream, dump], check=True, stdout=subprocess.DEVNULL)
    d = facade_util.read_dump(dump)
    _compare(d, _oracle_expectation(case, frames), CASES[case]["color"])
    assert d["remeshes"] == 2


@pytest.mark.gpu
@pytest.mark.parametrize("batch", [1, 10])
@pytest.mark.parametrize("case", sorted(CASES))
def test_facade_client_on_device(tmp_path, case, batch):
    """batch = 10: the facade queues frames (CHISEL_B200_BATCH) and sends them through the fused multi-frame path; the client
    is unchanged and must see the same map, dirty set and meshes."""
    exe = facade_util.build_facade_client()
    stream, frames = _stream(tmp_path, case)
    dump = str(tmp_path / "b200.dump")
    subprocess.run([exe, stream, dump], check=True, env=dict(os.environ, CHISEL_B200_BATCH=str(batch)))
    d = facade_util.read_dump(dump)
    drv = _oracle_expectation(case, frames)
    _compare(d, drv, CASES[case]["color"])
    assert d["remeshes"] == 2
    from oracle.pyoracle import OracleChisel
    lines = OracleChisel(16, 0.05, False).frustum(frames[-1][2], common.SMALL_CAM.as_array())[1]
    assert np.array_equal(d["lines"].view(np.uint32), lines.view(np.uint32))
    ply = open(dump + ".ply").read().split("\n")
    nverts = sum(len(m["vertices"]) for m in d["meshes"].values())
    assert ply[0] == "ply" and ("element vertex %d" % nverts) in ply[2]
    # Chisel::SaveAllMeshesToPLY against the bytes the REFERENCE wrote for the same client and stream (tests/golden/ply_golden.json,
    # made by tests/golden/make_ply_golden.py): header, every vertex line (position, colour) and every face line, byte for byte; only
    # the order of the per-chunk blocks (the reference iterates an unordered_map) is normalised
    import hashlib
    import json
    from tests.golden.make_ply_golden import canonical
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ply_golden.json")))[case]
    raw = open(dump + ".ply", "rb").read()
    header, nv, nf, canon = canonical(raw.decode())
    assert header == gold["header"] and nv == gold["vertices"] and nf == gold["faces"] and len(raw) == gold["bytes"]
    assert hashlib.sha256(canon).hexdigest() == gold["sha256_canonical"], "PLY export differs from the reference's"


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(CASES))
def test_facade_relaxed_reads_same_final_state(tmp_path, case):
    """CHISEL_B200_RELAXED_READS: the per-frame reads of chisel_ros (dirty set, HasChunk, GetChunk) are served from the mirrors of
    the last flushed state, so a batching facade stays on the fused path. Everything observable after a flush -- the map, the
    dirty set, every mesh -- is unchanged; only the client's own per-frame `remeshes` bookkeeping (it compares dirty-set sizes
    around UpdateMeshes) sees the lag."""
    exe = facade_util.build_facade_client()
    stream, frames = _stream(tmp_path, case)
    dump = str(tmp_path / "relaxed.dump")
    subprocess.run([exe, stream, dump], check=True, env=dict(os.environ, CHISEL_B200_BATCH="10", CHISEL_B200_RELAXED_READS="1"))
    d = facade_util.read_dump(dump)
    _compare(d, _oracle_expectation(case, frames), CASES[case]["color"])


@pytest.mark.gpu
def test_dropin_full_size(tmp_path):
    """The drop-in claim at full size (configs[1] shape: 752x480 depth + colour, 2 cm, 41 frames), through the reference's C++ API
    with chisel_ros's per-frame call sequence (integrate, PublishLatestChunkBoxes, frustum, UpdateMeshes): the facade's dump -- every
    voxel, the dirty set, every mesh array -- equals the oracle's, in all three modes."""
    import re
    cfg = scenes.CONFIG2
    setup = Setup(cfg.chunk, cfg.resolution, True)
    n = 41
    frames = [scenes.stream_frame(cfg, f) for f in range(n)]
    stream = str(tmp_path / "c2.stream")
    facade_util.write_stream(stream, setup, cfg.cam, frames, 3)
    exe = facade_util.build_facade_client()
    drv = common.Driver(setup, "oracle")
    for i, (depth, col, pose) in enumerate(frames):
        drv.integrate(depth, pose, cfg.cam.as_array(), col)
        if i % 10 == 0:
            drv.remesh()
    fps = {}
    for name, env in (("one_frame_per_call", {"CHISEL_B200_BATCH": "1"}), ("batch10_exact_reads", {"CHISEL_B200_BATCH": "10"}),
                      ("batch10_relaxed_reads", {"CHISEL_B200_BATCH": "10", "CHISEL_B200_RELAXED_READS": "1"})):
        dump = str(tmp_path / (name + ".dump"))
        best = 0.0
        for _ in range(2):
            p = subprocess.run([exe, stream, dump, "time"], capture_output=True, text=True, env=dict(os.environ, **env))
            assert p.returncode == 0, p.stderr[-2000:]
            best = max(best, float(re.search(r"fps ([0-9.]+)", p.stderr).group(1)))
        fps[name] = best
        d = facade_util.read_dump(dump)
        common.assert_state_equal(d["state"], drv.state(), name)
        assert np.array_equal(d["dirty"], drv.dirty()), name
        common.assert_meshes_equal(d["meshes"], drv.meshes())
    print("drop-in fps:", fps)
    # all three modes are bound by the caller's own per-frame host work (700-760 frames/s on the B200 boxes, within noise of each
    # other; the reference: 2 frames/s): a floor, not an ordering
    assert min(fps.values()) > 400.0, fps


def test_ply_writers_ascii_and_binary_agree(tmp_path):
    """Host-only: SaveMeshPLYASCII (the reference's layout, OC/src/io/PLY.cpp:29-88) and the binary_little_endian variant hold
    the same vertices, colours and faces."""
    exe = str(tmp_path / "ply_test")
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-I", os.path.join(facade_util.ROOT, "oracle", "eigen_shim"), "-I", os.path.join(facade_util.ROOT, "include"),
                           "-I", os.path.join(facade_util.ROOT, "cvids_b200", "include"), os.path.join(facade_util.ROOT, "tests", "cpp", "ply_test.cpp"), "-o", exe])
    a, b = str(tmp_path / "a.ply"), str(tmp_path / "b.ply")
    subprocess.check_call([exe, a, b])
    lines = open(a).read().split("\n")
    assert lines[0] == "ply" and lines[1] == "format ascii 1.0" and lines[2] == "element vertex 21"
    end = lines.index("end_header")
    verts = np.array([[float(x) for x in l.split()] for l in lines[end + 1:end + 22]])
    faces = np.array([[int(x) for x in l.split()] for l in lines[end + 22:end + 29]])
    raw = open(b, "rb").read()
    hdr_end = raw.index(b"end_header\n") + len(b"end_header\n")
    assert b"format binary_little_endian 1.0" in raw[:hdr_end] and b"element vertex 21" in raw[:hdr_end] and b"element face 7" in raw[:hdr_end]
    vt = np.dtype([("xyz", "<f4", 3), ("rgb", "u1", 3)])
    bv = np.frombuffer(raw, dtype=vt, count=21, offset=hdr_end)
    ft = np.dtype([("n", "u1"), ("idx", "<i4", 3)])
    bf = np.frombuffer(raw, dtype=ft, count=7, offset=hdr_end + 21 * vt.itemsize)
    assert len(raw) == hdr_end + 21 * 15 + 7 * 13
    assert np.allclose(bv["xyz"], verts[:, :3], rtol=1e-5, atol=1e-6)          # the ASCII file holds 6 significant digits
    assert np.array_equal(bv["rgb"], verts[:, 3:].astype(np.uint8))
    assert np.all(bf["n"] == 3) and np.array_equal(bf["idx"], faces[:, 1:])

"""CPU suite: the C-ABI library loads, exports every symbol include/chisel_b200.h declares, its host-side exact
geometry agrees bit for bit with the oracle, and it fails loudly (no fallback) without a CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

from cvids_b200 import capi, scenes
from tests import common

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "chisel_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(chs_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    names = _declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libchisel_b200.so does not export %s" % n
    assert sorted(capi.EXPORTS) == names
    assert lib.chs_abi_version() == 1


def test_library_is_sm100a_only():
    import subprocess
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_frustum_candidates_truncation_match_oracle():
    from oracle.pyoracle import OracleChisel
    o = OracleChisel(16, 0.02, False)
    for f in range(0, 50, 9):
        pose = scenes.orbit_pose(f, 50, 0.2)
        for cam in (scenes.KINECT_640.as_array(), scenes.EUROC_752.as_array()):
            for x, y in zip(capi.frustum(pose, cam), o.frustum(pose, cam)):
                assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
            assert np.array_equal(capi.candidate_ids(16, 0.02, pose, cam), o.candidate_ids(pose, cam))
    rng = np.random.RandomState(1)
    for d in rng.uniform(0.05, 20, 100).astype(np.float32):
        for kind, p in ((0, 0.2), (1, 2.0), (2, 8.0)):
            x, y = np.float32(capi.truncation(kind, p, float(d))), np.float32(o.truncation(kind, p, float(d)))
            assert x.view(np.uint32) == y.view(np.uint32)


def test_owner_hash_is_stable_and_balanced():
    ids = [(x, y, z) for x in range(-8, 8) for y in range(-8, 8) for z in range(-4, 4)]
    for world in (2, 4, 8):
        counts = np.bincount([capi.owner(*i) % world for i in ids], minlength=world)
        assert counts.min() > 0.8 * len(ids) / world, counts
    assert capi.owner(1, -2, 3) == capi.owner(1, -2, 3)


def test_argument_validation_needs_no_device():
    lib = capi.load_library()
    assert lib.chs_create(None, None) == capi.CHS_ERR_INVALID
    cfg = capi.chs_config(12, 0.05, 0, -1, 0, 1, 0, None)
    h = ctypes.c_void_p()
    assert lib.chs_create(ctypes.byref(cfg), ctypes.byref(h)) == capi.CHS_ERR_INVALID
    assert b"chunk_size" in lib.chs_last_error_string()


def test_no_cpu_fallback():
    """Without a device, creating a map must fail with CHS_ERR_CUDA; with one, this test is a no-op."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.ChiselError) as e:
        capi.Chisel(16, 0.05, False)
    assert e.value.code == capi.CHS_ERR_CUDA


def test_product_never_imports_oracle():
    """The product package must not reference oracle/ (the checker) in any way."""
    pkg = os.path.join(ROOT, "cvids_b200")
    bad = re.compile(r"^\s*(from\s+oracle|import\s+oracle)|pyoracle|chisel_oracle|libchisel_ref|#\s*include\s*[<\"][^>\"]*oracle", re.M)
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dp, f)).read()
                assert not bad.search(text), os.path.join(dp, f)


def test_color_integrate_integer_identity():
    """The CUDA kernel evaluates ColorVoxel::Integrate (ColorVoxel.h:65-85) in integer arithmetic. Exhaustive proof that
    (N * ceil(2^20 / D)) >> 20 equals the reference's uint8(saturate(float(w*old + new) / float(w + 1))) in binary32 for
    every colour weight the integrator lets through (w < 8, ProjectionIntegrator.h:153) and every old / new byte."""
    recip = np.array([0, 1048576, 524288, 349526, 262144, 209716, 174763, 149797, 131072], np.uint64)
    for d in range(1, 9):
        assert recip[d] == -(-(1 << 20) // d)
    old = np.arange(256, dtype=np.float32)[:, None]
    new = np.arange(256, dtype=np.float32)[None, :]
    for w in range(8):
        wf = np.float32(w)
        ref = np.clip((wf * old + new) / np.float32(w + 1), np.float32(0), np.float32(255)).astype(np.uint8)
        n = (w * np.arange(256, dtype=np.uint64)[:, None] + np.arange(256, dtype=np.uint64)[None, :])
        got = ((n * recip[w + 1]) >> np.uint64(20)).astype(np.uint8)
        assert np.array_equal(ref, got), w


def test_batch_entry_points_validate_arguments_without_a_device():
    """chs_integrate_batch / chs_wait_batch reject bad arguments before touching CUDA; the pinned allocator fails cleanly
    (NULL) where there is no device."""
    import ctypes as C
    from cvids_b200 import capi
    lib = capi.load_library()
    integ = capi.ProjectionIntegrator().as_struct()
    cam = capi.make_camera([100, 100, 32, 24, 64, 48, 0.05, 5.0])
    fr = (capi.chs_frame * 1)()
    assert lib.chs_integrate_batch(None, C.byref(integ), 1, fr, capi.MEM_HOST, C.byref(cam), 3, C.byref(cam)) == capi.CHS_ERR_INVALID
    n = C.c_int()
    assert lib.chs_wait_batch(None, 1, None, 0, C.byref(n)) == capi.CHS_ERR_INVALID
    t = C.c_int64()
    assert lib.chs_last_batch_ticket(None, C.byref(t)) == capi.CHS_ERR_INVALID
    import torch
    if not torch.cuda.is_available():
        assert not lib.chs_host_alloc(1024)
    lib.chs_host_free(None)


def test_chs_frame_layout_matches_the_header():
    """The ctypes mirror of chs_frame (four pointers, two 3x4 poses) has the C struct's size."""
    import ctypes as C
    from cvids_b200 import capi
    assert C.sizeof(capi.chs_frame) == 4 * C.sizeof(C.c_void_p) + 2 * 12 * C.sizeof(C.c_float)


def test_ctypes_mirrors_match_the_c_header(tmp_path):
    """Compile a C program against include/chisel_b200.h that prints sizeof / offsetof of every ABI struct and compare with the
    ctypes mirrors in cvids_b200/capi.py (catches silent ABI drift between the header and the Python harness)."""
    import ctypes as C
    import subprocess
    from cvids_b200 import capi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    structs = {"chs_config": capi.chs_config, "chs_camera": capi.chs_camera, "chs_integrator": capi.chs_integrator,
               "chs_frame_stats": capi.chs_frame_stats, "chs_mesh_counts": capi.chs_mesh_counts, "chs_timings": capi.chs_timings,
               "chs_frame": capi.chs_frame}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "chisel_b200.h"', 'int main(void) {']
    for name, cls in structs.items():
        lines.append('printf("%s sizeof %%zu\\n", sizeof(%s));' % (name, name))
        for field, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (name, field, name, field))
    lines += ['return 0; }']
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = str(tmp_path / "abi")
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(root, "include"), str(src), "-o", exe])
    got = dict(l.rsplit(" ", 1) for l in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.strip().split("\n"))
    for name, cls in structs.items():
        assert int(got["%s sizeof" % name]) == C.sizeof(cls), name
        for field, _ in cls._fields_:
            assert int(got["%s.%s" % (name, field)]) == getattr(cls, field).offset, "%s.%s" % (name, field)


def test_timeline_names_match_the_header_and_the_kernels():
    """chs_get_device_timeline: the header's stamp count, the kernels' enum and the Python names agree."""
    hdr = open(os.path.join(ROOT, "include", "chisel_b200.h")).read()
    n = int(re.search(r"#define CHS_TIMELINE_STAMPS (\d+)", hdr).group(1))
    kern = open(os.path.join(ROOT, "cvids_b200", "csrc", "kernels.h")).read()
    enum = re.search(r"enum TimelineStamp\s*\{(.*?)\}", kern, flags=re.S).group(1)
    names = [x for x in re.findall(r"\b(kT[A-Za-z]+)\b", re.sub(r"//.*", "", enum))]
    assert names[-1] == "kTimelineStamps" and len(names) - 1 == n
    assert len(capi.Chisel.TIMELINE) == n
    assert [x[3:].lower() for x in names[:-1]] == [x.replace("_", "") for x in capi.Chisel.TIMELINE]


def test_both_builds_of_the_fused_kernels_are_in_the_library():
    """integrate_batch_impl.cuh is compiled twice (half-brick tasks for depth-only batches, quarter-brick tasks for colour
    batches); both launchers and both fast brick kernels must be there, for sm_100a."""
    import subprocess
    syms = subprocess.run(["nm", "-C", capi.LIB_PATH], capture_output=True, text=True).stdout
    for ns in ("half", "quarter"):
        assert "chs::%s::launch_batch(" % ns in syms, ns
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-res-usage", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert re.search(r"4half24batch_bricks_fast_kernel", out) and re.search(r"7quarter24batch_bricks_fast_kernel", out)
    assert re.search(r"peer_push_kernel", out) and re.search(r"peer_wait_kernel", out)


def test_bench_timeline_summary():
    """bench.py's reading of the device timeline: medians in microseconds, gaps between consecutive steps, missing stamps."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    def step(t0, push=True):
        return {"push_start": t0 - 90000 if push else 0, "push_end": t0 - 30000 if push else 0, "wait_end": t0 - 1000 if push else 0,
                "hiz_start": t0, "hiz_end": t0 + 12000, "cand_start": t0 + 15000, "cand_end": t0 + 35000,
                "bricks_start": t0 + 36000, "bricks_end": t0 + 90000}
    s = bench.summarize_timeline([step(1000000 + 100000 * i) for i in range(5)])
    assert s["step"] == 100.0 and s["hiz"] == 12.0 and s["candidates"] == 20.0 and s["bricks"] == 54.0
    assert s["gap_bricks_to_next_hiz"] == 10.0 and s["gap_candidates_to_bricks"] == 1.0 and s["push"] == 60.0 and s["push_end_to_hiz_start"] == 30.0
    s = bench.summarize_timeline([step(1000000 + 100000 * i, push=False) for i in range(3)])
    assert s["push"] is None and s["arrival_to_hiz_start"] is None and len(s["steps"]) == 2

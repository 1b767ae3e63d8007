#!/usr/bin/env python3
"""Debug helper: the scenario of tests/test_parity_gpu.py::test_import_overwrite_resets_brick_flags_for_carving with a report of the
voxels that differ from the oracle (which path: fused / generic brick kernel / single-frame calls)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import common  # noqa: E402
from tests.common import Setup  # noqa: E402


def run(mode):
    setup = Setup(16, 0.05, False)
    cam = common.SMALL_CAM
    camv = cam.as_array()
    frames = list(common.carve_stream(cam, 3, 5))
    a, o = common.Driver(setup, "cuda"), common.Driver(setup, "oracle")
    for depth, _, pose in frames[3:5]:
        a.integrate(depth, pose, camv)
        o.integrate(depth, pose, camv)
    ids, sdf, w, rgbw = o.state()
    sdf, w = sdf.copy(), w.copy()
    rng = np.random.RandomState(3)
    planted = np.zeros(sdf.shape, bool)
    for c in range(0, len(ids), 3):
        v = rng.choice(sdf.shape[1], 400, replace=False)
        sdf[c, v] = np.float32(-0.02)
        w[c, v] = np.float32(3.0)
        planted[c, v] = True
    a.m.import_chunks(ids, sdf, w)
    o.m.import_chunks(ids, sdf, w)
    grp = frames[5:]
    if mode == "single":
        for depth, _, pose in grp:
            a.integrate(depth, pose, camv)
    else:
        a.m.integrate_batch(a.integ, [f[0] for f in grp], [f[2] for f in grp], camv)
    hist = []
    for depth, _, pose in grp:
        o.integrate(depth, pose, camv)
        st = o.state()
        hist.append((st[0].copy(), st[1].copy(), st[2].copy()))
    sa, so = a.state(), o.state()
    assert np.array_equal(sa[0], so[0])
    bad = np.argwhere(sa[2].view(np.uint32) != so[2].view(np.uint32))
    print(mode, "mismatching weights:", len(bad), "planted among them:", int(planted[bad[:, 0], bad[:, 1]].sum()) if len(bad) else 0)
    for c, v in bad[:12]:
        x, y, z = v % 16, (v // 16) % 16, v // 256
        print("  chunk", sa[0][c], "voxel", (x, y, z), "brick", (x // 8, y // 8, z // 8), "cuda sdf,w", sa[1][c, v], sa[2][c, v], "oracle", so[1][c, v], so[2][c, v],
              "imported", sdf[c, v], w[c, v])
        key = tuple(sa[0][c])
        for fi, (hid, hs, hw) in enumerate(hist):
            r = [i for i in range(len(hid)) if tuple(hid[i]) == key]
            if r:
                print("      oracle after frame", fi, ":", hs[r[0], v], hw[r[0], v])


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1])
    else:
        import subprocess
        for mode, env in (("batch", {}), ("batch", {"CHS_NO_FAST_BRICKS": "1"}), ("single", {})):
            print("==", mode, env, flush=True)
            subprocess.run([sys.executable, __file__, mode], env=dict(os.environ, **env))

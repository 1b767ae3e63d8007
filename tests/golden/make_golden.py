#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ from oracle/_ref -- the UNMODIFIED reference sources
compiled in place from /root/reference (oracle/Makefile, target `ref`). Run in the build container only:

    python tests/golden/make_golden.py

The reference ships no golden vectors of its own; these fixtures are what pins the C
restatement (tests/test_oracle_golden.py, everywhere) and the CUDA path (tests/test_parity_gpu.py, on the GPU box,
where /root/reference does not exist).

Fixtures:
  digests.json      sha256 digests + counts of the final state of several streams
  small_state.npz   full voxel state, dirty set and meshes of one tiny stream (8^3 chunks, 10 cm, 160x120)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from tests import common  # noqa: E402
from tests.golden import cases  # noqa: E402


def main():
    from oracle import pyoracle
    pyoracle.build("ref")
    out = {}
    for name, case in cases.CASES.items():
        drv = common.Driver(case["setup"], "ref")
        cases.run_case(drv, case)
        out[name] = dict(state=common.digest_state(drv.state()), meshes=common.digest_meshes(drv.meshes()),
                         dirty=int(len(drv.dirty())))
        print(name, out[name])
        if name == cases.FULL_STATE_CASE:
            ids, sdf, w, rgbw = drv.state()
            meshes = drv.meshes()
            arrays = dict(ids=ids, sdf=sdf, weight=w, rgbw=rgbw, dirty=drv.dirty(),
                          mesh_ids=np.asarray(sorted(meshes), np.int32).reshape(-1, 3))
            for i, k in enumerate(sorted(meshes)):
                for f in ("vertices", "normals", "colors", "grids"):
                    arrays["mesh%d_%s" % (i, f)] = meshes[k][f]
            np.savez_compressed(os.path.join(HERE, "small_state.npz"), **arrays)
    with open(os.path.join(HERE, "digests.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

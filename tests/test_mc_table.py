"""CPU suite: the packed marching-cubes table (tools/gen_mc_table.py) is structurally a valid Bourke table, both
copies are identical, and (in the build container) it equals the reference's array entry for entry."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EDGES = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]


def _load(path):
    text = open(path).read()
    packed = [int(x, 16) for x in re.findall(r"0x([0-9A-F]{16})ULL", text)]
    counts = [int(x) for x in re.search(r"MC_TRI_COUNT_INIT \{(.*?)\}", text, re.S).group(1).replace("\\", "").split(",") if x.strip()]
    return packed, counts


def test_two_copies_identical():
    a = open(os.path.join(ROOT, "oracle", "mc_table.inc")).read()
    b = open(os.path.join(ROOT, "cvids_b200", "csrc", "mc_table.inc")).read()
    assert a == b


def test_structure():
    packed, counts = _load(os.path.join(ROOT, "oracle", "mc_table.inc"))
    assert len(packed) == 256 and len(counts) == 256
    for cfg in range(256):
        ent = [(packed[cfg] >> (4 * k)) & 0xF for k in range(16)]
        n = ent.index(0xF)
        assert n % 3 == 0 and n // 3 == counts[cfg] and all(e == 0xF for e in ent[n:])
        # every referenced edge must have a sign change in this configuration
        for e in ent[:n]:
            a, b = EDGES[e]
            assert ((cfg >> a) & 1) != ((cfg >> b) & 1), (cfg, e)
        # all edges with a sign change are used (the surface is closed inside the cube)
        crossing = {i for i, (a, b) in enumerate(EDGES) if ((cfg >> a) & 1) != ((cfg >> b) & 1)}
        assert set(ent[:n]) == crossing, cfg
    assert counts[0] == 0 and counts[255] == 0 and max(counts) == 5


@pytest.mark.skipif(not os.path.exists("/root/reference/OpenChisel/open_chisel/src/marching_cubes/MarchingCubes.cpp"),
                    reason="reference not present")
def test_equals_reference_table():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_mc_table
    packed, counts = gen_mc_table.pack(gen_mc_table.read_reference_table())
    assert (packed, counts) == _load(os.path.join(ROOT, "cvids_b200", "csrc", "mc_table.inc"))

#!/usr/bin/env python3
"""Golden fixture for the PLY export: the bytes Chisel::SaveAllMeshesToPLY (OC/src/Chisel.cpp:69-105, OC/src/io/PLY.cpp:29-88)
writes when the REFERENCE's own sources run tests/cpp/chisel_client.cpp on the streams of tests/test_facade.py. Build container only
(needs /root/reference):

    python tests/golden/make_ply_golden.py

Writes tests/golden/ply_golden.json: per case the header lines, the vertex / face counts, sha256 of the raw file and sha256 of the
file's CANONICAL form -- header, then the triangles (three vertex lines each) sorted, then the face lines. The reference walks an
unordered_map of meshes, so the order of the per-chunk blocks is an accident of libstdc++'s bucket layout; everything else is
compared byte for byte."""
import hashlib
import json
import os
import pathlib
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from tests import facade_util  # noqa: E402


def canonical(text: str):
    """(header lines, n vertices, n faces, canonical bytes) of an ASCII PLY triangle soup."""
    lines = text.split("\n")
    end = lines.index("end_header")
    header = lines[:end + 1]
    nv = int([l for l in header if l.startswith("element vertex")][0].split()[-1])
    nf = int([l for l in header if l.startswith("element face")][0].split()[-1])
    verts = lines[end + 1:end + 1 + nv]
    faces = lines[end + 1 + nv:end + 1 + nv + nf]
    tris = sorted("\n".join(verts[3 * i:3 * i + 3]) for i in range(nv // 3))
    canon = "\n".join(header + tris + faces) + "\n"
    return header, nv, nf, canon.encode()


def main():
    from tests.test_facade import CASES, _stream
    exe = facade_util.build_reference_client()
    out = {}
    tmp = pathlib.Path(tempfile.mkdtemp())
    for case in sorted(CASES):
        stream, _ = _stream(tmp, case)
        dump = str(tmp / (case + ".dump"))
        subprocess.run(["taskset", "-c", "0", exe, stream, dump], check=True, stdout=subprocess.DEVNULL)
        raw = open(dump + ".ply", "rb").read()
        header, nv, nf, canon = canonical(raw.decode())
        out[case] = dict(header=header, vertices=nv, faces=nf, bytes=len(raw), sha256_raw=hashlib.sha256(raw).hexdigest(),
                         sha256_canonical=hashlib.sha256(canon).hexdigest())
        print(case, out[case]["vertices"], out[case]["faces"], out[case]["sha256_canonical"][:16])
    with open(os.path.join(HERE, "ply_golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

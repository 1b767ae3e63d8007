#!/usr/bin/env python3
"""Golden vectors for the frame-ingestion kernel (chs_ingest_depth): what cv2.resize (INTER_LINEAR) does to a float depth map in
the shape the collaborative server resizes (SPG/src/collaborative_server_system.cpp:213-214: sensor size -> 640x480; here the same
ratio at a quarter of the size, 188x120 -> 160x120, and a case that changes both axes). Needs the OpenCV Python wheel (build
container): python tests/golden/make_ingest_golden.py -> tests/golden/ingest_golden.npz"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.RandomState(5)
    out = {}
    for name, (sh, sw), (dh, dw) in (("euroc_quarter", (120, 188), (120, 160)), ("both_axes", (90, 150), (120, 160)), ("upscale", (60, 80), (120, 160))):
        src = rng.uniform(0.05, 25.0, size=(sh, sw)).astype(np.float32)
        src[sh // 4:sh // 3, sw // 5:sw // 3] = np.nan                       # invalid input pixels spread by the interpolation
        out[name + "_src"] = src
        out[name + "_resized"] = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
    np.savez_compressed(os.path.join(HERE, "ingest_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

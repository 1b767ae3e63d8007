#!/usr/bin/env python3
"""Build differently tuned variants of the library for A/B runs on the GPU box:
    python tools/ab_build.py name1:DEF=1,DEF2=3 name2:...
writes cvids_b200/_ab_<name>.so (git-ignored); select one with CHS_LIB_PATH."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cvids_b200 import build  # noqa: E402

for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    out = os.path.join(ROOT, "cvids_b200", "_ab_%s.so" % name)
    build.build(force=True, out=out, defines=[d for d in defs.split(",") if d])
    print(out)

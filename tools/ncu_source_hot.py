#!/usr/bin/env python3
"""Per-source-line executed-instruction counts and stall samples of one kernel: joins the per-SASS-address metrics of an ncu
report (--page source) with nvdisasm's line info of the shipped cubin.
usage: ncu_source_hot.py report.ncu-rep kernel-regex mangled-substring cubin [top]"""
import collections
import csv
import io
import re
import subprocess
import sys


def line_map(cubin, mangled):
    out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
    m, cur, on = {}, None, False
    for l in out:
        if l.startswith(".text."):
            on = mangled in l
            continue
        if not on:
            continue
        if "//## File" in l:
            g = re.search(r'File "([^"]+)", line (\d+)', l)
            if g:
                cur = (g.group(1).split("/")[-1], int(g.group(2)))
            continue
        g = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
        if g:
            m[int(g.group(1), 16)] = cur
    return m


def main():
    rep, kern, mangled, cubin = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    lm = line_map(cubin, mangled)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ci, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
    base = None
    inst, smp = collections.Counter(), collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) <= ci or not r[0].startswith("0x"):
            continue
        a = int(r[0], 16)
        if base is None:
            base = a
        key = lm.get(a - base, ("?", 0))
        inst[key] += int(r[ci])
        smp[key] += int(r[si])
    tot, tots = sum(inst.values()) or 1, sum(smp.values()) or 1
    print("total warp instructions %d, samples %d" % (tot, tots))
    for key, n in inst.most_common(top):
        print("%6.2f%% inst %6.2f%% smp  %s:%d" % (100.0 * n / tot, 100.0 * smp[key] / tots, key[0], key[1]))


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Static SASS instruction mix of one kernel of the shipped library (cuobjdump -sass), grouped by pipe.
usage: sass_mix.py <regex on the mangled function name> [library.so]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GROUPS = [("FMA pipe (FFMA/FMUL/FADD/IMAD...)", r"^(FFMA|FMUL|FADD|IMAD|FMNMX3?|FSEL|FCHK|HFMA2|HADD2|HMUL2)$"),
          ("ALU pipe (compare/select/logic/integer add/shift)", r"^(FSETP|FSET|ISETP|VIADD|LOP3|PLOP3|SEL|IADD3|IADD|LEA|SHF|SHL|SHR|PRMT|MOV|IABS|IMNMX|VIMNMX|BMSK|SGXT|POPC|FLO|BREV|P2R|R2P|CS2R|S2R|VOTE|VOTEU|I2FP|F2FP)"),
          ("XU (MUFU, conversions)", r"^(MUFU|F2I|I2F|F2F|FRND)"),
          ("memory: shared / constant", r"^(LDS|STS|ATOMS|LDC|LDCU|ULDC|LDSM)$"),
          ("memory: global / local", r"^(LDG|STG|LD|ST|LDL|STL|ATOMG|ATOM|RED|REDG)$"),
          ("warp: shuffle / redux / match", r"^(SHFL|REDUX|MATCH|WARPSYNC)"),
          ("uniform datapath", r"^(U[A-Z0-9]+|R2UR|S2UR)"),
          ("control", r"^(BRA|BRX|JMP|CALL|RET|EXIT|BSSY|BSYNC|BAR|BMOV|NANOSLEEP|YIELD|DEPBAR|ERRBAR|MEMBAR|NOP|WARPSYNC|ACQBULK|SYNCS|CCTL|FENCE)")]


def main():
    pat = re.compile(sys.argv[1])
    lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "cvids_b200", "libchisel_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    name, ops = None, collections.Counter()
    for line in out.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                break
            name = m.group(1) if pat.search(m.group(1)) else None
            continue
        if name:
            m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
            if m:
                ops[m.group(1)] += 1
    total = sum(ops.values())
    print("kernel:", name)
    print("static SASS instructions:", total)
    rest = dict(ops)
    for label, rx in GROUPS:
        sel = {k: v for k, v in rest.items() if re.match(rx, k)}
        for k in sel:
            rest.pop(k)
        n = sum(sel.values())
        print("  %-58s %6d  %5.1f %%   %s" % (label, n, 100.0 * n / max(total, 1), ", ".join("%s %d" % kv for kv in sorted(sel.items(), key=lambda kv: -kv[1])[:8])))
    n = sum(rest.values())
    print("  %-58s %6d  %5.1f %%   %s" % ("other", n, 100.0 * n / max(total, 1), ", ".join("%s %d" % kv for kv in sorted(rest.items(), key=lambda kv: -kv[1])[:8])))


if __name__ == "__main__":
    main()

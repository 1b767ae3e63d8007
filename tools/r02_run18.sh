#!/bin/bash
mkdir -p gpurun_out
CHS_BENCH_DEBUG_TIMELINE=1 timeout 600 python bench.py --no-cpu --no-side-lines --quick > gpurun_out/r02_18.json 2> gpurun_out/r02_18.err
grep -A30 "timed pass" gpurun_out/r02_18.err

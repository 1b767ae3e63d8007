#!/bin/bash
mkdir -p gpurun_out
export CHS_HOST_PROFILE=1
timeout 600 python bench.py --no-cpu --no-side-lines --quick > gpurun_out/r02_27_n1.json 2> gpurun_out/r02_27_n1.err; grep -A1 "host profile" gpurun_out/r02_27_n1.err

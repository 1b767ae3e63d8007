#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"batch_candidates" --launch-skip 8 --launch-count 1 -o gpurun_out/r02_7_cand python bench.py --steps 6 --warmup 3 --passes 1 --no-cpu --quick --no-side-lines --parity-steps 0 --time-steps-per-step 2 > gpurun_out/r02_7_ncu.log 2>&1
tail -2 gpurun_out/r02_7_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --launch-skip 60 --csv --log-file gpurun_out/r02_7_launches.csv python bench.py --steps 6 --warmup 3 --passes 1 --no-cpu --quick --no-side-lines --parity-steps 0 --time-steps-per-step 2 > /dev/null 2>&1
grep -c . gpurun_out/r02_7_launches.csv

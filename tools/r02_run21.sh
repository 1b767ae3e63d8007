#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 6 --warmup 3 --passes 1 --no-cpu --quick --no-side-lines --parity-steps 0"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"batch_candidates|batch_bricks_fast|batch_hiz_tma" --launch-skip 30 --launch-count 3 -o gpurun_out/r02_21_kernels $B > gpurun_out/r02_21_ncu.log 2>&1
tail -2 gpurun_out/r02_21_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 90 --launch-skip 40 --csv --log-file gpurun_out/r02_21_launches.csv $B > /dev/null 2>&1
tail -12 gpurun_out/r02_21_launches.csv | cut -c1-200

#!/bin/bash
mkdir -p gpurun_out
export CHS_HOST_PROFILE=1
timeout 600 python bench.py --no-cpu --no-side-lines --quick > gpurun_out/r02_26_n1.json 2> gpurun_out/r02_26_n1.err; grep -A1 "host profile" gpurun_out/r02_26_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-cpu --quick > gpurun_out/r02_26_n2.json 2> gpurun_out/r02_26_n2.err; grep -A1 "host profile" gpurun_out/r02_26_n2.err

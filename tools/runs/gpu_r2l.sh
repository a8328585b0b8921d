#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --clips 20000 --e2e-clips 2048 --no-cpu --no-extra > gpurun_out/plain_r2l.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:logmel512_kernel -s 3 -c 1 -f -o gpurun_out/prof_r2l python bench.py --steps 3 --warmup 3 --clips 20000 --e2e-clips 2048 --no-cpu --no-extra > gpurun_out/ncu_r2l.log 2>&1
tail -2 gpurun_out/ncu_r2l.log

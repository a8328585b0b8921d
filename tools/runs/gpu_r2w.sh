#!/bin/bash
# round 2, GPU run W: config 5 with the mapped features file pre-faulted in the background
mkdir -p gpurun_out
timeout 300 python bench_stage2.py --devices 0 > gpurun_out/stage2_r2w.json 2> gpurun_out/stage2_r2w.err; cut -c1-600 gpurun_out/stage2_r2w.json; tail -2 gpurun_out/stage2_r2w.err

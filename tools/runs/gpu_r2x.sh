#!/bin/bash
# round 2, GPU run X (8 GPUs): config 5 end to end, same-rate and 44.1 kHz files
mkdir -p gpurun_out
timeout 300 python bench_stage2.py --devices all > gpurun_out/stage2_8gpu_r2x.json 2> gpurun_out/stage2_8gpu_r2x.err; cut -c1-500 gpurun_out/stage2_8gpu_r2x.json; tail -2 gpurun_out/stage2_8gpu_r2x.err
timeout 300 python bench_stage2.py --devices all --file-rate 44100 > gpurun_out/stage2_8gpu_44k_r2x.json 2>> gpurun_out/stage2_8gpu_r2x.err; cut -c1-500 gpurun_out/stage2_8gpu_44k_r2x.json

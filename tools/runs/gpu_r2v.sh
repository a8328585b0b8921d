#!/bin/bash
# round 2, GPU run V: in-place features.npy + native loader probe (config 5), classical dataset tests, 405-clip classical parity
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_classical.py tests/test_gpu_extractors.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_r2v.log; tail -6 gpurun_out/pytest_r2v.log | cut -c1-300
timeout 300 python bench_stage2.py --devices 0 > gpurun_out/stage2_r2v.json 2> gpurun_out/stage2_r2v.err; cut -c1-700 gpurun_out/stage2_r2v.json; tail -2 gpurun_out/stage2_r2v.err
timeout 600 python tools/classical_check.py 405 > gpurun_out/cls_parity_r2v.jsonl 2> gpurun_out/cls_parity_r2v.err; tail -2 gpurun_out/cls_parity_r2v.err; cut -c1-300 gpurun_out/cls_parity_r2v.jsonl

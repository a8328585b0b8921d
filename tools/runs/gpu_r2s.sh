#!/bin/bash
# round 2, GPU run S: audio_classical tests, rate, profile
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_classical.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_r2s.log; tail -12 gpurun_out/pytest_r2s.log | cut -c1-300
timeout 120 python tools/prof_classical.py 2>&1 | tail -1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:classical -s 1 -c 1 -f -o gpurun_out/prof_cls_r2s python tools/prof_classical.py > gpurun_out/ncu_cls_r2s.log 2>&1
tail -1 gpurun_out/ncu_cls_r2s.log

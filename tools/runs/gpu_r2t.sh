#!/bin/bash
# round 2, GPU run T: audio_classical after the selection rewrite / 512-thread CTAs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_classical.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_r2t.log; tail -6 gpurun_out/pytest_r2t.log | cut -c1-300
timeout 300 python tools/classical_check.py 42 > gpurun_out/cls_check_r2t.jsonl 2> gpurun_out/cls_check_r2t.err; tail -2 gpurun_out/cls_check_r2t.err
timeout 120 python tools/prof_classical.py 2>&1 | tail -1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:classical -s 1 -c 1 -f -o gpurun_out/prof_cls_r2t python tools/prof_classical.py > gpurun_out/ncu_cls_r2t.log 2>&1
tail -1 gpurun_out/ncu_cls_r2t.log

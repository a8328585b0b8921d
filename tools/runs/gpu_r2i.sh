#!/bin/bash
# round 2, GPU run I: 1024 kernel at 96 registers per role (no setmaxnreg, no spills) + shifted staging
mkdir -p gpurun_out
timeout 90 python tools/prof_1024.py > gpurun_out/p1024_r2i.log 2>&1 || { echo "1024 mel run failed/hung"; tail -3 gpurun_out/p1024_r2i.log; exit 1; }
tail -1 gpurun_out/p1024_r2i.log
timeout 90 python tools/prof_1024.py mfcc > gpurun_out/p1024m_r2i.log 2>&1 || { echo "1024 mfcc run failed/hung"; tail -3 gpurun_out/p1024m_r2i.log; exit 1; }
tail -1 gpurun_out/p1024m_r2i.log
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_r2i.log; tail -6 gpurun_out/pytest_r2i.log
timeout 200 python tools/sweep.py 2> gpurun_out/sweep_r2i.err | head -8 > gpurun_out/sweep_r2i.jsonl; cat gpurun_out/sweep_r2i.jsonl | cut -c1-220

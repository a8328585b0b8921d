#!/bin/bash
mkdir -p gpurun_out
timeout 90 python tools/prof_1024.py > gpurun_out/p1024_r2k.log 2>&1 || { echo "1024 mel run failed/hung"; tail -3 gpurun_out/p1024_r2k.log; exit 1; }
tail -1 gpurun_out/p1024_r2k.log
timeout 90 python tools/prof_1024.py mfcc > gpurun_out/p1024m_r2k.log 2>&1 || { echo "1024 mfcc run failed/hung"; exit 1; }
tail -1 gpurun_out/p1024m_r2k.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4

#!/bin/bash
mkdir -p gpurun_out
timeout 90 python tools/prof_1024.py > gpurun_out/p1024_r2n.log 2>&1 || { echo "1024 mel run failed/hung"; tail -3 gpurun_out/p1024_r2n.log; exit 1; }
tail -1 gpurun_out/p1024_r2n.log
timeout 90 python tools/prof_1024.py mfcc > gpurun_out/p1024m_r2n.log 2>&1 || { echo "1024 mfcc run failed/hung"; tail -3 gpurun_out/p1024m_r2n.log; exit 1; }
tail -1 gpurun_out/p1024m_r2n.log
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_extractors.py -m gpu -x -q 2>&1 | tail -12
timeout 120 python tools/cqt_bench.py 8192 2>/dev/null | tail -1

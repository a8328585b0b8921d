#!/bin/bash
# round 2, GPU run P: tile-DCT rewrite of the 1024 kernel — parity, speed, profiles of both kinds
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_extractors.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_r2p.log; tail -3 gpurun_out/pytest_r2p.log
timeout 90 python tools/prof_1024.py mfcc 2>&1 | tail -1
timeout 90 python tools/prof_1024.py 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:logmel1024 -s 2 -c 1 -f -o gpurun_out/prof_1024m_r2p python tools/prof_1024.py mfcc > gpurun_out/ncu_1024m_r2p.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:logmel1024 -s 2 -c 1 -f -o gpurun_out/prof_1024_r2p python tools/prof_1024.py > gpurun_out/ncu_1024_r2p.log 2>&1
tail -1 gpurun_out/ncu_1024_r2p.log

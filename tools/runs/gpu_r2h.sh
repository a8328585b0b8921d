#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/prof_1024.py > gpurun_out/plain_1024_r2h.log 2>&1; cat gpurun_out/plain_1024_r2h.log | tail -2
timeout 120 python tools/prof_1024.py mfcc > gpurun_out/plain_1024m_r2h.log 2>&1; cat gpurun_out/plain_1024m_r2h.log | tail -2
timeout 120 python tools/prof_1024.py > gpurun_out/plain2_1024_r2h.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:logmel1024 -s 2 -c 1 -f -o gpurun_out/prof_1024_r2h python tools/prof_1024.py > gpurun_out/ncu_1024_r2h.log 2>&1
tail -2 gpurun_out/ncu_1024_r2h.log

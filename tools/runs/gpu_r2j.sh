#!/bin/bash
mkdir -p gpurun_out
timeout 90 python tools/prof_1024.py > gpurun_out/plain_1024_r2j.log 2>&1 &&
timeout 200 ncu --set full --clock-control none --import-source on -k regex:logmel1024 -s 2 -c 1 -f -o gpurun_out/prof_1024_r2j python tools/prof_1024.py > gpurun_out/ncu_1024_r2j.log 2>&1
tail -2 gpurun_out/ncu_1024_r2j.log

#!/bin/bash
# round 2, GPU run F: reworked CQT bank kernel (8 frames/lane, batched staging) + two-level decimator; chunk sweep; ncu
mkdir -p gpurun_out
python tools/cqt_floor.py 405 bank8_lvl1 > gpurun_out/cqt_floor_r2f.jsonl 2> gpurun_out/floor_r2f.err; cat gpurun_out/cqt_floor_r2f.jsonl; tail -3 gpurun_out/floor_r2f.err
: > gpurun_out/cqt_speed_r2f.jsonl
for c in 1024 256 192 128; do B2A_CQT_CHUNK=$c python tools/cqt_bench.py 8192 >> gpurun_out/cqt_speed_r2f.jsonl 2>> gpurun_out/floor_r2f.err; done
cat gpurun_out/cqt_speed_r2f.jsonl
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_r2f.log; tail -4 gpurun_out/pytest_r2f.log
python tools/cqt_bench.py 2048 > gpurun_out/plain_cqt_r2f.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:cqt_bank_kernel -s 2 -c 1 -f -o gpurun_out/prof_cqt_bank_r2f python tools/cqt_bench.py 2048 > gpurun_out/ncu_f_cqt_r2f.log 2>&1
tail -2 gpurun_out/ncu_f_cqt_r2f.log

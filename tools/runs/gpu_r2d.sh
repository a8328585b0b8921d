#!/bin/bash
# round 2, GPU run D: fast CQT bank kernel + decimator accumulation variants; augmentation tests
mkdir -p gpurun_out
V=audio_edge_ml_pipeline_b200/build
: > gpurun_out/cqt_floor_r2d.jsonl; : > gpurun_out/floor_r2d.err; : > gpurun_out/cqt_speed_r2d.jsonl
python tools/cqt_floor.py 405 product >> gpurun_out/cqt_floor_r2d.jsonl 2>> gpurun_out/floor_r2d.err
python tools/cqt_bench.py >> gpurun_out/cqt_speed_r2d.jsonl 2>> gpurun_out/floor_r2d.err
for t in lvl1 lvl2; do
  B2A_LIBRARY=$PWD/$V/libb2a_$t.so python tools/cqt_floor.py 405 $t >> gpurun_out/cqt_floor_r2d.jsonl 2>> gpurun_out/floor_r2d.err
  B2A_LIBRARY=$PWD/$V/libb2a_$t.so python tools/cqt_bench.py >> gpurun_out/cqt_speed_r2d.jsonl 2>> gpurun_out/floor_r2d.err
done
cat gpurun_out/cqt_floor_r2d.jsonl gpurun_out/cqt_speed_r2d.jsonl; tail -5 gpurun_out/floor_r2d.err
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_r2d.log; tail -8 gpurun_out/pytest_r2d.log
python bench.py --extractor cqt --clips 4096 --steps 2 --warmup 2 --no-cpu --e2e-clips 1024 > gpurun_out/plain_cqt_r2d.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_cqt_r2d.csv \
    python bench.py --extractor cqt --clips 4096 --steps 2 --warmup 2 --no-cpu --e2e-clips 1024 > gpurun_out/ncu_l_cqt_r2d.log 2>&1
tail -3 gpurun_out/ncu_l_cqt_r2d.log

#!/bin/bash
# round 2, GPU run A: regression tests, numerics A/B (mfcc / cqt variants), default bench with extras
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi_r2a.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_r2a.log; tail -5 gpurun_out/pytest_r2a.log
V=audio_edge_ml_pipeline_b200/build
: > gpurun_out/mfcc_floor_r2a.jsonl
for t in base exact log2f dct64 all; do
  B2A_LIBRARY=$PWD/$V/libb2a_$t.so python tools/mfcc_floor.py 2025 $t >> gpurun_out/mfcc_floor_r2a.jsonl 2>> gpurun_out/floor_r2a.err
done
cat gpurun_out/mfcc_floor_r2a.jsonl
: > gpurun_out/cqt_floor_r2a.jsonl
for t in base exact dec64 dec64x; do
  B2A_LIBRARY=$PWD/$V/libb2a_$t.so python tools/cqt_floor.py 405 $t >> gpurun_out/cqt_floor_r2a.jsonl 2>> gpurun_out/floor_r2a.err
done
cat gpurun_out/cqt_floor_r2a.jsonl
tail -5 gpurun_out/floor_r2a.err
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -c 6000 gpurun_out/bench_r2a.json; tail -5 gpurun_out/bench_r2a.err

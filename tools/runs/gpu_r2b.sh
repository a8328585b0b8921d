#!/bin/bash
# round 2, GPU run B: numerics A/B (mfcc / cqt variants) + new front-end tests
mkdir -p gpurun_out
V=audio_edge_ml_pipeline_b200/build
: > gpurun_out/mfcc_floor_r2b.jsonl; : > gpurun_out/floor_r2b.err
for t in base exact log2f dct64 all; do
  B2A_LIBRARY=$PWD/$V/libb2a_$t.so python tools/mfcc_floor.py 2025 $t >> gpurun_out/mfcc_floor_r2b.jsonl 2>> gpurun_out/floor_r2b.err
done
cat gpurun_out/mfcc_floor_r2b.jsonl
: > gpurun_out/cqt_floor_r2b.jsonl
for t in base exact dec64 dec64x; do
  B2A_LIBRARY=$PWD/$V/libb2a_$t.so python tools/cqt_floor.py 405 $t >> gpurun_out/cqt_floor_r2b.jsonl 2>> gpurun_out/floor_r2b.err
done
cat gpurun_out/cqt_floor_r2b.jsonl
tail -5 gpurun_out/floor_r2b.err
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_r2b.log; tail -8 gpurun_out/pytest_r2b.log

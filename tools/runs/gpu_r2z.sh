#!/bin/bash
# round 2, GPU run Z: augmentation tests incl. the vocoder steps and the staged default chain; full-chain Stage 1b + 2 timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_augment.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_r2z.log; tail -8 gpurun_out/pytest_r2z.log | cut -c1-300
timeout 600 python bench_stage2.py --device-augmented --full-chain > gpurun_out/stage2_fullchain_r2z.json 2> gpurun_out/stage2_fullchain_r2z.err; cut -c1-800 gpurun_out/stage2_fullchain_r2z.json; tail -3 gpurun_out/stage2_fullchain_r2z.err

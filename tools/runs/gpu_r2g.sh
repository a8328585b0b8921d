#!/bin/bash
# round 2, GPU run G: n_fft 1024 warp-specialised kernel (parity tests + sweep), CQT with slimmer staging
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_r2g.log; tail -12 gpurun_out/pytest_r2g.log
timeout 300 python tools/sweep.py > gpurun_out/sweep_r2g.jsonl 2> gpurun_out/sweep_r2g.err; cat gpurun_out/sweep_r2g.jsonl; tail -3 gpurun_out/sweep_r2g.err
timeout 300 python tools/cqt_bench.py 8192 > gpurun_out/cqt_speed_r2g.jsonl 2>> gpurun_out/sweep_r2g.err; cat gpurun_out/cqt_speed_r2g.jsonl

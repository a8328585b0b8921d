#!/bin/bash
# round 2, GPU run AH: HEAD — GPU suite, smoke, config 5 on one GPU
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_r2ah.log; tail -4 gpurun_out/pytest_r2ah.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench_stage2.py --devices 0 > gpurun_out/stage2_r2ah.json 2> gpurun_out/stage2_r2ah.err; cut -c1-520 gpurun_out/stage2_r2ah.json

#!/bin/bash
# round 2, GPU run U: full GPU suite, smoke, bench with extras (classical included)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_r2u.log; tail -6 gpurun_out/pytest_r2u.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2u.json 2> gpurun_out/bench_r2u.err; tail -c 400 gpurun_out/bench_r2u.json; tail -3 gpurun_out/bench_r2u.err

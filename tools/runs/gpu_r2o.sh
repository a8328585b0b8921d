#!/bin/bash
# round 2, GPU run O: full GPU test suite, smoke, final bench, sweep, 1024-kernel profile (mfcc defaults)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_r2o.log; tail -4 gpurun_out/pytest_r2o.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2o.json 2> gpurun_out/bench_r2o.err; tail -c 300 gpurun_out/bench_r2o.json; tail -3 gpurun_out/bench_r2o.err
timeout 300 python tools/sweep.py 2> gpurun_out/sweep_r2o.err > gpurun_out/sweep_r2o.jsonl; cut -c1-200 gpurun_out/sweep_r2o.jsonl
timeout 90 python tools/prof_1024.py mfcc > gpurun_out/plain_1024m_r2o.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:logmel1024 -s 2 -c 1 -f -o gpurun_out/prof_1024m_r2o python tools/prof_1024.py mfcc > gpurun_out/ncu_1024m_r2o.log 2>&1
tail -1 gpurun_out/ncu_1024m_r2o.log

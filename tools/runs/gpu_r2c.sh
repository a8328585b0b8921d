#!/bin/bash
# round 2, GPU run C: time-domain CQT bank kernel — parity per octave, tests, throughput
mkdir -p gpurun_out
python tools/cqt_floor.py 405 bank > gpurun_out/cqt_floor_r2c.jsonl 2> gpurun_out/floor_r2c.err; cat gpurun_out/cqt_floor_r2c.jsonl; tail -3 gpurun_out/floor_r2c.err
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_r2c.log; tail -8 gpurun_out/pytest_r2c.log
python bench.py --extractor cqt --clips 16384 --steps 5 --warmup 3 --no-cpu --e2e-clips 2048 > gpurun_out/bench_cqt_r2c.json 2> gpurun_out/bench_cqt_r2c.err; tail -c 1500 gpurun_out/bench_cqt_r2c.json; tail -3 gpurun_out/bench_cqt_r2c.err
python bench.py --extractor cqt --clips 4096 --steps 2 --warmup 2 --no-cpu --e2e-clips 1024 > gpurun_out/plain_cqt_r2c.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_cqt_r2c.csv \
    python bench.py --extractor cqt --clips 4096 --steps 2 --warmup 2 --no-cpu --e2e-clips 1024 > gpurun_out/ncu_l_cqt_r2c.log 2>&1
tail -3 gpurun_out/ncu_l_cqt_r2c.log

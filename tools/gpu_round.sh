#!/bin/bash
# usage: bash gpu_round.sh <tag>   — tests + bench + ncu launch list + one full capture
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi_$TAG.txt
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log; tail -5 gpurun_out/pytest_$TAG.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 3000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
python bench.py --steps 3 --warmup 3 --clips 20000 --e2e-clips 2048 --no-cpu > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --clips 20000 --e2e-clips 2048 --no-cpu > gpurun_out/ncu_l_$TAG.log 2>&1
python bench.py --steps 3 --warmup 3 --clips 20000 --e2e-clips 2048 --no-cpu > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:logmel512_kernel -s 3 -c 2 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 3 --warmup 3 --clips 20000 --e2e-clips 2048 --no-cpu > gpurun_out/ncu_f_$TAG.log 2>&1
ls -la gpurun_out/

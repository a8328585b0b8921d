#!/bin/bash
# round 2, GPU run AA: validation of HEAD — full GPU suite, smoke, bench (with extras), sweep, full-size parity incl. the
# reference-default shapes, launch list + full captures of the headline / 1024 / classical kernels, full-chain Stage 1b timing
TAG=r2aa
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi_$TAG.txt
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log; tail -5 gpurun_out/pytest_$TAG.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 300 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 300 python tools/sweep.py 2> gpurun_out/sweep_$TAG.err > gpurun_out/sweep_$TAG.jsonl; cut -c1-160 gpurun_out/sweep_$TAG.jsonl
timeout 900 python tools/full_parity.py 2025 defaults > gpurun_out/full_parity_$TAG.jsonl 2> gpurun_out/full_parity_$TAG.err; cut -c1-260 gpurun_out/full_parity_$TAG.jsonl; tail -2 gpurun_out/full_parity_$TAG.err
python bench.py --steps 3 --warmup 3 --clips 20000 --e2e-clips 2048 --no-cpu --no-extra > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --clips 20000 --e2e-clips 2048 --no-cpu --no-extra > gpurun_out/ncu_l_$TAG.log 2>&1
python bench.py --steps 3 --warmup 3 --clips 20000 --e2e-clips 2048 --no-cpu --no-extra > gpurun_out/plain2_$TAG.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:logmel512_kernel -s 3 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 3 --warmup 3 --clips 20000 --e2e-clips 2048 --no-cpu --no-extra > gpurun_out/ncu_f_$TAG.log 2>&1
timeout 90 python tools/prof_1024.py mfcc 2>&1 | tail -1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:logmel1024 -s 2 -c 1 -f -o gpurun_out/prof_1024m_$TAG python tools/prof_1024.py mfcc > gpurun_out/ncu_1024m_$TAG.log 2>&1
timeout 90 python tools/prof_1024.py 2>&1 | tail -1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:logmel1024 -s 2 -c 1 -f -o gpurun_out/prof_1024_$TAG python tools/prof_1024.py > gpurun_out/ncu_1024_$TAG.log 2>&1
timeout 120 python tools/prof_classical.py 2>&1 | tail -1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:classical -s 1 -c 1 -f -o gpurun_out/prof_cls_$TAG python tools/prof_classical.py > gpurun_out/ncu_cls_$TAG.log 2>&1
timeout 600 python bench_stage2.py --device-augmented --full-chain > gpurun_out/stage2_fullchain_$TAG.json 2> gpurun_out/stage2_fullchain_$TAG.err; cut -c1-500 gpurun_out/stage2_fullchain_$TAG.json
ls gpurun_out/*$TAG* | wc -l

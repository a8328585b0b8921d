#!/bin/bash
# round 2, GPU run M: numbers and profiles of the round's final state
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2m.json 2> gpurun_out/bench_r2m.err; tail -c 600 gpurun_out/bench_r2m.json; tail -3 gpurun_out/bench_r2m.err
python bench_stage2.py --devices 0 --file-rate 44100 --repeat 2 > gpurun_out/stage2_44k_r2m.json 2> gpurun_out/stage2_44k_r2m.err; cut -c1-600 gpurun_out/stage2_44k_r2m.json; tail -2 gpurun_out/stage2_44k_r2m.err
python bench_stage2.py --devices 0 --device-augmented > gpurun_out/stage2_aug_r2m.json 2> gpurun_out/stage2_aug_r2m.err; cut -c1-700 gpurun_out/stage2_aug_r2m.json; tail -2 gpurun_out/stage2_aug_r2m.err
python bench.py --steps 2 --warmup 3 --clips 20000 --e2e-clips 2048 --no-cpu --no-extra > gpurun_out/plain_r2m.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2m.csv \
    python bench.py --steps 2 --warmup 3 --clips 20000 --e2e-clips 2048 --no-cpu --no-extra > gpurun_out/ncu_l_r2m.log 2>&1
python tools/cqt_bench.py 2048 > gpurun_out/plain_cqt_r2m.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cqt_ -s 11 -c 8 -f -o gpurun_out/prof_cqt_r2m python tools/cqt_bench.py 2048 > gpurun_out/ncu_cqt_r2m.log 2>&1
tail -1 gpurun_out/ncu_cqt_r2m.log
timeout 90 python tools/prof_1024.py mfcc > gpurun_out/plain_1024m_r2m.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:logmel1024 -s 2 -c 1 -f -o gpurun_out/prof_1024m_r2m python tools/prof_1024.py mfcc > gpurun_out/ncu_1024m_r2m.log 2>&1
tail -1 gpurun_out/ncu_1024m_r2m.log

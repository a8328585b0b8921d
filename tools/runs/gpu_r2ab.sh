#!/bin/bash
# round 2, GPU run AB (N GPUs): the bench under torchrun exactly as the driver launches it, both arms
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_r2ab.json 2> gpurun_out/bench_n${N}_r2ab.err; tail -c 1500 gpurun_out/bench_n${N}_r2ab.json; tail -3 gpurun_out/bench_n${N}_r2ab.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n${N}_r2ab.json 2>> gpurun_out/bench_n${N}_r2ab.err; tail -c 600 gpurun_out/bench_ref_n${N}_r2ab.json

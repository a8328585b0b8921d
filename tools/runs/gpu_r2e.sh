#!/bin/bash
# round 2, GPU run E (8-GPU box): e2e scaling against the copy-only ceiling at N = 1, 2, 4, 8; config 5 at 1 and 8 GPUs
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_r2e.txt 2>&1
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" > gpurun_out/lscpu_r2e.txt 2>&1
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_r2e_n$n.json 2> gpurun_out/bench_r2e_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $n --steps 10 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_r2e_n$n.json 2> gpurun_out/bench_r2e_n$n.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_r2e_n$n.json").read().strip().splitlines()[-1])
    e=d["e2e"]; print("N=$n value %.3fM e2e %.3fM copy_ceiling %.3fM frac %.3f h2d %.1f GB/s" % (d["value"]/1e6, e["value"]/1e6, e["copy_ceiling"]["value"]/1e6, e["frac_of_copy_ceiling"], e["copy_ceiling"]["h2d_GBps_total"]))
except Exception as ex: print("N=$n failed", ex)
PY
done
python bench_stage2.py --devices 0 > gpurun_out/stage2_r2e_1gpu.json 2> gpurun_out/stage2_r2e_1gpu.err; cat gpurun_out/stage2_r2e_1gpu.json | cut -c1-700
python bench_stage2.py --devices all > gpurun_out/stage2_r2e_8gpu.json 2> gpurun_out/stage2_r2e_8gpu.err; cat gpurun_out/stage2_r2e_8gpu.json | cut -c1-700
tail -3 gpurun_out/stage2_r2e_8gpu.err

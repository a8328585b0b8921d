#!/bin/bash
# round 2, GPU run Q: 1024 kernel (uniform combination step, PRMT widening, padded band pairs); mbarrier suspend hint A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_extractors.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_r2q.log; tail -3 gpurun_out/pytest_r2q.log
for lib in "" audio_edge_ml_pipeline_b200/build/libb2a_hint.so; do
  echo "== lib=${lib:-product}"
  B2A_LIBRARY=$lib timeout 90 python tools/prof_1024.py mfcc 2>&1 | tail -1
  B2A_LIBRARY=$lib timeout 90 python tools/prof_1024.py 2>&1 | tail -1
  B2A_LIBRARY=$lib timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra --e2e-clips 2048 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('headline', d['value'], d['ms_per_step'])"
  B2A_LIBRARY=$lib timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra --e2e-clips 2048 --extractor mfcc --clips 50000 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mfcc', d['value'], d['ms_per_step'])"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:logmel1024 -s 2 -c 1 -f -o gpurun_out/prof_1024_r2q python tools/prof_1024.py > gpurun_out/ncu_1024_r2q.log 2>&1
tail -1 gpurun_out/ncu_1024_r2q.log

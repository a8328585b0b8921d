#!/bin/bash
# round 2, GPU run AJ: extract_dataset with page-locked output staging + helper-thread copies — dataset tests, config 5
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_extractors.py tests/test_gpu_classical.py tests/test_resample.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_r2aj.log; tail -3 gpurun_out/pytest_r2aj.log | cut -c1-300
for st in 1 0; do
  B2A_OUT_STAGING=$st timeout 300 python bench_stage2.py --devices 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('staging=$st', round(d['value']), d['seconds'])"
done

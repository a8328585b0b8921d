#!/bin/bash
# round 2, GPU run AE: headline-shape mfcc with the DCT deferred into the next clip's tiles — parity, rate
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_extractors.py -m gpu -x -q -k "mfcc or ragged or smoke or mel" 2>&1 | tail -12 > gpurun_out/pytest_r2ae.log; tail -5 gpurun_out/pytest_r2ae.log | cut -c1-300
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra --e2e-clips 2048 --extractor mfcc --clips 50000 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mfcc', d['value'], d['ms_per_step'])"
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra --e2e-clips 2048 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mel', d['value'], d['ms_per_step'])"
timeout 600 python tools/full_parity.py 2025 2>/dev/null | grep mfcc | cut -c1-400

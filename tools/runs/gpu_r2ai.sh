#!/bin/bash
# round 2, GPU run AI: mbarrier suspend-hint values A/B (0 = no hint, 20 us, 100 us = product, 1 ms)
for lib in audio_edge_ml_pipeline_b200/build/libb2a_h0.so audio_edge_ml_pipeline_b200/build/libb2a_h20k.so "" audio_edge_ml_pipeline_b200/build/libb2a_h1m.so; do
  echo "== lib=${lib:-product(100k)}"
  for rep in 1 2; do
  B2A_LIBRARY=$lib timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra --e2e-clips 2048 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('headline', round(d['value']))"
  done
  B2A_LIBRARY=$lib timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra --e2e-clips 2048 --extractor mfcc --clips 50000 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mfcc', round(d['value']))"
  B2A_LIBRARY=$lib timeout 90 python tools/prof_1024.py mfcc 2>&1 | tail -1
done

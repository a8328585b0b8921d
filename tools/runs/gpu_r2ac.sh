#!/bin/bash
# round 2, GPU run AC: GPU suite + smoke after the last host-side changes (length buckets, ragged single-file path)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_r2ac.log; tail -5 gpurun_out/pytest_r2ac.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2

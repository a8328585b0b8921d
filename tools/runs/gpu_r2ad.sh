#!/bin/bash
# round 2, GPU run AD: GPU suite + smoke from a from-scratch build (new classical case: n_fft 2048, odd hop)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_r2ad.log; tail -6 gpurun_out/pytest_r2ad.log | cut -c1-400
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2

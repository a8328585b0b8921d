#!/bin/bash
# round 2, GPU run Y: first run of the phase-vocoder effects against their oracle
mkdir -p gpurun_out
timeout 600 python tools/effects_check.py 12 > gpurun_out/effects_check_r2y.jsonl 2> gpurun_out/effects_check_r2y.err; tail -5 gpurun_out/effects_check_r2y.err; cut -c1-900 gpurun_out/effects_check_r2y.jsonl

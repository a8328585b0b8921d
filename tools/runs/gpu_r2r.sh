#!/bin/bash
# round 2, GPU run R: first run of the audio_classical kernel against its oracle
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck python tools/classical_check.py 3 22050 1024 512 1.0 > gpurun_out/cls_sanitize_r2r.log 2>&1; tail -5 gpurun_out/cls_sanitize_r2r.log | cut -c1-300
timeout 300 python tools/classical_check.py 21 > gpurun_out/cls_check_r2r.jsonl 2> gpurun_out/cls_check_r2r.err; tail -3 gpurun_out/cls_check_r2r.err; cut -c1-1500 gpurun_out/cls_check_r2r.jsonl
timeout 300 python tools/classical_check.py 14 16000 512 160 5.0 >> gpurun_out/cls_check_r2r.jsonl 2>> gpurun_out/cls_check_r2r.err; tail -1 gpurun_out/cls_check_r2r.jsonl | cut -c1-600

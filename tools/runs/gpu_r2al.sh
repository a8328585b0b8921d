#!/bin/bash
# round 2, GPU run AL: more parity evidence — classical at 16 kHz / 512 / 160 on 405 clips, vocoder effects on 48 clips
mkdir -p gpurun_out
timeout 900 python tools/classical_check.py 405 16000 512 160 5.0 > gpurun_out/cls_parity16k_r2al.jsonl 2> gpurun_out/cls_parity16k_r2al.err; tail -2 gpurun_out/cls_parity16k_r2al.err; cut -c1-200 gpurun_out/cls_parity16k_r2al.jsonl
timeout 900 python tools/effects_check.py 48 > gpurun_out/effects_check48_r2al.jsonl 2> gpurun_out/effects_check48_r2al.err; tail -2 gpurun_out/effects_check48_r2al.err; cut -c1-260 gpurun_out/effects_check48_r2al.jsonl

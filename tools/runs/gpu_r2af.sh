#!/bin/bash
# round 2, GPU run AF: breakdown of config 5's extract time (decode vs engine call per window)
timeout 300 python tools/stage2_breakdown.py 2>&1 | tail -4

#!/bin/bash
# round 2, GPU run AG: Python-level profile of one config-5 run
timeout 300 python tools/stage2_profile.py 2>&1 | tail -60

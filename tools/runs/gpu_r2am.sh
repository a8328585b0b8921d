#!/bin/bash
# round 2, GPU run AM: classical tests after restating the contrast bound
timeout 600 python -m pytest tests/test_gpu_classical.py -m gpu -q 2>&1 | tail -3 | cut -c1-300

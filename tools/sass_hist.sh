#!/bin/bash
# usage: tools/sass_hist.sh <regex on mangled kernel name> [lib]   — static opcode histogram of one kernel
LIB=${2:-audio_edge_ml_pipeline_b200/libb2a.so}
cuobjdump -sass "$LIB" | awk -v f="$1" '/Function :/{p=($0 ~ f)} p' | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | awk '{ if ($2 ~ /^@/) print $3; else print $2 }' | sed 's/\..*//; s/;//' | sort | uniq -c | sort -rn | head -${3:-30}

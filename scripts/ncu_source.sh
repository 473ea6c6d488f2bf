#!/bin/bash
# Per-instruction profile of one kernel (run on the GPU box): ncu --set full --import-source on, SASS page exported as
# CSV (gzip) into gpurun_out/; summarise here with scripts/ncu_source_blocks.py.
#   bash scripts/ncu_source.sh <kernel regex> <out name> <python script> [args...]
cd "$(dirname "$0")/.."
K=$1; NAME=$2; shift 2
timeout 400 ncu --set full --import-source on --clock-control none -k regex:$K -c 1 -o /tmp/$NAME -f python "$@" > /dev/null 2>&1
ncu -i /tmp/$NAME.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/$NAME.csv.gz
ls -la gpurun_out/$NAME.csv.gz

#!/bin/bash
# Per-op ncu captures of round 2 (run under gpurun).  Reports stay in /tmp on the box (they exceed gpurun's 64 MiB
# return limit); only the CSV exports come back in gpurun_out/.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# (1) every tcgen05 launch of one frame, full set (DRAM traffic, tensor pipe, issue, occupancy)
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:mlp_tc_kernel -o /tmp/r02_mlp python scripts/ncu_ops.py frame > gpurun_out/r02_ncu_mlp.log 2>&1
ncu -i /tmp/r02_mlp.ncu-rep --page raw --csv > gpurun_out/r02_ncu_mlp_raw.csv 2>> gpurun_out/r02_ncu_mlp.log
# (2) every other named kernel (geometry, glue, pose fit, eval, crop, cluster FPS, fused query+group, adjoints)
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:'^(?!.*mlp_tc_kernel).*$' -o /tmp/r02_ops python scripts/ncu_ops.py all > gpurun_out/r02_ncu_ops.log 2>&1
ncu -i /tmp/r02_ops.ncu-rep --page raw --csv > gpurun_out/r02_ncu_ops_raw.csv 2>> gpurun_out/r02_ncu_ops.log
# (3) launch list of one eager frame (cold-cache, serialised: shares only)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_frame.csv python scripts/ncu_ops.py frame > /dev/null 2>&1
ls -la gpurun_out/r02_ncu_* gpurun_out/r02_launches_frame.csv

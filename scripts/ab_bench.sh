#!/bin/bash
# A/B of the round-2 frame-level knobs on one box (box-to-box variance is +-3 %, so compare within one call)
cd "$(dirname "$0")/.."
for cfg in "1 1" "0 1" "1 0" "0 0"; do
  set -- $cfg
  echo "CAPTRA_TC_SMALL_WIDE=$1 CAPTRA_TWO_STREAM=$2"
  CAPTRA_TC_SMALL_WIDE=$1 CAPTRA_TWO_STREAM=$2 timeout 200 python bench.py --steps 30 --warmup 3 --no-extras --no-cpu-baseline ${AB_ARGS} 2>/dev/null \
    | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  frames/s %.1f  ms/step %.3f  e2e %.1f  launches/step %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']/d['steps'])); [print('    %-80s %7.1f us x%.0f' % (k['tag'], k['avg_us'], k['launches_per_step'])) for k in d['kernels'][:14]]"
done

#!/bin/bash
# Round-2 final measurement on ONE box: GPU tests, the four bench workloads, smoke, and the ncu captures of the
# tensor-core launches + launch list (CSV exports only; the .ncu-rep stays in /tmp on the box).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -4
timeout 400 python bench.py --steps 20 --warmup 3 --dump-kernels gpurun_out/r02_kernels_f.json > gpurun_out/r02_bench_f.json 2> gpurun_out/r02_bench_f.err
timeout 200 python bench.py --workload cfg4 --steps 10 --no-cpu-baseline --no-extras > gpurun_out/r02_bench_f_cfg4.json 2> gpurun_out/r02_bench_f_cfg4.err
timeout 300 python bench.py --workload cfg3 --steps 20 --no-extras > gpurun_out/r02_bench_f_cfg3.json 2> gpurun_out/r02_bench_f_cfg3.err
timeout 200 python bench.py --workload cfg5 --steps 5 > gpurun_out/r02_bench_f_cfg5.json 2> gpurun_out/r02_bench_f_cfg5.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:mlp_tc_kernel -o /tmp/r02_mlp python scripts/ncu_ops.py frame > gpurun_out/r02_ncu_mlp.log 2>&1
ncu -i /tmp/r02_mlp.ncu-rep --page raw --csv > gpurun_out/r02_ncu_mlp_raw.csv 2>> gpurun_out/r02_ncu_mlp.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_frame.csv python scripts/ncu_ops.py frame > /dev/null 2>&1
ls -la gpurun_out/r02_bench_f*.json gpurun_out/r02_ncu_mlp_raw.csv gpurun_out/r02_launches_frame.csv

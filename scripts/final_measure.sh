#!/bin/bash
# Round-end measurement on one B200 (run under gpurun): parity tests, the bench line, the kernel probes,
# the ncu launch list of the bench command and one full capture of the dominant launch, cfg3 and cfg5.
# Everything lands in gpurun_out/; the summaries to keep are copied into profiles/ afterwards.
# The knock-out probes need the probes build of the library (made HERE, before gpurun, it travels with the snapshot):
#   CAPTRA_EXTRA_NVCC_FLAGS=-DCAPTRA_TC_PROBES CAPTRA_LIB_OUT=captra_b200/libcaptra_ops_probes.so python -m captra_b200.build
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests -m gpu -q --timeout 60 > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 300 python bench.py --dump-kernels $O/kernels_v12.json > $O/bench_v12.json 2> $O/bench_v12.err; tail -2 $O/bench_v12.err; cut -c1-160 $O/bench_v12.json
CAPTRA_LIB_PATH=$PWD/captra_b200/libcaptra_ops_probes.so PROBE_DBG=0,1,2,8,64,128 PROBE_STAMPS=1 timeout 200 python scripts/tc_probe.py > $O/tc_probe_v12.txt 2>&1; tail -3 $O/tc_probe_v12.txt | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_v12.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $O/launches_bench.log 2>&1; tail -1 $O/launches_bench.log | cut -c1-120
PROBE_CASES=5 PROBE_DENSE=0 PROBE_N=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 2 -c 1 \
    -o $O/v12_sa2pre_k128 python scripts/tc_probe.py > $O/ncu_full.log 2>&1; tail -2 $O/ncu_full.log
PROBE_CASES=0 PROBE_DENSE=0 PROBE_N=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 2 -c 1 \
    -o $O/v12_sa1_k128 python scripts/tc_probe.py > $O/ncu_full2.log 2>&1; tail -2 $O/ncu_full2.log
timeout 200 python bench.py --workload cfg3 --no-cpu-baseline > $O/bench_v12_cfg3.json 2> $O/bench_v12_cfg3.err; cut -c1-160 $O/bench_v12_cfg3.json
STRESS_CLOUD=surface timeout 200 python scripts/stress_cfg5.py > $O/stress_cfg5_surface.jsonl 2>&1; tail -1 $O/stress_cfg5_surface.jsonl
STRESS_CLOUD=uniform timeout 200 python scripts/stress_cfg5.py > $O/stress_cfg5_uniform.jsonl 2>&1; tail -1 $O/stress_cfg5_uniform.jsonl

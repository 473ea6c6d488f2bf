cd /root/repo
timeout 700 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -4
timeout 400 python bench.py --steps 20 --warmup 3 --dump-kernels gpurun_out/r02_kernels_i.json > gpurun_out/r02_bench_i.json 2> gpurun_out/r02_bench_i.err
timeout 200 python bench.py --workload cfg4 --steps 10 --no-cpu-baseline --no-extras > gpurun_out/r02_bench_i_cfg4.json 2> gpurun_out/r02_bench_i_cfg4.err
timeout 300 python bench.py --workload cfg3 --steps 20 --no-extras > gpurun_out/r02_bench_i_cfg3.json 2> gpurun_out/r02_bench_i_cfg3.err
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_i_reference.json 2> gpurun_out/r02_bench_i_reference.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for f in r02_bench_i r02_bench_i_cfg4 r02_bench_i_cfg3 r02_bench_i_reference; do python -c "
import json
d=json.loads([l for l in open('gpurun_out/$f.json').read().splitlines() if l.startswith('{')][-1]); print('$f', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1))"; done

timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02c_bench_default.log 2>&1; tail -1 gpurun_out/r02c_bench_default.log | cut -c1-400
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02c_bench_reference.log 2>&1; tail -1 gpurun_out/r02c_bench_reference.log | cut -c1-300
for c in 3 4 5; do timeout 600 python bench.py --config $c --steps 3 --no-cpu > gpurun_out/r02c_bench_c$c.log 2>&1; tail -1 gpurun_out/r02c_bench_c$c.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(d['config']['workload'][:40], d['value'], d['ms_per_step'], d['stage_ms'], d['roofline']['achieved'], d['roofline']['frac_of_nominal'], d['parity_check']['ok'], d['e2e']['value'])"; done
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2

python -m pytest tests/test_gpu_ransac.py tests/test_gpu_api.py -x -q 2>&1 | tail -8
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --pairs 2000"
python bench.py --steps 5 --no-cpu > gpurun_out/r02g_bench.log 2>&1; tail -1 gpurun_out/r02g_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['stage_ms'], d['config']['valid_pairs_last_step'])"
$CMD > gpurun_out/r02g_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ransac|static|filter_m|match_|build_' -s 33 -c 11 --csv --log-file gpurun_out/r02g_launches.csv $CMD > gpurun_out/r02g_ncu1.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02g_launches.csv')) if len(r)>5]
h=rows[0]; ni=h.index('Kernel Name'); vi=h.index('Metric Value'); 
for r in rows[1:]:
    print(r[ni][:50], r[vi])
PY

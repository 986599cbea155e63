CMD="python bench.py --steps 2 --warmup 3 --no-cpu --pairs 2000"
python scripts/ransac_stats.py 512 > gpurun_out/r02f_stats.txt 2>&1
$CMD > gpurun_out/r02f_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ransac|static|filter_m|match_|build_|prod_|fill_|fixed_' -s 51 -c 34 --csv --log-file gpurun_out/r02f_launches.csv $CMD > gpurun_out/r02f_ncu1.log 2>&1
cat gpurun_out/r02f_stats.txt
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02f_launches.csv')) if len(r)>5]
h=rows[0]; ni=h.index('Kernel Name'); vi=h.index('Metric Value'); 
for r in rows[1:]:
    print(r[ni][:50], r[vi])
PY

CMD="python bench.py --steps 2 --warmup 3 --no-cpu --pairs 2000"
$CMD > gpurun_out/r02h_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'ransac_score' -s 6 -c 1 -o gpurun_out/r02h_prof_score $CMD > gpurun_out/r02h_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'ransac_refit' -s 6 -c 1 -o gpurun_out/r02h_prof_refit $CMD > gpurun_out/r02h_ncu3.log 2>&1

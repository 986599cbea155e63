set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --pairs 2000"
$CMD > gpurun_out/r02d_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'evz' -s 60 -c 60 --csv --log-file gpurun_out/r02d_launches.csv $CMD > gpurun_out/r02d_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'ransac_score' -s 6 -c 2 -o gpurun_out/r02d_prof_score $CMD > gpurun_out/r02d_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'ransac_refit' -s 6 -c 2 -o gpurun_out/r02d_prof_refit $CMD > gpurun_out/r02d_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'static_filter|filter_matches|match_fixup|match_prepare' -s 12 -c 4 -o gpurun_out/r02d_prof_small $CMD > gpurun_out/r02d_ncu4.log 2>&1
tail -2 gpurun_out/r02d_plain.log
ls -la gpurun_out/

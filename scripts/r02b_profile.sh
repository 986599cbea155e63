# Round-2 evidence, second pass (after the bulk staging / model cache / refit / filter changes): every command runs plain
# first (exit 0), then under ncu (B200_PROFILING.md recipe).
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --pairs 2000"
$CMD > gpurun_out/r02b_plain_bench2000.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ransac|static|filter_m|match_|build_|prod_|fill_|fixed_' -s 51 -c 34 --csv --log-file gpurun_out/r02b_launches_bench_pairs2000.csv $CMD > gpurun_out/r02b_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'ransac_score' -s 6 -c 2 -o gpurun_out/r02b_prof_score $CMD > gpurun_out/r02b_ncu_score.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'ransac_refit' -s 6 -c 2 -o gpurun_out/r02b_prof_refit $CMD > gpurun_out/r02b_ncu_refit.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'filter_matches|static_filter|match_fixup' -s 9 -c 3 -o gpurun_out/r02b_prof_glue $CMD > gpurun_out/r02b_ncu_glue.log 2>&1
tail -2 gpurun_out/r02b_ncu_glue.log | cut -c1-200

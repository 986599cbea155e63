set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
./scripts/microbench/mb2 all > gpurun_out/r02a_mb2.txt 2>&1
for dbg in 0 1 2; do
  python scripts/bench_match.py 2000 2048 0,7 $dbg > gpurun_out/r02a_bm_2048_dbg$dbg.txt 2>&1
  python scripts/bench_match.py 500 8192 0,7 $dbg > gpurun_out/r02a_bm_8192_dbg$dbg.txt 2>&1
done
tail -n 5 gpurun_out/r02a_bm_*.txt
cat gpurun_out/r02a_mb2.txt

python -m pytest tests/test_gpu_chain.py -x -q 2>&1 | tail -5
python bench.py --steps 3 --pairs 1500 > gpurun_out/r02i_bench_small.log 2>&1; tail -3 gpurun_out/r02i_bench_small.log | cut -c1-3000
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02i_ref.log 2>&1; tail -1 gpurun_out/r02i_ref.log | cut -c1-1200

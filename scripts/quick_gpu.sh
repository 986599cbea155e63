#!/bin/bash
# GPU check used while iterating: the GPU test-suite, then a short default bench line (stage split, parity check).
tag=${1:-quick}
timeout 240 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/${tag}_bench.log 2>&1
tail -1 gpurun_out/${tag}_bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stage_ms'].items()}, round(d['roofline']['kernel_ms'],3), d['parity_check']['ok'], round(d['e2e']['value']), d['clocks'])"

#!/bin/bash
# A/B of match-kernel experiment builds (build_variants/libevz_*.so) against the shipped library, same box, same run.
set -u
mkdir -p gpurun_out
OUT=gpurun_out/${1:-r02x_exp}.txt
: > $OUT
cp evenvizion_b200/libevz.so /tmp/libevz_base.so
run() {
  echo "== $1" >> $OUT
  python scripts/bench_match.py 2000 2048 0 >> $OUT 2>&1
  python scripts/bench_match.py 500 8192 0 >> $OUT 2>&1
}
run base
for f in build_variants/libevz_*.so; do
  cp $f evenvizion_b200/libevz.so
  run $f
done
cp /tmp/libevz_base.so evenvizion_b200/libevz.so
run base_again
cat $OUT

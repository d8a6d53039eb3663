#!/bin/bash
# same box, same run: every variant twice, interleaved
for rep in 1 2; do
for v in mixed_b1 mixed_b8 split_b1 split_b8; do
  RAPID_B200_LIB=$PWD/rapid_b200/variants/lib_$v.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-carried 2>&1 >/dev/null | grep "per step" | sed "s/^/c5 $v: /" | cut -c1-200
  RAPID_B200_LIB=$PWD/rapid_b200/variants/lib_$v.so python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "per step" | sed "s/^/c4 $v: /" | cut -c1-200
done
done

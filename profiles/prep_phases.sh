#!/bin/bash
# k_prepare phase timings (RAPID_B200_PREP_STAMPS prints them at every host collect) for C5 and C4, for a few grid sizes.
for g in 0 16 32 64 128; do
  for w in c5 c4; do
    echo "== workload $w grid override $g"
    if [ "$g" = "0" ]; then
      RAPID_B200_PREP_STAMPS=1 python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline --no-carried 2>&1 >/dev/null | grep "k_prepare phases" | tail -3
    else
      RAPID_B200_PREP_STAMPS=1 RAPID_B200_PREP_GRID=$g python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline --no-carried 2>&1 >/dev/null | grep "k_prepare phases" | tail -3
    fi
  done
done

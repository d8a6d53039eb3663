#!/bin/bash
# same box, same run: every variant twice, interleaved (run through gpurun from the repo root after profiles/ab_build.sh)
mkdir -p gpurun_out
for rep in 1 2; do
for v in memo0 memo1; do
  RAPID_B200_LIB=$PWD/rapid_b200/ab/lib_$v.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_c5_${v}_$rep.json 2> gpurun_out/ab_c5_${v}_$rep.err
  RAPID_B200_LIB=$PWD/rapid_b200/ab/lib_$v.so python bench.py --workload c4 --stream batches --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_c4b_${v}_$rep.json 2> gpurun_out/ab_c4b_${v}_$rep.err
  python - <<P
import json
for w in ("c5", "c4b"):
    try:
        d = json.loads(open("gpurun_out/ab_%s_${v}_$rep.json" % w).read().strip().splitlines()[-1])
        rc = d.get("roofline_carried") or {}
        print("%s ${v} rep$rep: ms/step %.3f  dominant %.3f ms frac %.3f  carried %s ms frac %s" % (
            w, d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], rc.get("kernel_ms"), rc.get("frac")))
    except Exception as e:
        print(w, "${v}", "failed", e)
P
done
done

#!/bin/bash
# Regenerates the round's profile artefacts on ONE B200 (run through gpurun from the repo root):
#   gpurun --timeout 1500 -- 'bash profiles/collect.sh'
# then, back in the build container:  python profiles/summarize.py ...  (see profiles/README in DESIGN.md §6)
set -x
mkdir -p gpurun_out
for w in c5 c4 c3 c2; do
  python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_${w}_1gpu.json 2> gpurun_out/bench_${w}_1gpu.err
done
python bench.py --workload c4 --stream batches --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_batches_1gpu.json 2> gpurun_out/bench_c4_batches_1gpu.err
# launch lists (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c5.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c4.csv \
    python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b.log 2>&1
# the dominant kernels, full sets: fresh path (C5 step), carried path (second half of the split batch), sequence pass (C4)
ncu --set full --clock-control none --import-source on -k regex:k_apply_uniform -s 3 -c 1 -o gpurun_out/prof_c5_fresh \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-carried > gpurun_out/b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_apply_uniform -s 31 -c 1 -o gpurun_out/prof_c5_carried \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_apply_uniform -s 3 -c 1 -o gpurun_out/prof_c4_seq \
    python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_prepare -s 3 -c 1 -o gpurun_out/prof_prepare \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-carried > gpurun_out/b.log 2>&1
python profiles/bench_view.py > gpurun_out/bench_view.json 2>/dev/null
python profiles/bench_classic_paxos.py > gpurun_out/bench_px.json 2>/dev/null
python profiles/bench_fd.py > gpurun_out/bench_fd.json 2>/dev/null
python profiles/bench_wire.py > gpurun_out/bench_wire.json 2>/dev/null
ls -la gpurun_out

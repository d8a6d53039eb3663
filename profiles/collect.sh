#!/bin/bash
# Regenerates the round's profile artefacts on a GPU box (run through gpurun from the repo root):
#   gpurun --timeout 1500 -- 'bash profiles/collect.sh'
# then, back in the build container:  python profiles/summarize.py launches gpurun_out/launches.csv > profiles/rNN_launches_c5_1gpu.txt
#                                     python profiles/summarize.py full gpurun_out/prof_uniform.ncu-rep > profiles/rNN_ncu_full_k_apply_uniform.txt
set -x
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_apply_uniform -s 2 -c 1 -o gpurun_out/prof_uniform \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_prepare -s 2 -c 1 -o gpurun_out/prof_prepare \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b3.log 2>&1
ls -la gpurun_out

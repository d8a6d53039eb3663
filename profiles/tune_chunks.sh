for c in 0 4 6 8 16; do
  if [ "$c" = "0" ]; then python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "per step" | sed "s/^/auto: /"
  else RAPID_B200_CHUNKS=$c python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "per step" | sed "s/^/chunks=$c: /"; fi
done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 >gpurun_out/bench_c5.json | grep "per step"
python -c "
import json; d=json.load(open('gpurun_out/bench_c5.json')); print('c5 frac', d['roofline']['frac'], 'carried', d['roofline_carried']['frac'], d['roofline_carried']['kernel_ms'])"
ncu --set full --clock-control none --import-source on -k regex:k_apply_uniform -s 12 -c 1 -o gpurun_out/prof_c5_carried python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b.log 2>&1
timeout 600 python -m pytest tests/test_gpu_cut_detection.py tests/test_gpu_fast_paxos.py tests/test_gpu_full_scale.py -m gpu -x -q 2>&1 | tail -4

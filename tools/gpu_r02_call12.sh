#!/bin/bash
# round-2 GPU call 12: FULL -m gpu suite as the driver runs it + the N=1 bench line
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/r02_tests12.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_tests12.log; tail -12 gpurun_out/r02_tests12.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench12.json 2> gpurun_out/r02_bench12.err
echo "bench rc=$?"; tail -2 gpurun_out/r02_bench12.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench12.json'))
print({k:d[k] for k in ('value','ms_per_step','phases_ms','gpu_launches')}); print('e2e',d['e2e']['value'])
print('b4', d['batch4_512']['value'], d['batch4_512']['ms_per_call'])
t=d['tiled2048']; print('tiled', t['value'], t['ms_per_image'], t['phases_ms_rank0'])
v=d['v21_1024_b4']; print('v21', v['value'], v['ms_per_batch'], v['phases_ms_rank0'])
r=d['roofline']; print('roof', r['frac'], r['achieved'], r['kernel_ms_per_forward'], r['attention'], r.get('tiled_regime'))
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['sample']); print('gpu torch', d['gpu_torch_baseline']['ms_per_image'])
PY
cp gpurun_out/kernel_census.csv gpurun_out/r02_kernel_census12.csv
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02_bench12_ref.json 2> gpurun_out/r02_bench12_ref.err; echo "ref rc=$?"; head -c 400 gpurun_out/r02_bench12_ref.json

#!/bin/bash
# round-2 GPU call 15: final state -- FULL -m gpu suite as the driver runs it, smoke, the N=1 bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r02_tests15.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_tests15.log; tail -9 gpurun_out/r02_tests15.log
timeout 200 python __graft_entry__.py --smoke > gpurun_out/r02_smoke15.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke15.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench15.json 2> gpurun_out/r02_bench15.err
echo "bench rc=$?"; tail -9 gpurun_out/r02_bench15.err | cut -c1-250; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench15.json'))
print({k:d.get(k) for k in ('value','ms_per_step','phases_ms','gpu_launches','incomplete')}); print('e2e',d['e2e']['value'])
print('b4', d['batch4_512']['value'], d['batch4_512']['ms_per_call'])
t=d['tiled2048']; print('tiled', t['value'], t['ms_per_image'], t['phases_ms_rank0'])
v=d['v21_1024_b4']; print('v21', v['value'], v['ms_per_batch'])
r=d['roofline']; print('roof', r['frac'], r['achieved'], r['kernel_ms_per_forward'], r['attention']['frac'], r['tiled_regime']['gemm_frac'])
print('cpu', d['cpu_baseline']); print('gpu torch', d['gpu_torch_baseline'])
PY
cp gpurun_out/kernel_census.csv gpurun_out/r02_kernel_census15.csv

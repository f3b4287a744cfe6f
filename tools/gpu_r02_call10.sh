#!/bin/bash
# round-2 GPU call 10: FULL -m gpu suite (as the driver runs it), the bench line, smoke, tile-class A/B
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r02_tests10.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_tests10.log; tail -14 gpurun_out/r02_tests10.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r02_smoke10.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke10.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench10.json 2> gpurun_out/r02_bench10.err
echo "bench rc=$?"; tail -2 gpurun_out/r02_bench10.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench10.json'))
print({k:d[k] for k in ('value','ms_per_step','phases_ms','gpu_launches')}); print('e2e',d['e2e']['value'])
print('b4', d['batch4_512'])
t=d['tiled2048']; print('tiled', t['value'], t['ms_per_image'], t['phases_ms_rank0'])
v=d['v21_1024_b4']; print('v21', v['value'], v['ms_per_batch'], v['phases_ms_rank0'])
r=d['roofline']; print('roof', r['frac'], r['kernel_ms_per_forward'], r['attention'], r.get('tiled_regime'))
print('cpu', d['cpu_baseline']); print('gpu torch', d['gpu_torch_baseline'])
PY
cp gpurun_out/kernel_census.csv gpurun_out/r02_kernel_census10.csv
AB=gpurun_out/r02_ab10.jsonl; : > $AB
run() { echo "== $*"; env "${@:2}" timeout 500 python tools/gpu_forward_ab.py $1 ${EXTRA} >> $AB 2>> gpurun_out/r02_ab10.err; tail -1 $AB | cut -c1-300; }
EXTRA="" run base X=1
run wide0 DBIR_GEMM_WIDE=0
run wide0_nopdl DBIR_GEMM_WIDE=0 DBIR_PDL=0
run geglu128 DBIR_GEGLU_TILE=128

#!/bin/bash
# round-2 GPU call 14: batch-invariance fix, where did the bench stall (phase log), then the rest of the suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/r02_tests14.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_tests14.log; tail -4 gpurun_out/r02_tests14.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench14.json 2> gpurun_out/r02_bench14.err
echo "bench rc=$?"; tail -12 gpurun_out/r02_bench14.err | cut -c1-300
[ -s gpurun_out/r02_bench14.json ] && python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench14.json'))
print({k:d[k] for k in ('value','ms_per_step','phases_ms','gpu_launches')}); print('e2e',d['e2e']['value'])
print('b4', d['batch4_512']['value'], d['batch4_512']['ms_per_call'])
t=d['tiled2048']; print('tiled', t['value'], t['ms_per_image'], t['phases_ms_rank0'])
v=d['v21_1024_b4']; print('v21', v['value'], v['ms_per_batch'], v['phases_ms_rank0'])
r=d['roofline']; print('roof', r['frac'], r['achieved'], r['kernel_ms_per_forward'], r['attention'], r.get('tiled_regime'))
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['sample']); print('gpu torch', d['gpu_torch_baseline']['ms_per_image'])
PY
cp gpurun_out/kernel_census.csv gpurun_out/r02_kernel_census14.csv 2>/dev/null
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_engines.py tests/test_gpu_zz_inference_cli.py -m gpu -x -q > gpurun_out/r02_tests14b.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_tests14b.log; tail -4 gpurun_out/r02_tests14b.log

#!/bin/bash
# round-2 GPU call 6: fused GroupNorm finalize+apply, BSRNet, 256-wide GEGLU; bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engines.py tests/test_gpu_multi.py -m gpu -q -s -k "not shapes" > gpurun_out/r02_tests6.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_tests6.log; grep -n "rrdbnet\|FAILED\|Error\|passed\|failed" gpurun_out/r02_tests6.log | tail -12
timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -s -k "bsrnet or small_pipeline_matches or full_config_50_step_psnr" > gpurun_out/r02_tests6b.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_tests6b.log; grep -n "rel-rms\|passed\|failed\|Error\|PSNR" gpurun_out/r02_tests6b.log | tail -12
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench6.json 2> gpurun_out/r02_bench6.err
echo "bench rc=$?"; tail -2 gpurun_out/r02_bench6.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench6.json'))
print({k:d[k] for k in ('value','ms_per_step','phases_ms')}); print('e2e',d['e2e']['value'])
t=d['tiled2048']; print('tiled', t['value'], t['ms_per_image'], t['phases_ms_rank0'])
v=d['v21_1024_b4']; print('v21', v['value'], v['ms_per_batch'], v['phases_ms_rank0'])
r=d['roofline']; print('roof', r['frac'], r['kernel_ms_per_forward'], r['attention'], r.get('tiled_regime'))
PY
cp gpurun_out/kernel_census.csv gpurun_out/r02_kernel_census6.csv
AB=gpurun_out/r02_ab6.jsonl; : > $AB
run() { echo "== $*"; env "${@:2}" timeout 500 python tools/gpu_forward_ab.py $1 ${EXTRA} >> $AB 2>> gpurun_out/r02_ab6.err; tail -1 $AB | cut -c1-300; }
EXTRA="" run fused X=1
run unfused DBIR_GN_FUSED=0
EXTRA="--nb=28" run det28_fused DBIR_DETERMINISTIC=1
EXTRA="--nb=28" run det28_unfused DBIR_DETERMINISTIC=1 DBIR_GN_FUSED=0

#!/bin/bash
# round-2 GPU call 5: norm-kernel changes + new sampler tests, bench with the v2.1 block, large-batch tile experiments
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engines.py tests/test_gpu_multi.py -m gpu -q -s -k "not shapes" > gpurun_out/r02_tests5.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_tests5.log; tail -3 gpurun_out/r02_tests5.log
timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -s -k "edm_dpm or small_pipeline_matches" > gpurun_out/r02_tests5b.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_tests5b.log; grep -n "rel-rms\|passed\|failed\|Error" gpurun_out/r02_tests5b.log | tail -14
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench5.json 2> gpurun_out/r02_bench5.err
echo "bench rc=$?"; tail -2 gpurun_out/r02_bench5.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench5.json'))
print({k:d[k] for k in ('value','ms_per_step','phases_ms')}); print('e2e',d['e2e']['value'])
t=d['tiled2048']; print('tiled', t['value'], t['ms_per_image'], t['phases_ms_rank0'])
v=d['v21_1024_b4']; print('v21', v['value'], v['ms_per_batch'], v['phases_ms_rank0'])
print('roof', d['roofline']['frac'], d['roofline']['kernel_ms_per_forward'], d['roofline']['attention'])
PY
cp gpurun_out/kernel_census.csv gpurun_out/r02_kernel_census5.csv
AB=gpurun_out/r02_ab5.jsonl; : > $AB
run() { echo "== $*"; env "${@:2}" timeout 500 python tools/gpu_forward_ab.py $1 ${EXTRA} >> $AB 2>> gpurun_out/r02_ab5.err; tail -1 $AB | cut -c1-420; }
EXTRA="--census --nb=28" run det28 DBIR_DETERMINISTIC=1
EXTRA="--census --nb=28" run det28_geglu256 DBIR_DETERMINISTIC=1 DBIR_GEGLU_TILE=256
EXTRA="--census --nb=28" run det28_ragged DBIR_DETERMINISTIC=1 DBIR_GEMM_RAGGED_WIDE=1
EXTRA="--nb=28" run det28_ptmem DBIR_DETERMINISTIC=1 DBIR_LIB_TAG=ptmem
EXTRA="--nb=14" run det14 DBIR_DETERMINISTIC=1
EXTRA="--nb=14" run det14_geglu256 DBIR_DETERMINISTIC=1 DBIR_GEGLU_TILE=256
EXTRA="" run nb2_geglu256 DBIR_GEGLU_TILE=256
EXTRA="--nb=4" run nb4 X=1
EXTRA="--nb=8" run nb8 X=1
